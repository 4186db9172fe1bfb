// host_batch.cuh -- batched gemm front end
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_gemm.cuh"
#include "gemm_grouped.cuh"
namespace b200 {

// ---- batched gemm --------------------------------------------------------
// ?gemm_batch_ / cblas_?gemm_batch (frame/compat/extra/bla_gemm_batch.c:44-131): group i holds group_size[i]
// independent problems with the same shape, transposition and scalars; the reference loops over them calling
// bli_?gemm_ex one after the other.  Here the problems of a batch whose operands are device resident (or pinned)
// are issued round-robin on a pool of streams, so that small problems, which cannot fill 148 SMs one at a time,
// run side by side; problems with pageable host operands take the ordinary (synchronous, staged) path.
// SMALL device-resident problems (m*n*k <= batch_grouped_max, default 128^3) do not get a launch each: they are collected
// into one array of GroupProb records and served by ONE launch of gemm_grouped_kernel (gemm_grouped.cuh) on the caller's
// stream, next to the pool (different problems write different C, so the two need no ordering between them).

template <typename T>
static int gemm_batch_front( int group_count, const int* group_size, const int* transa, const int* transb,
                             const int64_t* m, const int64_t* n, const int64_t* k, const T* alpha,
                             const T* const* a, const int64_t* rs_a, const int64_t* cs_a,
                             const T* const* b, const int64_t* rs_b, const int64_t* cs_b,
                             const T* beta, T* const* c, const int64_t* rs_c, const int64_t* cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( group_count < 0 ) return fail( "b200_gemm_batch: negative group count" );
	if ( group_count == 0 ) return kSuccess;
	if ( !group_size || !transa || !transb || !m || !n || !k || !alpha || !beta || !a || !b || !c ||
	     !rs_a || !cs_a || !rs_b || !cs_b || !rs_c || !cs_c ) return fail( "b200_gemm_batch: NULL argument array" );
	Context& cx = ctx();
	cudaStream_t st = cur_stream();
	std::lock_guard<std::mutex> lock( cx.batch_mu );          // one batch at a time owns the stream pool
	int rc = kSuccess;
	B200_CUDA( cudaEventRecord( cx.batch_fork, st ) );
	for ( int s = 0; s < Context::kBatchStreams; ++s ) B200_CUDA( cudaStreamWaitEvent( cx.batch_streams[s], cx.batch_fork, 0 ) );
	int64_t idx = 0; int next = 0;
	std::vector<GroupProb<T>> small;                          // problems for the grouped kernel
	int64_t small_tiles = 0;
	for ( int g = 0; g < group_count && rc == kSuccess; ++g )
	{
		if ( group_size[g] < 0 || m[g] < 0 || n[g] < 0 || k[g] < 0 ) { rc = fail( "b200_gemm_batch: negative size in group %d", g ); break; }
		int64_t ra = rs_a[g], ca = cs_a[g], rb = rs_b[g], cb = cs_b[g];
		if ( transa[g] & B200_TRANSPOSE ) std::swap( ra, ca );
		if ( transb[g] & B200_TRANSPOSE ) std::swap( rb, cb );
		const bool conja = Elem<T>::cplx && ( transa[g] & B200_CONJ_NO_TRANSPOSE ), conjb = Elem<T>::cplx && ( transb[g] & B200_CONJ_NO_TRANSPOSE );
		const bool need_ab = ( k[g] > 0 && !Scalar<T>::is_zero( alpha[g] ) );
		for ( int j = 0; j < group_size[g] && rc == kSuccess; ++j, ++idx )
		{
			if ( m[g] == 0 || n[g] == 0 ) continue;
			const bool on_device = classify( c[idx] ) == MemKind::Device &&
			                       ( !need_ab || ( classify( a[idx] ) == MemKind::Device && classify( b[idx] ) == MemKind::Device ) );
			const bool elem_aligned = ( (uintptr_t)a[idx] % sizeof(T) == 0 ) && ( (uintptr_t)b[idx] % sizeof(T) == 0 ) && ( (uintptr_t)c[idx] % sizeof(T) == 0 );
			if ( on_device && cx.batch_grouped && elem_aligned && (double)m[g] * (double)n[g] * (double)std::max<int64_t>( k[g], 1 ) <= (double)cx.batch_grouped_max &&
			     small_tiles < ( 1ll << 30 ) )
			{
				GroupProb<T> p;
				p.a = a[idx]; p.b = b[idx]; p.c = c[idx];
				p.rs_a = ra; p.cs_a = ca; p.rs_b = rb; p.cs_b = cb; p.rs_c = rs_c[g]; p.cs_c = cs_c[g];
				p.alpha = alpha[g]; p.beta = beta[g];
				p.m = (int)m[g]; p.n = (int)n[g]; p.k = need_ab ? (int)k[g] : 0;      // alpha == 0: C := beta*C, A and B are not read
				p.conja = conja; p.conjb = conjb; p.beta_is_zero = Scalar<T>::is_zero( beta[g] ) ? 1 : 0;
				p.tile0 = (int)small_tiles; p.tiles_n = (int)( ( n[g] + kGroupTile - 1 ) / kGroupTile );
				small_tiles += (int64_t)( ( m[g] + kGroupTile - 1 ) / kGroupTile ) * p.tiles_n;
				small.push_back( p );
			}
			else if ( on_device )
			{
				cudaStream_t bs = cx.batch_streams[next]; next = ( next + 1 ) % Context::kBatchStreams;
				rc = gemm_dev<T>( conja, conjb, m[g], n[g], k[g], alpha[g], a[idx], ra, ca, b[idx], rb, cb, beta[g], c[idx], rs_c[g], cs_c[g], bs );
			}
			else
				rc = gemm_front<T>( transa[g], transb[g], m[g], n[g], k[g], alpha + g, a[idx], rs_a[g], cs_a[g], b[idx], rs_b[g], cs_b[g],
				                    beta + g, c[idx], rs_c[g], cs_c[g] );
		}
	}
	if ( rc == kSuccess && !small.empty() )
	{
		// the records go up through a pinned buffer of the context (the batch lock is held; the buffer is reused once the
		// previous batch's upload has completed), so the call stays asynchronous
		void* dprobs = nullptr;
		const size_t bytes = small.size() * sizeof( GroupProb<T> );
		if ( cx.batch_desc_done ) cudaEventSynchronize( cx.batch_desc_done );
		else B200_CUDA( cudaEventCreateWithFlags( &cx.batch_desc_done, cudaEventDisableTiming ) );
		if ( bytes > cx.batch_desc_bytes )
		{
			if ( cx.batch_desc ) cudaFreeHost( cx.batch_desc );
			cx.batch_desc = nullptr; cx.batch_desc_bytes = 0;
			B200_CUDA( cudaMallocHost( &cx.batch_desc, 2 * bytes ) );
			cx.batch_desc_bytes = 2 * bytes;
		}
		memcpy( cx.batch_desc, small.data(), bytes );
		if ( dev_alloc( &dprobs, bytes, st ) != kSuccess ) rc = kFailure;
		else
		{
			if ( cudaMemcpyAsync( dprobs, cx.batch_desc, bytes, cudaMemcpyHostToDevice, st ) != cudaSuccess ) rc = fail( "b200_gemm_batch: descriptor upload failed" );
			else
			{
				cudaEventRecord( cx.batch_desc_done, st );
				static const std::string kname = kfmt( "gemm_grouped_kernel<%s>", tname<T>() );
				const int grid = (int)std::min<int64_t>( small_tiles, (int64_t)cx.num_sms * 8 );
				gemm_grouped_kernel<T><<<grid, 256, 0, st>>>( (const GroupProb<T>*)dprobs, (int)small.size(), (int)small_tiles );
				if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm_batch: grouped launch failed" );
				else note_launch( kname.c_str() );
			}
			dev_free( dprobs, st );
		}
	}
	// join: the caller's stream continues after every pool stream has drained
	for ( int s = 0; s < Context::kBatchStreams; ++s )
	{
		cudaEventRecord( cx.batch_join[s], cx.batch_streams[s] );
		cudaStreamWaitEvent( st, cx.batch_join[s], 0 );
	}
	return rc;
}

} // namespace b200
