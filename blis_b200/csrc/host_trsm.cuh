// host_trsm.cuh -- trsm: recursive blocked solve on top of gemm_dev + the block-solve kernel
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_gemm.cuh"
#include "trsm.cuh"
#include "trsm_panel.cuh"
#include <functional>
namespace b200 {

// ---- trsm -----------------------------------------------------------------------------
template <typename T> struct TrsmBlk;
template <> struct TrsmBlk<float>   { static constexpr int NB = 64, CN = 64; };
template <> struct TrsmBlk<double>  { static constexpr int NB = 64, CN = 64; };
template <> struct TrsmBlk<float2>  { static constexpr int NB = 32, CN = 64; };
template <> struct TrsmBlk<double2> { static constexpr int NB = 32, CN = 64; };

template <typename T>
struct TrsmPlan
{
	const T* A; int64_t rs_a, cs_a;      // effective triangular matrix (trans folded into strides)
	T*       B; int64_t rs_b, cs_b;
	int64_t  n;
	bool     upper, unit, conj;
	cudaStream_t st;
	// A arriving from the host while the solve runs (trsm_host_rowpipe): one event per launch of the recursion, in the
	// recursion's own order; nullptr: A is resident
	const std::vector<cudaEvent_t>* a_ready = nullptr;
	size_t*  a_next = nullptr;
	// rows [i0, i0+mb) of X are final when their sub-solve returns: told once per subtree of at most notify_rows rows
	// (trsm_host_rowpipe sends them home while the rest of the solve runs); empty: nobody listens
	std::function<void( int64_t, int64_t )> on_final;
	int64_t  notify_rows = 0;
};
template <typename T>
static void trsm_wait_a( const TrsmPlan<T>& p )
{
	if ( p.a_ready && *p.a_next < p.a_ready->size() ) cudaStreamWaitEvent( p.st, ( *p.a_ready )[( *p.a_next )++], 0 );
}

template <typename T>
static int trsm_base( const TrsmPlan<T>& p, int64_t i0, int mb, T alpha )
{
	constexpr int NB = TrsmBlk<T>::NB, CN = TrsmBlk<T>::CN;
	TrsmBaseArgs<T> a;
	a.A = p.A + i0 * ( p.rs_a + p.cs_a ); a.rs_a = p.rs_a; a.cs_a = p.cs_a;
	a.B = p.B + i0 * p.rs_b;              a.rs_b = p.rs_b; a.cs_b = p.cs_b;
	a.n = p.n; a.mb = mb; a.upper = p.upper; a.unit = p.unit; a.conj = p.conj; a.alpha = alpha;
	constexpr int NT = 256;
	auto kern = trsm_base_kernel<T, NB, CN, NT>;
	constexpr int smem = trsm_base_smem<T, NB, CN>();
	static std::atomic<bool> attr{ false };
	if ( !attr ) { if ( set_smem( kern, smem ) != kSuccess ) return kFailure; attr = true; }
	const int64_t grid = ( p.n + CN - 1 ) / CN;
	kern<<<(unsigned)grid, NT, smem, p.st>>>( a );
	B200_CUDA( cudaGetLastError() );
	note_launch( "trsm_base_kernel" );
	return kSuccess;
}

// Fused diagonal-panel solve (trsm_panel.cuh): rows [i0, i0+mb), mb <= 256, all n columns, one launch.
static int trsm_panel( const TrsmPlan<double>& p, int64_t i0, int mb, double alpha )
{
	TrsmPanelArgs a;
	a.A = p.A + i0 * ( p.rs_a + p.cs_a ); a.rs_a = p.rs_a; a.cs_a = p.cs_a;
	a.B = p.B + i0 * p.rs_b;              a.rs_b = p.rs_b; a.cs_b = p.cs_b;
	a.n = p.n; a.pb = mb; a.upper = p.upper; a.unit = p.unit; a.alpha = alpha;
	const bool ai = ( p.rs_a <= p.cs_a ), bk = ( p.rs_b <= p.cs_b );
	a.a_vec = ( ( ai ? p.rs_a : p.cs_a ) == 1 && ( ( ai ? p.cs_a : p.rs_a ) % 2 ) == 0 && ( (uintptr_t)a.A % 16 ) == 0 ) ? 1 : 0;
	const int64_t grid = ( p.n + TrsmPanelCfg::CN - 1 ) / TrsmPanelCfg::CN;
	auto go = [&]( auto Uc, auto Ac, auto Bc ) -> int
	{
		constexpr bool U = decltype( Uc )::value, AI = decltype( Ac )::value, BK = decltype( Bc )::value;
		auto kern = trsm_panel_kernel<U, AI, BK>;
		static const std::string kname = kfmt( "trsm_panel_kernel<double,256x64,UPPER=%d,AI=%d,BK=%d>", U, AI, BK );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, TrsmPanelCfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<(unsigned)grid, TrsmPanelCfg::NT, TrsmPanelCfg::SMEM_BYTES, p.st>>>( a );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	switch ( ( p.upper ? 4 : 0 ) | ( ai ? 2 : 0 ) | ( bk ? 1 : 0 ) )
	{
		case 0: return go( Ff{}, Ff{}, Ff{} );  case 1: return go( Ff{}, Ff{}, Tt{} );
		case 2: return go( Ff{}, Tt{}, Ff{} );  case 3: return go( Ff{}, Tt{}, Tt{} );
		case 4: return go( Tt{}, Ff{}, Ff{} );  case 5: return go( Tt{}, Ff{}, Tt{} );
		case 6: return go( Tt{}, Tt{}, Ff{} );  default: return go( Tt{}, Tt{}, Tt{} );
	}
}
template <typename T> static int trsm_leaf_rows() { return TrsmBlk<T>::NB; }
template <> int trsm_leaf_rows<double>() { return ctx().trsm_fused ? TrsmPanelCfg::PB : TrsmBlk<double>::NB; }
template <typename T> static int trsm_leaf( const TrsmPlan<T>& p, int64_t i0, int mb, T alpha ) { return trsm_base( p, i0, mb, alpha ); }
template <> int trsm_leaf<double>( const TrsmPlan<double>& p, int64_t i0, int mb, double alpha )
{
	return mb > TrsmBlk<double>::NB ? trsm_panel( p, i0, mb, alpha ) : trsm_base( p, i0, mb, alpha );
}

// Recursive blocked solve of rows [i0, i0+mb): solve one half, rank-k update of
// the other half with the gemm kernel, solve the other half.  alpha is applied
// exactly once to every row (either by the base kernel or as the update's beta,
// as bli_trsm_ex passes alpha as beta: bli_l3_oapi_ex.c:778-789).
template <typename T>
static int trsm_rec( const TrsmPlan<T>& p, int64_t i0, int64_t mb, T alpha )
{
	if ( p.on_final && mb <= p.notify_rows )
	{
		TrsmPlan<T> q = p; q.on_final = nullptr;
		const int rc = trsm_rec( q, i0, mb, alpha );
		if ( rc == kSuccess ) p.on_final( i0, mb );
		return rc;
	}
	const int NB = trsm_leaf_rows<T>();
	if ( mb <= NB ) { trsm_wait_a( p ); return trsm_leaf<T>( p, i0, (int)mb, alpha ); }
	const int64_t nblk = ( mb + NB - 1 ) / NB;
	const int64_t m1 = ( ( nblk + 1 ) / 2 ) * NB, m2 = mb - m1;
	const T one = Scalar<T>::make( 1.0, 0.0 ), mone = Scalar<T>::make( -1.0, 0.0 );
	if ( !p.upper )
	{
		if ( trsm_rec( p, i0, m1, alpha ) != kSuccess ) return kFailure;
		// B2 := alpha*B2 - A21 * X1
		trsm_wait_a( p );
		if ( gemm_dev<T>( p.conj, false, m2, p.n, m1, mone,
		                  p.A + ( i0 + m1 ) * p.rs_a + i0 * p.cs_a, p.rs_a, p.cs_a,
		                  p.B + i0 * p.rs_b, p.rs_b, p.cs_b,
		                  alpha, p.B + ( i0 + m1 ) * p.rs_b, p.rs_b, p.cs_b, p.st ) != kSuccess ) return kFailure;
		return trsm_rec( p, i0 + m1, m2, one );
	}
	else
	{
		// upper: the trailing block is solved first; split so the LAST block is the ragged one's partner
		if ( trsm_rec( p, i0 + m2, m1, alpha ) != kSuccess ) return kFailure;
		// B1 := alpha*B1 - A12 * X2
		trsm_wait_a( p );
		if ( gemm_dev<T>( p.conj, false, m2, p.n, m1, mone,
		                  p.A + i0 * p.rs_a + ( i0 + m2 ) * p.cs_a, p.rs_a, p.cs_a,
		                  p.B + ( i0 + m2 ) * p.rs_b, p.rs_b, p.cs_b,
		                  alpha, p.B + i0 * p.rs_b, p.rs_b, p.cs_b, p.st ) != kSuccess ) return kFailure;
		return trsm_rec( p, i0, m2, one );
	}
}

// ---- host operands: pipeline ------------------------------------------------------------------------------------
// A triangular solve with everything in host memory moves 8(m^2/2 + 2mn) bytes over PCIe; done in sequence (upload B,
// upload A, solve, download X) none of it overlaps the kernels: T1 through the reference's dtrsm_ took 417 ms for 251 ms
// of kernels.  trsm_host_rowpipe (below) hides the transfers behind the solve; its building blocks:
//   * A travels IN THE ORDER THE SOLVE READS IT: trsm_upload_plan walks trsm_rec's own tree over a diagonal block, uploads
//     each diagonal square (<= 1024 rows) and each update block A21 (lower) / A12 (upper) as one 2-D copy on the copy
//     stream and one event PER LAUNCH of the recursion is recorded, in its order; trsm_rec waits for the next event before
//     every launch.  Only the stored triangle (plus the unstored half of the small diagonal squares, ~1.5 %) travels.
//   * X goes home in ROW chunks: the rows of a finished sub-solve are final (trsm_rec tells through on_final).
struct TrsmPiece { int64_t r0, r1, c0, c1; int launches; };      // rows x columns of the effective view; launches of the solve that read it
// The pieces in the order trsm_rec reads them (pure index arithmetic; b200_trsm_upload_plan exposes it to the CPU tests):
// the split is trsm_rec's own; a diagonal block of at most max(leaf_rows, 1024) rows travels as one square and serves all
// 2*ceil(mb/leaf_rows) - 1 launches of its subtree.
static void trsm_upload_plan( int leaf_rows, bool upper, int64_t i0, int64_t mb, std::vector<TrsmPiece>& out )
{
	const int64_t NB = leaf_rows, nblk = ( mb + NB - 1 ) / NB;
	if ( mb <= std::max<int64_t>( NB, 1024 ) ) { out.push_back( { i0, i0 + mb, i0, i0 + mb, (int)( 2 * nblk - 1 ) } ); return; }
	const int64_t m1 = ( ( nblk + 1 ) / 2 ) * NB, m2 = mb - m1;      // the split of trsm_rec
	if ( !upper )
	{
		trsm_upload_plan( leaf_rows, upper, i0, m1, out );
		out.push_back( { i0 + m1, i0 + mb, i0, i0 + m1, 1 } );       // A21 of the update
		trsm_upload_plan( leaf_rows, upper, i0 + m1, m2, out );
	}
	else
	{
		trsm_upload_plan( leaf_rows, upper, i0 + m2, m1, out );
		out.push_back( { i0, i0 + m2, i0 + m2, i0 + mb, 1 } );       // A12 of the update
		trsm_upload_plan( leaf_rows, upper, i0, m2, out );
	}
}
// ---- host operands: row blocks (left-looking top level) ---------------------------------------------------------------
// Run as one recursion over all m rows, the first big update reads ALL of B and three quarters of A after a quarter of the
// flops, so B's upload (and A's big block) stands exposed before it (round 2's first pipeline: 298 ms for 251 ms of
// kernels, with B in one to three column blocks).  Cutting the TOP level into row blocks of rb rows and running it left-looking (the lazy order: block row j first
// receives all its updates in ONE gemm, B_j := alpha*B_j - A[j, 0:j] * X[0:j], then is solved by the recursion) makes every
// byte's deadline proportional to the flops before it:
//   * B travels in row blocks and B_j is not touched before step j,
//   * A travels in row panels: A[j, 0:j] for the update, then the diagonal block in the recursion's own order
//     (trsm_upload_plan of that block), one event per launch as before,
//   * the rows of X_j are final after step j and go home under steps j+1.. (in quarters of a block, so that of the last block
//     only a quarter is exposed).
// Same flops, same kernels (the updates above the block size merge into one gemm per block row; below it the recursion is
// unchanged); only the first block's B_0 and diagonal squares (and the cheap first steps, which compute less than the next
// block's bytes take) stay exposed.  Upper triangular: the mirror image, block rows from the bottom up.
// Rows per block: `opt` > 0 explicit (rounded up to whole leaves), 0 the engine's choice for this shape, < 0 none.
// Returns false when the call is not cut into block rows (one block is no pipeline).
// [B200] (tools/trsm_e2e_sweep.py, profiles/r02b_trsm_e2e_sweep.md) T1 end to end, kernels alone 250.5 ms:
//   rb = 8192 / 4096 / 2048 / 1024 / 512  ->  272.9 / 263.5 / 259.4 / 256.7 / 256.1 ms   (round 2's column blocks: 296.8)
// m = 16384 (n = 8192): best at 512 (69.8 ms; 256: 75.7, 1024: 70.8, column blocks 87.9); m = 8192: 256 and 512 alike (22.3);
// m = 4096: 256 (10.0 ms; 512: 10.9, column blocks 11.8) -- so the engine takes m/32 in whole 256 rows.  Graded schedules
// (small first blocks, large later ones) measured like the uniform schedule of their LARGE size, and the same top level
// with everything resident on the device measured like the plain recursion (250.1-251.2 vs 250.8 ms; worse for narrow B:
// its block-row gemms have too few tiles), so the device-resident solve keeps the recursion.
struct TrsmRowSched { int64_t rb; };
static bool trsm_row_sched( int64_t m, int64_t n, int leaf_rows, long long opt, TrsmRowSched& s, bool pageable = false )
{
	if ( opt < 0 ) return false;
	const Context& c = ctx();
	int64_t rb = opt;
	if ( rb == 0 )
	{
		if ( m < c.trsm_host_rb_min_m || n < 1024 ) return false;
		rb = std::max<int64_t>( 256, ( m / std::max<long long>( 2, c.trsm_host_rb_div ) + 255 ) / 256 * 256 );
		// pageable operands are packed by host threads line by line (a block row of a column-stored matrix = lines of rb
		// elements): at least 16 KiB lines.  [B200] T1 with pageable A and B: sequential 486 ms, rb = 1024 / 2048 / 8192 ->
		// 510-550 / 437 / 431 ms; the packing rate (36 GB/s alone, ~25 GB/s beside the DMA engines) bounds it, not the GPU
		if ( pageable ) rb = std::max<int64_t>( rb, 2048 );
	}
	s.rb = ( rb + leaf_rows - 1 ) / leaf_rows * leaf_rows;
	return s.rb < m;
}
static void trsm_row_blocks( int64_t m, int leaf_rows, const TrsmRowSched& s, bool upper, std::vector<std::pair<int64_t, int64_t>>& blk )
{
	// boundaries are multiples of the leaf size (counted from row 0 for both triangles, so that the blocks stay aligned)
	const int64_t mr = ( m + leaf_rows - 1 ) / leaf_rows * leaf_rows;
	for ( int64_t cum = 0; cum < mr; )                                // in processing order: top down (lower), bottom up (upper)
	{
		int64_t sz = std::min( s.rb, mr - cum );
		if ( mr - cum - sz < s.rb / 2 ) sz = mr - cum;                // no sliver at the end
		if ( !upper ) blk.push_back( { cum, std::min( m, cum + sz ) } );
		else          blk.push_back( { mr - cum - sz, std::min( m, mr - cum ) } );
		cum += sz;
	}
}
// What travels up, in the order of its deadlines (b200_trsm_rowblock_plan exposes it to the CPU tests): per block row the
// rows of B (a piece with c0 = c1 = -1 that no launch waits for: the stream's order covers it), the block A[j, 0:j]
// (A[j, j+1:] for upper) its update gemm reads, then the diagonal block's pieces in the recursion's order.
static void trsm_rowblock_plan( int leaf_rows, bool upper, int64_t m, const TrsmRowSched& s, std::vector<TrsmPiece>& out )
{
	std::vector<std::pair<int64_t, int64_t>> blk;
	trsm_row_blocks( m, leaf_rows, s, upper, blk );
	for ( const auto& [r0, r1] : blk )
	{
		out.push_back( { r0, r1, -1, -1, 0 } );                             // B_j
		if ( !upper && r0 > 0 ) out.push_back( { r0, r1, 0, r0, 1 } );      // A[j, 0:j] of the block row's update
		if ( upper && r1 < m )  out.push_back( { r0, r1, r1, m, 1 } );
		trsm_upload_plan( leaf_rows, upper, r0, r1 - r0, out );
	}
}

// a: effective m x m view (host or device), b: m x n host, pinned, column-stored (rs_b == 1)
template <typename T>
static int trsm_host_rowpipe( int64_t m, int64_t n, T al, const T* a, int64_t rs_a, int64_t cs_a, bool a_host, bool upper, bool unit, bool conj,
                              T* b, int64_t cs_b, bool b_pageable, const TrsmRowSched& sched, cudaStream_t st )
{
	constexpr size_t ES = sizeof(T);
	Context& cx = ctx();
	cudaStream_t s_in = cx.copy_stream, s_out = cx.d2h_stream;
	const int leaf = trsm_leaf_rows<T>();
	std::vector<std::pair<int64_t, int64_t>> blk;
	trsm_row_blocks( m, leaf, sched, upper, blk );
	const int nblk = (int)blk.size();
	void *da = nullptr, *db = nullptr;
	if ( dev_alloc( &db, (size_t)m * n * ES, st ) != kSuccess ) return kFailure;
	if ( a_host && dev_alloc( &da, (size_t)m * m * ES, st ) != kSuccess ) { dev_free( db, st ); return kFailure; }
	std::vector<cudaEvent_t> ev_b( nblk, nullptr ), ev_done, ev_a, ev_own;
	cudaEvent_t ev_alloc = nullptr, ev_out = nullptr;
	cudaEventCreateWithFlags( &ev_alloc, cudaEventDisableTiming ); cudaEventCreateWithFlags( &ev_out, cudaEventDisableTiming );
	cudaEventRecord( ev_alloc, st );
	cudaStreamWaitEvent( s_in, ev_alloc, 0 ); cudaStreamWaitEvent( s_out, ev_alloc, 0 );
	int rc = kSuccess;
	T* adev = const_cast<T*>( a ); int64_t rs_ad = rs_a, cs_ad = cs_a;
	if ( a_host ) { adev = (T*)da; rs_ad = ( rs_a == 1 ? 1 : m ); cs_ad = ( rs_a == 1 ? m : 1 ); }
	// everything that travels up, in the order of its deadlines (trsm_rowblock_plan): per block row B_j, A[j, 0:j], the
	// diagonal block's pieces.  Uploads, kernels and downloads of a block row are queued together, block row by block row:
	// from page-locked memory every copy is asynchronous, so the host runs ahead and the streams' events do the ordering;
	// from PAGEABLE memory (what a legacy dtrsm_ caller passes) a copy occupies this thread while it packs into / unpacks
	// from the pinned ring, so the kernels of block row j are queued before block row j+1 is packed, and the rows of X
	// are unpacked one block row late (`pending`): waiting for X_j right after queueing step j would idle the GPU.
	std::vector<TrsmPiece> plan;
	trsm_rowblock_plan( leaf, upper, m, sched, plan );
	struct Home { int64_t i0, mb; cudaEvent_t e; int j; };
	std::vector<Home> pending;
	auto send_home = [&]( int before_block )
	{
		size_t k = 0;
		for ( ; k < pending.size() && pending[k].j < before_block && rc == kSuccess; ++k )
		{
			cudaStreamWaitEvent( s_out, pending[k].e, 0 );
			rc = stage_block_to_host( b + pending[k].i0, 1, cs_b, (T*)db + pending[k].i0, m, pending[k].mb, n, ES, s_out );
		}
		pending.erase( pending.begin(), pending.begin() + k );
	};
	size_t pi = 0, next = 0;
	const T one = Scalar<T>::make( 1.0, 0.0 ), mone = Scalar<T>::make( -1.0, 0.0 );
	for ( int j = 0; j < nblk && rc == kSuccess; ++j )
	{
		const int64_t r0 = blk[j].first, r1 = blk[j].second, rows = r1 - r0;
		// -- up: B_j, then the pieces of A this block row reads
		if ( pi >= plan.size() || plan[pi].c0 >= 0 || plan[pi].r0 != r0 || plan[pi].r1 != r1 ) { rc = fail( "b200_trsm: row-block plan and block rows disagree" ); break; }
		++pi;
		rc = stage_block_to_device( (T*)db + r0, m, b + r0, rows, n, 1, cs_b, ES, s_in );
		if ( rc != kSuccess ) break;
		if ( cudaEventCreateWithFlags( &ev_b[j], cudaEventDisableTiming ) != cudaSuccess ) { rc = fail( "trsm: event creation failed" ); break; }
		cudaEventRecord( ev_b[j], s_in );
		for ( ; pi < plan.size() && plan[pi].c0 >= 0 && rc == kSuccess; ++pi )
		{
			if ( !a_host ) continue;
			const TrsmPiece& q = plan[pi];
			// the device image keeps the host's orientation (column- or row-stored)
			if ( rs_a == 1 ) rc = stage_block_to_device( (T*)da + q.r0 + q.c0 * m, m, a + q.r0 + q.c0 * cs_a, q.r1 - q.r0, q.c1 - q.c0, 1, cs_a, ES, s_in );
			else             rc = stage_block_to_device( (T*)da + q.c0 + q.r0 * m, m, a + q.c0 + q.r0 * rs_a, q.c1 - q.c0, q.r1 - q.r0, 1, rs_a, ES, s_in );
			if ( rc != kSuccess ) break;
			cudaEvent_t e;
			if ( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) != cudaSuccess ) { rc = fail( "trsm: event creation failed" ); break; }
			cudaEventRecord( e, s_in );
			ev_own.push_back( e );
			for ( int l = 0; l < q.launches; ++l ) ev_a.push_back( e );
		}
		if ( rc != kSuccess ) break;
		// -- the step: one update gemm over everything solved so far, then the recursion on the diagonal block
		cudaStreamWaitEvent( st, ev_b[j], 0 );
		TrsmPlan<T> p{ adev, rs_ad, cs_ad, (T*)db, 1, m, n, upper, unit, conj, st };
		if ( a_host ) { p.a_ready = &ev_a; p.a_next = &next; }
		T al_solve = al;
		if ( !upper && r0 > 0 )
		{
			// B_j := alpha*B_j - A[j, 0:j] * X[0:j]
			trsm_wait_a( p );
			if ( gemm_dev<T>( conj, false, rows, n, r0, mone, adev + r0 * rs_ad, rs_ad, cs_ad, (T*)db, 1, m, al, (T*)db + r0, 1, m, st ) != kSuccess ) { rc = kFailure; break; }
			al_solve = one;
		}
		if ( upper && r1 < m )
		{
			trsm_wait_a( p );
			if ( gemm_dev<T>( conj, false, rows, n, m - r1, mone, adev + r0 * rs_ad + r1 * cs_ad, rs_ad, cs_ad, (T*)db + r1, 1, m, al, (T*)db + r0, 1, m, st ) != kSuccess ) { rc = kFailure; break; }
			al_solve = one;
		}
		// -- home: the rows of X_j are final as their sub-solves finish (in quarters of a large block, so that of the last
		// block row only a quarter is exposed)
		p.notify_rows = std::max<int64_t>( 1024, ( sched.rb / 4 + 255 ) / 256 * 256 );
		p.on_final = [&]( int64_t i0, int64_t mb )
		{
			if ( rc != kSuccess ) return;
			cudaEvent_t e;
			if ( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) != cudaSuccess ) { rc = fail( "trsm: event creation failed" ); return; }
			ev_done.push_back( e );
			cudaEventRecord( e, st );
			pending.push_back( { i0, mb, e, j } );
			if ( !b_pageable ) send_home( j + 1 );
		};
		const int rs = trsm_rec( p, r0, rows, al_solve );
		if ( rc == kSuccess ) rc = rs;
		if ( rc == kSuccess ) send_home( j );                 // (pageable B: what block row j-1 left)
	}
	if ( rc == kSuccess ) send_home( nblk );
	if ( rc == kSuccess && pi != plan.size() ) rc = fail( "b200_trsm: row-block plan and block rows disagree" );
	if ( rc == kSuccess && a_host && next != ev_a.size() ) rc = fail( "b200_trsm: row-block upload plan (%zu events) and solve (%zu launches) disagree", ev_a.size(), next );
	cudaEventRecord( ev_out, s_out );
	cudaStreamWaitEvent( st, ev_out, 0 );
	if ( cudaStreamSynchronize( st ) != cudaSuccess && rc == kSuccess ) rc = fail( "b200_trsm: stream sync failed" );
	if ( cudaStreamSynchronize( s_in ) != cudaSuccess && rc == kSuccess ) rc = fail( "b200_trsm: copy stream sync failed" );
	for ( auto e : ev_b ) if ( e ) cudaEventDestroy( e );
	for ( auto e : ev_done ) cudaEventDestroy( e );
	for ( auto e : ev_own ) cudaEventDestroy( e );
	cudaEventDestroy( ev_alloc ); cudaEventDestroy( ev_out );
	dev_free( da, st ); dev_free( db, st );
	return rc;
}

template <typename T>
static int trsm_front( int side, int uplo, int transa, int diag, int64_t m, int64_t n,
                       const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                       T* b, int64_t rs_b, int64_t cs_b )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || n < 0 ) return fail( "b200_trsm: negative dimension" );
	if ( !alpha ) return fail( "b200_trsm: alpha must be a non-NULL host pointer" );
	if ( uplo != B200_LOWER && uplo != B200_UPPER ) return fail( "b200_trsm: uplo must be BLIS_LOWER or BLIS_UPPER" );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	constexpr size_t ES = sizeof(T);
	const T al = *alpha;

	// right side: X * op(A) = alpha*B  <=>  op(A)^T * X^T = alpha * B^T   (bli_l3_oapi_ex.c:748-759)
	if ( side == B200_RIGHT )
	{
		std::swap( m, n ); std::swap( rs_b, cs_b );
		transa ^= B200_TRANSPOSE;
	}
	bool upper = ( uplo == B200_UPPER );
	if ( transa & B200_TRANSPOSE ) { std::swap( rs_a, cs_a ); upper = !upper; }
	const bool conj = Elem<T>::cplx && ( transa & B200_CONJ_NO_TRANSPOSE );
	// now: A is m x m (effective uplo `upper`), B is m x n

	void *da = nullptr, *db = nullptr;
	int rc = kSuccess;
	const MemKind kind_b = classify( b );
	const bool b_host = ( kind_b != MemKind::Device );
	const bool zero_alpha = Scalar<T>::is_zero( al );
	if ( ctx().trsm_host_pipe && !zero_alpha && b_host && rs_b == 1 && cs_b >= m && m >= 4096 && n >= 1024 )
	{
		// host B (and possibly A), column-stored, large: transfers run under the solve (trsm_host_rowpipe); page-locked
		// operands are copied by the DMA engines directly, pageable ones through the pinned ring
		const MemKind kind_a = classify( a );
		const bool a_lines = ( rs_a == 1 && cs_a >= m ) || ( cs_a == 1 && rs_a >= m );
		if ( kind_a == MemKind::Device || a_lines )
		{
			TrsmRowSched sched;
			const bool pageable = ( kind_b == MemKind::HostPageable || kind_a == MemKind::HostPageable );
			if ( trsm_row_sched( m, n, trsm_leaf_rows<T>(), ctx().trsm_host_rb, sched, pageable ) )
				return trsm_host_rowpipe<T>( m, n, al, a, rs_a, cs_a, kind_a != MemKind::Device, upper, diag == B200_UNIT_DIAG, conj, b, cs_b,
				                             kind_b == MemKind::HostPageable, sched, st );
		}
	}
	T* bdev = b; int64_t rs_bd = rs_b, cs_bd = cs_b;
	if ( b_host )
	{
		if ( dev_alloc( &db, (size_t)m * n * ES, st ) != kSuccess ) return kFailure;
		if ( !zero_alpha ) rc = stage_to_device( db, b, m, n, rs_b, cs_b, ES, st );
		bdev = (T*)db; rs_bd = 1; cs_bd = m;
	}
	if ( zero_alpha )
	{
		// bli_l3_return_early_if_trivial( alpha, a, b, &BLIS_ZERO, b ):  B := 0
		if ( rc == kSuccess ) rc = scal2d( bdev, rs_bd, cs_bd, m, n, al, st );
	}
	else
	{
		if ( rc == kSuccess && classify( a ) != MemKind::Device )
		{
			if ( dev_alloc( &da, (size_t)m * m * ES, st ) != kSuccess ) rc = kFailure;
			else rc = stage_tri_to_device( da, a, m, rs_a, cs_a, upper, ES, st );     // (rs_a, cs_a, upper): the effective view
			a = (const T*)da; rs_a = 1; cs_a = m;
		}
		if ( rc == kSuccess )
		{
			TrsmPlan<T> p{ a, rs_a, cs_a, bdev, rs_bd, cs_bd, n, upper, diag == B200_UNIT_DIAG, conj, st };
			rc = trsm_rec( p, 0, m, al );
		}
	}
	if ( rc == kSuccess && b_host )
	{
		rc = stage_to_host( b, rs_b, cs_b, db, m, n, ES, st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "b200_trsm: stream sync failed" );
	}
	dev_free( da, st ); dev_free( db, st );
	return rc;
}

} // namespace b200
