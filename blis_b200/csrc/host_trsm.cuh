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
	// A arriving from the host while the solve runs (trsm_host_pipe): one event per launch of the recursion, in the
	// recursion's own order (trsm_upload builds the list); nullptr: A is resident
	const std::vector<cudaEvent_t>* a_ready = nullptr;
	size_t*  a_next = nullptr;
	// rows [i0, i0+mb) of X are final when their sub-solve returns: told once per subtree of at most notify_rows rows
	// (trsm_host_pipeline sends them home while the rest of the solve runs); empty: nobody listens
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
// upload A, solve, download X) none of it overlaps the kernels: T1 through the reference's dtrsm_ took 417 ms for 253 ms
// of kernels.  Two things make the transfers disappear behind the solve:
//   * A travels IN THE ORDER THE RECURSION READS IT.  trsm_rec touches A11 (recursively), then the block A21 (lower) /
//     A12 (upper) of the update, then A22 -- and spends its time in the same proportion (a quarter, a half, a quarter of
//     the flops for a quarter, a half, a quarter of the triangle's bytes).  trsm_upload_plan walks the same tree, uploads
//     each diagonal block (<= 1024 rows, as a square) and each update block as one 2-D copy on the copy stream and records
//     one event PER LAUNCH of the solve's recursion, in its order; trsm_rec waits for the next event before every launch.
//     Only the stored triangle (plus the unstored half of the small diagonal squares, ~1.5 %) travels.
//   * B travels in column blocks (the reference's own parallel dimension, bli_trsm_cntl.c:446-451: columns are
//     independent): block 0 goes up first and is solved while A streams in; block j+1 goes up and block j-1 comes down
//     (d2h stream) under the solve of block j.  Every block repeats the latency-bound diagonal panels (5.8 ms at T1), so
//     there are only a few blocks (trsm_host_pipe = their maximal number; one for tall systems).
//   * X goes home in ROW chunks: the rows of a finished sub-solve are final (trsm_rec tells, eighths of m), so only the
//     last eighth's download is exposed.
// [B200] T1 through dtrsm_: 417 ms (sequential) -> 298 ms = 29.5 TFLOP/s end to end (kernels alone: 251 ms).  What is left is
// structural: the top-level update needs all of B and three quarters of A (99 ms of PCIe time) after a quarter of the flops.
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
template <typename T>
static int trsm_upload( T* da, int64_t m, const T* a, int64_t rs_a, int64_t cs_a, bool upper, cudaStream_t s_in, std::vector<cudaEvent_t>& ev )
{
	constexpr size_t ES = sizeof(T);
	std::vector<TrsmPiece> plan;
	trsm_upload_plan( trsm_leaf_rows<T>(), upper, 0, m, plan );
	for ( const TrsmPiece& q : plan )
	{
		// the device image keeps the host's orientation (column- or row-stored)
		int rc;
		if ( rs_a == 1 ) rc = stage_block_to_device( da + q.r0 + q.c0 * m, m, a + q.r0 + q.c0 * cs_a, q.r1 - q.r0, q.c1 - q.c0, 1, cs_a, ES, s_in );
		else             rc = stage_block_to_device( da + q.c0 + q.r0 * m, m, a + q.c0 + q.r0 * rs_a, q.c1 - q.c0, q.r1 - q.r0, 1, rs_a, ES, s_in );
		if ( rc != kSuccess ) return rc;
		cudaEvent_t e;
		if ( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) != cudaSuccess ) return fail( "trsm: event creation failed" );
		cudaEventRecord( e, s_in );
		for ( int l = 0; l < q.launches; ++l ) ev.push_back( e );    // the same event serves every launch inside this piece
	}
	return kSuccess;
}

// a: effective m x m view (host or device), b: m x n host, pinned, column-stored (rs_b == 1)
template <typename T>
static int trsm_host_pipeline( int64_t m, int64_t n, T al, const T* a, int64_t rs_a, int64_t cs_a, bool a_host, bool upper, bool unit, bool conj,
                               T* b, int64_t cs_b, cudaStream_t st )
{
	constexpr size_t ES = sizeof(T);
	Context& cx = ctx();
	cudaStream_t s_in = cx.copy_stream, s_out = cx.d2h_stream;
	// column blocks: the first one is only as wide as A's journey lasts (its solve is gated by A anyway), the others share the rest
	// Every block repeats the width-independent part of the solve (diagonal panels and the small updates: ~9 % of T1), which
	// grows with m, while what a block saves is B's journey, which does not.  [B200] T1 (m = 32768): 1 / 2 / 3 blocks ->
	// 301 / 317 / 325 ms, so tall systems keep one block (A streaming and X leaving in row chunks do the overlapping there).
	int nblk = (int)std::min<int64_t>( std::max( 1, cx.trsm_host_pipe ), std::max<int64_t>( 1, n / 1024 ) );
	nblk = std::min( nblk, m >= 24576 ? 1 : ( m >= 12288 ? 2 : 3 ) );
	std::vector<int64_t> col{ 0 };
	if ( nblk > 1 )
	{
		const int64_t first = std::max<int64_t>( 512, ( (int64_t)( 0.7 * (double)n / nblk ) + 127 ) / 128 * 128 );
		col.push_back( std::min( n, first ) );
		const int64_t rest = ( ( n - col.back() + nblk - 2 ) / ( nblk - 1 ) + 127 ) / 128 * 128;
		while ( col.back() < n ) col.push_back( std::min( n, col.back() + rest ) );
	}
	else col.push_back( n );
	nblk = (int)col.size() - 1;
	void *da = nullptr, *db = nullptr;
	if ( dev_alloc( &db, (size_t)m * n * ES, st ) != kSuccess ) return kFailure;
	if ( a_host && dev_alloc( &da, (size_t)m * m * ES, st ) != kSuccess ) { dev_free( db, st ); return kFailure; }
	std::vector<cudaEvent_t> ev_b( nblk ), ev_done, ev_a;
	cudaEvent_t ev_alloc = nullptr, ev_out = nullptr;
	for ( auto& e : ev_b ) cudaEventCreateWithFlags( &e, cudaEventDisableTiming );
	cudaEventCreateWithFlags( &ev_alloc, cudaEventDisableTiming ); cudaEventCreateWithFlags( &ev_out, cudaEventDisableTiming );
	cudaEventRecord( ev_alloc, st );
	cudaStreamWaitEvent( s_in, ev_alloc, 0 ); cudaStreamWaitEvent( s_out, ev_alloc, 0 );
	int rc = kSuccess;
	auto send_b = [&]( int j ) -> int
	{
		const int64_t j0 = col[j], w = col[j + 1] - j0;
		const int r = stage_block_to_device( (T*)db + j0 * m, m, b + j0 * cs_b, m, w, 1, cs_b, ES, s_in );
		cudaEventRecord( ev_b[j], s_in );
		return r;
	};
	rc = send_b( 0 );
	T* adev = const_cast<T*>( a ); int64_t rs_ad = rs_a, cs_ad = cs_a;
	if ( rc == kSuccess && a_host )
	{
		rc = trsm_upload<T>( (T*)da, m, a, rs_a, cs_a, upper, s_in, ev_a );
		adev = (T*)da; rs_ad = ( rs_a == 1 ? 1 : m ); cs_ad = ( rs_a == 1 ? m : 1 );
	}
	for ( int j = 1; j < nblk && rc == kSuccess; ++j ) rc = send_b( j );
	for ( int j = 0; j < nblk && rc == kSuccess; ++j )
	{
		const int64_t j0 = col[j], w = col[j + 1] - j0;
		cudaStreamWaitEvent( st, ev_b[j], 0 );
		size_t next = 0;
		TrsmPlan<T> p{ adev, rs_ad, cs_ad, (T*)db + j0 * m, 1, m, w, upper, unit, conj, st };
		if ( a_host && j == 0 ) { p.a_ready = &ev_a; p.a_next = &next; }     // later blocks run after block 0: all of A is there
		// rows of X go home as soon as their sub-solve is done (eighths of the block), under the rest of the solve
		p.notify_rows = std::max<int64_t>( 1024, ( m / 8 + 255 ) / 256 * 256 );
		p.on_final = [&]( int64_t i0, int64_t mb )
		{
			if ( rc != kSuccess ) return;
			cudaEvent_t e;
			if ( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) != cudaSuccess ) { rc = fail( "trsm: event creation failed" ); return; }
			ev_done.push_back( e );
			cudaEventRecord( e, st );
			cudaStreamWaitEvent( s_out, e, 0 );
			rc = stage_block_to_host( b + j0 * cs_b + i0, 1, cs_b, (T*)db + j0 * m + i0, m, mb, w, ES, s_out );
		};
		const int rs = trsm_rec( p, 0, m, al );
		if ( rc == kSuccess ) rc = rs;
		if ( rc == kSuccess && p.a_ready && next != ev_a.size() ) rc = fail( "b200_trsm: upload plan (%zu events) and solve recursion (%zu launches) disagree", ev_a.size(), next );
	}
	cudaEventRecord( ev_out, s_out );
	cudaStreamWaitEvent( st, ev_out, 0 );
	if ( cudaStreamSynchronize( st ) != cudaSuccess && rc == kSuccess ) rc = fail( "b200_trsm: stream sync failed" );
	if ( cudaStreamSynchronize( s_in ) != cudaSuccess && rc == kSuccess ) rc = fail( "b200_trsm: copy stream sync failed" );
	for ( auto e : ev_b ) cudaEventDestroy( e );
	for ( auto e : ev_done ) cudaEventDestroy( e );
	{ cudaEvent_t last = nullptr; for ( auto e : ev_a ) { if ( e != last ) cudaEventDestroy( e ); last = e; } }
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
	if ( ctx().trsm_host_pipe && !zero_alpha && kind_b == MemKind::HostPinned && rs_b == 1 && cs_b >= m && m >= 4096 && n >= 1024 )
	{
		// pinned host B (and possibly A), large: transfers run under the solve (trsm_host_pipeline)
		const MemKind kind_a = classify( a );
		const bool a_lines = ( rs_a == 1 && cs_a >= m ) || ( cs_a == 1 && rs_a >= m );
		if ( kind_a == MemKind::Device || ( kind_a == MemKind::HostPinned && a_lines ) )
			return trsm_host_pipeline<T>( m, n, al, a, rs_a, cs_a, kind_a != MemKind::Device, upper, diag == B200_UNIT_DIAG, conj, b, cs_b, st );
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
