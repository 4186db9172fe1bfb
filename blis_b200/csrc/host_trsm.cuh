// host_trsm.cuh -- trsm: recursive blocked solve on top of gemm_dev + the block-solve kernel
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_gemm.cuh"
#include "trsm.cuh"
#include "trsm_panel.cuh"
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
};

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
	static bool attr = false;
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
		static bool attr = false;
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
	const int NB = trsm_leaf_rows<T>();
	if ( mb <= NB ) return trsm_leaf<T>( p, i0, (int)mb, alpha );
	const int64_t nblk = ( mb + NB - 1 ) / NB;
	const int64_t m1 = ( ( nblk + 1 ) / 2 ) * NB, m2 = mb - m1;
	const T one = Scalar<T>::make( 1.0, 0.0 ), mone = Scalar<T>::make( -1.0, 0.0 );
	if ( !p.upper )
	{
		if ( trsm_rec( p, i0, m1, alpha ) != kSuccess ) return kFailure;
		// B2 := alpha*B2 - A21 * X1
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
		if ( gemm_dev<T>( p.conj, false, m2, p.n, m1, mone,
		                  p.A + i0 * p.rs_a + ( i0 + m2 ) * p.cs_a, p.rs_a, p.cs_a,
		                  p.B + ( i0 + m2 ) * p.rs_b, p.rs_b, p.cs_b,
		                  alpha, p.B + i0 * p.rs_b, p.rs_b, p.cs_b, p.st ) != kSuccess ) return kFailure;
		return trsm_rec( p, i0, m2, one );
	}
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
	const bool b_host = ( classify( b ) != MemKind::Device );
	const bool zero_alpha = Scalar<T>::is_zero( al );
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
