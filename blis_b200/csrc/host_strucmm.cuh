// host_strucmm.cuh -- hemm, symm, trmm3, trmm front end
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_gemm.cuh"
namespace b200 {

// ---- hemm, symm, trmm, trmm3 -----------------------------------------------------------------
// bli_hemm_ex / bli_symm_ex / bli_trmm3_ex / bli_trmm_ex (frame/3/bli_l3_oapi_ex.c:349-689): the gemm control tree
// with a structured A.  The reference resolves the structure while PACKING (bli_packm_struc_cxk.c:146-301: the
// unstored side of a Hermitian/symmetric matrix is read from its mirror image, conjugated for Hermitian; the unstored
// side of a triangular matrix is packed as explicit zeros; ref_kernels/1m/bli_packm_cxc_diag_ref.c:36-98: a unit
// diagonal is packed as one, a Hermitian diagonal loses its imaginary part) and then runs gemm-shaped macrokernels
// (trmm ones skip the zero k range).  Here the structure is resolved ONCE into a dense m x m device matrix
// (O(m^2) traffic against O(m^2 n) flops) and the product is the gemm kernel, trimmed in k for trmm.

enum { kStrucTri = 0, kStrucSym = 1, kStrucHerm = 2 };

template <typename R, int NC>
__global__ void densify_kernel( R* dst, int64_t ldd, const R* src, int64_t rs, int64_t cs,
                                int64_t m, int struc, int lower, int unit )
{
	const int64_t total = m * m;
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		const int64_t i = e % m, j = e / m;
		const bool stored = lower ? ( i >= j ) : ( i <= j );
		R re = (R)0, im = (R)0;
		if ( i == j )
		{
			if ( struc == kStrucTri && unit ) re = (R)1;
			else
			{
				const R* p = src + ( i * rs + j * cs ) * NC;
				re = p[0];
				if ( NC == 2 && struc != kStrucHerm ) im = p[NC - 1];
			}
		}
		else if ( stored )
		{
			const R* p = src + ( i * rs + j * cs ) * NC;
			re = p[0]; if ( NC == 2 ) im = p[NC - 1];
		}
		else if ( struc != kStrucTri )
		{
			const R* p = src + ( j * rs + i * cs ) * NC;         // mirror image
			re = p[0]; if ( NC == 2 ) im = ( struc == kStrucHerm ) ? -p[NC - 1] : p[NC - 1];
		}
		R* d = dst + ( i + j * ldd ) * NC;
		if ( src == dst && stored && i != j ) continue;       // in place (staged copy): stored elements stay
		d[0] = re; if ( NC == 2 ) d[NC - 1] = im;
	}
}

// Dense, structure-resolved copy of the ma x ma matrix A (host or device) in `*da` (column-major, ld = ma).
template <typename T>
static int densify_operand( void** da, const T* a, int64_t rs_a, int64_t cs_a, int64_t ma, int struc, int uplo, bool unit, cudaStream_t st )
{
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	*da = nullptr;
	if ( dev_alloc( da, (size_t)ma * ma * sizeof(T), st ) != kSuccess ) return kFailure;
	const T* src = a; int64_t rs = rs_a, cs = cs_a;
	if ( classify( a ) != MemKind::Device )
	{
		// the whole array is staged (the unstored triangle may hold anything; it is never used) and resolved in place
		if ( stage_to_device( *da, a, ma, ma, rs_a, cs_a, sizeof(T), st ) != kSuccess ) return kFailure;
		src = (const T*)*da; rs = 1; cs = ma;
	}
	const int64_t total = ma * ma;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	densify_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)*da, ma, (const R*)src, rs, cs, ma, struc, uplo == B200_LOWER ? 1 : 0, unit ? 1 : 0 );
	B200_CUDA( cudaGetLastError() );
	note_launch( "densify_kernel" );
	return kSuccess;
}

// op: 0 hemm, 1 symm, 2 trmm3, 3 trmm (C == B, beta ignored)
template <typename T>
static int struc_mm_front( int op, int side, int uploa, int transa, int diaga, int transb, int64_t m, int64_t n,
                           const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                           const T* b, int64_t rs_b, int64_t cs_b,
                           const T* beta, T* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || n < 0 ) return fail( "%s: negative dimension", name );
	if ( !alpha || ( op != 3 && !beta ) ) return fail( "%s: alpha/beta must be non-NULL host pointers", name );
	if ( uploa != B200_LOWER && uploa != B200_UPPER ) return fail( "%s: uplo must be BLIS_LOWER or BLIS_UPPER", name );
	if ( side != B200_LEFT && side != B200_RIGHT ) return fail( "%s: side must be BLIS_LEFT or BLIS_RIGHT", name );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	const int64_t ma = ( side == B200_LEFT ) ? m : n;
	const T zero = Scalar<T>::make( 0.0, 0.0 );
	const T al = *alpha, be = ( op == 3 ) ? zero : *beta;
	const int struc = ( op == 0 ) ? kStrucHerm : ( op == 1 ) ? kStrucSym : kStrucTri;

	void *da = nullptr, *dt = nullptr;
	int rc = kSuccess;
	if ( !Scalar<T>::is_zero( al ) )
		rc = densify_operand<T>( &da, a, rs_a, cs_a, ma, struc, uploa, diaga == B200_UNIT_DIAG, st );
	const T* ad = (const T*)da;

	const T* bsrc = b; int64_t rs_bs = rs_b, cs_bs = cs_b; int transb_use = transb;
	if ( op == 3 )
	{
		// trmm is in place: B := alpha * transa(A) * B.  The product reads a copy of B (bli_trmm_ex aliases C = B and
		// relies on the macrokernel's loop order; a copy costs O(mn) against O(m^2 n)).
		transb_use = B200_NO_TRANSPOSE;
		if ( rc == kSuccess && !Scalar<T>::is_zero( al ) )
		{
			if ( dev_alloc( &dt, (size_t)m * n * sizeof(T), st ) != kSuccess ) rc = kFailure;
			else if ( classify( b ) != MemKind::Device ) rc = stage_to_device( dt, b, m, n, rs_b, cs_b, sizeof(T), st );
			else rc = copy2d( (T*)dt, (int64_t)1, m, b, rs_b, cs_b, m, n, st );
			bsrc = (const T*)dt; rs_bs = 1; cs_bs = m;
		}
	}
	if ( rc == kSuccess )
	{
		// effective triangle of transa(A): transposition mirrors it
		int tri_operand = 0;
		if ( struc == kStrucTri )
		{
			const bool lower_eff = ( uploa == B200_LOWER ) != ( ( transa & B200_TRANSPOSE ) != 0 );
			tri_operand = ( side == B200_LEFT ? kTriA : kTriB ) | ( lower_eff ? kTriLower : kTriUpper );
		}
		const int ta = ( struc == kStrucTri ) ? transa : ( transa & B200_CONJ_NO_TRANSPOSE );   // hemm/symm: conja only
		if ( side == B200_LEFT )
			rc = gemm_front<T>( ta, transb_use, m, n, m, &al, ad, 1, ma, bsrc, rs_bs, cs_bs, &be, c, rs_c, cs_c, tri_operand );
		else
			rc = gemm_front<T>( transb_use, ta, m, n, n, &al, bsrc, rs_bs, cs_bs, ad, 1, ma, &be, c, rs_c, cs_c, tri_operand );
	}
	dev_free( da, st ); dev_free( dt, st );
	return rc;
}

} // namespace b200
