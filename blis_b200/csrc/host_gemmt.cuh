// host_gemmt.cuh -- gemmt family front end (gemmt, syrk, herk, syr2k, her2k)
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_gemm.cuh"
namespace b200 {

// ---- gemmt family: gemmt, syrk, herk, syr2k, her2k ---------------------------------------
// bli_gemmt_ex / bli_syrk_ex / bli_herk_ex / bli_syr2k_ex / bli_her2k_ex (frame/3/bli_l3_oapi_ex.c:151-346):
// every one of them is one or two gemmt's, C := beta*C + alpha*A*B restricted to the stored triangle of the
// m x m matrix C (macrokernels frame/3/gemmt/bli_gemmt_{l,u}_ker_var2.c); herk/her2k then zero the imaginary
// part of the diagonal (bli_setid).  Here a gemmt is the gemm kernel with a triangular tile schedule.
enum { kOpGemmt = 0, kOpSyrk = 1, kOpHerk = 2, kOpSyr2k = 3, kOpHer2k = 4 };

template <typename R>
__global__ void zero_diag_imag_kernel( R* c, int64_t inc, int64_t m )
{
	for ( int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x )
		c[2 * i * inc + 1] = (R)0;
}

// Device view of a host or device operand: host data is staged into a dense column-major temporary.
template <typename T>
static int operand_to_device( const T*& p, int64_t& rs, int64_t& cs, int64_t m, int64_t n, void** tmp, cudaStream_t st )
{
	*tmp = nullptr;
	if ( m <= 0 || n <= 0 || classify( p ) == MemKind::Device ) return kSuccess;
	if ( dev_alloc( tmp, (size_t)m * n * sizeof(T), st ) != kSuccess ) return kFailure;
	if ( stage_to_device( *tmp, p, m, n, rs, cs, sizeof(T), st ) != kSuccess ) return kFailure;
	p = (const T*)*tmp; rs = 1; cs = m;
	return kSuccess;
}

template <typename T>
static int gemmt_family_front( int op, int uploc, int transa, int transb, int64_t m, int64_t k,
                               const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                               const T* b, int64_t rs_b, int64_t cs_b,
                               const T* beta, T* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || k < 0 ) return fail( "%s: negative dimension", name );
	if ( !alpha || !beta ) return fail( "%s: alpha/beta must be non-NULL host pointers", name );
	if ( uploc != B200_LOWER && uploc != B200_UPPER ) return fail( "%s: uplo must be BLIS_LOWER or BLIS_UPPER", name );
	if ( m == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	constexpr bool CPLX = Elem<T>::cplx;
	const bool two_operands = ( op == kOpGemmt || op == kOpSyr2k || op == kOpHer2k );
	const bool hermitian    = ( op == kOpHerk || op == kOpHer2k ); (void)hermitian;
	if ( !two_operands ) { b = a; rs_b = rs_a; cs_b = cs_a; transb = transa; }

	// op(A): m x k.  op(B): k x m for gemmt, m x k for the rank-2k operations (bli_l3_tapi_ex.c:251-252).
	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool ca = CPLX && ( transa & B200_CONJ_NO_TRANSPOSE );
	const bool cb = CPLX && ( transb & B200_CONJ_NO_TRANSPOSE );
	const T al = *alpha, be = *beta;
	const bool need_ab = ( k > 0 && !Scalar<T>::is_zero( al ) );

	void *da = nullptr, *db = nullptr, *dc = nullptr;
	int rc = kSuccess;
	const bool shared_ab = ( op != kOpGemmt && a == b && rs_a == rs_b && cs_a == cs_b );
	if ( need_ab )
	{
		const T* a0 = a;
		rc = operand_to_device( a, rs_a, cs_a, m, k, &da, st );
		if ( rc == kSuccess )
		{
			if ( !two_operands || ( shared_ab && a0 != a ) ) { b = a; rs_b = rs_a; cs_b = cs_a; }
			else if ( op == kOpGemmt ) rc = operand_to_device( b, rs_b, cs_b, k, m, &db, st );
			else                       rc = operand_to_device( b, rs_b, cs_b, m, k, &db, st );
		}
	}
	const bool c_host = ( classify( c ) != MemKind::Device );
	T* cdev = c; int64_t rs_cd = rs_c, cs_cd = cs_c;
	if ( rc == kSuccess && c_host )
	{
		// the whole array travels both ways, so the triangle that is not stored returns unchanged
		const T* cc = c;
		rc = operand_to_device( cc, rs_cd, cs_cd, m, m, &dc, st );
		cdev = (T*)dc;
	}

	const T one = Scalar<T>::make( 1.0, 0.0 );
	if ( rc == kSuccess )
	{
		switch ( op )
		{
			case kOpGemmt:          // C := beta*C + alpha * op(A) * op(B)
				rc = gemm_dev<T>( ca, cb, m, m, k, al, a, rs_a, cs_a, b, rs_b, cs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpSyrk:           // C := beta*C + alpha * op(A) * op(A)^T
				rc = gemm_dev<T>( ca, ca, m, m, k, al, a, rs_a, cs_a, a, cs_a, rs_a, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpHerk:           // C := beta*C + alpha * op(A) * op(A)^H   (alpha, beta real)
				rc = gemm_dev<T>( ca, !ca && CPLX, m, m, k, al, a, rs_a, cs_a, a, cs_a, rs_a, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpSyr2k:          // C := beta*C + alpha * op(A) * op(B)^T + alpha * op(B) * op(A)^T
				rc = gemm_dev<T>( ca, cb, m, m, k, al, a, rs_a, cs_a, b, cs_b, rs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				if ( rc == kSuccess )
				rc = gemm_dev<T>( cb, ca, m, m, k, al, b, rs_b, cs_b, a, cs_a, rs_a, one, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpHer2k:          // C := beta*C + alpha * op(A) * op(B)^H + conj(alpha) * op(B) * op(A)^H   (beta real)
			{
				T alh = al;
				if constexpr ( CPLX ) alh.y = -alh.y;
				rc = gemm_dev<T>( ca, !cb && CPLX, m, m, k, al, a, rs_a, cs_a, b, cs_b, rs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				if ( rc == kSuccess )
				rc = gemm_dev<T>( cb, !ca && CPLX, m, m, k, alh, b, rs_b, cs_b, a, cs_a, rs_a, one, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			}
			default: rc = fail( "%s: unknown operation", name );
		}
	}
	if constexpr ( CPLX )
	{
		if ( rc == kSuccess && hermitian )
		{
			using R = typename Elem<T>::real;
			const int blocks = (int)std::min<int64_t>( ( m + 255 ) / 256, (int64_t)ctx().num_sms * 4 );
			zero_diag_imag_kernel<R><<<blocks, 256, 0, st>>>( (R*)cdev, rs_cd + cs_cd, m );
			if ( cudaGetLastError() != cudaSuccess ) rc = fail( "%s: launch failed", name );
			note_launch( "zero_diag_imag_kernel" );
		}
	}
	if ( rc == kSuccess && c_host )
	{
		rc = stage_to_host( c, rs_c, cs_c, dc, m, m, sizeof(T), st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "%s: stream sync failed", name );
	}
	dev_free( da, st ); dev_free( db, st ); dev_free( dc, st );
	return rc;
}

} // namespace b200
