// host_md.cuh -- mixed-datatype gemm front end
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_gemm.cuh"
namespace b200 {

// ---- mixed-datatype gemm ---------------------------------------------
// bli_gemm_ex with operands of different domain and/or precision (docs/MixedDatatypes.md; frame/3/gemm/bli_gemm_cntl.c:
// 87-392): A and B are typecast to the computation precision while they are packed, the product runs in the computation
// precision in the smallest domain that holds it (table of MixedDatatypes.md: "R += C*C" keeps only the real part and
// costs 4mnk, "C += R*C" treats the complex operand as a real matrix with twice the rows, ...), and the result is
// typecast and accumulated into C with beta in C's own datatype (ref_kernels/3/bli_gemm_ref.c:318-385,
// ref_kernels/ind/bli_gemm_{ccr,crr,rcc}_ref.c).  Here: one conversion pass per operand (typecast, transposition,
// conjugation, alpha where it has to act before a projection), ONE homogeneous real or complex gemm of the computation
// precision with the kernels above, one combine pass into C.

struct MdElem { double r, i; };

__device__ __forceinline__ MdElem md_load( const void* base, int dt, int64_t off )
{
	MdElem e; e.i = 0.0;
	switch ( dt )
	{
		case B200_FLOAT:    e.r = ( (const float*)base )[off]; break;
		case B200_DOUBLE:   e.r = ( (const double*)base )[off]; break;
		case B200_SCOMPLEX: { const float2 v = ( (const float2*)base )[off]; e.r = v.x; e.i = v.y; break; }
		default:            { const double2 v = ( (const double2*)base )[off]; e.r = v.x; e.i = v.y; break; }
	}
	return e;
}
__device__ __forceinline__ void md_store( void* base, int dt, int64_t off, MdElem e )
{
	switch ( dt )
	{
		case B200_FLOAT:    ( (float*)base )[off] = (float)e.r; break;
		case B200_DOUBLE:   ( (double*)base )[off] = e.r; break;
		case B200_SCOMPLEX: ( (float2*)base )[off] = make_float2( (float)e.r, (float)e.i ); break;
		default:            ( (double2*)base )[off] = make_double2( e.r, e.i ); break;
	}
}

// dst(i,j) [dense, strides rs_d/cs_d, datatype dt_d] := f( src(i,j) ) for an m x n view of src:
// typecast to the precision of dt_d, optional conjugation, optional multiplication by kappa (after the cast, as
// packm's scal2s does), real projection when dt_d is real, optional negation of the imaginary part afterwards.
__global__ void md_convert_kernel( void* dst, int dt_d, int64_t rs_d, int64_t cs_d, const void* src, int dt_s, int64_t rs_s, int64_t cs_s,
                                   int64_t m, int64_t n, int conj, int use_kappa, double kr, double ki, int neg_imag, int single_prec )
{
	const int64_t total = m * n;
	const bool inner_row = ( rs_d <= cs_d );
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		MdElem v = md_load( src, dt_s, i * rs_s + j * cs_s );
		if ( single_prec ) { v.r = (double)(float)v.r; v.i = (double)(float)v.i; }
		if ( conj ) v.i = -v.i;
		if ( use_kappa )
		{
			MdElem w;
			if ( single_prec ) { w.r = (double)( (float)kr * (float)v.r - (float)ki * (float)v.i ); w.i = (double)( (float)kr * (float)v.i + (float)ki * (float)v.r ); }
			else               { w.r = kr * v.r - ki * v.i; w.i = kr * v.i + ki * v.r; }
			v = w;
		}
		if ( neg_imag ) v.i = -v.i;
		md_store( dst, dt_d, i * rs_d + j * cs_d, v );
	}
}

// C(i,j) := beta * C(i,j) + alpha * T(i,j), evaluated in C's precision on the typecast T (bli_txpbys / bli_taxpbys
// with the C datatype as computation type); beta == 0 does not read C; a real C keeps the real part.
__global__ void md_combine_kernel( void* c, int dt_c, int64_t rs_c, int64_t cs_c, const void* t, int dt_t, int64_t rs_t, int64_t cs_t,
                                   int64_t m, int64_t n, double ar, double ai, double br, double bi, int beta_is_zero )
{
	const int64_t total = m * n;
	const bool inner_row = ( ( rs_c < 0 ? -rs_c : rs_c ) <= ( cs_c < 0 ? -cs_c : cs_c ) );
	const bool c_single = ( dt_c == B200_FLOAT || dt_c == B200_SCOMPLEX );
	const bool c_real = ( dt_c == B200_FLOAT || dt_c == B200_DOUBLE );
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		MdElem x; x.r = 0.0; x.i = 0.0;
		if ( ar != 0.0 || ai != 0.0 ) x = md_load( t, dt_t, i * rs_t + j * cs_t );       // alpha == 0: T is not read
		MdElem y; y.r = 0.0; y.i = 0.0;
		if ( !beta_is_zero ) y = md_load( c, dt_c, i * rs_c + j * cs_c );
		MdElem o;
		if ( c_single )
		{
			const float xr = (float)x.r, xi = (float)x.i, far_ = (float)ar, fai = (float)ai, fbr = (float)br, fbi = (float)bi;
			float pr = far_ * xr - fai * xi, pi = far_ * xi + fai * xr;
			if ( !beta_is_zero ) { pr += fbr * (float)y.r - fbi * (float)y.i; pi += fbr * (float)y.i + fbi * (float)y.r; }
			o.r = pr; o.i = pi;
		}
		else
		{
			double pr = ar * x.r - ai * x.i, pi = ar * x.i + ai * x.r;
			if ( !beta_is_zero ) { pr += br * y.r - bi * y.i; pi += br * y.i + bi * y.r; }
			o.r = pr; o.i = pi;
		}
		if ( c_real ) o.i = 0.0;
		md_store( c, dt_c, i * rs_c + j * cs_c, o );
	}
}

static inline bool dt_is_real( int dt ) { return dt == B200_FLOAT || dt == B200_DOUBLE; }
static inline size_t dt_size( int dt ) { return dt == B200_FLOAT ? 4 : dt == B200_DCOMPLEX ? 16 : 8; }
static inline int dt_make( bool real, bool single ) { return real ? ( single ? B200_FLOAT : B200_DOUBLE ) : ( single ? B200_SCOMPLEX : B200_DCOMPLEX ); }

static int md_convert( void* dst, int dt_d, int64_t rs_d, int64_t cs_d, const void* src, int dt_s, int64_t rs_s, int64_t cs_s,
                       int64_t m, int64_t n, bool conj, bool use_kappa, double kr, double ki, bool neg_imag, cudaStream_t st )
{
	const int64_t total = m * n;
	if ( total <= 0 ) return kSuccess;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	const int single = ( dt_d == B200_FLOAT || dt_d == B200_SCOMPLEX ) ? 1 : 0;
	md_convert_kernel<<<blocks, 256, 0, st>>>( dst, dt_d, rs_d, cs_d, src, dt_s, rs_s, cs_s, m, n, conj ? 1 : 0, use_kappa ? 1 : 0, kr, ki, neg_imag ? 1 : 0, single );
	B200_CUDA( cudaGetLastError() );
	note_launch( "md_convert_kernel" );
	return kSuccess;
}

// Real or complex homogeneous product T := X * Y (beta = 0) of the computation precision on dense device operands.
static int md_gemm( bool real, bool single, int64_t m, int64_t n, int64_t k, const void* x, int64_t rs_x, int64_t cs_x,
                    const void* y, int64_t rs_y, int64_t cs_y, void* t, int64_t rs_t, int64_t cs_t, cudaStream_t st )
{
	if ( real && single )   return gemm_dev<float>  ( false, false, m, n, k, 1.f, (const float*)x, rs_x, cs_x, (const float*)y, rs_y, cs_y, 0.f, (float*)t, rs_t, cs_t, st );
	if ( real )             return gemm_dev<double> ( false, false, m, n, k, 1.0, (const double*)x, rs_x, cs_x, (const double*)y, rs_y, cs_y, 0.0, (double*)t, rs_t, cs_t, st );
	if ( single )           return gemm_dev<float2> ( false, false, m, n, k, make_float2( 1.f, 0.f ), (const float2*)x, rs_x, cs_x, (const float2*)y, rs_y, cs_y, make_float2( 0.f, 0.f ), (float2*)t, rs_t, cs_t, st );
	return gemm_dev<double2>( false, false, m, n, k, make_double2( 1.0, 0.0 ), (const double2*)x, rs_x, cs_x, (const double2*)y, rs_y, cs_y, make_double2( 0.0, 0.0 ), (double2*)t, rs_t, cs_t, st );
}

static int gemm_md_front( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb, int64_t m, int64_t n, int64_t k,
                          const double* alpha, const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                          const double* beta, void* c, int64_t rs_c, int64_t cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	for ( int dt : { dt_a, dt_b, dt_c } ) if ( dt < 0 || dt > 3 ) return fail( "b200_gemm_md: unsupported datatype %d", dt );
	if ( comp_prec != 0 && comp_prec != 2 ) return fail( "b200_gemm_md: computation precision must be BLIS_SINGLE_PREC (0) or BLIS_DOUBLE_PREC (2)" );
	if ( m < 0 || n < 0 || k < 0 ) return fail( "b200_gemm_md: negative dimension" );
	if ( !alpha || !beta ) return fail( "b200_gemm_md: alpha/beta must be non-NULL host pointers (dcomplex)" );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	const bool a_real = dt_is_real( dt_a ), b_real = dt_is_real( dt_b ), c_real = dt_is_real( dt_c ), single = ( comp_prec == 0 );
	// alpha lives in the computation precision, complex if any operand is; beta in C's datatype (bli_gemm_cntl.c:174-189)
	double ar = alpha[0], ai = ( a_real && b_real && c_real ) ? 0.0 : alpha[1];
	double br = beta[0],  bi = c_real ? 0.0 : beta[1];
	if ( single ) { ar = (double)(float)ar; ai = (double)(float)ai; }
	if ( dt_c == B200_FLOAT || dt_c == B200_SCOMPLEX ) { br = (double)(float)br; bi = (double)(float)bi; }
	const bool beta_zero = ( br == 0.0 && bi == 0.0 );

	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool conja = !a_real && ( transa & B200_CONJ_NO_TRANSPOSE ), conjb = !b_real && ( transb & B200_CONJ_NO_TRANSPOSE );

	void *da = nullptr, *db = nullptr, *dc = nullptr, *pa = nullptr, *pb = nullptr, *pt = nullptr;
	int rc = kSuccess;
	// host operands: raw bytes to the device first
	const bool c_host = ( classify( c ) != MemKind::Device );
	const bool trivial = ( k == 0 || ( ar == 0.0 && ai == 0.0 ) );
	void* cdev = c; int64_t rs_cd = rs_c, cs_cd = cs_c;
	if ( c_host )
	{
		if ( dev_alloc( &dc, (size_t)m * n * dt_size( dt_c ), st ) != kSuccess ) return kFailure;
		if ( !beta_zero ) rc = stage_to_device( dc, c, m, n, rs_c, cs_c, dt_size( dt_c ), st );
		cdev = dc; rs_cd = 1; cs_cd = m;
	}
	if ( trivial )
	{
		// bli_l3_return_early_if_trivial: C := beta * C
		if ( rc == kSuccess )
		{
			const int blocks = (int)std::min<int64_t>( ( m * n + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
			md_combine_kernel<<<blocks, 256, 0, st>>>( cdev, dt_c, rs_cd, cs_cd, cdev, dt_c, rs_cd, cs_cd, m, n, 0.0, 0.0, br, bi, beta_zero ? 1 : 0 );
			if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm_md: launch failed" );
			note_launch( "md_combine_kernel" );
		}
	}
	else
	{
		if ( rc == kSuccess && classify( a ) != MemKind::Device )
		{
			if ( dev_alloc( &da, (size_t)m * k * dt_size( dt_a ), st ) != kSuccess ) rc = kFailure;
			else rc = stage_to_device( da, a, m, k, rs_a, cs_a, dt_size( dt_a ), st );
			a = da; rs_a = 1; cs_a = m;
		}
		if ( rc == kSuccess && classify( b ) != MemKind::Device )
		{
			if ( dev_alloc( &db, (size_t)k * n * dt_size( dt_b ), st ) != kSuccess ) rc = kFailure;
			else rc = stage_to_device( db, b, k, n, rs_b, cs_b, dt_size( dt_b ), st );
			b = db; rs_b = 1; cs_b = k;
		}
		const size_t es_r = single ? 4 : 8, es_z = 2 * es_r;
		const int dt_r = dt_make( true, single ), dt_z = dt_make( false, single );
		int dt_t = dt_r; int64_t rs_t = 1, cs_t = m; double car = ar, cai = ai;      // T layout and the alpha left for the combine step
		if ( rc == kSuccess && ( dev_alloc( &pa, (size_t)m * k * es_z, st ) != kSuccess || dev_alloc( &pb, (size_t)k * n * es_z, st ) != kSuccess ||
		                         dev_alloc( &pt, (size_t)m * n * es_z, st ) != kSuccess ) ) rc = kFailure;
		if ( rc == kSuccess )
		{
			if ( a_real && b_real )
			{
				// R*R (C real or complex): real product, alpha (complex when C is) applied by the combine step
				rc = md_convert( pa, dt_r, 1, m, a, dt_a, rs_a, cs_a, m, k, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_r, 1, k, b, dt_b, rs_b, cs_b, k, n, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, m, n, k, pa, 1, m, pb, 1, k, pt, 1, m, st );
			}
			else if ( !a_real && !b_real && !c_real )
			{
				rc = md_convert( pa, dt_z, 1, m, a, dt_a, rs_a, cs_a, m, k, conja, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_z, 1, k, b, dt_b, rs_b, cs_b, k, n, conjb, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( false, single, m, n, k, pa, 1, m, pb, 1, k, pt, 1, m, st );
				dt_t = dt_z;
			}
			else if ( !c_real && !a_real && b_real )
			{
				// C += C*R: the complex A is a real matrix with 2m rows (interleaved re/im), T likewise: 4mnk flops
				rc = md_convert( pa, dt_z, 1, m, a, dt_a, rs_a, cs_a, m, k, conja, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_r, 1, k, b, dt_b, rs_b, cs_b, k, n, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, 2 * m, n, k, pa, 1, 2 * m, pb, 1, k, pt, 1, 2 * m, st );
				dt_t = dt_z;
			}
			else if ( !c_real && a_real && !b_real )
			{
				// C += R*C: transposed, T^T = B^T A^T with B^T a real matrix with 2n rows; T comes out row-major
				rc = md_convert( pb, dt_z, 1, n, b, dt_b, cs_b, rs_b, n, k, conjb, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pa, dt_r, 1, k, a, dt_a, cs_a, rs_a, k, m, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, 2 * n, m, k, pb, 1, 2 * n, pa, 1, k, pt, 1, 2 * n, st );
				dt_t = dt_z; rs_t = n; cs_t = 1;
			}
			else if ( c_real && !a_real && !b_real )
			{
				// R += C*C: T = Re( alpha*A * B ) = [ Re | -Im ]( alpha*A ) * [ Re ; Im ]( B ): a real product with 2k inner
				// dimension (the reference's 1r packing with one operand conjugated, bli_gemm_cntl.c:349-366): 4mnk flops
				rc = md_convert( pa, dt_z, k, 1, a, dt_a, rs_a, cs_a, m, k, conja, true, ar, ai, true, st );      // row-major m x k
				if ( rc == kSuccess ) rc = md_convert( pb, dt_z, 1, k, b, dt_b, rs_b, cs_b, k, n, conjb, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, m, n, 2 * k, pa, 2 * k, 1, pb, 1, 2 * k, pt, 1, m, st );
				car = 1.0; cai = 0.0;
			}
			else
			{
				// R += C*R or R += R*C: only the real part of ( alpha * the complex operand ) takes part
				// (BLIS_PACKED_PANELS_RO, bli_gemm_cntl.c:374-389)
				rc = md_convert( pa, dt_r, 1, m, a, dt_a, rs_a, cs_a, m, k, conja, !a_real, ar, ai, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_r, 1, k, b, dt_b, rs_b, cs_b, k, n, conjb, !b_real, ar, ai, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, m, n, k, pa, 1, m, pb, 1, k, pt, 1, m, st );
				car = 1.0; cai = 0.0;
			}
		}
		if ( rc == kSuccess )
		{
			const int blocks = (int)std::min<int64_t>( ( m * n + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
			md_combine_kernel<<<blocks, 256, 0, st>>>( cdev, dt_c, rs_cd, cs_cd, pt, dt_t, rs_t, cs_t, m, n, car, cai, br, bi, beta_zero ? 1 : 0 );
			if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm_md: launch failed" );
			note_launch( "md_combine_kernel" );
		}
	}
	if ( rc == kSuccess && c_host )
	{
		rc = stage_to_host( c, rs_c, cs_c, dc, m, n, dt_size( dt_c ), st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "b200_gemm_md: stream sync failed" );
	}
	dev_free( da, st ); dev_free( db, st ); dev_free( dc, st ); dev_free( pa, st ); dev_free( pb, st ); dev_free( pt, st );
	return rc;
}

} // namespace b200
