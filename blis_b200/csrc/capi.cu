// capi.cu -- the extern "C" entry points of include/blis_b200.h (the drop-in boundary).
//
// Every entry point is a thin datatype dispatch onto a host-side front end that mirrors one of the reference's
// object-API front ends (frame/3/bli_l3_oapi_ex.c):
//   host_gemm.cuh     bli_gemm_ex (:48-148): trivial cases, storage based operand swap (bli_gemm_cntl.c:98-161),
//                     kernel form, host-operand pipelines
//   host_trsm.cuh     bli_trsm_ex (:692-801) + bli_trsm_blk_var1 (frame/3/trsm/bli_trsm_blk_var1.c:40-188)
//   host_gemmt.cuh    bli_gemmt_ex / syrk / herk / syr2k / her2k (:151-346)
//   host_strucmm.cuh  bli_hemm_ex / symm / trmm3 / trmm (:349-689)
//   host_md.cuh       mixed-datatype gemm (frame/3/gemm/bli_gemm_cntl.c:87-392)
//   host_batch.cuh    ?gemm_batch_ (frame/compat/extra/bla_gemm_batch.c)
//   host_dist.cuh     multi-GPU gemm / trsm (one process per GPU, NCCL): the reference's jc x ic partitioning across GPUs
//   host_util.cuh     scalar traits, strided copy / scale / transpose kernels
// The kernels themselves are launched from gemm_{d,z,s,c}.cu (gemm_launch.cuh).

#include "host_util.cuh"
#include "host_gemm.cuh"
#include "host_trsm.cuh"
#include "host_gemmt.cuh"
#include "host_strucmm.cuh"
#include "host_md.cuh"
#include "host_batch.cuh"
#include "host_dist.cuh"


// ---- C ABI --------------------------------------------------------------------------
using namespace b200;

#define B200_DEF_GEMM( ch, ctype, T ) \
extern "C" b200_err_t b200_##ch##gemm( int transa, int transb, b200_dim_t m, b200_dim_t n, b200_dim_t k, \
	const ctype* alpha, const ctype* a, b200_inc_t rs_a, b200_inc_t cs_a, \
	const ctype* b, b200_inc_t rs_b, b200_inc_t cs_b, const ctype* beta, \
	ctype* c, b200_inc_t rs_c, b200_inc_t cs_c ) \
{ \
	return gemm_front<T>( transa, transb, m, n, k, (const T*)alpha, (const T*)a, rs_a, cs_a, \
	                      (const T*)b, rs_b, cs_b, (const T*)beta, (T*)c, rs_c, cs_c ); \
}
B200_DEF_GEMM( s, float, float )
B200_DEF_GEMM( d, double, double )
B200_DEF_GEMM( c, b200_scomplex, float2 )
B200_DEF_GEMM( z, b200_dcomplex, double2 )

extern "C" b200_err_t b200_gemm( int dt, int transa, int transb, b200_dim_t m, b200_dim_t n, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
	const void* b, b200_inc_t rs_b, b200_inc_t cs_b, const void* beta,
	void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{
	switch ( dt )
	{
		case B200_FLOAT:    return gemm_front<float>  ( transa, transb, m, n, k, (const float*)alpha,   (const float*)a,   rs_a, cs_a, (const float*)b,   rs_b, cs_b, (const float*)beta,   (float*)c,   rs_c, cs_c );
		case B200_DOUBLE:   return gemm_front<double> ( transa, transb, m, n, k, (const double*)alpha,  (const double*)a,  rs_a, cs_a, (const double*)b,  rs_b, cs_b, (const double*)beta,  (double*)c,  rs_c, cs_c );
		case B200_SCOMPLEX: return gemm_front<float2> ( transa, transb, m, n, k, (const float2*)alpha,  (const float2*)a,  rs_a, cs_a, (const float2*)b,  rs_b, cs_b, (const float2*)beta,  (float2*)c,  rs_c, cs_c );
		case B200_DCOMPLEX: return gemm_front<double2>( transa, transb, m, n, k, (const double2*)alpha, (const double2*)a, rs_a, cs_a, (const double2*)b, rs_b, cs_b, (const double2*)beta, (double2*)c, rs_c, cs_c );
	}
	return fail( "b200_gemm: unsupported datatype %d (mixed-datatype gemm is out of scope)", dt );
}

#define B200_DEF_TRSM( ch, ctype, T ) \
extern "C" b200_err_t b200_##ch##trsm( int side, int uploa, int transa, int diaga, b200_dim_t m, b200_dim_t n, \
	const ctype* alpha, const ctype* a, b200_inc_t rs_a, b200_inc_t cs_a, \
	ctype* b, b200_inc_t rs_b, b200_inc_t cs_b ) \
{ \
	return trsm_front<T>( side, uploa, transa, diaga, m, n, (const T*)alpha, (const T*)a, rs_a, cs_a, (T*)b, rs_b, cs_b ); \
}
B200_DEF_TRSM( s, float, float )
B200_DEF_TRSM( d, double, double )
B200_DEF_TRSM( c, b200_scomplex, float2 )
B200_DEF_TRSM( z, b200_dcomplex, double2 )

extern "C" b200_err_t b200_trsm( int dt, int side, int uploa, int transa, int diaga, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
	void* b, b200_inc_t rs_b, b200_inc_t cs_b )
{
	switch ( dt )
	{
		case B200_FLOAT:    return trsm_front<float>  ( side, uploa, transa, diaga, m, n, (const float*)alpha,   (const float*)a,   rs_a, cs_a, (float*)b,   rs_b, cs_b );
		case B200_DOUBLE:   return trsm_front<double> ( side, uploa, transa, diaga, m, n, (const double*)alpha,  (const double*)a,  rs_a, cs_a, (double*)b,  rs_b, cs_b );
		case B200_SCOMPLEX: return trsm_front<float2> ( side, uploa, transa, diaga, m, n, (const float2*)alpha,  (const float2*)a,  rs_a, cs_a, (float2*)b,  rs_b, cs_b );
		case B200_DCOMPLEX: return trsm_front<double2>( side, uploa, transa, diaga, m, n, (const double2*)alpha, (const double2*)a, rs_a, cs_a, (double2*)b, rs_b, cs_b );
	}
	return fail( "b200_trsm: unsupported datatype %d", dt );
}

// k-panel accumulation: C := beta*C + alpha * sum_{s<npanels} op(A_s) * op(B_s), every panel k wide, all
// A panels (resp. B panels) with the same strides.  Device-resident operands, d and z only.

static int struc_mm_dt( int dt, int op, int side, int uploa, int transa, int diaga, int transb, int64_t m, int64_t n,
                        const void* alpha, const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                        const void* beta, void* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	switch ( dt )
	{
		case B200_FLOAT:    return struc_mm_front<float>  ( op, side, uploa, transa, diaga, transb, m, n, (const float*)alpha,   (const float*)a,   rs_a, cs_a, (const float*)b,   rs_b, cs_b, (const float*)beta,   (float*)c,   rs_c, cs_c, name );
		case B200_DOUBLE:   return struc_mm_front<double> ( op, side, uploa, transa, diaga, transb, m, n, (const double*)alpha,  (const double*)a,  rs_a, cs_a, (const double*)b,  rs_b, cs_b, (const double*)beta,  (double*)c,  rs_c, cs_c, name );
		case B200_SCOMPLEX: return struc_mm_front<float2> ( op, side, uploa, transa, diaga, transb, m, n, (const float2*)alpha,  (const float2*)a,  rs_a, cs_a, (const float2*)b,  rs_b, cs_b, (const float2*)beta,  (float2*)c,  rs_c, cs_c, name );
		case B200_DCOMPLEX: return struc_mm_front<double2>( op, side, uploa, transa, diaga, transb, m, n, (const double2*)alpha, (const double2*)a, rs_a, cs_a, (const double2*)b, rs_b, cs_b, (const double2*)beta, (double2*)c, rs_c, cs_c, name );
	}
	return fail( "%s: unsupported datatype %d", name, dt );
}

extern "C" b200_err_t b200_hemm( int dt, int side, int uploa, int conja, int transb, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return struc_mm_dt( dt, 0, side, uploa, conja, B200_NONUNIT_DIAG, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_hemm" ); }

extern "C" b200_err_t b200_symm( int dt, int side, int uploa, int conja, int transb, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return struc_mm_dt( dt, 1, side, uploa, conja, B200_NONUNIT_DIAG, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_symm" ); }

extern "C" b200_err_t b200_trmm3( int dt, int side, int uploa, int transa, int diaga, int transb, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return struc_mm_dt( dt, 2, side, uploa, transa, diaga, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_trmm3" ); }

extern "C" b200_err_t b200_trmm( int dt, int side, int uploa, int transa, int diaga, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, void* b, b200_inc_t rs_b, b200_inc_t cs_b )
{ return struc_mm_dt( dt, 3, side, uploa, transa, diaga, B200_NO_TRANSPOSE, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, nullptr, b, rs_b, cs_b, "b200_trmm" ); }

// ---- gemmt family C ABI ------------------------------------------------------------------
// Real-typed scalars of herk (alpha, beta) and her2k (beta) are widened to the matrix datatype (imaginary part 0),
// as bli_obj_init_finish_1x1( dt_r, ... ) + typecast does in bli_l3_tapi_ex.c:184-185,259-260.
template <typename T> static T widen_real( const void* p )
{
	using R = typename Elem<T>::real;
	return Scalar<T>::make( (double)*(const R*)p, 0.0 );
}

template <typename T>
static int gemmt_family_any( int op, int uploc, int transa, int transb, int64_t m, int64_t k, const void* alpha,
                             const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                             const void* beta, void* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	if ( !alpha || !beta ) return fail( "%s: alpha/beta must be non-NULL host pointers", name );
	const T al = ( op == kOpHerk )                    ? widen_real<T>( alpha ) : *(const T*)alpha;
	const T be = ( op == kOpHerk || op == kOpHer2k ) ? widen_real<T>( beta )  : *(const T*)beta;
	return gemmt_family_front<T>( op, uploc, transa, transb, m, k, &al, (const T*)a, rs_a, cs_a, (const T*)b, rs_b, cs_b,
	                              &be, (T*)c, rs_c, cs_c, name );
}

static int gemmt_family_dt( int dt, int op, int uploc, int transa, int transb, int64_t m, int64_t k, const void* alpha,
                            const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                            const void* beta, void* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	switch ( dt )
	{
		case B200_FLOAT:    return gemmt_family_any<float>  ( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
		case B200_DOUBLE:   return gemmt_family_any<double> ( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
		case B200_SCOMPLEX: return gemmt_family_any<float2> ( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
		case B200_DCOMPLEX: return gemmt_family_any<double2>( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
	}
	return fail( "%s: unsupported datatype %d", name, dt );
}

extern "C" b200_err_t b200_gemmt( int dt, int uploc, int transa, int transb, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpGemmt, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_gemmt" ); }

extern "C" b200_err_t b200_syr2k( int dt, int uploc, int transa, int transb, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpSyr2k, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_syr2k" ); }

extern "C" b200_err_t b200_her2k( int dt, int uploc, int transa, int transb, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta_real, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpHer2k, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta_real, c, rs_c, cs_c, "b200_her2k" ); }

extern "C" b200_err_t b200_syrk( int dt, int uploc, int transa, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpSyrk, uploc, transa, transa, m, k, alpha, a, rs_a, cs_a, a, rs_a, cs_a, beta, c, rs_c, cs_c, "b200_syrk" ); }

extern "C" b200_err_t b200_herk( int dt, int uploc, int transa, b200_dim_t m, b200_dim_t k,
	const void* alpha_real, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* beta_real, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpHerk, uploc, transa, transa, m, k, alpha_real, a, rs_a, cs_a, a, rs_a, cs_a, beta_real, c, rs_c, cs_c, "b200_herk" ); }


extern "C" b200_err_t b200_gemm_md( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb,
	b200_dim_t m, b200_dim_t n, b200_dim_t k, const double* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
	const void* b, b200_inc_t rs_b, b200_inc_t cs_b, const double* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{
	return gemm_md_front( dt_a, dt_b, dt_c, comp_prec, transa, transb, m, n, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c );
}


extern "C" b200_err_t b200_gemm_batch( int dt, int group_count, const int* group_size, const int* transa, const int* transb,
	const b200_dim_t* m, const b200_dim_t* n, const b200_dim_t* k, const void* alpha,
	const void* const* a, const b200_inc_t* rs_a, const b200_inc_t* cs_a,
	const void* const* b, const b200_inc_t* rs_b, const b200_inc_t* cs_b,
	const void* beta, void* const* c, const b200_inc_t* rs_c, const b200_inc_t* cs_c )
{
	switch ( dt )
	{
		case B200_FLOAT:    return gemm_batch_front<float>  ( group_count, group_size, transa, transb, m, n, k, (const float*)alpha,   (const float* const*)a,   rs_a, cs_a, (const float* const*)b,   rs_b, cs_b, (const float*)beta,   (float* const*)c,   rs_c, cs_c );
		case B200_DOUBLE:   return gemm_batch_front<double> ( group_count, group_size, transa, transb, m, n, k, (const double*)alpha,  (const double* const*)a,  rs_a, cs_a, (const double* const*)b,  rs_b, cs_b, (const double*)beta,  (double* const*)c,  rs_c, cs_c );
		case B200_SCOMPLEX: return gemm_batch_front<float2> ( group_count, group_size, transa, transb, m, n, k, (const float2*)alpha,  (const float2* const*)a,  rs_a, cs_a, (const float2* const*)b,  rs_b, cs_b, (const float2*)beta,  (float2* const*)c,  rs_c, cs_c );
		case B200_DCOMPLEX: return gemm_batch_front<double2>( group_count, group_size, transa, transb, m, n, k, (const double2*)alpha, (const double2* const*)a, rs_a, cs_a, (const double2* const*)b, rs_b, cs_b, (const double2*)beta, (double2* const*)c, rs_c, cs_c );
	}
	return fail( "b200_gemm_batch: unsupported datatype %d", dt );
}

template <typename T>
static int kpanels_front( int transa, int transb, int64_t m, int64_t n, int64_t k, int npanels, const T* alpha,
                          const T* const* a, int64_t rs_a, int64_t cs_a, const T* const* b, int64_t rs_b, int64_t cs_b,
                          const T* beta, T* c, int64_t rs_c, int64_t cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( npanels < 1 || npanels > 8 ) return fail( "b200_gemm_kpanels: 1..8 panels per call" );
	if ( m < 0 || n < 0 || k < 0 ) return fail( "b200_gemm_kpanels: negative dimension" );
	if ( m == 0 || n == 0 ) return kSuccess;
	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool conja = Elem<T>::cplx && ( transa & B200_CONJ_NO_TRANSPOSE );
	const bool conjb = Elem<T>::cplx && ( transb & B200_CONJ_NO_TRANSPOSE );
	for ( int sgm = 0; sgm < npanels; ++sgm )
		if ( classify( a[sgm] ) != MemKind::Device || classify( b[sgm] ) != MemKind::Device )
			return fail( "b200_gemm_kpanels: panels must be device resident" );
	if ( classify( c ) != MemKind::Device ) return fail( "b200_gemm_kpanels: C must be device resident" );
	return gemm_dev<T>( conja, conjb, m, n, k, *alpha, a[0], rs_a, cs_a, b[0], rs_b, cs_b, *beta, c, rs_c, cs_c,
	                    cur_stream(), npanels, a + 1, b + 1 );
}

extern "C" b200_err_t b200_gemm_kpanels( int dt, int transa, int transb, b200_dim_t m, b200_dim_t n, b200_dim_t k, int npanels,
	const void* alpha, const void* const* a, b200_inc_t rs_a, b200_inc_t cs_a,
	const void* const* b, b200_inc_t rs_b, b200_inc_t cs_b, const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{
	if ( dt == B200_DOUBLE )
		return kpanels_front<double>( transa, transb, m, n, k, npanels, (const double*)alpha, (const double* const*)a, rs_a, cs_a,
		                              (const double* const*)b, rs_b, cs_b, (const double*)beta, (double*)c, rs_c, cs_c );
	if ( dt == B200_DCOMPLEX )
		return kpanels_front<double2>( transa, transb, m, n, k, npanels, (const double2*)alpha, (const double2* const*)a, rs_a, cs_a,
		                               (const double2* const*)b, rs_b, cs_b, (const double2*)beta, (double2*)c, rs_c, cs_c );
	return fail( "b200_gemm_kpanels: only d and z are supported" );
}

extern "C" int b200_trsm_rowblock_plan( b200_dim_t m, b200_dim_t n, int leaf_rows, int upper, b200_dim_t rb, b200_dim_t* out, int cap )
{
	std::vector<TrsmPiece> plan;
	TrsmRowSched sched;
	if ( m <= 0 || leaf_rows <= 0 || !trsm_row_sched( m, n, leaf_rows, rb, sched ) ) return 0;      // 0: the column-block pipeline serves this shape
	trsm_rowblock_plan( leaf_rows, upper != 0, m, sched, plan );
	for ( size_t i = 0; i < plan.size() && (int)i < cap && out; ++i )
	{
		out[5 * i] = plan[i].r0; out[5 * i + 1] = plan[i].r1; out[5 * i + 2] = plan[i].c0; out[5 * i + 3] = plan[i].c1; out[5 * i + 4] = plan[i].launches;
	}
	return (int)plan.size();
}

extern "C" int b200_trsm_upload_plan( b200_dim_t m, int leaf_rows, int upper, b200_dim_t* out, int cap )
{
	std::vector<TrsmPiece> plan;
	if ( m > 0 && leaf_rows > 0 ) trsm_upload_plan( leaf_rows, upper != 0, 0, m, plan );
	for ( size_t i = 0; i < plan.size() && (int)i < cap && out; ++i )
	{
		out[5 * i] = plan[i].r0; out[5 * i + 1] = plan[i].r1; out[5 * i + 2] = plan[i].c0; out[5 * i + 3] = plan[i].c1; out[5 * i + 4] = plan[i].launches;
	}
	return (int)plan.size();
}

// ---- multi-GPU (host_dist.cuh) ------------------------------------------------------------------------------------
extern "C" void b200_partition_2x2( b200_dim_t n_thread, b200_dim_t work1, b200_dim_t work2, b200_dim_t* nt1, b200_dim_t* nt2 )
{ partition_2x2( n_thread, work1, work2, nt1, nt2 ); }
extern "C" void b200_range_sub( b200_dim_t work_id, b200_dim_t n_way, b200_dim_t n, b200_dim_t bf, int handle_edge_low, b200_dim_t* start, b200_dim_t* end )
{ range_sub( work_id, n_way, n, bf, handle_edge_low != 0, start, end ); }
extern "C" b200_err_t b200_dist_plan( int world, int rank, b200_dim_t m, b200_dim_t n, b200_dim_t k, b200_dim_t kb, b200_dist_plan_t* plan )
{ return dist_plan( world, rank, m, n, k, kb, plan ); }
extern "C" b200_err_t b200_dist_unique_id( void* id ) { return dist_unique_id( id ); }
extern "C" b200_err_t b200_dist_init( int world, int rank, const void* id ) { return dist_init( world, rank, id ); }
extern "C" b200_err_t b200_dist_finalize( void ) { return dist_finalize(); }
extern "C" double b200_dist_last_wait_ms( void ) { return dist_last_wait_ms(); }
extern "C" b200_err_t b200_dist_register( const void* a_loc, const void* b_loc ) { return dist_register( a_loc, b_loc ); }
extern "C" b200_err_t b200_dist_unregister( const void* a_loc, const void* b_loc ) { return dist_unregister( a_loc, b_loc ); }
extern "C" int b200_dist_transport( void ) { return dist().last_transport; }
extern "C" b200_err_t b200_dist_gemm( int dt, b200_dim_t m, b200_dim_t n, b200_dim_t k, b200_dim_t kb, const void* alpha,
	const void* a_loc, const void* b_loc, const void* beta, void* c_loc, b200_inc_t rs_c, b200_inc_t cs_c, int flags )
{
	if ( !alpha || !beta ) return fail( "b200_dist_gemm: alpha and beta must be non-NULL host pointers" );
	if ( dt == B200_DOUBLE )
		return dist_gemm<double>( m, n, k, kb, (const double*)alpha, (const double*)a_loc, (const double*)b_loc, (const double*)beta, (double*)c_loc, rs_c, cs_c, flags );
	if ( dt == B200_DCOMPLEX )
		return dist_gemm<double2>( m, n, k, kb, (const double2*)alpha, (const double2*)a_loc, (const double2*)b_loc, (const double2*)beta, (double2*)c_loc, rs_c, cs_c, flags );
	return fail( "b200_dist_gemm: only d and z are supported (k-panel accumulation, b200_gemm_kpanels)" );
}
extern "C" b200_err_t b200_dist_gemm_1d( int dt, int split, int root, b200_dim_t m, b200_dim_t n, b200_dim_t k, const void* alpha,
	void* a, b200_inc_t rs_a, b200_inc_t cs_a, void* b, b200_inc_t rs_b, b200_inc_t cs_b, const void* beta, void* c_loc, b200_inc_t rs_c, b200_inc_t cs_c )
{
	if ( !alpha || !beta ) return fail( "b200_dist_gemm_1d: alpha and beta must be non-NULL host pointers" );
	switch ( dt )
	{
		case B200_FLOAT:    return dist_gemm_1d<float>  ( split, root, m, n, k, (const float*)alpha,   (float*)a,   rs_a, cs_a, (float*)b,   rs_b, cs_b, (const float*)beta,   (float*)c_loc,   rs_c, cs_c );
		case B200_DOUBLE:   return dist_gemm_1d<double> ( split, root, m, n, k, (const double*)alpha,  (double*)a,  rs_a, cs_a, (double*)b,  rs_b, cs_b, (const double*)beta,  (double*)c_loc,  rs_c, cs_c );
		case B200_SCOMPLEX: return dist_gemm_1d<float2> ( split, root, m, n, k, (const float2*)alpha,  (float2*)a,  rs_a, cs_a, (float2*)b,  rs_b, cs_b, (const float2*)beta,  (float2*)c_loc,  rs_c, cs_c );
		case B200_DCOMPLEX: return dist_gemm_1d<double2>( split, root, m, n, k, (const double2*)alpha, (double2*)a, rs_a, cs_a, (double2*)b, rs_b, cs_b, (const double2*)beta, (double2*)c_loc, rs_c, cs_c );
	}
	return fail( "b200_dist_gemm_1d: unsupported datatype %d", dt );
}
extern "C" b200_err_t b200_dist_trsm( int dt, int side, int uploa, int transa, int diaga, int root, b200_dim_t m, b200_dim_t n, const void* alpha,
	void* a, b200_inc_t rs_a, b200_inc_t cs_a, void* b_loc, b200_inc_t rs_b, b200_inc_t cs_b )
{
	switch ( dt )
	{
		case B200_FLOAT:    return dist_trsm<float>  ( side, uploa, transa, diaga, root, m, n, (const float*)alpha,   (float*)a,   rs_a, cs_a, (float*)b_loc,   rs_b, cs_b );
		case B200_DOUBLE:   return dist_trsm<double> ( side, uploa, transa, diaga, root, m, n, (const double*)alpha,  (double*)a,  rs_a, cs_a, (double*)b_loc,  rs_b, cs_b );
		case B200_SCOMPLEX: return dist_trsm<float2> ( side, uploa, transa, diaga, root, m, n, (const float2*)alpha,  (float2*)a,  rs_a, cs_a, (float2*)b_loc,  rs_b, cs_b );
		case B200_DCOMPLEX: return dist_trsm<double2>( side, uploa, transa, diaga, root, m, n, (const double2*)alpha, (double2*)a, rs_a, cs_a, (double2*)b_loc, rs_b, cs_b );
	}
	return fail( "b200_dist_trsm: unsupported datatype %d", dt );
}

extern "C" b200_dim_t b200_blksz( int dt, int bs )
{
	auto pick = [&]( int mr, int nr, int mc, int kc, int nc ) -> b200_dim_t
	{
		switch ( bs ) { case B200_BS_MR: return mr; case B200_BS_NR: return nr; case B200_BS_MC: return mc;
		                case B200_BS_KC: return kc; case B200_BS_NC: return nc; }
		return -1;
	};
	switch ( dt )
	{
		case B200_FLOAT:    return pick( Tiles<float>::MR,   Tiles<float>::NR,   Tiles<float>::BQ,   Tiles<float>::BK,   Tiles<float>::BP );
		case B200_DOUBLE:   return pick( Tiles<double>::MR,  Tiles<double>::NR,  Tiles<double>::BQ,  Tiles<double>::BK,  Tiles<double>::BP );
		case B200_SCOMPLEX: return pick( Tiles<float2>::MR,  Tiles<float2>::NR,  Tiles<float2>::BQ,  Tiles<float2>::BK,  Tiles<float2>::BP );
		case B200_DCOMPLEX: return pick( Tiles<double2>::MR, Tiles<double2>::NR, Tiles<double2>::BQ, Tiles<double2>::BK, Tiles<double2>::BP );
	}
	return -1;
}

extern "C" unsigned long long b200_launch_count( void ) { return ctx().launches.load(); }

// Tuning knobs for sweeps (not part of the reference surface).
extern "C" b200_err_t b200_set_option( const char* key, long long value )
{
	Context& c = ctx();
	if      ( !strcmp( key, "dgemm_cfg" ) ) c.dgemm_cfg = (int)value;
	else if ( !strcmp( key, "zgemm_cfg" ) ) c.zgemm_cfg = (int)value;
	else if ( !strcmp( key, "sgemm_cfg" ) ) c.sgemm_cfg = (int)value;
	else if ( !strcmp( key, "cgemm_cfg" ) ) c.cgemm_cfg = (int)value;
	else if ( !strcmp( key, "grid_mult" ) ) c.grid_mult = (int)std::max<long long>( 1, value );
	else if ( !strcmp( key, "dynamic_tiles" ) ) c.dynamic_tiles = (int)value;
	else if ( !strcmp( key, "transpose_y" ) ) c.transpose_y = (int)value;
	else if ( !strcmp( key, "ktri_skip" ) ) c.ktri_skip = (int)value;
	else if ( !strcmp( key, "host_kpipe" ) ) c.host_kpipe = (int)value;
	else if ( !strcmp( key, "host_trace" ) ) c.host_trace = (int)value;
	else if ( !strcmp( key, "dmma_cst" ) ) c.dmma_cst = (int)value;
	else if ( !strcmp( key, "dmma_pp" ) ) c.dmma_pp = (int)value;
	else if ( !strcmp( key, "dgemm_splitk" ) ) c.dgemm_splitk = (int)value;
	else if ( !strcmp( key, "trsm_host_pipe" ) ) c.trsm_host_pipe = (int)value;
	else if ( !strcmp( key, "trsm_host_rb" ) ) c.trsm_host_rb = value;
	else if ( !strcmp( key, "trsm_host_rb_min_m" ) ) c.trsm_host_rb_min_m = std::max<long long>( 0, value );
	else if ( !strcmp( key, "trsm_host_rb_div" ) ) c.trsm_host_rb_div = std::max<long long>( 2, value );
	else if ( !strcmp( key, "batch_grouped" ) ) c.batch_grouped = (int)value;
	else if ( !strcmp( key, "batch_grouped_max" ) ) c.batch_grouped_max = std::max<long long>( 0, value );
	else if ( !strcmp( key, "trsm_fused" ) ) c.trsm_fused = (int)value;
	else if ( !strcmp( key, "dist_ab_static" ) ) dist().ab_static = (int)value;
	else if ( !strcmp( key, "tma_l2_promotion" ) ) c.tma_l2_promotion = (int)std::min<long long>( 3, std::max<long long>( 0, value ) );
	else if ( !strcmp( key, "raster_group" ) ) c.raster_group = (int)std::max<long long>( 1, value );
	else if ( !strcmp( key, "reserve_sms" ) )
	{
		// leave SMs free for concurrently running communication kernels (multi-GPU overlap)
		cudaDeviceProp prop; int dev = 0; cudaGetDevice( &dev ); cudaGetDeviceProperties( &prop, dev );
		c.num_sms = std::max( 1, prop.multiProcessorCount - (int)std::max<long long>( 0, value ) );
	}
	else return fail( "b200_set_option: unknown key %s", key );
	return kSuccess;
}
