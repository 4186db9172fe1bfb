/* ref_shim.c -- TEST INFRASTRUCTURE: calls the REAL reference's object API for operations that have no typed API
   (mixed-datatype gemm: bli_gemm on objects of different datatypes with an explicit computation precision).
   Compiled by tests/glue_build.py against the reference headers into oracle/_ref/libref_shim.so. */
#include "blis.h"

void ref_gemm_md( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb, dim_t m, dim_t n, dim_t k,
                  double* alpha, void* a, inc_t rs_a, inc_t cs_a, void* b, inc_t rs_b, inc_t cs_b,
                  double* beta, void* c, inc_t rs_c, inc_t cs_c )
{
	obj_t ao, bo, co, alphao, betao;
	const dim_t ma = ( transa & BLIS_TRANS_BIT ) ? k : m, na = ( transa & BLIS_TRANS_BIT ) ? m : k;
	const dim_t mb = ( transb & BLIS_TRANS_BIT ) ? n : k, nb = ( transb & BLIS_TRANS_BIT ) ? k : n;
	bli_init();
	bli_obj_create_with_attached_buffer( ( num_t )dt_a, ma, na, a, rs_a, cs_a, &ao );
	bli_obj_create_with_attached_buffer( ( num_t )dt_b, mb, nb, b, rs_b, cs_b, &bo );
	bli_obj_create_with_attached_buffer( ( num_t )dt_c, m,  n,  c, rs_c, cs_c, &co );
	bli_obj_set_conjtrans( ( trans_t )transa, &ao );
	bli_obj_set_conjtrans( ( trans_t )transb, &bo );
	bli_obj_set_comp_prec( ( prec_t )comp_prec, &co );
	bli_obj_create_1x1_with_attached_buffer( BLIS_DCOMPLEX, alpha, &alphao );
	bli_obj_create_1x1_with_attached_buffer( BLIS_DCOMPLEX, beta,  &betao );
	bli_gemm( &alphao, &ao, &bo, &betao, &co );
}
