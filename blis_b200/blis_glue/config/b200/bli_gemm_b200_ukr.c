/*
   bli_gemm_b200_ukr.c -- the BLIS_GEMM_UKR slot of the b200 context, backed by the engine.

   A BLIS microkernel is a host-callable function invoked once per MR x NR tile on PACKED micropanels
   (frame/3/bli_l3_ukr_ft.h:46-57, frame/3/gemm/bli_gemm_ker_var2.c:258).  With the b200 configuration no
   level-3 operation of the library gets that far -- the bli_<op>_ex entry points hand whole problems to the engine --
   but the slot is still reachable: applications and the reference testsuite call it directly
   (bli_gemm_ukernel, testsuite/src/test_gemm_ukr.c), the reference gemmtrsm kernel obtains its gemm part from it
   (ref_kernels/3/bli_gemmtrsm_ref.c:76-118), and bli_info reports it (frame/base/bli_gks.c:399-441).
   These functions make that slot truthful: c := beta*c + alpha * a * b with
       a : MR x k micropanel, column-stored, leading dimension PACKMR     (bli_packm_cxk_ref.c:80-145)
       b : k x NR micropanel, row-stored,    leading dimension PACKNR
   is one b200_gemm call (m <= MR, n <= NR edge tiles included; beta == 0 never reads c).  It is correct for any
   registered MR/NR; it is not fast (one PCIe round trip per tile) and not meant to be: the fast path is the
   whole-operation one.
*/
#include "blis.h"
#include "blis_b200.h"

#define BLI_B200_GEMM_UKR( ch, dtconst ) \
void bli_##ch##gemm_b200_ukr \
     ( \
             dim_t      m, \
             dim_t      n, \
             dim_t      k, \
       const void*      alpha, \
       const void*      a, \
       const void*      b, \
       const void*      beta, \
             void*      c, inc_t rs_c, inc_t cs_c, \
       const auxinfo_t* data, \
       const cntx_t*    cntx  \
     ) \
{ \
	( void )data; \
	const inc_t packmr = bli_cntx_get_blksz_max_dt( dtconst, BLIS_MR, cntx ); \
	const inc_t packnr = bli_cntx_get_blksz_max_dt( dtconst, BLIS_NR, cntx ); \
	const err_t r = b200_gemm( ( int )dtconst, B200_NO_TRANSPOSE, B200_NO_TRANSPOSE, m, n, k, \
	                           alpha, a, 1, packmr, b, packnr, 1, beta, c, rs_c, cs_c ); \
	if ( r != BLIS_SUCCESS ) \
	{ \
		fprintf( stderr, "libblis (b200): gemm microkernel failed: %s\n", b200_last_error() ); \
		bli_abort(); \
	} \
}

BLI_B200_GEMM_UKR( s, BLIS_FLOAT )
BLI_B200_GEMM_UKR( d, BLIS_DOUBLE )
BLI_B200_GEMM_UKR( c, BLIS_SCOMPLEX )
BLI_B200_GEMM_UKR( z, BLIS_DCOMPLEX )
