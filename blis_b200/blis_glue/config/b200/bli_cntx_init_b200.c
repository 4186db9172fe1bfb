/*
   bli_cntx_init_b200.c -- context initialisation of the b200 sub-configuration.

   Pattern: config/zen3/bli_cntx_init_zen3.c:37-258 -- start from the reference
   context, then override.  Called by bli_gks_init through
   bli_gks_register_cntx( BLIS_ARCH_B200, bli_cntx_init_b200, bli_cntx_init_b200_ref )
   (frame/base/bli_gks.c:58-93,174-272).

   What is registered:
     * blocksizes MR/NR/MC/KC/NC = the engine's warp/CTA tile shapes
       (b200_blksz); they satisfy the invariants bli_gks_register_cntx checks
       (MC%MR = NC%NR = MC%NR = NC%MR = 0; frame/base/bli_gks.c:248-271) and
       bli_family_b200.h raises BLIS_STACK_BUF_MAX_SIZE for the MR x NR tile.
       bli_info / bli_cntx queries therefore describe the GPU tiling truthfully.
     * the whole-operation gemm handler + thresholds (bli_b200_install).
     * the BLIS_GEMM_UKR slots: engine-backed microkernels (bli_gemm_b200_ukr.c), so that bli_info, direct
       microkernel calls (bli_gemm_ukernel, the testsuite's gemm_ukr / gemmtrsm_ukr modules) and the reference
       gemmtrsm kernel's gemm part all end in the engine.  The trsm/gemmtrsm/packm slots keep the reference kernels in
       their run-time-blocksize form (bli_kernel_defs_b200.h sets BLIS_MR_x/BLIS_NR_x to -1), which is consistent with
       whatever MR/NR the context holds.
   No level-3 operation of a b200 build reaches a control tree: the bli_<op>_ex entry points of the library are the
   glue's (frame/3/bli_l3_oapi_ex.c steps aside under BLIS_CONFIG_B200).
*/
#include "blis.h"
#include "blis_b200.h"

void bli_b200_install( cntx_t* cntx );

// engine-backed gemm microkernels (bli_gemm_b200_ukr.c); signature of gemm_ukr_ft (frame/3/bli_l3_ukr_ft.h:46-57)
#define BLI_B200_UKR_PROT( ch ) \
void bli_##ch##gemm_b200_ukr( dim_t m, dim_t n, dim_t k, const void* alpha, const void* a, const void* b, const void* beta, \
                              void* c, inc_t rs_c, inc_t cs_c, const auxinfo_t* data, const cntx_t* cntx );
BLI_B200_UKR_PROT( s ) BLI_B200_UKR_PROT( d ) BLI_B200_UKR_PROT( c ) BLI_B200_UKR_PROT( z )

void bli_cntx_init_b200( cntx_t* cntx )
{
	blksz_t blkszs[ BLIS_NUM_BLKSZS ];

	// Set default kernel blocksizes and functions.
	bli_cntx_init_b200_ref( cntx );

	// -------------------------------------------------------------------------

	// Tile shapes of the engine, queried from the library so that the numbers
	// registered here can never drift from the kernels'.
	//                                                      s                          d                          c                          z
	bli_blksz_init_easy( &blkszs[ BLIS_MR ], b200_blksz( 0, B200_BS_MR ), b200_blksz( 2, B200_BS_MR ), b200_blksz( 1, B200_BS_MR ), b200_blksz( 3, B200_BS_MR ) );
	bli_blksz_init_easy( &blkszs[ BLIS_NR ], b200_blksz( 0, B200_BS_NR ), b200_blksz( 2, B200_BS_NR ), b200_blksz( 1, B200_BS_NR ), b200_blksz( 3, B200_BS_NR ) );
	bli_blksz_init_easy( &blkszs[ BLIS_MC ], b200_blksz( 0, B200_BS_MC ), b200_blksz( 2, B200_BS_MC ), b200_blksz( 1, B200_BS_MC ), b200_blksz( 3, B200_BS_MC ) );
	bli_blksz_init_easy( &blkszs[ BLIS_KC ], b200_blksz( 0, B200_BS_KC ), b200_blksz( 2, B200_BS_KC ), b200_blksz( 1, B200_BS_KC ), b200_blksz( 3, B200_BS_KC ) );
	bli_blksz_init_easy( &blkszs[ BLIS_NC ], b200_blksz( 0, B200_BS_NC ), b200_blksz( 2, B200_BS_NC ), b200_blksz( 1, B200_BS_NC ), b200_blksz( 3, B200_BS_NC ) );

	bli_cntx_set_blkszs
	(
	  cntx,

	  // level-3
	  BLIS_NC, &blkszs[ BLIS_NC ], BLIS_NR,
	  BLIS_KC, &blkszs[ BLIS_KC ], BLIS_KR,
	  BLIS_MC, &blkszs[ BLIS_MC ], BLIS_MR,
	  BLIS_NR, &blkszs[ BLIS_NR ], BLIS_NR,
	  BLIS_MR, &blkszs[ BLIS_MR ], BLIS_MR,

	  BLIS_VA_END
	);

	// -------------------------------------------------------------------------

	// Engine-backed gemm microkernels.  They accept any storage of C; "column preference" mirrors the engine's
	// own choice (D = C^T for column-stored C costs nothing).
	bli_cntx_set_ukrs
	(
	  cntx,
	  BLIS_GEMM_UKR, BLIS_FLOAT,    bli_sgemm_b200_ukr,
	  BLIS_GEMM_UKR, BLIS_DOUBLE,   bli_dgemm_b200_ukr,
	  BLIS_GEMM_UKR, BLIS_SCOMPLEX, bli_cgemm_b200_ukr,
	  BLIS_GEMM_UKR, BLIS_DCOMPLEX, bli_zgemm_b200_ukr,
	  BLIS_VA_END
	);
	bli_cntx_set_ukr_prefs
	(
	  cntx,
	  BLIS_GEMM_UKR_ROW_PREF, BLIS_FLOAT,    FALSE,
	  BLIS_GEMM_UKR_ROW_PREF, BLIS_DOUBLE,   FALSE,
	  BLIS_GEMM_UKR_ROW_PREF, BLIS_SCOMPLEX, FALSE,
	  BLIS_GEMM_UKR_ROW_PREF, BLIS_DCOMPLEX, FALSE,
	  BLIS_VA_END
	);

	// -------------------------------------------------------------------------

	// Whole-operation gemm handler and its thresholds.
	bli_b200_install( cntx );
}
