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
   The gemm/gemmtrsm microkernel slots keep the reference kernels, which serve
   only operations outside this engine's scope (level-1/2, 1m): with the
   handler installed no homogeneous gemm reaches them, and mixed-datatype gemm,
   trsm and the other level-3 operations are taken by the bli_*_ex_b200
   overrides of the glue before any control tree is built.
*/
#include "blis.h"
#include "blis_b200.h"

void bli_b200_install( cntx_t* cntx );

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

	// Whole-operation gemm handler and its thresholds.
	bli_b200_install( cntx );
}
