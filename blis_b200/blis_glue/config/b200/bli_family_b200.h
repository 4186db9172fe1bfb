/*
   bli_family_b200.h -- family header of the b200 sub-configuration
   (docs/ConfigurationHowTo.md:195-207; defaults in
   frame/include/bli_kernel_macro_defs.h:112-151).
*/
#ifndef BLIS_FAMILY_B200_H
#define BLIS_FAMILY_B200_H

// The registered MR x NR (the warp tile, up to 64 x 32 doubles) must fit the
// microkernel stack buffer that bli_gks_register_cntx checks:
// MR*NR*sizeof(dcomplex) <= BLIS_STACK_BUF_MAX_SIZE (frame/base/bli_check.c:820-844).
#define BLIS_STACK_BUF_MAX_SIZE ( 64 * 64 * 16 )

// Matrices created through bli_obj_create() are allocated in page-locked host
// memory so that the engine's host<->device copies run at full PCIe speed
// without the extra pinned staging pass (frame/base/bli_obj.c:190).
void* b200_malloc_pinned( size_t size );
void  b200_free_pinned( void* p );
#define BLIS_MALLOC_USER b200_malloc_pinned
#define BLIS_FREE_USER   b200_free_pinned

// The b200 engine sits in the sup-handler slot: sup handling must stay enabled.
#ifdef BLIS_DISABLE_SUP_HANDLING
#error "config/b200 requires sup handling (do not configure with --disable-sup-handling)"
#endif

#endif
