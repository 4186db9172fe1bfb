/*
   bli_kernel_defs_b200.h -- compile-time register blocksizes for the reference
   kernels compiled for the b200 sub-configuration
   (config/haswell/bli_kernel_defs_haswell.h is the pattern).

   The b200 context registers the ENGINE's tile shapes as MR/NR (bli_cntx_init_b200.c), which are far larger than
   anything the reference kernels unroll for.  -1 makes every reference kernel that is still in a slot (trsm,
   gemmtrsm, packm) take its general-purpose form, which reads MR/NR/PACKMR/PACKNR from the context at run time
   (ref_kernels/3/bli_gemm_ref.c:157-186: "If compile-time MR/NR are not available (indicated by BLIS_[MN]R_x = -1),
   then the non-unrolled version is used"; bli_trsm_ref.c:59-72 and bli_gemmtrsm_ref.c:67-73 always query the context).
   Without this the reference gemm kernel would be compiled for the defaults MR_d = 4, NR_d = 8
   (frame/include/bli_kernel_macro_defs.h:260-353) while packm and the macrokernel use the context's tile sizes.
*/
#ifndef BLIS_KERNEL_DEFS_B200_H
#define BLIS_KERNEL_DEFS_B200_H

#define BLIS_MR_s -1
#define BLIS_MR_d -1
#define BLIS_MR_c -1
#define BLIS_MR_z -1

#define BLIS_NR_s -1
#define BLIS_NR_d -1
#define BLIS_NR_c -1
#define BLIS_NR_z -1

#endif
