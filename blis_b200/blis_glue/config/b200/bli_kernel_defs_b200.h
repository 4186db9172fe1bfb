/*
   bli_kernel_defs_b200.h -- compile-time register blocksizes for the reference
   kernels compiled for the b200 sub-configuration
   (config/haswell/bli_kernel_defs_haswell.h is the pattern).  The reference
   kernels only serve out-of-scope operations here, so the defaults of
   frame/include/bli_kernel_macro_defs.h:260-353 are kept.
*/
#ifndef BLIS_KERNEL_DEFS_B200_H
#define BLIS_KERNEL_DEFS_B200_H
#endif
