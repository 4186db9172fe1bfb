// common.cuh -- device helpers shared by the B200 level-3 kernels.
//
// Target: sm_100a only (nvcc -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// BLIS err_t values (frame/include/bli_type_defs.h:1502-1607).
constexpr int kSuccess = -1;
constexpr int kFailure = -2;

// ---- element traits ---------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float>   { using real = float;  static constexpr bool cplx = false; };
template <> struct Elem<double>  { using real = double; static constexpr bool cplx = false; };
template <> struct Elem<float2>  { using real = float;  static constexpr bool cplx = true;  };
template <> struct Elem<double2> { using real = double; static constexpr bool cplx = true;  };

// ---- shared memory / cp.async ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32( const void* p )
{
	return (uint32_t)__cvta_generic_to_shared( p );
}

// Asynchronous global->shared copy of CPB bytes of which the first src_bytes
// come from memory and the rest is zero-filled (LDGSTS in SASS).
template <int CPB>
__device__ __forceinline__ void cp_async( uint32_t dst, const void* src, int src_bytes )
{
	static_assert( CPB == 4 || CPB == 8 || CPB == 16, "cp.async size" );
	if constexpr ( CPB == 16 )
		asm volatile( "cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
		              :: "r"(dst), "l"(src), "r"(src_bytes) : "memory" );
	else if constexpr ( CPB == 8 )
		asm volatile( "cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
		              :: "r"(dst), "l"(src), "r"(src_bytes) : "memory" );
	else
		asm volatile( "cp.async.ca.shared.global [%0], [%1], 4, %2;\n"
		              :: "r"(dst), "l"(src), "r"(src_bytes) : "memory" );
}
__device__ __forceinline__ void cp_async_commit()
{
	asm volatile( "cp.async.commit_group;\n" ::: "memory" );
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
	asm volatile( "cp.async.wait_group %0;\n" :: "n"(N) : "memory" );
}

// ---- FP64 tensor-core tile ----------------------------------------------------
// D(8x8) += A(8x4) * B(4x8); lane = 4*g + t holds A[g][t], B[t][g] and
// D[g][2t], D[g][2t+1].  SASS: DMMA.8x8x4 (the only FP64 MMA shape sm_100 has;
// the m16n8k{4,8,16} PTX forms are split into it by ptxas).
__device__ __forceinline__ void dmma884( double& d0, double& d1, double a, double b )
{
	asm( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
	     : "+d"(d0), "+d"(d1) : "d"(a), "d"(b) );
}

// Flip the sign of x when s is true (used for run-time conjugation).
__device__ __forceinline__ double flip_sign( double x, bool s ) { return s ? -x : x; }
__device__ __forceinline__ float  flip_sign( float  x, bool s ) { return s ? -x : x; }

// ---- complex epilogue arithmetic with a FIXED rounding order --------------------
// ab *= alpha (bli_tscals) and c := ab + beta*c (bli_txpbys).  Written with explicit mul/fma so that the
// compiler cannot contract the expressions differently in different epilogue paths: the vectorised interior path,
// the bounds-checked edge path and the triangular (gemmt) path of every kernel round identically.
__device__ __forceinline__ double mul_rn( double a, double b ) { return __dmul_rn( a, b ); }
__device__ __forceinline__ float  mul_rn( float  a, float  b ) { return __fmul_rn( a, b ); }
template <typename R>
__device__ __forceinline__ void cscal( R ar, R ai, R xr, R xi, R& rr, R& ri )      // (ar + i ai) * (xr + i xi)
{
	rr = fma( -ai, xi, mul_rn( ar, xr ) );
	ri = fma(  ai, xr, mul_rn( ar, xi ) );
}
template <typename R>
__device__ __forceinline__ void cxpby( R br, R bi, R yr, R yi, R& rr, R& ri )      // (rr + i ri) += (br + i bi) * (yr + i yi)
{
	rr = fma( br, yr, rr ); rr = fma( -bi, yi, rr );
	ri = fma( br, yi, ri ); ri = fma(  bi, yr, ri );
}

// ---- complex arithmetic in the reference's operation order -------------------
// (frame/include/level0: bli_tdots / bli_tscals / bli_txpbys for c,z)
template <typename R> struct Cx { R r, i; };

} // namespace b200
