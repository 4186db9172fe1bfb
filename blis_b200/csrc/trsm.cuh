// trsm.cuh -- diagonal-block triangular solve (the "gemmtrsm" base case).
//
// Replaces, for one NB x NB diagonal block of A and all n right-hand sides:
//   bli_trsm_ll_ker_var2 / _lu_   frame/3/trsm/bli_trsm_l{l,u}_ker_var2.c:38-335
//   gemmtrsm / trsm microkernels  ref_kernels/3/bli_gemmtrsm_ref.c:43-196,
//                                 ref_kernels/3/bli_trsm_ref.c:44-128,140-224
//   packm of the triangular block frame/1m/packm/bli_packm_struc_cxk.c:155-301,
//                                 ref_kernels/1m/bli_packm_cxc_diag_ref.c:161-236
//
// What is kept from the reference: the diagonal is PRE-INVERTED while the
// block is staged (BLIS_ENABLE_TRSM_PREINVERSION, build/bli_config.h.in:175;
// bli_trsm_ref.c:130-134 multiplies by 1/alpha11), a unit diagonal is stored
// as one, only the stored triangle of A is ever read, conjugation is applied
// while staging, and row i is  x_i = ( alpha*b_i - sum_{l<i} a_il x_l ) * inv(a_ii)
// evaluated in increasing l.  The reference's parallelism is also kept: columns
// of B are independent (jr loop), rows are sequential.
//
// B200 mapping: the block of A lives in shared memory (broadcast reads), each
// thread owns one right-hand-side column whose solved entries stay in
// registers; upper-triangular blocks are solved as lower ones by reversing
// the index order while staging.  The rank-k updates between blocks are done
// by the gemm kernels (see b200_trsm in capi.cu).
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T>
struct TrsmBaseArgs
{
	const T* A;  int64_t rs_a, cs_a;   // diagonal block, mb x mb
	T*       B;  int64_t rs_b, cs_b;   // mb x n
	int64_t  n;
	int      mb;
	int      upper, unit, conj;
	T        alpha;
};

// 1/x with the reference's scaling for complex (bli_tinverts, frame/include/level0).
__device__ __forceinline__ float   recip( float x )  { return 1.0f / x; }
__device__ __forceinline__ double  recip( double x ) { return 1.0 / x; }
__device__ __forceinline__ float2  recip( float2 a )
{
	const float s = fmaxf( fabsf( a.x ), fabsf( a.y ) );
	const float ar = a.x / s, ai = a.y / s;
	const float t = ar * a.x + ai * a.y;
	return make_float2( ar / t, -ai / t );
}
__device__ __forceinline__ double2 recip( double2 a )
{
	const double s = fmax( fabs( a.x ), fabs( a.y ) );
	const double ar = a.x / s, ai = a.y / s;
	const double t = ar * a.x + ai * a.y;
	return make_double2( ar / t, -ai / t );
}

__device__ __forceinline__ float   one_of( float )   { return 1.0f; }
__device__ __forceinline__ double  one_of( double )  { return 1.0; }
__device__ __forceinline__ float2  one_of( float2 )  { return make_float2( 1.f, 0.f ); }
__device__ __forceinline__ double2 one_of( double2 ) { return make_double2( 1.0, 0.0 ); }
__device__ __forceinline__ float   zero_of( float )   { return 0.0f; }
__device__ __forceinline__ double  zero_of( double )  { return 0.0; }
__device__ __forceinline__ float2  zero_of( float2 )  { return make_float2( 0.f, 0.f ); }
__device__ __forceinline__ double2 zero_of( double2 ) { return make_double2( 0.0, 0.0 ); }

__device__ __forceinline__ float   cmul( float a, float b )   { return a * b; }
__device__ __forceinline__ double  cmul( double a, double b ) { return a * b; }
__device__ __forceinline__ float2  cmul( float2 a, float2 b )   { return make_float2( a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x ); }
__device__ __forceinline__ double2 cmul( double2 a, double2 b ) { return make_double2( a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x ); }
// acc - a*b
__device__ __forceinline__ float   msub( float acc, float a, float b )    { return fmaf( -a, b, acc ); }
__device__ __forceinline__ double  msub( double acc, double a, double b ) { return fma( -a, b, acc ); }
__device__ __forceinline__ float2  msub( float2 acc, float2 a, float2 b )
{
	acc.x = fmaf( -a.x, b.x, acc.x ); acc.x = fmaf( a.y, b.y, acc.x );
	acc.y = fmaf( -a.x, b.y, acc.y ); acc.y = fmaf( -a.y, b.x, acc.y );
	return acc;
}
__device__ __forceinline__ double2 msub( double2 acc, double2 a, double2 b )
{
	acc.x = fma( -a.x, b.x, acc.x ); acc.x = fma( a.y, b.y, acc.x );
	acc.y = fma( -a.x, b.y, acc.y ); acc.y = fma( -a.y, b.x, acc.y );
	return acc;
}
__device__ __forceinline__ float   conj_if( float a, bool )   { return a; }
__device__ __forceinline__ double  conj_if( double a, bool )  { return a; }
__device__ __forceinline__ float2  conj_if( float2 a, bool c )  { if ( c ) a.y = -a.y; return a; }
__device__ __forceinline__ double2 conj_if( double2 a, bool c ) { if ( c ) a.y = -a.y; return a; }

template <typename T, int NB, int CN>
constexpr int trsm_base_smem() { return (int)sizeof(T) * ( NB * ( NB + 2 ) + NB * ( CN + 1 ) ); }

// NT threads stage and write back; the first CN threads each solve one column.
// The solve is RIGHT-LOOKING: as soon as x_l is known every remaining row gets its update
// b_i -= a_il * x_l.  Per element this is the reference's sequence (alpha*b_i - a_i0 x_0 - a_i1 x_1 ...,
// increasing l, then * inv(a_ii)), so the bits are those of the row-by-row form, but the NB-l-1 updates
// of a step are independent of each other: the kernel runs at FMA throughput, not at FMA latency.
template <typename T, int NB, int CN, int NT>
__global__ void __launch_bounds__( NT )
trsm_base_kernel( const TrsmBaseArgs<T> a )
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	// At[l][i] = A(i,l): column l of the block is contiguous, rows padded to NB+2 (16-byte aligned pairs)
	T (*At)[NB + 2] = reinterpret_cast<T (*)[NB + 2]>( smem_raw );
	T (*Bs)[CN + 1] = reinterpret_cast<T (*)[CN + 1]>( smem_raw + sizeof(T) * NB * ( NB + 2 ) );

	const int tid = threadIdx.x;
	const int mb  = a.mb;
	const int64_t j0 = (int64_t)blockIdx.x * CN;
	const int nc = (int)min( (int64_t)CN, a.n - j0 );
	const bool a_row_fast = ( a.rs_a <= a.cs_a );
	const bool b_row_fast = ( a.rs_b <= a.cs_b );

	// ---- stage the triangular block: lower form, conj applied, diagonal inverted, identity
	// extension of a ragged block.  Index reversal turns an upper block into a lower one.
	// Loads are issued in batches of 4 so that several global requests are in flight per thread.
	constexpr int A_ITERS = ( NB * NB + NT - 1 ) / NT;
	#pragma unroll 1
	for ( int it = 0; it < A_ITERS; it += 4 )
	{
		T v[4]; int ii[4], ll[4];
		#pragma unroll
		for ( int u = 0; u < 4; ++u )
		{
			const int e = tid + ( it + u ) * NT;
			int i, l;
			if ( a_row_fast ) { i = e % NB; l = e / NB; } else { l = e % NB; i = e / NB; }
			ii[u] = i; ll[u] = ( e < NB * NB ) ? l : -1;
			v[u] = zero_of( T{} );
			if ( e < NB * NB && i < mb && l <= i && !( l == i && a.unit ) )
			{
				const int si = a.upper ? mb - 1 - i : i;
				const int sl = a.upper ? mb - 1 - l : l;
				v[u] = a.A[si * a.rs_a + sl * a.cs_a];
			}
		}
		#pragma unroll
		for ( int u = 0; u < 4; ++u )
		{
			if ( ll[u] < 0 ) continue;
			const int i = ii[u], l = ll[u];
			T w = zero_of( T{} );
			if ( i < mb )
			{
				if ( l < i ) w = conj_if( v[u], a.conj );
				else if ( l == i ) w = a.unit ? one_of( T{} ) : recip( conj_if( v[u], a.conj ) );
			}
			else if ( i == l ) w = one_of( T{} );
			At[l][i] = w;
		}
	}
	// ---- stage the right-hand sides
	constexpr int B_ITERS = ( NB * CN + NT - 1 ) / NT;
	#pragma unroll 1
	for ( int it = 0; it < B_ITERS; it += 4 )
	{
		T v[4]; int ii[4], jj[4];
		#pragma unroll
		for ( int u = 0; u < 4; ++u )
		{
			const int e = tid + ( it + u ) * NT;
			int i, j;
			if ( b_row_fast ) { i = e % NB; j = e / NB; } else { j = e % CN; i = e / CN; }
			ii[u] = i; jj[u] = ( e < NB * CN ) ? j : -1;
			v[u] = zero_of( T{} );
			if ( e < NB * CN && i < mb && j < nc )
			{
				const int si = a.upper ? mb - 1 - i : i;
				v[u] = a.B[si * a.rs_b + ( j0 + j ) * a.cs_b];
			}
		}
		#pragma unroll
		for ( int u = 0; u < 4; ++u ) if ( jj[u] >= 0 ) Bs[ii[u]][jj[u]] = v[u];
	}
	__syncthreads();

	// ---- right-looking substitution, one column per thread, the column lives in registers
	if ( tid < CN )
	{
		T bc[NB];
		#pragma unroll
		for ( int i = 0; i < NB; ++i ) bc[i] = cmul( a.alpha, Bs[i][tid] );
		#pragma unroll
		for ( int l = 0; l < NB; ++l )
		{
			const T x = cmul( bc[l], At[l][l] );
			bc[l] = x;
			#pragma unroll
			for ( int i = l + 1; i < NB; ++i ) bc[i] = msub( bc[i], At[l][i], x );
		}
		#pragma unroll
		for ( int i = 0; i < NB; ++i ) Bs[i][tid] = bc[i];
	}
	__syncthreads();

	// ---- write back
	#pragma unroll 4
	for ( int e = tid; e < NB * CN; e += NT )
	{
		int i, j;
		if ( b_row_fast ) { i = e % NB; j = e / NB; } else { j = e % CN; i = e / CN; }
		if ( i < mb && j < nc )
		{
			const int si = a.upper ? mb - 1 - i : i;
			a.B[si * a.rs_b + ( j0 + j ) * a.cs_b] = Bs[i][j];
		}
	}
}

} // namespace b200
