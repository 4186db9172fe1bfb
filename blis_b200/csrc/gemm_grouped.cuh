// gemm_grouped.cuh -- ONE launch for a whole batch of small gemm problems.
//
// The reference's ?gemm_batch_ (frame/compat/extra/bla_gemm_batch.c:62-131) loops over the groups and calls bli_?gemm for
// every problem, one after the other; test/test_gemm_batch.c times exactly that.  On the GPU a problem of 32^3 or 64^3 is
// far too small for a launch of its own (a launch costs more than the arithmetic, and one problem cannot occupy 148 SMs),
// so the device-resident small problems of a batch are described by an array of GroupProb records in HBM and ONE
// persistent kernel walks the concatenated list of their 32 x 32 output tiles: a CTA finds the problem of its tile by
// binary search over the tile prefix, stages 32-wide k slabs of op(A) and op(B) through shared memory straight from the
// caller's strides (any rs/cs, transposition = swapped strides, conjugation applied while staging, zero fill at the
// edges: what packm does for the reference's microkernel) and keeps a 2 x 2 register tile per thread.
// Arithmetic as in the reference microkernel (ref_kernels/3/bli_gemm_ref.c:250-314): ab = sum_l a*b in k order, ab *= alpha,
// C := beta*C + ab, beta == 0 never reads C; alpha == 0 arrives here as k = 0 (C := beta*C without touching A or B).
// Plain FMA pipes (DFMA / FFMA), no tensor cores: these problems are latency bound, not throughput bound; anything
// larger goes to the stream pool and the tiled kernels (host_batch.cuh).
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T>
struct GroupProb
{
	const T* a; const T* b; T* c;
	int64_t  rs_a, cs_a, rs_b, cs_b, rs_c, cs_c;      // of op(A) (m x k), op(B) (k x n), C (m x n)
	T        alpha, beta;
	int      m, n, k;
	int      conja, conjb, beta_is_zero;
	int      tile0, tiles_n;                          // first tile of this problem in the batch-wide list; tiles per row of tiles
};

template <typename T> struct GOps;
template <> struct GOps<float>
{
	static __device__ __forceinline__ float  zero() { return 0.f; }
	static __device__ __forceinline__ float  conj( float x, int ) { return x; }
	static __device__ __forceinline__ void   mac( float& acc, float a, float b ) { acc = fmaf( a, b, acc ); }
	static __device__ __forceinline__ float  finish( float alpha, float ab, float beta, const float* c, int bz )
	{ float r = alpha * ab; if ( !bz ) r = fmaf( beta, *c, r ); return r; }
};
template <> struct GOps<double>
{
	static __device__ __forceinline__ double zero() { return 0.0; }
	static __device__ __forceinline__ double conj( double x, int ) { return x; }
	static __device__ __forceinline__ void   mac( double& acc, double a, double b ) { acc = fma( a, b, acc ); }
	static __device__ __forceinline__ double finish( double alpha, double ab, double beta, const double* c, int bz )
	{ double r = alpha * ab; if ( !bz ) r = fma( beta, *c, r ); return r; }
};
template <> struct GOps<float2>
{
	static __device__ __forceinline__ float2 zero() { return make_float2( 0.f, 0.f ); }
	static __device__ __forceinline__ float2 conj( float2 x, int cj ) { return cj ? make_float2( x.x, -x.y ) : x; }
	static __device__ __forceinline__ void   mac( float2& acc, float2 a, float2 b )
	{ acc.x = fmaf( a.x, b.x, acc.x ); acc.x = fmaf( -a.y, b.y, acc.x ); acc.y = fmaf( a.y, b.x, acc.y ); acc.y = fmaf( a.x, b.y, acc.y ); }
	static __device__ __forceinline__ float2 finish( float2 alpha, float2 ab, float2 beta, const float2* c, int bz )
	{
		float rr, ri; cscal( alpha.x, alpha.y, ab.x, ab.y, rr, ri );
		if ( !bz ) { const float2 o = *c; cxpby( beta.x, beta.y, o.x, o.y, rr, ri ); }
		return make_float2( rr, ri );
	}
};
template <> struct GOps<double2>
{
	static __device__ __forceinline__ double2 zero() { return make_double2( 0.0, 0.0 ); }
	static __device__ __forceinline__ double2 conj( double2 x, int cj ) { return cj ? make_double2( x.x, -x.y ) : x; }
	static __device__ __forceinline__ void    mac( double2& acc, double2 a, double2 b )
	{ acc.x = fma( a.x, b.x, acc.x ); acc.x = fma( -a.y, b.y, acc.x ); acc.y = fma( a.y, b.x, acc.y ); acc.y = fma( a.x, b.y, acc.y ); }
	static __device__ __forceinline__ double2 finish( double2 alpha, double2 ab, double2 beta, const double2* c, int bz )
	{
		double rr, ri; cscal( alpha.x, alpha.y, ab.x, ab.y, rr, ri );
		if ( !bz ) { const double2 o = *c; cxpby( beta.x, beta.y, o.x, o.y, rr, ri ); }
		return make_double2( rr, ri );
	}
};

constexpr int kGroupTile = 32;        // output tile and k slab

template <typename T>
__global__ void __launch_bounds__( 256 )
gemm_grouped_kernel( const GroupProb<T>* __restrict__ probs, int nprob, int total_tiles )
{
	constexpr int TS = kGroupTile;
	__shared__ T As[TS][TS + 1];        // As[l][i] = op(A)(i0 + i, l0 + l)
	__shared__ T Bs[TS][TS + 1];        // Bs[l][j] = op(B)(l0 + l, j0 + j)
	const int tid = threadIdx.x;
	const int tx = tid & 15, ty = tid >> 4;           // rows 2tx, 2tx+1; columns 2ty, 2ty+1

	for ( int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x )
	{
		int lo = 0, hi = nprob - 1;
		while ( lo < hi )
		{
			const int mid = ( lo + hi + 1 ) >> 1;
			if ( probs[mid].tile0 <= tile ) lo = mid; else hi = mid - 1;
		}
		const GroupProb<T>* __restrict__ P = probs + lo;
		const int m = P->m, n = P->n, k = P->k;
		const int t = tile - P->tile0;
		const int i0 = ( t / P->tiles_n ) * TS, j0 = ( t % P->tiles_n ) * TS;
		const T* __restrict__ a = P->a; const T* __restrict__ b = P->b;
		const int64_t rs_a = P->rs_a, cs_a = P->cs_a, rs_b = P->rs_b, cs_b = P->cs_b;
		const int cja = P->conja, cjb = P->conjb;

		T acc[2][2] = { { GOps<T>::zero(), GOps<T>::zero() }, { GOps<T>::zero(), GOps<T>::zero() } };
		for ( int l0 = 0; l0 < k; l0 += TS )
		{
			// stage: thread e covers (x = e % 32, y = e / 32); the index that is contiguous in memory runs along x
			const bool a_rows_fast = ( rs_a == 1 || cs_a != 1 ), b_cols_fast = ( cs_b == 1 || rs_b != 1 );
			#pragma unroll
			for ( int r = 0; r < TS * TS / 256; ++r )
			{
				const int e = tid + r * 256, x = e & ( TS - 1 ), y = e / TS;
				{
					const int i = a_rows_fast ? x : y, l = a_rows_fast ? y : x;
					T v = GOps<T>::zero();
					if ( i0 + i < m && l0 + l < k ) v = GOps<T>::conj( a[( i0 + i ) * rs_a + ( l0 + l ) * cs_a], cja );
					As[l][i] = v;
				}
				{
					const int j = b_cols_fast ? x : y, l = b_cols_fast ? y : x;
					T v = GOps<T>::zero();
					if ( j0 + j < n && l0 + l < k ) v = GOps<T>::conj( b[( l0 + l ) * rs_b + ( j0 + j ) * cs_b], cjb );
					Bs[l][j] = v;
				}
			}
			__syncthreads();
			#pragma unroll 8
			for ( int l = 0; l < TS; ++l )
			{
				const T a0 = As[l][2 * tx], a1 = As[l][2 * tx + 1], b0 = Bs[l][2 * ty], b1 = Bs[l][2 * ty + 1];
				GOps<T>::mac( acc[0][0], a0, b0 ); GOps<T>::mac( acc[0][1], a0, b1 );
				GOps<T>::mac( acc[1][0], a1, b0 ); GOps<T>::mac( acc[1][1], a1, b1 );
			}
			__syncthreads();
		}
		const T alpha = P->alpha, beta = P->beta;
		const int bz = P->beta_is_zero;
		T* __restrict__ c = P->c; const int64_t rs_c = P->rs_c, cs_c = P->cs_c;
		#pragma unroll
		for ( int di = 0; di < 2; ++di )
			#pragma unroll
			for ( int dj = 0; dj < 2; ++dj )
			{
				const int i = i0 + 2 * tx + di, j = j0 + 2 * ty + dj;
				if ( i < m && j < n )
				{
					T* cp = c + i * rs_c + j * cs_c;
					*cp = GOps<T>::finish( alpha, acc[di][dj], beta, cp, bz );
				}
			}
	}
}

} // namespace b200
