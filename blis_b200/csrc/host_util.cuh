// host_util.cuh -- scalar traits, strided copy / scale / transpose kernels
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "gemm_launch.cuh"
#include <vector>
namespace b200 {

// ---- small helpers -------------------------------------------------------------
template <typename T> struct Scalar;
template <> struct Scalar<float>
{
	static float   make( double r, double ) { return (float)r; }
	static bool    is_zero( float a ) { return a == 0.0f; }
	static bool    is_one( float a )  { return a == 1.0f; }
};
template <> struct Scalar<double>
{
	static double  make( double r, double ) { return r; }
	static bool    is_zero( double a ) { return a == 0.0; }
	static bool    is_one( double a )  { return a == 1.0; }
};
template <> struct Scalar<float2>
{
	static float2  make( double r, double i ) { return make_float2( (float)r, (float)i ); }
	static bool    is_zero( float2 a ) { return a.x == 0.0f && a.y == 0.0f; }
	static bool    is_one( float2 a )  { return a.x == 1.0f && a.y == 0.0f; }
};
template <> struct Scalar<double2>
{
	static double2 make( double r, double i ) { return make_double2( r, i ); }
	static bool    is_zero( double2 a ) { return a.x == 0.0 && a.y == 0.0; }
	static bool    is_one( double2 a )  { return a.x == 1.0 && a.y == 0.0; }
};

// ---- strided copy / scale kernels (component-wise, so complex data only
//      needs the alignment of its real type) -----------------------------------
template <typename R, int NC>
__global__ void copy2d_kernel( R* __restrict__ dst, int64_t rsd, int64_t csd,
                               const R* __restrict__ src, int64_t rss, int64_t css,
                               int64_t m, int64_t n, int inner_is_row, int tri )
{
	const int64_t total = m * n;
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_is_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		if ( ( tri == 1 && i < j ) || ( tri == 2 && i > j ) ) continue;      // only the stored triangle (1: lower, 2: upper)
		const R* s = src + ( i * rss + j * css ) * NC;
		R*       d = dst + ( i * rsd + j * csd ) * NC;
		#pragma unroll
		for ( int c = 0; c < NC; ++c ) d[c] = s[c];
	}
}

// C := beta * C  (beta == 0 stores zeros without reading C: bli_scalm / bli_setm)
template <typename R, int NC>
__global__ void scal2d_kernel( R* __restrict__ c, int64_t rs, int64_t cs, int64_t m, int64_t n,
                               R br, R bi, int beta_is_zero, int inner_is_row, int tri )
{
	const int64_t total = m * n;
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_is_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		if ( ( tri == 1 && i < j ) || ( tri == 2 && i > j ) ) continue;      // only the stored triangle (1: lower, 2: upper)
		R* p = c + ( i * rs + j * cs ) * NC;
		if ( beta_is_zero ) { for ( int q = 0; q < NC; ++q ) p[q] = (R)0; }
		else if ( NC == 1 ) p[0] = br * p[0];
		else { const R xr = p[0], xi = p[1]; p[0] = br * xr - bi * xi; p[1] = br * xi + bi * xr; }
	}
}

static inline int64_t iabs64( int64_t x ) { return x < 0 ? -x : x; }

// dst[c*ldd + r] = src[r*lds + c]: 32 x 32 tiles through shared memory, both sides coalesced.
template <typename T>
__global__ void __launch_bounds__( 256 ) transpose2d_kernel( T* __restrict__ dst, int64_t ldd, const T* __restrict__ src, int64_t lds,
                                                             int64_t R, int64_t Cn, int tiles_c )
{
	__shared__ T tile[32][33];
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const int64_t ntiles = ( ( R + 31 ) / 32 ) * tiles_c;
	for ( int64_t t = blockIdx.x; t < ntiles; t += gridDim.x )
	{
		const int64_t r0 = ( t / tiles_c ) * 32, c0 = ( t % tiles_c ) * 32;
		#pragma unroll
		for ( int i = 0; i < 4; ++i )
		{
			const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
			if ( r < R && c < Cn ) tile[ty + 8 * i][tx] = src[r * lds + c];
		}
		__syncthreads();
		#pragma unroll
		for ( int i = 0; i < 4; ++i )
		{
			const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
			if ( r < R && c < Cn ) dst[c * ldd + r] = tile[tx][ty + 8 * i];
		}
		__syncthreads();
	}
}

template <typename T>
static int transpose2d( T* dst, int64_t ldd, const T* src, int64_t lds, int64_t R, int64_t Cn, cudaStream_t st )
{
	const int64_t tiles_c = ( Cn + 31 ) / 32, ntiles = ( ( R + 31 ) / 32 ) * tiles_c;
	if ( ntiles <= 0 ) return kSuccess;
	if ( tiles_c >= ( 1ll << 31 ) ) return fail( "transpose2d: matrix too wide" );
	const int blocks = (int)std::min<int64_t>( ntiles, (int64_t)ctx().num_sms * 32 );
	transpose2d_kernel<T><<<blocks, 256, 0, st>>>( dst, ldd, src, lds, R, Cn, (int)tiles_c );
	B200_CUDA( cudaGetLastError() );
	note_launch( "transpose2d_kernel" );
	return kSuccess;
}

template <typename T>
static int copy2d( T* dst, int64_t rsd, int64_t csd, const T* src, int64_t rss, int64_t css,
                   int64_t m, int64_t n, cudaStream_t st, int uplo = 0 )
{
	if ( m <= 0 || n <= 0 ) return kSuccess;
	const int tri = ( uplo == B200_LOWER ) ? 1 : ( uplo == B200_UPPER ) ? 2 : 0;
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	const int inner_is_row = ( iabs64( rss ) + iabs64( rsd ) <= iabs64( css ) + iabs64( csd ) );
	const int64_t total = m * n;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	copy2d_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)dst, rsd, csd, (const R*)src, rss, css, m, n, inner_is_row, tri );
	B200_CUDA( cudaGetLastError() );
	note_launch( "copy2d_kernel" );
	return kSuccess;
}

template <typename T>
static int scal2d( T* c, int64_t rs, int64_t cs, int64_t m, int64_t n, T beta, cudaStream_t st, int uplo = 0 )
{
	if ( m <= 0 || n <= 0 || Scalar<T>::is_one( beta ) ) return kSuccess;
	const int tri = ( uplo == B200_LOWER ) ? 1 : ( uplo == B200_UPPER ) ? 2 : 0;
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	R br, bi;
	if constexpr ( Elem<T>::cplx ) { br = beta.x; bi = beta.y; } else { br = beta; bi = 0; }
	const int inner_is_row = ( iabs64( rs ) <= iabs64( cs ) );
	const int64_t total = m * n;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	scal2d_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)c, rs, cs, m, n, br, bi, Scalar<T>::is_zero( beta ) ? 1 : 0, inner_is_row, tri );
	B200_CUDA( cudaGetLastError() );
	note_launch( "scal2d_kernel" );
	return kSuccess;
}

// Tile shapes per datatype = the "blocksizes" this engine registers
// (MR/NR become the warp tile, MC/NC the CTA tile, KC the staged k slab).
template <typename T> struct Tiles;
template <> struct Tiles<double>  { static constexpr int BP = 128, BQ = 128, BK = 16, MR = 32, NR = 64; };
template <> struct Tiles<double2> { static constexpr int BP = 64,  BQ = 128, BK = 8,  MR = 32, NR = 32; };
template <> struct Tiles<float>   { static constexpr int BP = 128, BQ = 128, BK = 16, MR = 8,  NR = 8;  };
template <> struct Tiles<float2>  { static constexpr int BP = 64,  BQ = 128, BK = 16, MR = 4,  NR = 8;  };

} // namespace b200
