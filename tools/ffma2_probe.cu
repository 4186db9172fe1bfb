// ffma2_probe.cu -- clocks per FFMA2 (fma.rn.f32x2) per scheduler for the operand patterns of an outer-product sgemm
// register tile (8 scalars x 4 pairs -> 32 accumulator pairs... here 8 x 4 and 8 x 8), no memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ffma2_probe.cu -o /tmp/ffma2_probe && /tmp/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2( float lo, float hi ) { u64 r; asm( "mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi) ); return r; }
__device__ __forceinline__ u64 ffma2( u64 a, u64 b, u64 c ) { u64 d; asm volatile( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c) ); return d; }

__device__ __forceinline__ void ffma2_ip( u64& c, u64 a, u64 b ) { asm volatile( "fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b) ); }
// MODE 3 / 4: as 0 / 1 with the accumulator updated IN PLACE in the PTX (fma d, a, b, d)
// MODE 0: i outer, j inner (scalar reused)   MODE 1: j outer, i inner (pair reused)   MODE 2: pair x pair (no scalar operand)
template <int MODE, int NI, int NJ>
__global__ void probe( const float* in, float* out, long long* clk, int iters )
{
	u64 acc[NI][NJ];
	float x[NI]; u64 y[NJ];
	for ( int i = 0; i < NI; ++i ) x[i] = in[threadIdx.x + i];
	for ( int j = 0; j < NJ; ++j ) y[j] = pack2( in[threadIdx.x + 8 + j], in[threadIdx.x + 40 + j] );
	for ( int i = 0; i < NI; ++i ) for ( int j = 0; j < NJ; ++j ) acc[i][j] = 0ull;
	__syncthreads();
	const long long t0 = clock64();
	for ( int it = 0; it < iters; ++it )
	{
		if ( MODE == 0 )
		{
			#pragma unroll
			for ( int i = 0; i < NI; ++i ) { const u64 x2 = pack2( x[i], x[i] );
				#pragma unroll
				for ( int j = 0; j < NJ; ++j ) acc[i][j] = ffma2( x2, y[j], acc[i][j] ); }
		}
		else if ( MODE == 1 )
		{
			#pragma unroll
			for ( int j = 0; j < NJ; ++j )
				#pragma unroll
				for ( int i = 0; i < NI; ++i ) acc[i][j] = ffma2( pack2( x[i], x[i] ), y[j], acc[i][j] );
		}
		else if ( MODE == 2 )
		{
			#pragma unroll
			for ( int j = 0; j < NJ; ++j )
				#pragma unroll
				for ( int i = 0; i < NI; ++i ) acc[i][j] = ffma2( y[( i + j ) % NJ], y[j], acc[i][j] );
		}
		else if ( MODE == 3 )
		{
			#pragma unroll
			for ( int i = 0; i < NI; ++i ) { const u64 x2 = pack2( x[i], x[i] );
				#pragma unroll
				for ( int j = 0; j < NJ; ++j ) ffma2_ip( acc[i][j], x2, y[j] ); }
		}
		else
		{
			#pragma unroll
			for ( int j = 0; j < NJ; ++j )
				#pragma unroll
				for ( int i = 0; i < NI; ++i ) ffma2_ip( acc[i][j], pack2( x[i], x[i] ), y[j] );
		}
	}
	const long long t1 = clock64();
	__syncthreads();
	float s = 0.f;
	for ( int i = 0; i < NI; ++i ) for ( int j = 0; j < NJ; ++j ) { float lo, hi; asm( "mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i][j]) ); s += lo + hi; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if ( threadIdx.x == 0 && blockIdx.x == 0 ) clk[0] = t1 - t0;
}
template <int MODE, int NI, int NJ> void run( const char* name, int warps, const float* in, float* out, long long* clk )
{
	const int iters = 4000;
	for ( int r = 0; r < 2; ++r ) { probe<MODE, NI, NJ><<<1, warps * 32>>>( in, out, clk, iters ); cudaDeviceSynchronize(); }
	long long c; cudaMemcpy( &c, clk, 8, cudaMemcpyDeviceToHost );
	const double per = (double)c / ( (double)iters * NI * NJ * ( warps / 4.0 ) );
	printf( "%-44s %2d warps  %5.2f clk per FFMA2 per scheduler   %s\n", name, warps, per, cudaGetErrorString( cudaGetLastError() ) );
}
int main()
{
	float *in, *out; long long* clk;
	cudaMalloc( &in, 4096 * 4 ); cudaMemset( in, 0, 4096 * 4 ); cudaMalloc( &out, 4096 * 4 ); cudaMalloc( &clk, 8 );
	for ( int w : { 8 } )
	{
		run<0, 8, 4>( "8x4 pairs, scalar reused (i outer)", w, in, out, clk );
		run<1, 8, 4>( "8x4 pairs, pair reused (j outer)", w, in, out, clk );
		run<2, 8, 4>( "8x4 pairs, pair x pair", w, in, out, clk );
	}
	run<3, 8, 4>( "8x4 pairs, scalar reused, in place", 8, in, out, clk );
	run<4, 8, 4>( "8x4 pairs, pair reused, in place", 8, in, out, clk );
	run<4, 4, 4>( "4x4 pairs, pair reused, in place", 8, in, out, clk );
	run<4, 8, 2>( "8x2 pairs, pair reused, in place", 8, in, out, clk );
	run<4, 16, 4>( "16x4 pairs, pair reused, in place", 8, in, out, clk );
	run<1, 16, 4>( "16x4 pairs, pair reused", 8, in, out, clk );
	run<0, 8, 8>( "8x8 pairs, scalar reused (i outer)", 8, in, out, clk );
	run<1, 8, 8>( "8x8 pairs, pair reused (j outer)", 8, in, out, clk );
	return 0;
}
