// lds_probe.cu -- how many shared-memory wavefronts does one warp-wide LDS cost on sm_100a for the address patterns an
// outer-product (FFMA2) sgemm inner loop produces?  One CTA of 8 warps issues back-to-back independent loads; the
// cycles per load instruction per SM are the crossbar cost (128 B/clk): 4 = every quarter-warp on its own, 1 = merged.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/lds_probe.cu -o /tmp/lds_probe && /tmp/lds_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int W>   // bytes per lane: 4, 8, 16
__global__ void probe( const int* lane_off, long long* out, int iters )
{
	extern __shared__ __align__(16) unsigned char sm[];
	for ( int i = threadIdx.x; i < 16384 / 4; i += blockDim.x ) reinterpret_cast<int*>( sm )[i] = i;
	__syncthreads();
	const unsigned base = (unsigned)__cvta_generic_to_shared( sm ) + lane_off[threadIdx.x & 31];
	unsigned acc = 0;
	__syncthreads();
	const long long t0 = clock64();
	for ( int it = 0; it < iters; ++it )
	{
		#pragma unroll
		for ( int u = 0; u < 16; ++u )
		{
			const unsigned a = base + u * 512;          // 16 independent loads, same lane pattern
			if constexpr ( W == 16 ) { unsigned x, y, z, w; asm volatile( "ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a) ); acc ^= x ^ y ^ z ^ w; }
			if constexpr ( W == 8 )  { unsigned x, y; asm volatile( "ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a) ); acc ^= x ^ y; }
			if constexpr ( W == 4 )  { unsigned x; asm volatile( "ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a) ); acc ^= x; }
		}
	}
	__syncthreads();
	const long long t1 = clock64();
	if ( threadIdx.x == 0 ) { out[0] = t1 - t0; }
	if ( acc == 0x12345678u ) out[1] = acc;
}

struct Pat { const char* name; int width; int ( *f )( int lane ); };

int main()
{
	Pat pats[] = {
		{ "LDS.128 lane*16 (512 B distinct)", 16, []( int l ) { return l * 16; } },
		{ "LDS.128 (lane&7)*16 (Y: every quarter-warp the same 128 B)", 16, []( int l ) { return ( l & 7 ) * 16; } },
		{ "LDS.128 (lane>>3)*16 (X: one chunk per quarter-warp, 64 B)", 16, []( int l ) { return ( l >> 3 ) * 16; } },
		{ "LDS.128 all lanes one chunk", 16, []( int l ) { return 0; } },
		{ "LDS.128 (lane>>1)*16 (256 B, pairs)", 16, []( int l ) { return ( l >> 1 ) * 16; } },
		{ "LDS.128 (lane&15)*16 (256 B, halves the same)", 16, []( int l ) { return ( l & 15 ) * 16; } },
		{ "LDS.128 (lane>>2)*16 (128 B, 4 lanes per chunk)", 16, []( int l ) { return ( l >> 2 ) * 16; } },
		{ "LDS.128 ((lane&3)+4*(lane>>4))*16 (128 B: tx 4 x ty 2 interleaved)", 16, []( int l ) { return ( ( l & 3 ) + 4 * ( l >> 4 ) ) * 16; } },
		{ "LDS.64 lane*8 (256 B distinct)", 8, []( int l ) { return l * 8; } },
		{ "LDS.64 (lane&7)*8 (64 B)", 8, []( int l ) { return ( l & 7 ) * 8; } },
		{ "LDS.64 (lane>>3)*8 (32 B)", 8, []( int l ) { return ( l >> 3 ) * 8; } },
		{ "LDS.64 (lane&15)*8 (128 B, halves the same)", 8, []( int l ) { return ( l & 15 ) * 8; } },
		{ "LDS.32 lane*4 (128 B distinct)", 4, []( int l ) { return l * 4; } },
		{ "LDS.32 (lane>>3)*4 (4 words)", 4, []( int l ) { return ( l >> 3 ) * 4; } },
		{ "LDS.32 (lane&7)*4 (8 words)", 4, []( int l ) { return ( l & 7 ) * 4; } },
	};
	int* d_off; long long* d_out;
	cudaMalloc( &d_off, 32 * sizeof(int) ); cudaMalloc( &d_out, 2 * sizeof(long long) );
	const int iters = 2000, warps = 8;
	for ( const Pat& p : pats )
	{
		int off[32]; for ( int l = 0; l < 32; ++l ) off[l] = p.f( l );
		cudaMemcpy( d_off, off, sizeof( off ), cudaMemcpyHostToDevice );
		for ( int rep = 0; rep < 2; ++rep )
		{
			if ( p.width == 16 ) probe<16><<<1, warps * 32, 16384>>>( d_off, d_out, iters );
			if ( p.width == 8 )  probe<8><<<1, warps * 32, 16384>>>( d_off, d_out, iters );
			if ( p.width == 4 )  probe<4><<<1, warps * 32, 16384>>>( d_off, d_out, iters );
			cudaDeviceSynchronize();
		}
		long long o[2]; cudaMemcpy( o, d_out, sizeof( o ), cudaMemcpyDeviceToHost );
		printf( "%-72s %6.2f clk per warp-LDS per SM\n", p.name, (double)o[0] / ( (double)iters * 16 * warps ) );
	}
	printf( "%s\n", cudaGetErrorString( cudaGetLastError() ) );
	return 0;
}
