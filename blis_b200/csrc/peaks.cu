// peaks.cu -- pipe-peak microbenchmarks: the FP64/FP32 roofline denominators.
//
// MEASURED_PEAKS.json (driver-written) holds only HBM GB/s and bf16 TFLOP/s;
// SURVEY.md section 8d asks round 1 to measure DFMA / DMMA / FFMA so that
// "% of FP64 peak" has a denominator.  Each kernel keeps many independent
// accumulator chains per warp in registers and touches no memory in the loop.
#include "common.cuh"
#include "../../include/blis_b200.h"

namespace b200 {

constexpr int kPeakIters = 4096;

__global__ void __launch_bounds__(256) peak_dfma( double* out, double a, double b )
{
	double acc[16];
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) acc[i] = (double)( threadIdx.x + i );
	for ( int it = 0; it < kPeakIters; ++it )
	{
		#pragma unroll
		for ( int i = 0; i < 16; ++i ) acc[i] = fma( acc[i], a, b );
	}
	double s = 0;
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) s += acc[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) peak_ffma( float* out, float a, float b )
{
	float acc[16];
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) acc[i] = (float)( threadIdx.x + i );
	for ( int it = 0; it < kPeakIters; ++it )
	{
		#pragma unroll
		for ( int i = 0; i < 16; ++i ) acc[i] = fmaf( acc[i], a, b );
	}
	float s = 0;
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) s += acc[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) peak_ffma2( float* out, float a, float b )
{
	unsigned long long acc[16];
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) acc[i] = (unsigned long long)( threadIdx.x + i );
	unsigned long long a2, b2;
	asm( "mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a) );
	asm( "mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b) );
	for ( int it = 0; it < kPeakIters; ++it )
	{
		#pragma unroll
		for ( int i = 0; i < 16; ++i ) asm( "fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2) );
	}
	unsigned long long s = 0;
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) s ^= acc[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = (float)( s & 0xffff );
}

__global__ void __launch_bounds__(256) peak_dmma( double* out, double a, double b )
{
	double acc[16][2];
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
	double af = a + threadIdx.x, bf = b - threadIdx.x;
	for ( int it = 0; it < kPeakIters; ++it )
	{
		#pragma unroll
		for ( int i = 0; i < 16; ++i ) dmma884( acc[i][0], acc[i][1], af, bf );
	}
	double s = 0;
	#pragma unroll
	for ( int i = 0; i < 16; ++i ) s += acc[i][0] + acc[i][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace b200

extern "C" double b200_measure_peak( int kind, int millis )
{
	using namespace b200;
	if ( kind < 0 || kind > 3 ) return -1.0;
	const int threads = 256;
	const int blocks  = kNumSMs * 4;            // 32 warps per SM
	void* out = nullptr;
	if ( cudaMalloc( &out, (size_t)blocks * threads * sizeof(double) ) != cudaSuccess ) return -1.0;
	cudaEvent_t e0, e1;
	cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
	auto launch = [&]()
	{
		if      ( kind == 0 ) peak_dfma<<<blocks, threads>>>( (double*)out, 1.0000001, 1e-9 );
		else if ( kind == 1 ) peak_dmma<<<blocks, threads>>>( (double*)out, 1.0000001, 1e-9 );
		else if ( kind == 2 ) peak_ffma<<<blocks, threads>>>( (float*)out, 1.0000001f, 1e-9f );
		else                  peak_ffma2<<<blocks, threads>>>( (float*)out, 1.0000001f, 1e-9f );
	};
	// flop per launch
	double flop;
	if ( kind == 1 ) flop = (double)blocks * ( threads / 32 ) * kPeakIters * 16.0 * ( 2.0 * 8 * 8 * 4 );
	else if ( kind == 3 ) flop = (double)blocks * threads * kPeakIters * 16.0 * 4.0;
	else             flop = (double)blocks * threads * kPeakIters * 16.0 * 2.0;
	launch(); launch();
	cudaDeviceSynchronize();
	// calibrate, then run for about `millis`
	cudaEventRecord( e0 ); launch(); cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
	float ms = 0; cudaEventElapsedTime( &ms, e0, e1 );
	int reps = (int)( millis / ( ms > 1e-3f ? ms : 1e-3f ) ); if ( reps < 1 ) reps = 1; if ( reps > 100000 ) reps = 100000;
	cudaEventRecord( e0 );
	for ( int r = 0; r < reps; ++r ) launch();
	cudaEventRecord( e1 ); cudaEventSynchronize( e1 );
	cudaEventElapsedTime( &ms, e0, e1 );
	const bool ok = ( cudaGetLastError() == cudaSuccess );
	cudaEventDestroy( e0 ); cudaEventDestroy( e1 ); cudaFree( out );
	if ( !ok ) return -1.0;
	return flop * reps / ( ms * 1e-3 ) / 1e12;
}
