// gemm_ffma_tma.cuh -- sgemm on the FP32 FMA pipe with TMA tensor-map staging (aligned operands).
//
// Same contract as the other gemm kernels (D = alpha*X*Y + beta*D, q contiguous; see gemm_dmma.cuh for the
// reference files replaced) and the same skeleton as gemm_dmma_tma.cuh: one producer thread drives the TMA unit,
// 8 consumer warps, full/empty mbarrier ring, dynamic tile scheduler, setmaxnreg.  No TF32: every product is an
// IEEE fp32 FFMA.
//
// The FP32 pipe is ISSUE bound (one FFMA per scheduler per clock), so the design goal is the fewest non-FFMA
// instructions per FFMA (ncu on gemm_ffma*.cuh: 16 % of the issue slots went to the element-wise transposing
// loader, address arithmetic and multi-wavefront LDS.128):
//   * no transposition at all: a k-contiguous operand stays k-contiguous in shared memory and is read with 16-byte
//     loads ALONG k (one LDS.128 = one row, four k steps); a p/q-contiguous operand is read with 16-byte loads
//     along p/q (one LDS.128 = four rows, one k step).  Either way: 16 LDS.128 per 256 FFMA;
//   * the row/column ownership of a lane depends on the operand's orientation so that every LDS.128 of a warp
//     touches distinct 16-byte chunks of distinct banks under the 128-byte TMA swizzle:
//         k-contiguous X: rows  ty + 4*i          p-contiguous X: rows  16*(i/4) + 4*ty + i%4
//         k-contiguous Y: cols  tx + 8*j          q-contiguous Y: cols  32*(j/4) + 4*tx + j%4
//     (lane = 8*ty + tx, 4 x 8 lanes per warp, warp tile 32 x 64, CTA tile 128 x 128, BK = 32);
//   * swizzled offsets are precomputed once per lane ((chunk ^ line%8) << 4 for the 8 possible line%8), so the k
//     loop adds one register to the stage base per group of loads and uses immediate offsets for the rest.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_ws.cuh"
#include "gemm_dmma_tma.cuh"

namespace b200 {

// Packed FP32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): two independent IEEE fp32 FMAs per instruction on an aligned
// register pair.  Same pipe throughput as two FFMA but half the issue slots and 64-bit operand reads, which is what
// an issue-bound outer-product loop needs (ncu on the scalar version: 13.8 % dispatch stalls, FMA pipe 72.9 %).
__device__ __forceinline__ unsigned long long pack2( float lo, float hi )
{
	unsigned long long r;
	asm( "mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi) );
	return r;
}
__device__ __forceinline__ void unpack2( unsigned long long v, float& lo, float& hi )
{
	asm( "mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v) );
}
__device__ __forceinline__ unsigned long long ffma2( unsigned long long a, unsigned long long b, unsigned long long c )
{
	unsigned long long d;
	asm( "fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c) );
	return d;
}

struct FfmaTmaCfg
{
	static constexpr int BP = 128, BQ = 128, BK = 32, STAGES = 6;
	static constexpr int OPER_BYTES  = 128 * 128;                 // 128 rows x 32 floats, or 4 boxes of 32 k x 32 rows
	static constexpr int STAGE_BYTES = 2 * OPER_BYTES;
	static constexpr int NCONS = 256, NPROD = 128, NT_ALL = NCONS + NPROD;
	static constexpr int BAR_BYTES  = 2 * STAGES * 8 + 4 * 8 + 16;
	static constexpr int SMEM_BYTES = STAGE_BYTES * STAGES + BAR_BYTES + 1024;
};

// CST: as in gemm_dmma_tma.cuh, small-k problems stage the D tile through the ring: two extra stages per tile, each
// 64 rows x 128 columns of D as four 128B-swizzled {32 floats, 64 rows} boxes; warp row-group r (rows 32r..32r+31) takes
// its rows from extra stage r / 2.  Only the q-contiguous ownership (!YK) has the vector epilogue that uses it.
template <bool XK, bool YK, bool TRI = false, bool CST = false>
__global__ void __launch_bounds__( 384, 1 )
gemm_ffma_tma_kernel( const GemmArgs<float> g, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                      const __grid_constant__ CUtensorMap tmd )
{
	using Cfg = FfmaTmaCfg;
	constexpr int BP = Cfg::BP, BQ = Cfg::BQ, BK = Cfg::BK, STAGES = Cfg::STAGES;

	extern __shared__ unsigned char smem_unaligned[];
	const uint32_t raw = smem_u32( smem_unaligned );
	const uint32_t sbase = ( raw + 1023u ) & ~1023u;
	unsigned char* const smem = smem_unaligned + ( sbase - raw );
	const uint32_t bar_base = sbase + (uint32_t)Cfg::STAGE_BYTES * STAGES;
	auto full_bar    = [&]( int s ) { return bar_base + (uint32_t)s * 8u; };
	auto empty_bar   = [&]( int s ) { return bar_base + (uint32_t)( STAGES + s ) * 8u; };
	auto sched_full  = [&]( int s ) { return bar_base + (uint32_t)( 2 * STAGES + s ) * 8u; };
	auto sched_empty = [&]( int s ) { return bar_base + (uint32_t)( 2 * STAGES + 2 + s ) * 8u; };
	volatile int* const sched_tile = reinterpret_cast<volatile int*>( smem + (size_t)Cfg::STAGE_BYTES * STAGES + ( 2 * STAGES + 4 ) * 8 );

	const int tid = threadIdx.x;
	if ( tid == 0 )
	{
		#pragma unroll
		for ( int s = 0; s < STAGES; ++s ) { mbar_init( full_bar( s ), 1 ); mbar_init( empty_bar( s ), Cfg::NCONS / 32 ); }
		#pragma unroll
		for ( int s = 0; s < 2; ++s ) { mbar_init( sched_full( s ), 1 ); mbar_init( sched_empty( s ), Cfg::NCONS / 32 ); }
		asm volatile( "fence.mbarrier_init.release.cluster;\n" ::: "memory" );
	}
	__syncthreads();

	const int64_t KT = ( g.K + BK - 1 ) / BK;
	const int num_tiles = g.tiles_p * g.tiles_q;

	if ( tid >= Cfg::NCONS )
	{
		// ============ PRODUCER warpgroup: one thread drives the TMA unit ============
		setmaxnreg_dec<40>();
		if ( tid != Cfg::NCONS ) return;
		asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmx) : "memory" );
		asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmy) : "memory" );
		int stage = 0; uint32_t phase = 0;
		for ( int it = 0; ; ++it )
		{
			const int slot = it & 1;
			mbar_wait( sched_empty( slot ), ( ( it >> 1 ) & 1 ) ^ 1u );
			const int tile = g.tile_counter ? atomicAdd( g.tile_counter, 1 ) : (int)( blockIdx.x + (unsigned)it * gridDim.x );
			sched_tile[slot] = tile;
			mbar_arrive( sched_full( slot ) );
			if ( tile >= num_tiles ) break;
			int tp, tq;
			tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
			const int p0 = tp * BP, q0 = tq * BQ;
			if ( TRI && tri_skip_tile( g, p0, q0, (int)min( (int64_t)BP, g.P - p0 ), (int)min( (int64_t)BQ, g.Q - q0 ) ) ) continue;
			int64_t kt0 = 0, kt1 = KT;
			if constexpr ( TRI ) tile_k_range( g, p0, (int)min( (int64_t)BP, g.P - p0 ), q0, (int)min( (int64_t)BQ, g.Q - q0 ), BK, KT, kt0, kt1 );
			if constexpr ( !CST ) prefetch_d_tile_l2( g, p0, q0, BP, BQ );
			for ( int64_t kt = kt0; kt < kt1; ++kt )
			{
				mbar_wait( empty_bar( stage ), phase ^ 1u );
				const uint32_t xs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES, ys = xs + Cfg::OPER_BYTES;
				const uint32_t fb = full_bar( stage );
				const int k0 = (int)( kt * BK );
				mbar_arrive_expect_tx( fb, 2u * Cfg::OPER_BYTES );
				if constexpr ( XK ) tma_load_2d( xs, &tmx, k0, p0, fb );               // box {32 k, 128 rows}
				else
				{
					#pragma unroll
					for ( int b = 0; b < BP / 32; ++b ) tma_load_2d( xs + b * 4096, &tmx, p0 + b * 32, k0, fb );   // boxes {32 rows, 32 k}
				}
				if constexpr ( YK ) tma_load_2d( ys, &tmy, k0, q0, fb );
				else
				{
					#pragma unroll
					for ( int b = 0; b < BQ / 32; ++b ) tma_load_2d( ys + b * 4096, &tmy, q0 + b * 32, k0, fb );
				}
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
			if constexpr ( CST )
			{
				#pragma unroll 1
				for ( int qd = 0; qd < BP / 64; ++qd )
				{
					mbar_wait( empty_bar( stage ), phase ^ 1u );
					const uint32_t cs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES;
					const uint32_t fb = full_bar( stage );
					mbar_arrive_expect_tx( fb, (uint32_t)Cfg::STAGE_BYTES );
					#pragma unroll
					for ( int b = 0; b < BQ / 32; ++b ) tma_load_2d( cs + b * 8192, &tmd, q0 + b * 32, p0 + qd * 64, fb );
					if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
				}
			}
		}
		if ( g.tile_counter )
		{
			if ( atomicAdd( g.tile_counter + 1, 1 ) == (int)gridDim.x - 1 ) { g.tile_counter[0] = 0; g.tile_counter[1] = 0; __threadfence(); }
		}
		return;
	}

	// =============================== CONSUMER warps ===============================
	setmaxnreg_inc<224>();
	const int lane = tid & 31, warp = tid >> 5;
	const int ty = lane >> 3, tx = lane & 7;
	const int wr0 = ( warp >> 1 ) * 32;          // 4 x 2 warps: warp tile 32 rows x 64 cols
	const int wc0 = ( warp & 1 ) * 64;

	// ownership maps
	auto row_of = [&]( int i ) { return XK ? wr0 + ty + 4 * i : wr0 + 16 * ( i >> 2 ) + 4 * ty + ( i & 3 ); };
	auto col_of = [&]( int j ) { return YK ? wc0 + tx + 8 * j : wc0 + 32 * ( j >> 2 ) + 4 * tx + ( j & 3 ); };

	// precomputed swizzle terms: ((chunk ^ c) << 4) for c = 0..7
	//   k-contiguous X: chunk = k group (compile time), line = row -> row%8 = ty + 4*(i&1): term[c] uses c = k group
	//   p-contiguous X: chunk = (row%32)>>2 = 4*(i>>2) + ty,  line = k  -> c = k%8
	int xsw[2][8], ysw[2][8];
	#pragma unroll
	for ( int c = 0; c < 8; ++c )
	{
		if constexpr ( XK ) { xsw[0][c] = ( c ^ ty ) << 4;        xsw[1][c] = ( c ^ ( ty + 4 ) ) << 4; }
		else                { xsw[0][c] = ( ty ^ c ) << 4;        xsw[1][c] = ( ( ty + 4 ) ^ c ) << 4; }
		if constexpr ( YK ) { ysw[0][c] = ( c ^ tx ) << 4;        ysw[1][c] = 0; }
		else                { ysw[0][c] = ( tx ^ c ) << 4;        ysw[1][c] = 0; }
	}
	// fixed byte offsets inside an operand buffer
	const int xfix = XK ? ( wr0 + ty ) * 128 : ( wr0 >> 5 ) * 4096;      // + i*512 (XK) | + k*128 (PK)
	const int yfix = YK ? ( wc0 + tx ) * 128 : ( wc0 >> 5 ) * 4096;      // + j*1024 (YK) | + (j>>2)*4096 + k*128 (QK)

	int stage = 0; uint32_t phase = 0;

	for ( int it = 0; ; ++it )
	{
		const int slot = it & 1;
		mbar_wait( sched_full( slot ), ( it >> 1 ) & 1 );
		const int tile = sched_tile[slot];
		__syncwarp();
		if ( lane == 0 ) mbar_arrive( sched_empty( slot ) );
		if ( tile >= num_tiles ) break;
		int tp, tq;
		tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
		const int64_t p0 = (int64_t)tp * BP, q0 = (int64_t)tq * BQ;
		const int p_lim = (int)min( (int64_t)BP, g.P - p0 );
		const int q_lim = (int)min( (int64_t)BQ, g.Q - q0 );
		if ( TRI && tri_skip_tile( g, p0, q0, p_lim, q_lim ) ) continue;

		unsigned long long acc2[8][4];             // acc2[i][j2] = ( acc[i][2*j2], acc[i][2*j2+1] )
		#pragma unroll
		for ( int i = 0; i < 8; ++i )
			#pragma unroll
			for ( int j = 0; j < 4; ++j ) acc2[i][j] = 0ull;

		// operand vectors of one group of 4 k steps: [element][k step]; two buffers (a, b) are alternated so the
		// LDS of group n+1 are in flight under the FFMA2 of group n
		auto load_group = [&]( float ( &xv )[8][4], float ( &yv )[8][4], int st, int kg )
		{
			const unsigned char* xs = smem + (size_t)st * Cfg::STAGE_BYTES + xfix;
			const unsigned char* ys = smem + (size_t)st * Cfg::STAGE_BYTES + Cfg::OPER_BYTES + yfix;
			if constexpr ( XK )
			{
				#pragma unroll
				for ( int i = 0; i < 8; ++i )          // one LDS.128 per owned row: rows ty+4i, k chunk kg
				{
					const float4 v = *reinterpret_cast<const float4*>( xs + i * 512 + xsw[i & 1][kg] );
					xv[i][0] = v.x; xv[i][1] = v.y; xv[i][2] = v.z; xv[i][3] = v.w;
				}
			}
			else
			{
				#pragma unroll
				for ( int kk = 0; kk < 4; ++kk )
				{
					const int k = kg * 4 + kk;
					#pragma unroll
					for ( int h = 0; h < 2; ++h )
					{
						const float4 v = *reinterpret_cast<const float4*>( xs + k * 128 + xsw[h][k & 7] );
						xv[h * 4 + 0][kk] = v.x; xv[h * 4 + 1][kk] = v.y; xv[h * 4 + 2][kk] = v.z; xv[h * 4 + 3][kk] = v.w;
					}
				}
			}
			if constexpr ( YK )
			{
				#pragma unroll
				for ( int j = 0; j < 8; ++j )
				{
					const float4 v = *reinterpret_cast<const float4*>( ys + j * 1024 + ysw[0][kg] );
					yv[j][0] = v.x; yv[j][1] = v.y; yv[j][2] = v.z; yv[j][3] = v.w;
				}
			}
			else
			{
				#pragma unroll
				for ( int kk = 0; kk < 4; ++kk )
				{
					const int k = kg * 4 + kk;
					#pragma unroll
					for ( int h = 0; h < 2; ++h )
					{
						const float4 v = *reinterpret_cast<const float4*>( ys + h * 4096 + k * 128 + ysw[0][k & 7] );
						yv[h * 4 + 0][kk] = v.x; yv[h * 4 + 1][kk] = v.y; yv[h * 4 + 2][kk] = v.z; yv[h * 4 + 3][kk] = v.w;
					}
				}
			}
		};
		// Issue order: j outer, i inner, so that the PAIR operand (y2, 64 bits) is the one consecutive FFMA2 share and sits
		// in the operand reuse latch.  tools/ffma2_probe.cu (no memory traffic, this register tile): 2.24 clocks per FFMA2
		// per scheduler with the pair reused, 2.43 with the scalar reused (i outer), against 2.02 for a few accumulators
		// with both operands fixed -- the FP32 pipe is bound by register-file operand delivery as ptxas allocates and
		// orders it, not by issue slots or occupancy (a 192 x 128 variant with three warps per scheduler measured 58.5
		// against 58.9 TFLOP/s).  ptxas keeps this order for about half of the k loop (SASS: 462 of 1024 FFMA2 with
		// .reuse on the pair); [B200] 16384^3: 58.9 -> 59.5 TFLOP/s.
		auto fma_group = [&]( const float ( &xv )[8][4], const float ( &yv )[8][4] )
		{
			#pragma unroll
			for ( int kk = 0; kk < 4; ++kk )
			{
				#pragma unroll
				for ( int j = 0; j < 4; ++j )
				{
					const unsigned long long y2 = pack2( yv[2 * j][kk], yv[2 * j + 1][kk] );
					#pragma unroll
					for ( int i = 0; i < 8; ++i )
						acc2[i][j] = ffma2( pack2( xv[i][kk], xv[i][kk] ), y2, acc2[i][j] );     // scalar-broadcast operand of FFMA2
				}
			}
		};

		float xa[8][4], ya[8][4], xb[8][4], yb[8][4];
		mbar_wait( full_bar( stage ), phase );
		load_group( xa, ya, stage, 0 );
		int64_t kt0 = 0, kt1 = KT;
		if constexpr ( TRI ) tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
		for ( int64_t kt = kt0; kt < kt1; ++kt )
		{
			#pragma unroll
			for ( int kg = 0; kg < BK / 4; kg += 2 )
			{
				load_group( xb, yb, stage, kg + 1 );
				fma_group( xa, ya );
				if ( kg + 2 < BK / 4 )
				{
					load_group( xa, ya, stage, kg + 2 );
					fma_group( xb, yb );
				}
				else
				{
					int ns = stage + 1; uint32_t nph = phase;
					if ( ns == STAGES ) { ns = 0; nph ^= 1u; }
					if ( kt + 1 < kt1 )
					{
						mbar_wait( full_bar( ns ), nph );
						load_group( xa, ya, ns, 0 );
					}
					fma_group( xb, yb );
					__syncwarp();
					if ( lane == 0 ) mbar_arrive( empty_bar( stage ) );
					stage = ns; phase = nph;
				}
			}
		}

		// ---- epilogue: D = alpha*acc + beta*D (beta == 0: D is not read)
		const bool interior = ( !TRI || tri_tile_interior( g, p0, q0, p_lim, q_lim ) );
		int dlo = 0, dhi = 0;
		if constexpr ( TRI ) tri_band( g, p0, q0, dlo, dhi );
		auto keep = [&]( int d ) { if constexpr ( TRI ) return in_band( d, dlo, dhi ); else return true; };
		float acc[8][8];
		#pragma unroll
		for ( int i = 0; i < 8; ++i )
			#pragma unroll
			for ( int j = 0; j < 4; ++j ) unpack2( acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1] );
		if constexpr ( CST )
		{
			// the two D stages of this tile; every warp walks the ring, row-group wr0 / 64 reads its rows
			const bool fast = ( !YK && g.d_vec_ok && q_lim == BQ && interior );
			#pragma unroll 1
			for ( int qd = 0; qd < BP / 64; ++qd )
			{
				mbar_wait( full_bar( stage ), phase );
				if ( fast && qd == ( wr0 >> 6 ) )
				{
					const unsigned char* cs = smem + (size_t)stage * Cfg::STAGE_BYTES + ( wc0 >> 5 ) * 8192;
					#pragma unroll
					for ( int i = 0; i < 8; ++i )
					{
						const int pl = row_of( i );
						if ( pl >= p_lim ) continue;
						const int r = pl & 63;
						const unsigned char* rowp = cs + r * 128 + ( ( tx ^ ( r & 7 ) ) << 4 );
						const float4 o0 = *reinterpret_cast<const float4*>( rowp ), o1 = *reinterpret_cast<const float4*>( rowp + 8192 );
						float4* dp0 = reinterpret_cast<float4*>( g.D + ( p0 + pl ) * g.ldd + q0 + col_of( 0 ) );
						float4* dp1 = reinterpret_cast<float4*>( g.D + ( p0 + pl ) * g.ldd + q0 + col_of( 4 ) );
						float4 r0, r1;
						r0.x = fmaf( g.beta, o0.x, g.alpha * acc[i][0] ); r0.y = fmaf( g.beta, o0.y, g.alpha * acc[i][1] );
						r0.z = fmaf( g.beta, o0.z, g.alpha * acc[i][2] ); r0.w = fmaf( g.beta, o0.w, g.alpha * acc[i][3] );
						r1.x = fmaf( g.beta, o1.x, g.alpha * acc[i][4] ); r1.y = fmaf( g.beta, o1.y, g.alpha * acc[i][5] );
						r1.z = fmaf( g.beta, o1.z, g.alpha * acc[i][6] ); r1.w = fmaf( g.beta, o1.w, g.alpha * acc[i][7] );
						__stcs( dp0, r0 ); __stcs( dp1, r1 );
					}
				}
				__syncwarp();
				if ( lane == 0 ) mbar_arrive( empty_bar( stage ) );
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
			if ( fast ) continue;
		}
		if ( !YK && g.d_vec_ok && q_lim == BQ && interior )
		{
			// Interior tile, q-contiguous ownership: two 16-byte accesses per row; the loads of row i+1 are in flight
			// while row i is scaled and stored (the producer has already asked L2 for the tile).
			float4 o[2][2];
			auto ptr = [&]( int i, int h ) { return reinterpret_cast<float4*>( g.D + ( p0 + row_of( i ) ) * g.ldd + q0 + col_of( 4 * h ) ); };
			auto load_row = [&]( int i )
			{
				if ( g.beta_is_zero || row_of( i ) >= p_lim ) return;
				o[i & 1][0] = __ldcs( ptr( i, 0 ) ); o[i & 1][1] = __ldcs( ptr( i, 1 ) );
			};
			load_row( 0 );
			#pragma unroll
			for ( int i = 0; i < 8; ++i )
			{
				if ( i + 1 < 8 ) load_row( i + 1 );
				if ( row_of( i ) >= p_lim ) continue;
				float4 r0 = make_float4( g.alpha * acc[i][0], g.alpha * acc[i][1], g.alpha * acc[i][2], g.alpha * acc[i][3] );
				float4 r1 = make_float4( g.alpha * acc[i][4], g.alpha * acc[i][5], g.alpha * acc[i][6], g.alpha * acc[i][7] );
				if ( !g.beta_is_zero )
				{
					const float4 o0 = o[i & 1][0], o1 = o[i & 1][1];
					r0.x = fmaf( g.beta, o0.x, r0.x ); r0.y = fmaf( g.beta, o0.y, r0.y ); r0.z = fmaf( g.beta, o0.z, r0.z ); r0.w = fmaf( g.beta, o0.w, r0.w );
					r1.x = fmaf( g.beta, o1.x, r1.x ); r1.y = fmaf( g.beta, o1.y, r1.y ); r1.z = fmaf( g.beta, o1.z, r1.z ); r1.w = fmaf( g.beta, o1.w, r1.w );
				}
				__stcs( ptr( i, 0 ), r0 ); __stcs( ptr( i, 1 ), r1 );
			}
			continue;
		}
		#pragma unroll
		for ( int i = 0; i < 8; ++i )
		{
			const int pl = row_of( i );
			if ( pl >= p_lim ) continue;
			float* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			float o[8];
			#pragma unroll
			for ( int j = 0; j < 8; ++j )
			{
				const int ql = col_of( j );
				o[j] = ( !g.beta_is_zero && ql < q_lim && keep( ql - pl ) ) ? drow[ql] : 0.f;
			}
			#pragma unroll
			for ( int j = 0; j < 8; ++j )
			{
				const int ql = col_of( j );
				if ( ql >= q_lim || !keep( ql - pl ) ) continue;
				float r = g.alpha * acc[i][j];
				if ( !g.beta_is_zero ) r = fmaf( g.beta, o[j], r );
				drow[ql] = r;
			}
		}
	}
}

} // namespace b200
