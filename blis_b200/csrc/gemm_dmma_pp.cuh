// gemm_dmma_pp.cuh -- dgemm for SMALL k: two consumer groups taking turns on the tensor pipe ("ping-pong").
//
// Regime: the rank-k updates of blocked factorizations and the shapes the reference's sup path serves
// (frame/3/bli_l3_sup.c:37-135; BASELINE configs[2], k = 64).  Per 128 x 128 tile and SM the k loop is
// 2*128*128*k / (37.1 TFLOP/s / 148) = 8.4 us at k = 64 and the read-modify-write of D is 256 KiB / (6.46 TB/s / 148)
// = 6.0 us: both bounds are within 30 % of each other, so the kernel is only as fast as they OVERLAP.  In
// gemm_dmma_tma.cuh all eight consumer warps walk the same ring in lockstep, reach their epilogues together, and the
// DMMA pipe idles while D is read, updated and stored (ncu, k = 64: tensor pipe 78 % busy, 3.6 TB/s).
//
// Here the tile is split into its two q-halves (128 x 64 each), one per consumer GROUP of four warps (one warp per SM
// sub-partition -- enough to saturate the pipe: a warp owns 32 independent 8x8 accumulator tiles), and the groups
// alternate:            group 0:  M(i) E(i)....  M(i+1) E(i+1)...
//                       group 1:  ......  M(i) E(i)....  M(i+1)
// Named barriers enforce the order of the k loops M (never two at once), so every epilogue E runs under the other
// group's k loop and the pipe is handed over without a bubble.  Each group has its OWN ring (4 stages of X 128x16 +
// Y-half 64x16 = 24 KiB) fed by its OWN producer thread, so neither group's progress depends on slots the other one
// holds (a shared ring deadlocks as soon as a k loop is longer than the ring).  X is staged twice (once per group, from
// L2); D is read straight from global memory in the epilogue -- its latency is what the other group's k loop hides --
// after an L2 prefetch issued by the producer when the tile is drawn.
//
// Same contract (GemmArgs), same TMA tensor maps / 128-byte swizzle / permuted-k fragment addressing, same tile
// scheduler and the same arithmetic per accumulator (bit-identical results) as gemm_dmma_tma.cuh.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_ws.cuh"
#include "gemm_dmma_tma.cuh"

namespace b200 {

__device__ __forceinline__ void named_bar_sync( int id, int count )   { asm volatile( "bar.sync %0, %1;\n" :: "r"(id), "r"(count) : "memory" ); }
__device__ __forceinline__ void named_bar_arrive( int id, int count ) { asm volatile( "bar.arrive %0, %1;\n" :: "r"(id), "r"(count) : "memory" ); }

struct DmmaPpCfg
{
	static constexpr int BP = 128, BQ = 128, BQH = 64, BK = 16, STAGES = 4, SCHED = 4;
	static constexpr int MT = 4, NTL = 8;                              // per warp: 32 rows x 64 columns of 8x8 tiles
	static constexpr int X_BYTES = 128 * 128, Y_BYTES = 64 * 128;      // 16 KiB + 8 KiB per stage
	static constexpr int STAGE_BYTES = X_BYTES + Y_BYTES;
	static constexpr int RING_BYTES  = STAGE_BYTES * STAGES;           // 96 KiB per group
	static constexpr int NGROUP = 128, NCONS = 256, NT_ALL = 384;
	static constexpr int NBAR = 2 * 2 * STAGES + 2 * SCHED;
	static constexpr int SMEM_BYTES = 2 * RING_BYTES + NBAR * 8 + SCHED * 4 + 16 + 1024;
};

template <bool XK, bool YK>
__global__ void __launch_bounds__( 384, 1 )
gemm_dmma_pp_kernel( const GemmArgs<double> g, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy )
{
	using Cfg = DmmaPpCfg;
	constexpr int BP = Cfg::BP, BQ = Cfg::BQ, BQH = Cfg::BQH, BK = Cfg::BK, STAGES = Cfg::STAGES, SCHED = Cfg::SCHED;
	constexpr int MT = Cfg::MT, NTL = Cfg::NTL, KS = BK / 4;

	extern __shared__ unsigned char smem_unaligned[];
	const uint32_t raw = smem_u32( smem_unaligned );
	const uint32_t sbase = ( raw + 1023u ) & ~1023u;
	unsigned char* const smem = smem_unaligned + ( sbase - raw );
	const uint32_t bar_base = sbase + 2u * Cfg::RING_BYTES;
	auto full_bar    = [&]( int grp, int s ) { return bar_base + (uint32_t)( grp * 2 * STAGES + s ) * 8u; };
	auto empty_bar   = [&]( int grp, int s ) { return bar_base + (uint32_t)( grp * 2 * STAGES + STAGES + s ) * 8u; };
	auto sched_full  = [&]( int s ) { return bar_base + (uint32_t)( 4 * STAGES + s ) * 8u; };
	auto sched_empty = [&]( int s ) { return bar_base + (uint32_t)( 4 * STAGES + SCHED + s ) * 8u; };
	volatile int* const sched_tile = reinterpret_cast<volatile int*>( smem + 2 * (size_t)Cfg::RING_BYTES + Cfg::NBAR * 8 );

	const int tid = threadIdx.x;
	if ( tid == 0 )
	{
		for ( int grp = 0; grp < 2; ++grp )
			for ( int s = 0; s < STAGES; ++s )
			{
				mbar_init( full_bar( grp, s ), 1 );
				mbar_init( empty_bar( grp, s ), Cfg::NGROUP / 32 );
			}
		for ( int s = 0; s < SCHED; ++s )
		{
			mbar_init( sched_full( s ), 1 );
			mbar_init( sched_empty( s ), Cfg::NCONS / 32 + 1 );        // eight consumer warps + the second producer
		}
		asm volatile( "fence.mbarrier_init.release.cluster;\n" ::: "memory" );
	}
	__syncthreads();

	const int64_t KT = ( g.K + BK - 1 ) / BK;
	const int num_tiles = g.tiles_p * g.tiles_q;

	if ( tid >= Cfg::NCONS )
	{
		// ============ PRODUCERS: thread 256 draws tiles and feeds group 0, thread 288 feeds group 1 ============
		setmaxnreg_dec<40>();
		const int ptid = tid - Cfg::NCONS;
		if ( ptid != 0 && ptid != 32 ) return;
		const int grp = ptid >> 5;
		asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmx) : "memory" );
		asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmy) : "memory" );
		const uint32_t ring = sbase + (uint32_t)grp * Cfg::RING_BYTES;
		int stage = 0; uint32_t phase = 0;
		for ( int it = 0; ; ++it )
		{
			const int slot = it % SCHED;
			const uint32_t par = (uint32_t)( ( it / SCHED ) & 1 );
			int tile;
			if ( grp == 0 )
			{
				mbar_wait( sched_empty( slot ), par ^ 1u );
				tile = g.tile_counter ? atomicAdd( g.tile_counter, 1 ) : (int)( blockIdx.x + (unsigned)it * gridDim.x );
				sched_tile[slot] = tile;
				mbar_arrive( sched_full( slot ) );
			}
			else
			{
				mbar_wait( sched_full( slot ), par );
				tile = sched_tile[slot];
				mbar_arrive( sched_empty( slot ) );
			}
			if ( tile >= num_tiles ) break;
			int tp, tq;
			tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
			const int p0 = tp * BP, q0 = tq * BQ + grp * BQH;
			if ( grp == 0 ) prefetch_d_tile_l2( g, p0, tq * BQ, BP, BQ );
			for ( int64_t kt = 0; kt < KT; ++kt )
			{
				mbar_wait( empty_bar( grp, stage ), phase ^ 1u );
				const uint32_t xs = ring + (uint32_t)stage * Cfg::STAGE_BYTES, ys = xs + Cfg::X_BYTES;
				const uint32_t fb = full_bar( grp, stage );
				const int k0 = (int)( kt * BK );
				mbar_arrive_expect_tx( fb, (uint32_t)Cfg::STAGE_BYTES );
				if constexpr ( XK ) tma_load_2d( xs, &tmx, k0, p0, fb );
				else
				{
					#pragma unroll
					for ( int b = 0; b < BP / 16; ++b ) tma_load_2d( xs + b * 2048, &tmx, p0 + b * 16, k0, fb );
				}
				if constexpr ( YK ) tma_load_2d( ys, &tmy, k0, q0, fb );           // box {16 k, 64 rows}
				else
				{
					#pragma unroll
					for ( int b = 0; b < BQH / 16; ++b ) tma_load_2d( ys + b * 2048, &tmy, q0 + b * 16, k0, fb );
				}
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
		}
		if ( grp == 0 && g.tile_counter )
		{
			if ( atomicAdd( g.tile_counter + 1, 1 ) == (int)gridDim.x - 1 ) { g.tile_counter[0] = 0; g.tile_counter[1] = 0; __threadfence(); }
		}
		return;
	}

	// =============================== CONSUMER groups ===============================
	setmaxnreg_inc<224>();
	const int lane = tid & 31, warp = tid >> 5;
	const int grp = warp >> 2;                                     // warps 0-3: q-half 0, warps 4-7: q-half 1
	const int gq = lane >> 2, t4 = lane & 3;
	const int wp0 = ( warp & 3 ) * 32;
	const unsigned char* const ring = smem + (size_t)grp * Cfg::RING_BYTES;

	auto frag_off = [&]( bool kmajor, int w0, int i, int s ) -> int
	{
		const int ts = ( t4 + s ) & 3;
		if ( kmajor )
		{
			const int chunk = ( ( t4 >> 1 ) << 2 ) | ts;
			return ( w0 + i * 8 + gq ) * 128 + ( ( chunk ^ gq ) << 4 ) + ( t4 & 1 ) * 8;
		}
		const int k  = ( t4 & 1 ) | ( ( t4 >> 1 ) << 3 ) | ( ts << 1 );
		const int k7 = ( t4 & 1 ) | ( ts << 1 );
		const int chunk = ( ( i & 1 ) << 2 ) | ( gq >> 1 );
		return ( ( w0 >> 4 ) + ( i >> 1 ) ) * 2048 + k * 128 + ( ( chunk ^ k7 ) << 4 ) + ( gq & 1 ) * 8;
	};
	auto load_frags = [&]( double ( &xf )[MT], double ( &yf )[NTL], int st, int s )
	{
		const unsigned char* xs = ring + (size_t)st * Cfg::STAGE_BYTES;
		const unsigned char* ys = xs + Cfg::X_BYTES;
		#pragma unroll
		for ( int i = 0; i < MT; ++i ) xf[i] = *reinterpret_cast<const double*>( xs + frag_off( XK, wp0, i, s ) );
		#pragma unroll
		for ( int j = 0; j < NTL; ++j ) yf[j] = *reinterpret_cast<const double*>( ys + frag_off( YK, 0, j, s ) );
	};

	// barrier 1: "group 0 may run its k loop", barrier 2: "group 1 may"; each completes with 128 arrivals + 128 waiters
	if ( grp == 1 ) named_bar_arrive( 1, Cfg::NCONS );

	int stage = 0; uint32_t phase = 0;
	for ( int it = 0; ; ++it )
	{
		const int slot = it % SCHED;
		mbar_wait( sched_full( slot ), (uint32_t)( ( it / SCHED ) & 1 ) );
		const int tile = sched_tile[slot];
		__syncwarp();
		if ( lane == 0 ) mbar_arrive( sched_empty( slot ) );
		if ( tile >= num_tiles ) break;
		int tp, tq;
		tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
		const int64_t p0 = (int64_t)tp * BP, q0 = (int64_t)tq * BQ + grp * BQH;
		const int p_lim = (int)min( (int64_t)BP, g.P - p0 );
		const int q_lim = (int)max( (int64_t)0, min( (int64_t)BQH, g.Q - q0 ) );      // 0: this half lies outside D (the k loop still runs, on zero fill)

		double acc[MT][NTL][2];
		#pragma unroll
		for ( int i = 0; i < MT; ++i )
			#pragma unroll
			for ( int j = 0; j < NTL; ++j ) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
		auto mma_step = [&]( double ( &xf )[MT], double ( &yf )[NTL] )
		{
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
				#pragma unroll
				for ( int j = 0; j < NTL; ++j )
					dmma884( acc[i][j][0], acc[i][j][1], xf[i], yf[j] );
		};

		double xa[MT], ya[NTL], xb[MT], yb[NTL];
		mbar_wait( full_bar( grp, stage ), phase );
		load_frags( xa, ya, stage, 0 );
		named_bar_sync( 1 + grp, Cfg::NCONS );                                   // my turn on the tensor pipe
		for ( int64_t kt = 0; kt < KT; ++kt )
		{
			#pragma unroll
			for ( int kk = 0; kk < KS; kk += 2 )
			{
				load_frags( xb, yb, stage, kk + 1 );
				mma_step( xa, ya );
				if ( kk + 2 < KS )
				{
					load_frags( xa, ya, stage, kk + 2 );
					mma_step( xb, yb );
				}
				else
				{
					int ns = stage + 1; uint32_t nph = phase;
					if ( ns == STAGES ) { ns = 0; nph ^= 1u; }
					if ( kt + 1 < KT )
					{
						mbar_wait( full_bar( grp, ns ), nph );
						load_frags( xa, ya, ns, 0 );
					}
					mma_step( xb, yb );
					__syncwarp();
					if ( lane == 0 ) mbar_arrive( empty_bar( grp, stage ) );
					stage = ns; phase = nph;
				}
			}
		}
		named_bar_arrive( 2 - grp, Cfg::NCONS );                                 // the other group's turn; my epilogue runs under its k loop

		// ---- epilogue: D = alpha*acc + beta*D   (beta == 0: D is not read)
		if ( g.d_vec_ok && q_lim == BQH )
		{
			// two rows of 8x8 tiles in flight: the fragment registers are dead here
			#pragma unroll
			for ( int i = 0; i < MT; i += 2 )
			{
				double2* __restrict__ dp0 = reinterpret_cast<double2*>( g.D + ( p0 + wp0 + i * 8 + gq ) * g.ldd + q0 + 2 * t4 );
				double2* __restrict__ dp1 = dp0 + 4 * g.ldd;                     // 8 rows further = 8*ldd doubles = 4*ldd double2
				const bool r0 = ( wp0 + i * 8 + gq ) < p_lim, r1 = ( wp0 + i * 8 + 8 + gq ) < p_lim;
				double2 o0[NTL], o1[NTL];
				if ( !g.beta_is_zero )
				{
					if ( r0 ) {
						#pragma unroll
						for ( int j = 0; j < NTL; ++j ) o0[j] = __ldcs( dp0 + j * 4 );
					}
					if ( r1 ) {
						#pragma unroll
						for ( int j = 0; j < NTL; ++j ) o1[j] = __ldcs( dp1 + j * 4 );
					}
				}
				if ( r0 )
				{
					#pragma unroll
					for ( int j = 0; j < NTL; ++j )
					{
						double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
						if ( !g.beta_is_zero ) { v0 = fma( g.beta, o0[j].x, v0 ); v1 = fma( g.beta, o0[j].y, v1 ); }
						__stcs( dp0 + j * 4, make_double2( v0, v1 ) );
					}
				}
				if ( r1 )
				{
					#pragma unroll
					for ( int j = 0; j < NTL; ++j )
					{
						double v0 = g.alpha * acc[i + 1][j][0], v1 = g.alpha * acc[i + 1][j][1];
						if ( !g.beta_is_zero ) { v0 = fma( g.beta, o1[j].x, v0 ); v1 = fma( g.beta, o1[j].y, v1 ); }
						__stcs( dp1 + j * 4, make_double2( v0, v1 ) );
					}
				}
			}
			continue;
		}
		#pragma unroll
		for ( int i = 0; i < MT; ++i )
		{
			const int pl = wp0 + i * 8 + gq;
			if ( pl >= p_lim ) continue;
			double* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			#pragma unroll
			for ( int j = 0; j < NTL; ++j )
			{
				const int ql = j * 8 + 2 * t4;
				double r0 = g.alpha * acc[i][j][0], r1 = g.alpha * acc[i][j][1];
				if ( ql < q_lim )
				{
					if ( !g.beta_is_zero ) r0 = fma( g.beta, drow[ql], r0 );
					drow[ql] = r0;
				}
				if ( ql + 1 < q_lim )
				{
					if ( !g.beta_is_zero ) r1 = fma( g.beta, drow[ql + 1], r1 );
					drow[ql + 1] = r1;
				}
			}
		}
	}
}

} // namespace b200
