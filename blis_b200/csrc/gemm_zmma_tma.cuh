// gemm_zmma_tma.cuh -- zgemm with TMA tensor-map staging (16-byte aligned operands).
//
// Same contract, consumer math and tile scheduler as gemm_dmma_ws_kernel<double2> (gemm_dmma_ws.cuh: four DMMA.8x8x4 per
// complex 8x8x4 product, conjugation as a sign flip of the fragment's imaginary part) and the same producer as
// gemm_dmma_tma.cuh: ONE thread issues cp.async.bulk.tensor per operand and k slab into 128-byte-swizzled shared memory.
// Replaces, for datatype z, the jc/ic/pc loops + packm + bli_gemm_ker_var2 + gemm ukr
// (frame/3/gemm/bli_gemm_blk_var{2,3,1}.c, frame/1m/packm/bli_packm_blk_var1.c, ref_kernels/3/bli_gemm_ref.c).
//
// A complex element is 16 bytes, so one 128-byte swizzle row holds 8 complex numbers: BK = 8, tensor maps are typed
// FLOAT64 with two elements per complex number.
//   k-contiguous operand  : ONE box {8 k, rows}: row r = 128 bytes = 8 complex k values
//   p/q-contiguous operand: rows/8 boxes {8 rows, 8 k}: inside a box, line k = 128 bytes = 8 consecutive p (or q)
// The 16-byte fragment load of lane (g = lane/4, t = lane%4) is LDS.128.  A quarter warp (8 lanes: two values of g, four
// of t) is conflict free when its 8 chunks differ; with the natural k = 4s + t the two rows would hit the same four chunks,
// so the k index a lane uses in k4-step s is k(s,t) = 2t + s (A and B fragments use the same k, the product is unchanged):
// k-contiguous: chunk = k ^ g -> rows g, g+1 take the even / odd chunk cosets; p/q-contiguous: chunk = g ^ k, lines k differ.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_ws.cuh"
#include "gemm_dmma_tma.cuh"

namespace b200 {

struct ZmmaTmaCfg
{
	static constexpr int BP = 64, BQ = 128, BK = 8, WP = 2, WQ = 4, STAGES = 8;
	static constexpr int WTP = BP / WP, WTQ = BQ / WQ, MT = WTP / 8, NTL = WTQ / 8;      // warp tile 32 x 32: 4 x 4 DMMA tiles
	static constexpr int X_BYTES = BP * 128, Y_BYTES = BQ * 128;
	static constexpr int STAGE_BYTES = X_BYTES + Y_BYTES;                                  // 24 KiB
	static constexpr int NCONS = WP * WQ * 32, NPROD = 128, NT_ALL = NCONS + NPROD;
	static constexpr int BAR_BYTES  = 2 * STAGES * 8 + 4 * 8 + 16;
	static constexpr int SMEM_BYTES = STAGE_BYTES * STAGES + BAR_BYTES + 1024;
};

template <bool XK, bool YK, bool TRI = false>
__global__ void __launch_bounds__( 384, 1 )
gemm_zmma_tma_kernel( const GemmArgs<double2> g, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy )
{
	using Cfg = ZmmaTmaCfg;
	constexpr int BP = Cfg::BP, BQ = Cfg::BQ, BK = Cfg::BK, WQ = Cfg::WQ, STAGES = Cfg::STAGES;
	constexpr int MT = Cfg::MT, NTL = Cfg::NTL;

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
			prefetch_d_tile_l2( g, p0, q0, BP, BQ );
			for ( int64_t kt = kt0; kt < kt1; ++kt )
			{
				mbar_wait( empty_bar( stage ), phase ^ 1u );
				const uint32_t xs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES, ys = xs + Cfg::X_BYTES;
				const uint32_t fb = full_bar( stage );
				const int k0 = (int)( kt * BK );
				mbar_arrive_expect_tx( fb, (uint32_t)Cfg::STAGE_BYTES );
				// coordinates are in doubles along the contiguous dimension (two per complex element)
				if constexpr ( XK ) tma_load_2d( xs, &tmx, 2 * k0, p0, fb );                     // box {8 k, 64 rows}
				else
				{
					#pragma unroll
					for ( int b = 0; b < BP / 8; ++b ) tma_load_2d( xs + b * 1024, &tmx, 2 * ( p0 + b * 8 ), k0, fb );   // boxes {8 rows, 8 k}
				}
				if constexpr ( YK ) tma_load_2d( ys, &tmy, 2 * k0, q0, fb );                     // box {8 k, 128 rows}
				else
				{
					#pragma unroll
					for ( int b = 0; b < BQ / 8; ++b ) tma_load_2d( ys + b * 1024, &tmy, 2 * ( q0 + b * 8 ), k0, fb );
				}
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
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
	const int gq = lane >> 2, t4 = lane & 3;
	const int wp0 = ( warp / WQ ) * Cfg::WTP;
	const int wq0 = ( warp % WQ ) * Cfg::WTQ;
	const bool cjx = g.conjx != 0, cjy = g.conjy != 0;

	// byte offset of the fragment element of 8x8 tile `i` (rows w0 + 8i + g) in k4-step s: k = 2t + s
	auto frag_off = [&]( bool kmajor, int w0, int i, int s ) -> int
	{
		const int k = 2 * t4 + s;
		if ( kmajor ) return ( w0 + i * 8 + gq ) * 128 + ( ( k ^ gq ) << 4 );
		return ( ( w0 >> 3 ) + i ) * 1024 + k * 128 + ( ( gq ^ k ) << 4 );
	};

	int stage = 0; uint32_t phase = 0;
	auto load_frags = [&]( double2 ( &xf )[MT], double2 ( &yf )[NTL], int st, int s )
	{
		const unsigned char* xs = smem + (size_t)st * Cfg::STAGE_BYTES;
		const unsigned char* ys = xs + Cfg::X_BYTES;
		#pragma unroll
		for ( int i = 0; i < MT; ++i ) xf[i] = *reinterpret_cast<const double2*>( xs + frag_off( XK, wp0, i, s ) );
		#pragma unroll
		for ( int j = 0; j < NTL; ++j ) yf[j] = *reinterpret_cast<const double2*>( ys + frag_off( YK, wq0, j, s ) );
	};

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

		double acc[2][MT][NTL][2];
		#pragma unroll
		for ( int c = 0; c < 2; ++c )
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
				#pragma unroll
				for ( int j = 0; j < NTL; ++j ) { acc[c][i][j][0] = 0.0; acc[c][i][j][1] = 0.0; }

		double2 xa[MT], ya[NTL], xb[MT], yb[NTL];
		mbar_wait( full_bar( stage ), phase );
		load_frags( xa, ya, stage, 0 );

		auto mma_step = [&]( double2 ( &xf )[MT], double2 ( &yf )[NTL] )
		{
			// re += xr*yr - xi*yi, im += xr*yi + xi*yr.  The two DMMAs into one accumulator are issued a whole pass apart
			// (32 independent DMMAs in between) so that the second never waits for the first; per accumulator the order
			// of the additions is unchanged.
			double xi[MT], yi[NTL];
			#pragma unroll
			for ( int i = 0; i < MT; ++i ) xi[i] = flip_sign( xf[i].y, cjx );
			#pragma unroll
			for ( int j = 0; j < NTL; ++j ) yi[j] = flip_sign( yf[j].y, cjy );
			#pragma unroll
			for ( int j = 0; j < NTL; ++j )
				#pragma unroll
				for ( int i = 0; i < MT; ++i )
				{
					dmma884( acc[0][i][j][0], acc[0][i][j][1], xf[i].x, yf[j].x );
					dmma884( acc[1][i][j][0], acc[1][i][j][1], xf[i].x, yi[j] );
				}
			#pragma unroll
			for ( int j = 0; j < NTL; ++j )
				#pragma unroll
				for ( int i = 0; i < MT; ++i )
				{
					dmma884( acc[0][i][j][0], acc[0][i][j][1], -xi[i], yi[j] );
					dmma884( acc[1][i][j][0], acc[1][i][j][1], xi[i],  yf[j].x );
				}
		};

		int64_t kt0 = 0, kt1 = KT;
		if constexpr ( TRI ) tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
		for ( int64_t kt = kt0; kt < kt1; ++kt )
		{
			// two k4-steps per stage, fragments double-buffered (a <-> b)
			load_frags( xb, yb, stage, 1 );
			mma_step( xa, ya );
			int ns = stage + 1; uint32_t nph = phase;
			if ( ns == STAGES ) { ns = 0; nph ^= 1u; }
			if ( kt + 1 < kt1 )
			{
				mbar_wait( full_bar( ns ), nph );
				load_frags( xa, ya, ns, 0 );
			}
			mma_step( xb, yb );
			__syncwarp();
			if ( lane == 0 ) mbar_arrive( empty_bar( stage ) );
			stage = ns; phase = nph;
		}

		// ---- epilogue: D = alpha*acc + beta*D (beta == 0: D is not read); complex scalars as bli_tscals / bli_txpbys
		const bool interior = ( !TRI || tri_tile_interior( g, p0, q0, p_lim, q_lim ) );
		int dlo = 0, dhi = 0;
		if constexpr ( TRI ) tri_band( g, p0, q0, dlo, dhi );
		auto keep = [&]( int d ) { if constexpr ( TRI ) return in_band( d, dlo, dhi ); else return true; };
		if ( g.d_vec_ok && q_lim == BQ && interior )
		{
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
			{
				const int pl = wp0 + i * 8 + gq;
				if ( pl >= p_lim ) continue;
				double2* __restrict__ dp = g.D + ( p0 + pl ) * g.ldd + q0 + wq0 + 2 * t4;
				double2 o[NTL][2];
				if ( !g.beta_is_zero )
				{
					#pragma unroll
					for ( int j = 0; j < NTL; ++j ) { o[j][0] = __ldcs( dp + j * 8 ); o[j][1] = __ldcs( dp + j * 8 + 1 ); }
				}
				#pragma unroll
				for ( int j = 0; j < NTL; ++j )
					#pragma unroll
					for ( int e = 0; e < 2; ++e )
					{
						double rr, ri;
						cscal( g.alpha.x, g.alpha.y, acc[0][i][j][e], acc[1][i][j][e], rr, ri );
						if ( !g.beta_is_zero ) cxpby( g.beta.x, g.beta.y, o[j][e].x, o[j][e].y, rr, ri );
						__stcs( dp + j * 8 + e, make_double2( rr, ri ) );
					}
			}
			continue;
		}
		#pragma unroll
		for ( int i = 0; i < MT; ++i )
		{
			const int pl = wp0 + i * 8 + gq;
			if ( pl >= p_lim ) continue;
			double2* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			#pragma unroll
			for ( int j = 0; j < NTL; ++j )
			{
				const int ql = wq0 + j * 8 + 2 * t4;
				#pragma unroll
				for ( int e = 0; e < 2; ++e )
				{
					if ( ql + e >= q_lim || !keep( ql + e - pl ) ) continue;
					double rr, ri;
					cscal( g.alpha.x, g.alpha.y, acc[0][i][j][e], acc[1][i][j][e], rr, ri );
					if ( !g.beta_is_zero )
					{
						const double2 o = drow[ql + e];
						cxpby( g.beta.x, g.beta.y, o.x, o.y, rr, ri );
					}
					drow[ql + e] = make_double2( rr, ri );
				}
			}
		}
	}
}

} // namespace b200
