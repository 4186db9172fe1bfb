// gemm_cfma_tma.cuh -- cgemm on the FP32 FMA pipe with TMA staging and packed FFMA2 (aligned operands).
//
// Same contract/skeleton as gemm_ffma_tma.cuh (see there and gemm_dmma.cuh for the reference mapping).  A complex
// multiply-accumulate is two packed FMAs with the complex y operand used exactly as it lies in memory:
//      P = (P.x, P.y) += xr * ( yr, yi )     FFMA2, scalar xr broadcast to both lanes
//      Q = (Q.x, Q.y) += xi * ( yr, yi )     FFMA2, scalar xi broadcast
// and the four real sums are combined once per output element in the epilogue:
//      re = P.x - sx*sy*Q.y ,  im = sy*P.y + sx*Q.x        (sx/sy = -1 for conj(x)/conj(y), else +1)
// i.e. the reference's bli_tdots for c (re += xr*yr - xi*yi ; im += xr*yi + xi*yr; frame/include/level0) with the
// four partial sums kept apart, so that the k loop contains no negation, no register shuffle and no select:
// FFMA2 and LDS.128 only.  No TF32, every product is an IEEE fp32 FMA.
//
// Lane tile 4 (p) x 8 (q) complex, lanes 4 x 8 per warp, warp tile 16 x 64, 4 x 2 warps, CTA tile 64 x 128, BK = 16
// (one 128-byte line = 16 complex).  Ownership by orientation (conflict-free LDS.128 under the 128B swizzle):
//      k-contiguous X: rows ty + 4*i            p-contiguous X: rows 8*(i/2) + 2*ty + i%2
//      k-contiguous Y: cols tx + 8*j            q-contiguous Y: cols 16*(j/2) + 2*tx + j%2
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_ws.cuh"
#include "gemm_dmma_tma.cuh"
#include "gemm_ffma_tma.cuh"

namespace b200 {

struct CfmaTmaCfg
{
	static constexpr int BP = 64, BQ = 128, BK = 16, STAGES = 8;
	static constexpr int X_BYTES = BP * 128, Y_BYTES = BQ * 128;
	static constexpr int STAGE_BYTES = X_BYTES + Y_BYTES;             // 24 KiB
	static constexpr int NCONS = 256, NPROD = 128, NT_ALL = NCONS + NPROD;
	static constexpr int BAR_BYTES  = 2 * STAGES * 8 + 4 * 8 + 16;
	static constexpr int SMEM_BYTES = STAGE_BYTES * STAGES + BAR_BYTES + 1024;
};

// CST: small-k problems stage the D tile through the ring as in gemm_dmma_tma.cuh: four extra stages per tile, each
// 16 rows x 128 columns of D as eight 128B-swizzled {16 complex, 16 rows} boxes (16 KiB of the 24 KiB stage); warp
// row-group r (rows 16r..16r+15) takes its rows from extra stage r.  Used with the q-contiguous ownership (!YK).
template <bool XK, bool YK, bool TRI = false, bool CST = false>
__global__ void __launch_bounds__( 384, 1 )
gemm_cfma_tma_kernel( const GemmArgs<float2> g, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                      const __grid_constant__ CUtensorMap tmd )
{
	using Cfg = CfmaTmaCfg;
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
				const uint32_t xs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES, ys = xs + Cfg::X_BYTES;
				const uint32_t fb = full_bar( stage );
				const int k0 = (int)( kt * BK );
				mbar_arrive_expect_tx( fb, (uint32_t)Cfg::STAGE_BYTES );
				if constexpr ( XK ) tma_load_2d( xs, &tmx, k0, p0, fb );                  // box {16 k, 64 rows}
				else
				{
					#pragma unroll
					for ( int b = 0; b < BP / 16; ++b ) tma_load_2d( xs + b * 2048, &tmx, p0 + b * 16, k0, fb );   // boxes {16 rows, 16 k}
				}
				if constexpr ( YK ) tma_load_2d( ys, &tmy, k0, q0, fb );                  // box {16 k, 128 rows}
				else
				{
					#pragma unroll
					for ( int b = 0; b < BQ / 16; ++b ) tma_load_2d( ys + b * 2048, &tmy, q0 + b * 16, k0, fb );
				}
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
			if constexpr ( CST )
			{
				#pragma unroll 1
				for ( int qd = 0; qd < BP / 16; ++qd )
				{
					mbar_wait( empty_bar( stage ), phase ^ 1u );
					const uint32_t cs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES;
					const uint32_t fb = full_bar( stage );
					mbar_arrive_expect_tx( fb, 16u * 1024u );
					#pragma unroll
					for ( int b = 0; b < BQ / 16; ++b ) tma_load_2d( cs + b * 2048, &tmd, q0 + b * 16, p0 + qd * 16, fb );
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
	const int wr0 = ( warp >> 1 ) * 16;          // 4 x 2 warps: warp tile 16 rows x 64 cols
	const int wc0 = ( warp & 1 ) * 64;
	const bool cjx = g.conjx != 0, cjy = g.conjy != 0;

	auto row_of = [&]( int i ) { return XK ? wr0 + ty + 4 * i : wr0 + 8 * ( i >> 1 ) + 2 * ty + ( i & 1 ); };
	auto col_of = [&]( int j ) { return YK ? wc0 + tx + 8 * j : wc0 + 16 * ( j >> 1 ) + 2 * tx + ( j & 1 ); };

	// swizzle terms ((chunk ^ c) << 4), c = 0..7
	int xsw[2][8], ysw[8];
	#pragma unroll
	for ( int c = 0; c < 8; ++c )
	{
		if constexpr ( XK ) { xsw[0][c] = ( c ^ ty ) << 4;  xsw[1][c] = ( c ^ ( ty + 4 ) ) << 4; }    // line = row, row%8 = ty + 4*(i&1)
		else                { xsw[0][c] = ( ty ^ c ) << 4;  xsw[1][c] = ( ( ty + 4 ) ^ c ) << 4; }    // chunk = 4*(i>>1) + ty, line = k
		ysw[c] = ( tx ^ c ) << 4;                                                                     // both orientations
	}
	const int xfix = XK ? ( wr0 + ty ) * 128 : ( wr0 >> 4 ) * 2048;
	const int yfix = YK ? ( wc0 + tx ) * 128 : ( wc0 >> 4 ) * 2048;

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

		unsigned long long accP[4][8], accQ[4][8];  // P = sum xr*(yr,yi), Q = sum xi*(yr,yi)
		#pragma unroll
		for ( int i = 0; i < 4; ++i )
			#pragma unroll
			for ( int j = 0; j < 8; ++j ) { accP[i][j] = 0ull; accQ[i][j] = 0ull; }

		// operands of one group of 2 k steps: x[row][k step], y[col][k step] (complex)
		auto load_group = [&]( float2 ( &xv )[4][2], float2 ( &yv )[8][2], int st, int kg )
		{
			const unsigned char* xs = smem + (size_t)st * Cfg::STAGE_BYTES + xfix;
			const unsigned char* ys = smem + (size_t)st * Cfg::STAGE_BYTES + Cfg::X_BYTES + yfix;
			if constexpr ( XK )
			{
				#pragma unroll
				for ( int i = 0; i < 4; ++i )
				{
					const float4 v = *reinterpret_cast<const float4*>( xs + i * 512 + xsw[i & 1][kg] );
					xv[i][0] = make_float2( v.x, v.y ); xv[i][1] = make_float2( v.z, v.w );
				}
			}
			else
			{
				#pragma unroll
				for ( int kk = 0; kk < 2; ++kk )
				{
					const int k = kg * 2 + kk;
					#pragma unroll
					for ( int h = 0; h < 2; ++h )
					{
						const float4 v = *reinterpret_cast<const float4*>( xs + k * 128 + xsw[h][k & 7] );
						xv[h * 2 + 0][kk] = make_float2( v.x, v.y ); xv[h * 2 + 1][kk] = make_float2( v.z, v.w );
					}
				}
			}
			if constexpr ( YK )
			{
				#pragma unroll
				for ( int j = 0; j < 8; ++j )
				{
					const float4 v = *reinterpret_cast<const float4*>( ys + j * 1024 + ysw[kg] );
					yv[j][0] = make_float2( v.x, v.y ); yv[j][1] = make_float2( v.z, v.w );
				}
			}
			else
			{
				#pragma unroll
				for ( int kk = 0; kk < 2; ++kk )
				{
					const int k = kg * 2 + kk;
					#pragma unroll
					for ( int h = 0; h < 4; ++h )
					{
						const float4 v = *reinterpret_cast<const float4*>( ys + h * 2048 + k * 128 + ysw[k & 7] );
						yv[h * 2 + 0][kk] = make_float2( v.x, v.y ); yv[h * 2 + 1][kk] = make_float2( v.z, v.w );
					}
				}
			}
		};
		auto fma_group = [&]( const float2 ( &xv )[4][2], const float2 ( &yv )[8][2] )
		{
			#pragma unroll
			for ( int kk = 0; kk < 2; ++kk )
			{
				unsigned long long y2[8];
				#pragma unroll
				for ( int j = 0; j < 8; ++j ) y2[j] = pack2( yv[j][kk].x, yv[j][kk].y );
				#pragma unroll
				for ( int i = 0; i < 4; ++i )
				{
					const unsigned long long xr2 = pack2( xv[i][kk].x, xv[i][kk].x ), xi2 = pack2( xv[i][kk].y, xv[i][kk].y );
					#pragma unroll
					for ( int j = 0; j < 8; ++j )
					{
						accP[i][j] = ffma2( xr2, y2[j], accP[i][j] );
						accQ[i][j] = ffma2( xi2, y2[j], accQ[i][j] );
					}
				}
			}
		};

		float2 xa[4][2], ya[8][2];
		int64_t kt0 = 0, kt1 = KT;
		if constexpr ( TRI ) tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
		for ( int64_t kt = kt0; kt < kt1; ++kt )
		{
			mbar_wait( full_bar( stage ), phase );
			#pragma unroll
			for ( int kg = 0; kg < BK / 2; ++kg )
			{
				load_group( xa, ya, stage, kg );
				fma_group( xa, ya );
			}
			__syncwarp();
			if ( lane == 0 ) mbar_arrive( empty_bar( stage ) );
			if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
		}

		// ---- epilogue: D = alpha*acc + beta*D (beta == 0: D is not read); complex scalars as bli_tscals / bli_txpbys
		int dlo = 0, dhi = 0;
		if constexpr ( TRI ) tri_band( g, p0, q0, dlo, dhi );
		auto keep = [&]( int d ) { if constexpr ( TRI ) return in_band( d, dlo, dhi ); else return true; };
		if constexpr ( CST )
		{
			// the four D stages of this tile; every warp walks the ring, row-group wr0 / 16 reads its rows
			const bool fast = ( !YK && g.d_vec_ok && q_lim == BQ );
			const float sx = cjx ? -1.f : 1.f, sy = cjy ? -1.f : 1.f;
			#pragma unroll 1
			for ( int qd = 0; qd < BP / 16; ++qd )
			{
				mbar_wait( full_bar( stage ), phase );
				if ( fast && qd == ( wr0 >> 4 ) )
				{
					const unsigned char* cs = smem + (size_t)stage * Cfg::STAGE_BYTES + ( wc0 >> 4 ) * 2048;
					#pragma unroll
					for ( int i = 0; i < 4; ++i )
					{
						const int pl = row_of( i );
						if ( pl >= p_lim ) continue;
						const int r = pl & 15;
						const unsigned char* rowp = cs + r * 128 + ( ( tx ^ ( r & 7 ) ) << 4 );
						float4* dp = reinterpret_cast<float4*>( g.D + ( p0 + pl ) * g.ldd + q0 + col_of( 0 ) );
						#pragma unroll
						for ( int h = 0; h < 4; ++h )
						{
							const float4 o = *reinterpret_cast<const float4*>( rowp + h * 2048 );      // two complex elements
							float4 res;
							#pragma unroll
							for ( int e = 0; e < 2; ++e )
							{
								float px, py, qx, qy;
								unpack2( accP[i][2 * h + e], px, py ); unpack2( accQ[i][2 * h + e], qx, qy );
								const float ar = px - sx * sy * qy, ai = sy * py + sx * qx;
								float rr, ri;
								cscal( g.alpha.x, g.alpha.y, ar, ai, rr, ri );
								cxpby( g.beta.x, g.beta.y, e ? o.z : o.x, e ? o.w : o.y, rr, ri );
								if ( e ) { res.z = rr; res.w = ri; } else { res.x = rr; res.y = ri; }
							}
							__stcs( dp + h * 8, res );       // 16 complex columns = 8 float4 further
						}
					}
				}
				__syncwarp();
				if ( lane == 0 ) mbar_arrive( empty_bar( stage ) );
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
			if ( fast ) continue;
		}
		#pragma unroll
		for ( int i = 0; i < 4; ++i )
		{
			const int pl = row_of( i );
			if ( pl >= p_lim ) continue;
			float2* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			float2 o[8];
			#pragma unroll
			for ( int j = 0; j < 8; ++j )
			{
				const int ql = col_of( j );
				o[j] = ( !g.beta_is_zero && ql < q_lim && keep( ql - pl ) ) ? drow[ql] : make_float2( 0.f, 0.f );
			}
			#pragma unroll
			for ( int j = 0; j < 8; ++j )
			{
				const int ql = col_of( j );
				if ( ql >= q_lim || !keep( ql - pl ) ) continue;
				float px, py, qx, qy;
				unpack2( accP[i][j], px, py ); unpack2( accQ[i][j], qx, qy );
				const float sx = cjx ? -1.f : 1.f, sy = cjy ? -1.f : 1.f;
				const float ar = px - sx * sy * qy, ai = sy * py + sx * qx;
				float rr, ri;
				cscal( g.alpha.x, g.alpha.y, ar, ai, rr, ri );
				if ( !g.beta_is_zero )
				{
					cxpby( g.beta.x, g.beta.y, o[j].x, o[j].y, rr, ri );
				}
				drow[ql] = make_float2( rr, ri );
			}
		}
	}
}

} // namespace b200
