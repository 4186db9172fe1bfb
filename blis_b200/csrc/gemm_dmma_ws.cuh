// gemm_dmma_ws.cuh -- warp-specialised FP64 tensor-core gemm (d and z), the
// production path for large problems.
//
// Same contract and reference mapping as gemm_dmma.cuh (see there).  What
// changes is how the reference's pack -> barrier -> macrokernel sequence
// (frame/1m/packm/bli_packm_int.c:52,64: two thread barriers per packed block)
// is realised on the SM:
//   * one PRODUCER warpgroup streams X/Y k-slabs from the caller's layout into
//     a ring of shared-memory stages with cp.async and signals "full" through
//     an mbarrier (cp.async.mbarrier.arrive), i.e. packm becomes an
//     asynchronous copy engine that never blocks the math warps;
//   * two CONSUMER warpgroups (8 warps) own the 8x8 DMMA tiles; each warp waits
//     only on the stage it needs and releases it through an "empty" mbarrier,
//     so there is no CTA-wide barrier in the k loop at all;
//   * registers move from the producer to the consumers with setmaxnreg, which
//     pays for double-buffered operand fragments next to 128 accumulators;
//   * the ring keeps running across output tiles: the producer prefetches the
//     next tile's first k-slabs while the consumers run the epilogue.
// ncu on the single-role kernel (profiles/r01_dgemm_ncu_summary.md) showed the
// DMMA pipe 84% busy with the rest lost to barrier / short-scoreboard /
// long-scoreboard stalls; this structure removes exactly those.
#pragma once
#include "common.cuh"
#include "gemm_dmma.cuh"

namespace b200 {

// ---- mbarrier helpers (shared::cta) ----------------------------------------------
__device__ __forceinline__ void mbar_init( uint32_t bar, int count )
{
	asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory" );
}
__device__ __forceinline__ void mbar_arrive( uint32_t bar )
{
	asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory" );
}
__device__ __forceinline__ void mbar_wait( uint32_t bar, uint32_t parity )
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" :: "r"(bar), "r"(parity) : "memory" );
}
// The mbarrier receives one arrival when all cp.async issued so far by this
// thread have landed (count pre-accounted at init: .noinc).
__device__ __forceinline__ void cp_async_arrive_noinc( uint32_t bar )
{
	asm volatile( "cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" :: "r"(bar) : "memory" );
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile( "setmaxnreg.inc.sync.aligned.u32 %0;\n" :: "n"(N) ); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile( "setmaxnreg.dec.sync.aligned.u32 %0;\n" :: "n"(N) ); }

template <typename T, int BP, int BQ, int BK, int WP, int WQ, int STAGES, bool XK, bool YK, bool AL>
struct DmmaWsCfg : DmmaCfg<T, BP, BQ, BK, WP, WQ, STAGES, XK, YK, AL>
{
	using Base = DmmaCfg<T, BP, BQ, BK, WP, WQ, STAGES, XK, YK, AL>;
	static constexpr int NCONS  = WP * WQ * 32;        // consumer threads (8 warps)
	static constexpr int NPROD  = 128;                 // one producer warpgroup
	static constexpr int NT_ALL = NCONS + NPROD;
	static constexpr int BAR_BYTES  = 2 * STAGES * 8 + 4 * 8 + 16;   // stage ring + tile-scheduler ring (2 slots)
	static constexpr int SMEM_BYTES = Base::STAGE_BYTES * STAGES + BAR_BYTES;
	static_assert( NCONS == 256 || NCONS == 128, "one or two consumer warpgroups" );
	// one consumer warpgroup: two CTAs share an SM (one CTA's epilogue overlaps the other's MMAs)
	static constexpr int CTAS_PER_SM = ( NCONS == 128 ) ? 2 : 1;
	static constexpr int REG_PROD = ( NCONS == 128 ) ? 40 : 56;
	static constexpr int REG_CONS = ( NCONS == 128 ) ? 216 : 224;
};

template <typename T, int BP, int BQ, int BK, int WP, int WQ, int STAGES, bool XK, bool YK, bool AL, bool TRI = false>
__global__ void __launch_bounds__( WP * WQ * 32 + 128, ( WP * WQ == 4 ) ? 2 : 1 )
gemm_dmma_ws_kernel( const GemmArgs<T> g )
{
	using Cfg = DmmaWsCfg<T, BP, BQ, BK, WP, WQ, STAGES, XK, YK, AL>;
	constexpr bool CPLX = Cfg::CPLX;
	constexpr int  MT = Cfg::MT, NTL = Cfg::NTL, KS = BK / 4;
	constexpr int  SXK = BK + Cfg::PADK, SXP = BP + Cfg::PADC, SYQ = BQ + Cfg::PADC;
	constexpr int  STAGE_ELEMS = Cfg::XS_ELEMS + Cfg::YS_ELEMS;

	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* const smem = reinterpret_cast<T*>( smem_raw );
	const uint32_t bar_base = smem_u32( smem_raw + (size_t)Cfg::STAGE_BYTES * STAGES );
	auto full_bar  = [&]( int s ) { return bar_base + (uint32_t)s * 8u; };
	auto empty_bar = [&]( int s ) { return bar_base + (uint32_t)( STAGES + s ) * 8u; };
	// Tile scheduler: producer thread 0 draws the next output tile (dynamically from a global counter
	// when g.tile_counter is set: CTAs that start late or share their SM simply draw fewer tiles) and
	// publishes it to the consumers through a 2-slot ring guarded by mbarriers.
	auto sched_full  = [&]( int s ) { return bar_base + (uint32_t)( 2 * STAGES + s ) * 8u; };
	auto sched_empty = [&]( int s ) { return bar_base + (uint32_t)( 2 * STAGES + 2 + s ) * 8u; };
	volatile int* const sched_tile = reinterpret_cast<volatile int*>( smem_raw + (size_t)Cfg::STAGE_BYTES * STAGES + ( 2 * STAGES + 4 ) * 8 );

	const int tid = threadIdx.x;
	if ( tid == 0 )
	{
		#pragma unroll
		for ( int s = 0; s < STAGES; ++s )
		{
			mbar_init( full_bar( s ),  Cfg::NPROD );      // one cp.async-completion arrival per producer thread
			mbar_init( empty_bar( s ), WP * WQ );         // one arrival per consumer warp
		}
		#pragma unroll
		for ( int s = 0; s < 2; ++s )
		{
			mbar_init( sched_full( s ),  1 );             // the scheduling thread
			mbar_init( sched_empty( s ), WP * WQ );       // one arrival per consumer warp
		}
	}
	__syncthreads();

	const int64_t KT_SEG = ( g.K + BK - 1 ) / BK;          // k tiles per panel
	const int64_t KT = KT_SEG * g.nseg;                     // the consumers see one long k loop
	const int num_tiles = g.tiles_p * g.tiles_q;

	if ( tid >= Cfg::NCONS )
	{
		// =========================== PRODUCER warpgroup ===========================
		setmaxnreg_dec<Cfg::REG_PROD>();
		const int ptid = tid - Cfg::NCONS;
		int stage = 0; uint32_t phase = 0;
		for ( int it = 0; ; ++it )
		{
			const int slot = it & 1;
			if ( ptid == 0 )
			{
				mbar_wait( sched_empty( slot ), ( ( it >> 1 ) & 1 ) ^ 1u );
				const int t = g.tile_counter ? atomicAdd( g.tile_counter, 1 ) : (int)( blockIdx.x + (unsigned)it * gridDim.x );
				sched_tile[slot] = t;
				mbar_arrive( sched_full( slot ) );
			}
			asm volatile( "bar.sync 1, 128;\n" ::: "memory" );      // producer warpgroup only
			const int tile = sched_tile[slot];
			if ( tile >= num_tiles ) break;
			int tp, tq;
			tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
			const int64_t p0 = (int64_t)tp * BP, q0 = (int64_t)tq * BQ;
			const int p_lim = (int)min( (int64_t)BP, g.P - p0 );
			const int q_lim = (int)min( (int64_t)BQ, g.Q - q0 );
			if ( TRI && tri_skip_tile( g, p0, q0, p_lim, q_lim ) ) continue;
			const int64_t xo = XK ? p0 * g.ldx : p0, yo = YK ? q0 * g.ldy : q0;
			if ( !g.beta_is_zero )
			{
				// pull this tile of D towards L2 while the k loop runs; the epilogue reads it
				constexpr int LINES_PER_ROW = ( BQ * (int)sizeof(T) + 127 ) / 128;
				for ( int e = ptid; e < BP * LINES_PER_ROW; e += Cfg::NPROD )
				{
					const int r = e / LINES_PER_ROW, l = e % LINES_PER_ROW;
					if ( r < p_lim && l * ( 128 / (int)sizeof(T) ) < q_lim )
					{
						const T* pd = g.D + ( p0 + r ) * g.ldd + q0 + l * ( 128 / (int)sizeof(T) );
						asm volatile( "prefetch.global.L2 [%0];\n" :: "l"(pd) );
					}
				}
			}
			for ( int seg = 0; seg < g.nseg; ++seg )
			{
			const T* gx = ( seg == 0 ? g.X : g.Xseg[seg - 1] ) + xo;
			const T* gy = ( seg == 0 ? g.Y : g.Yseg[seg - 1] ) + yo;
			int64_t kt0 = 0, kt1 = KT_SEG;
			if constexpr ( TRI ) tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT_SEG, kt0, kt1 );    // nseg == 1 with a triangular operand
			for ( int64_t kt = kt0; kt < kt1; ++kt )
			{
				mbar_wait( empty_bar( stage ), phase ^ 1u );
				const int k_lim = (int)min( (int64_t)BK, g.K - kt * BK );
				T* xs = smem + (size_t)stage * STAGE_ELEMS;
				T* ys = xs + Cfg::XS_ELEMS;
				if constexpr ( XK ) load_tile<T, BP, BK, Cfg::PADK, Cfg::NPROD, AL>( smem_u32( xs ), gx + kt * BK, g.ldx, p_lim, k_lim, ptid );
				else                load_tile<T, BK, BP, Cfg::PADC, Cfg::NPROD, AL>( smem_u32( xs ), gx + kt * BK * g.ldx, g.ldx, k_lim, p_lim, ptid );
				if constexpr ( YK ) load_tile<T, BQ, BK, Cfg::PADK, Cfg::NPROD, AL>( smem_u32( ys ), gy + kt * BK, g.ldy, q_lim, k_lim, ptid );
				else                load_tile<T, BK, BQ, Cfg::PADC, Cfg::NPROD, AL>( smem_u32( ys ), gy + kt * BK * g.ldy, g.ldy, k_lim, q_lim, ptid );
				cp_async_arrive_noinc( full_bar( stage ) );
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
			}
		}
		cp_async_wait<0>();
		if ( ptid == 0 && g.tile_counter )
		{
			// the last CTA to finish re-arms the counter pair for the next launch
			if ( atomicAdd( g.tile_counter + 1, 1 ) == (int)gridDim.x - 1 ) { g.tile_counter[0] = 0; g.tile_counter[1] = 0; __threadfence(); }
		}
		return;
	}

	// =============================== CONSUMER warps ===============================
	setmaxnreg_inc<Cfg::REG_CONS>();
	const int lane = tid & 31, warp = tid >> 5;
	const int gq   = lane >> 2, t4 = lane & 3;
	const int wp0  = ( warp / WQ ) * Cfg::WTP;
	const int wq0  = ( warp % WQ ) * Cfg::WTQ;
	const bool cjx = CPLX && g.conjx, cjy = CPLX && g.conjy;

	// per-lane element offsets of the k4-step-0 fragments inside a stage
	const int xoff = XK ? ( wp0 + gq ) * SXK + t4 : t4 * SXP + wp0 + gq;
	const int yoff = YK ? ( wq0 + gq ) * SXK + t4 : t4 * SYQ + wq0 + gq;
	constexpr int XI = XK ? 8 * SXK : 8;            // next 8x8 tile along p
	constexpr int XS = XK ? 4 : 4 * SXP;            // next k4 step
	constexpr int YJ = YK ? 8 * SXK : 8;
	constexpr int YS = YK ? 4 : 4 * SYQ;

	int stage = 0; uint32_t phase = 0;

	auto load_frags = [&]( T ( &xf )[MT], T ( &yf )[NTL], int st, int kk )
	{
		const T* xs = smem + (size_t)st * STAGE_ELEMS + xoff + kk * XS;
		const T* ys = smem + (size_t)st * STAGE_ELEMS + Cfg::XS_ELEMS + yoff + kk * YS;
		#pragma unroll
		for ( int i = 0; i < MT; ++i ) xf[i] = xs[i * XI];
		#pragma unroll
		for ( int j = 0; j < NTL; ++j ) yf[j] = ys[j * YJ];
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

		double acc[CPLX ? 2 : 1][MT][NTL][2];
		#pragma unroll
		for ( int c = 0; c < ( CPLX ? 2 : 1 ); ++c )
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
				#pragma unroll
				for ( int j = 0; j < NTL; ++j ) { acc[c][i][j][0] = 0.0; acc[c][i][j][1] = 0.0; }

		T xa[MT], ya[NTL], xb[MT], yb[NTL];
		mbar_wait( full_bar( stage ), phase );
		load_frags( xa, ya, stage, 0 );

		auto mma_step = [&]( T ( &xf )[MT], T ( &yf )[NTL] )
		{
			if constexpr ( !CPLX )
			{
				#pragma unroll
				for ( int i = 0; i < MT; ++i )
					#pragma unroll
					for ( int j = 0; j < NTL; ++j )
						dmma884( acc[0][i][j][0], acc[0][i][j][1], xf[i], yf[j] );
			}
			else
			{
				// the two DMMAs into one accumulator are issued a whole pass apart (see gemm_zmma_tma.cuh)
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
			}
		};

		int64_t kt0 = 0, kt1 = KT;
		if constexpr ( TRI ) tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
		for ( int64_t kt = kt0; kt < kt1; ++kt )
		{
			// KS k4-steps per stage, fragments double-buffered (a <-> b); KS is even
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
					// last step of this stage: prefetch step 0 of the next stage first
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
			}
		}

		// ---- epilogue: D = alpha*acc + beta*D   (beta == 0: D is not read)
		if ( g.d_vec_ok && q_lim == BQ && ( !TRI || tri_tile_interior( g, p0, q0, p_lim, q_lim ) ) )
		{
			// Interior tile: all loads of a tile row are issued before the first store, so a lane has
			// NTL (2*NTL for complex) 16-byte loads in flight instead of one (small-k problems are bound
			// by exactly this read-modify-write of C).
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
			{
				const int pl = wp0 + i * 8 + gq;
				if ( pl >= p_lim ) continue;
				if constexpr ( !CPLX )
				{
					double2* __restrict__ dp = reinterpret_cast<double2*>( g.D + ( p0 + pl ) * g.ldd + q0 + wq0 + 2 * t4 );
					double2 o[NTL];
					if ( !g.beta_is_zero )
					{
						#pragma unroll
						for ( int j = 0; j < NTL; ++j ) o[j] = __ldcs( dp + j * 4 );
					}
					#pragma unroll
					for ( int j = 0; j < NTL; ++j )
					{
						double r0 = g.alpha * acc[0][i][j][0];
						double r1 = g.alpha * acc[0][i][j][1];
						if ( !g.beta_is_zero ) { r0 = fma( g.beta, o[j].x, r0 ); r1 = fma( g.beta, o[j].y, r1 ); }
						__stcs( dp + j * 4, make_double2( r0, r1 ) );
					}
				}
				else
				{
					double2* __restrict__ dp = reinterpret_cast<double2*>( g.D + ( p0 + pl ) * g.ldd + q0 + wq0 + 2 * t4 );
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
							const double ar = acc[0][i][j][e], ai = acc[1][i][j][e];
							double rr, ri;
							cscal( g.alpha.x, g.alpha.y, ar, ai, rr, ri );
							if ( !g.beta_is_zero )
							{
								cxpby( g.beta.x, g.beta.y, o[j][e].x, o[j][e].y, rr, ri );
							}
							__stcs( dp + j * 8 + e, make_double2( rr, ri ) );
						}
				}
			}
			continue;
		}
		int dlo = 0, dhi = 0;
		if constexpr ( TRI ) tri_band( g, p0, q0, dlo, dhi );
		auto keep = [&]( int d ) { if constexpr ( TRI ) return in_band( d, dlo, dhi ); else return true; };
		#pragma unroll
		for ( int i = 0; i < MT; ++i )
		{
			const int pl = wp0 + i * 8 + gq;
			if ( pl >= p_lim ) continue;
			T* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			#pragma unroll
			for ( int j = 0; j < NTL; ++j )
			{
				const int ql = wq0 + j * 8 + 2 * t4;
				const bool one = ( ql < q_lim && keep( ql - pl ) );
				const bool two = ( ql + 1 < q_lim && keep( ql + 1 - pl ) );
				if ( !one && !two ) continue;
				if constexpr ( !CPLX )
				{
					double r0 = g.alpha * acc[0][i][j][0];
					double r1 = g.alpha * acc[0][i][j][1];
					if ( one && two && g.d_vec_ok )
					{
						double2* dp = reinterpret_cast<double2*>( drow + ql );
						if ( !g.beta_is_zero ) { const double2 o = *dp; r0 = fma( g.beta, o.x, r0 ); r1 = fma( g.beta, o.y, r1 ); }
						*dp = make_double2( r0, r1 );
					}
					else
					{
						if ( one )
						{
							if ( !g.beta_is_zero ) r0 = fma( g.beta, drow[ql], r0 );
							drow[ql] = r0;
						}
						if ( two )
						{
							if ( !g.beta_is_zero ) r1 = fma( g.beta, drow[ql + 1], r1 );
							drow[ql + 1] = r1;
						}
					}
				}
				else
				{
					#pragma unroll
					for ( int e = 0; e < 2; ++e )
					{
						if ( e == 0 ? !one : !two ) continue;
						const double ar = acc[0][i][j][e], ai = acc[1][i][j][e];
						double rr, ri;
						cscal( g.alpha.x, g.alpha.y, ar, ai, rr, ri );
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
}

} // namespace b200
