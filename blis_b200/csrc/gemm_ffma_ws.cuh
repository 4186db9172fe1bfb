// gemm_ffma_ws.cuh -- warp-specialised FP32 FMA-pipe gemm for s and c.
//
// Same contract and reference mapping as gemm_dmma.cuh / gemm_ffma.cuh; same
// producer/consumer structure as gemm_dmma_ws.cuh (cp.async producer warpgroup,
// full/empty mbarrier ring, setmaxnreg, ring running across output tiles).
// The FP32 pipe is ISSUE bound (one FFMA per scheduler per clock), so what
// matters here is the number of non-FFMA instructions and shared-memory
// wavefronts per FFMA:
//   * each lane owns a TP x TQ register tile fed by 16-byte LDS;
//   * lanes are laid out 4 (rows) x 8 (columns) inside a warp, so one LDS.128 of
//     the X operand touches 64 contiguous bytes and one of the Y operand 128
//     contiguous bytes: every operand load is a single shared-memory wavefront
//     (the 16x2 layout of gemm_ffma.cuh needed 3x as many);
//   * operand vectors are double buffered across k steps.
// No TF32 anywhere: s/c results are IEEE fp32 FMA results.
#pragma once
#include "common.cuh"
#include "gemm_dmma.cuh"      // GemmArgs, tile_coords, load_tile
#include "gemm_ffma.cuh"      // load_tile_t
#include "gemm_dmma_ws.cuh"   // mbarrier helpers, setmaxnreg

namespace b200 {

template <typename T, int BP, int BQ, int BK, int TP, int TQ, int STAGES>
struct FfmaWsCfg
{
	static constexpr bool CPLX = Elem<T>::cplx;
	static constexpr int  VE   = 16 / (int)sizeof(T);     // elements per 16-byte vector (4 float, 2 float2)
	static constexpr int  GP   = TP / VE, GQ = TQ / VE;   // vector groups per lane
	static constexpr int  WR   = GP * 4 * VE;             // warp tile rows    (4 lane rows)
	static constexpr int  WC   = GQ * 8 * VE;             // warp tile columns (8 lane columns)
	static constexpr int  WARPS_P = BP / WR, WARPS_Q = BQ / WC;
	static constexpr int  NCONS = WARPS_P * WARPS_Q * 32;
	static constexpr int  NPROD = 128;
	static constexpr int  NT_ALL = NCONS + NPROD;
	static constexpr int  PAD  = VE;
	static constexpr int  XS_ELEMS = BK * ( BP + PAD );
	static constexpr int  YS_ELEMS = BK * ( BQ + PAD );
	static constexpr int  STAGE_BYTES = ( XS_ELEMS + YS_ELEMS ) * (int)sizeof(T);
	static constexpr int  BAR_BYTES  = 2 * STAGES * 8;
	static constexpr int  SMEM_BYTES = STAGE_BYTES * STAGES + BAR_BYTES;
	static_assert( NCONS == 256, "two consumer warpgroups" );
	static_assert( TP % VE == 0 && TQ % VE == 0 && BP % WR == 0 && BQ % WC == 0, "tile shape" );
};

template <typename T, int BP, int BQ, int BK, int TP, int TQ, int STAGES, bool XK, bool YK, bool AL>
__global__ void __launch_bounds__( 384, 1 )
gemm_ffma_ws_kernel( const GemmArgs<T> g )
{
	using Cfg = FfmaWsCfg<T, BP, BQ, BK, TP, TQ, STAGES>;
	constexpr bool CPLX = Cfg::CPLX;
	constexpr int  VE = Cfg::VE, GP = Cfg::GP, GQ = Cfg::GQ, PAD = Cfg::PAD;
	constexpr int  SXP = BP + PAD, SYQ = BQ + PAD;
	constexpr int  STAGE_ELEMS = Cfg::XS_ELEMS + Cfg::YS_ELEMS;

	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* const smem = reinterpret_cast<T*>( smem_raw );
	const uint32_t bar_base = smem_u32( smem_raw + (size_t)Cfg::STAGE_BYTES * STAGES );
	auto full_bar  = [&]( int s ) { return bar_base + (uint32_t)s * 8u; };
	auto empty_bar = [&]( int s ) { return bar_base + (uint32_t)( STAGES + s ) * 8u; };

	const int tid = threadIdx.x;
	if ( tid == 0 )
	{
		#pragma unroll
		for ( int s = 0; s < STAGES; ++s )
		{
			mbar_init( full_bar( s ),  Cfg::NPROD );
			mbar_init( empty_bar( s ), Cfg::NCONS / 32 );
		}
	}
	__syncthreads();

	const int64_t KT = ( g.K + BK - 1 ) / BK;
	const int num_tiles = g.tiles_p * g.tiles_q;

	if ( tid >= Cfg::NCONS )
	{
		// =========================== PRODUCER warpgroup ===========================
		setmaxnreg_dec<56>();
		const int ptid = tid - Cfg::NCONS;
		int stage = 0; uint32_t phase = 0;
		for ( int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x )
		{
			int tp, tq;
			tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
			const int64_t p0 = (int64_t)tp * BP, q0 = (int64_t)tq * BQ;
			const int p_lim = (int)min( (int64_t)BP, g.P - p0 );
			const int q_lim = (int)min( (int64_t)BQ, g.Q - q0 );
			if ( tri_skip_tile( g, p0, q0, p_lim, q_lim ) ) continue;
			const T* gx = XK ? g.X + p0 * g.ldx : g.X + p0;
			const T* gy = YK ? g.Y + q0 * g.ldy : g.Y + q0;
			int64_t kt0, kt1;
			tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
			for ( int64_t kt = kt0; kt < kt1; ++kt )
			{
				mbar_wait( empty_bar( stage ), phase ^ 1u );
				const int k_lim = (int)min( (int64_t)BK, g.K - kt * BK );
				T* xs = smem + (size_t)stage * STAGE_ELEMS;
				T* ys = xs + Cfg::XS_ELEMS;
				if constexpr ( XK ) load_tile_t<T, BP, BK, PAD, Cfg::NPROD>( smem_u32( xs ), gx + kt * BK, g.ldx, p_lim, k_lim, ptid );
				else                load_tile<T, BK, BP, PAD, Cfg::NPROD, AL>( smem_u32( xs ), gx + kt * BK * g.ldx, g.ldx, k_lim, p_lim, ptid );
				if constexpr ( YK ) load_tile_t<T, BQ, BK, PAD, Cfg::NPROD>( smem_u32( ys ), gy + kt * BK, g.ldy, q_lim, k_lim, ptid );
				else                load_tile<T, BK, BQ, PAD, Cfg::NPROD, AL>( smem_u32( ys ), gy + kt * BK * g.ldy, g.ldy, k_lim, q_lim, ptid );
				cp_async_arrive_noinc( full_bar( stage ) );
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
		}
		cp_async_wait<0>();
		return;
	}

	// =============================== CONSUMER warps ===============================
	setmaxnreg_inc<224>();
	const int lane = tid & 31, warp = tid >> 5;
	const int ty = lane >> 3, tx = lane & 7;
	const int wr0 = ( warp / Cfg::WARPS_Q ) * Cfg::WR;
	const int wc0 = ( warp % Cfg::WARPS_Q ) * Cfg::WC;
	const bool cjx = CPLX && g.conjx, cjy = CPLX && g.conjy;
	// lane's first element in each vector group: rows wr0 + gp*(4*VE) + ty*VE, cols wc0 + gq*(8*VE) + tx*VE
	const int xoff = wr0 + ty * VE;
	const int yoff = wc0 + tx * VE;

	int stage = 0; uint32_t phase = 0;

	auto load_vecs = [&]( float4 ( &xv )[GP], float4 ( &yv )[GQ], int st, int k )
	{
		const T* xs = smem + (size_t)st * STAGE_ELEMS + k * SXP + xoff;
		const T* ys = smem + (size_t)st * STAGE_ELEMS + Cfg::XS_ELEMS + k * SYQ + yoff;
		#pragma unroll
		for ( int gp = 0; gp < GP; ++gp ) xv[gp] = *reinterpret_cast<const float4*>( xs + gp * 4 * VE );
		#pragma unroll
		for ( int gq = 0; gq < GQ; ++gq ) yv[gq] = *reinterpret_cast<const float4*>( ys + gq * 8 * VE );
	};

	for ( int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x )
	{
		int tp, tq;
		tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
		const int64_t p0 = (int64_t)tp * BP, q0 = (int64_t)tq * BQ;
		const int p_lim = (int)min( (int64_t)BP, g.P - p0 );
		const int q_lim = (int)min( (int64_t)BQ, g.Q - q0 );
		if ( tri_skip_tile( g, p0, q0, p_lim, q_lim ) ) continue;

		T acc[TP][TQ];
		#pragma unroll
		for ( int i = 0; i < TP; ++i )
			#pragma unroll
			for ( int j = 0; j < TQ; ++j )
			{
				if constexpr ( CPLX ) acc[i][j] = make_float2( 0.f, 0.f );
				else                  acc[i][j] = 0.f;
			}

		auto fma_step = [&]( const float4 ( &xv )[GP], const float4 ( &yv )[GQ] )
		{
			if constexpr ( !CPLX )
			{
				float x[TP], y[TQ];
				#pragma unroll
				for ( int gp = 0; gp < GP; ++gp ) { x[gp * 4] = xv[gp].x; x[gp * 4 + 1] = xv[gp].y; x[gp * 4 + 2] = xv[gp].z; x[gp * 4 + 3] = xv[gp].w; }
				#pragma unroll
				for ( int gq = 0; gq < GQ; ++gq ) { y[gq * 4] = yv[gq].x; y[gq * 4 + 1] = yv[gq].y; y[gq * 4 + 2] = yv[gq].z; y[gq * 4 + 3] = yv[gq].w; }
				#pragma unroll
				for ( int i = 0; i < TP; ++i )
					#pragma unroll
					for ( int j = 0; j < TQ; ++j )
						acc[i][j] = fmaf( x[i], y[j], acc[i][j] );
			}
			else
			{
				float xr[TP], xi[TP], yr[TQ], yi[TQ];
				#pragma unroll
				for ( int gp = 0; gp < GP; ++gp )
				{
					xr[gp * 2] = xv[gp].x; xi[gp * 2] = flip_sign( xv[gp].y, cjx );
					xr[gp * 2 + 1] = xv[gp].z; xi[gp * 2 + 1] = flip_sign( xv[gp].w, cjx );
				}
				#pragma unroll
				for ( int gq = 0; gq < GQ; ++gq )
				{
					yr[gq * 2] = yv[gq].x; yi[gq * 2] = flip_sign( yv[gq].y, cjy );
					yr[gq * 2 + 1] = yv[gq].z; yi[gq * 2 + 1] = flip_sign( yv[gq].w, cjy );
				}
				#pragma unroll
				for ( int i = 0; i < TP; ++i )
				{
					const float nxi = -xi[i];
					#pragma unroll
					for ( int j = 0; j < TQ; ++j )
					{
						acc[i][j].x = fmaf( xr[i], yr[j], acc[i][j].x );
						acc[i][j].x = fmaf( nxi,   yi[j], acc[i][j].x );
						acc[i][j].y = fmaf( xr[i], yi[j], acc[i][j].y );
						acc[i][j].y = fmaf( xi[i], yr[j], acc[i][j].y );
					}
				}
			}
		};

		float4 xa[GP], ya[GQ], xb[GP], yb[GQ];
		mbar_wait( full_bar( stage ), phase );
		load_vecs( xa, ya, stage, 0 );

		int64_t kt0, kt1;
		tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
		for ( int64_t kt = kt0; kt < kt1; ++kt )
		{
			#pragma unroll
			for ( int k = 0; k < BK; k += 2 )
			{
				load_vecs( xb, yb, stage, k + 1 );
				fma_step( xa, ya );
				if ( k + 2 < BK )
				{
					load_vecs( xa, ya, stage, k + 2 );
					fma_step( xb, yb );
				}
				else
				{
					int ns = stage + 1; uint32_t nph = phase;
					if ( ns == STAGES ) { ns = 0; nph ^= 1u; }
					if ( kt + 1 < kt1 )
					{
						mbar_wait( full_bar( ns ), nph );
						load_vecs( xa, ya, ns, 0 );
					}
					fma_step( xb, yb );
					__syncwarp();
					if ( lane == 0 ) mbar_arrive( empty_bar( stage ) );
					stage = ns; phase = nph;
				}
			}
		}

		// ---- epilogue: D = alpha*acc + beta*D (beta == 0: D is not read)
		int dlo, dhi;
		tri_band( g, p0, q0, dlo, dhi );
		#pragma unroll
		for ( int i = 0; i < TP; ++i )
		{
			const int pl = xoff + ( i / VE ) * 4 * VE + ( i % VE );
			if ( pl >= p_lim ) continue;
			T* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			#pragma unroll
			for ( int gq = 0; gq < GQ; ++gq )
			{
				const int ql = yoff + gq * 8 * VE;
				if ( ql >= q_lim ) continue;
				T r[VE];
				#pragma unroll
				for ( int e = 0; e < VE; ++e )
				{
					const T a = acc[i][gq * VE + e];
					if constexpr ( CPLX ) cscal( g.alpha.x, g.alpha.y, a.x, a.y, r[e].x, r[e].y );
					else                  r[e] = g.alpha * a;
				}
				const bool full = ( ql + VE <= q_lim && in_band( ql - pl, dlo, dhi ) && in_band( ql + VE - 1 - pl, dlo, dhi ) );
				if ( full && g.d_vec_ok )
				{
					float4* dp = reinterpret_cast<float4*>( drow + ql );
					if ( !g.beta_is_zero )
					{
						const float4 o = *dp;
						if constexpr ( CPLX )
						{
							cxpby( g.beta.x, g.beta.y, o.x, o.y, r[0].x, r[0].y );
							cxpby( g.beta.x, g.beta.y, o.z, o.w, r[1].x, r[1].y );
						}
						else
						{
							r[0] = fmaf( g.beta, o.x, r[0] ); r[1] = fmaf( g.beta, o.y, r[1] );
							r[2] = fmaf( g.beta, o.z, r[2] ); r[3] = fmaf( g.beta, o.w, r[3] );
						}
					}
					if constexpr ( CPLX ) *dp = make_float4( r[0].x, r[0].y, r[1].x, r[1].y );
					else                  *dp = make_float4( r[0], r[1], r[2], r[3] );
				}
				else
				{
					#pragma unroll
					for ( int e = 0; e < VE; ++e )
					{
						if ( ql + e >= q_lim || !in_band( ql + e - pl, dlo, dhi ) ) continue;
						if ( !g.beta_is_zero )
						{
							const T o = drow[ql + e];
							if constexpr ( CPLX ) cxpby( g.beta.x, g.beta.y, o.x, o.y, r[e].x, r[e].y );
							else                  r[e] = fmaf( g.beta, o, r[e] );
						}
						drow[ql + e] = r[e];
					}
				}
			}
		}
	}
}

} // namespace b200
