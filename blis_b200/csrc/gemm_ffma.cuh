// gemm_ffma.cuh -- FP32 FMA-pipe gemm for s and c (no TF32: results stay IEEE fp32).
//
// Same role and same D/X/Y convention as gemm_dmma.cuh (see the header comment
// there for the reference files it replaces).  The microtile is a TP x TQ
// register block per thread fed by 16-byte shared-memory loads; shared memory
// holds X as [k][p] and Y as [k][q].  A k-contiguous operand is transposed on
// the fly by element-wise cp.async, so no HBM repack is needed for any
// storage/transpose combination.
#pragma once
#include "common.cuh"
#include "gemm_dmma.cuh"   // GemmArgs, tile_coords, load_tile

namespace b200 {

// Transposing tile copy: global [LS rows][LC contiguous] -> shared [c][s]
// (rows of LS+PAD), one element per cp.async.
template <typename T, int LS, int LC, int PAD, int NT>
__device__ __forceinline__ void load_tile_t( uint32_t sbase, const T* __restrict__ g, int64_t ld,
                                             int s_lim, int c_lim, int tid )
{
	constexpr int CPB   = (int)sizeof(T) >= 8 ? 8 : 4;
	constexpr int PER   = (int)sizeof(T) / CPB;          // 1 (float, float2) or 2 (double2)
	constexpr int TOTAL = LS * LC;
	constexpr int ITERS = ( TOTAL + NT - 1 ) / NT;
	constexpr int UNROLL = ITERS <= 8 ? ITERS : 4;
	#pragma unroll UNROLL
	for ( int i = 0; i < ITERS; ++i )
	{
		const int id = tid + i * NT;
		if ( TOTAL % NT != 0 && id >= TOTAL ) break;
		const int s = id / LC, c = id % LC;
		const bool ok = ( s < s_lim && c < c_lim );
		const char* src = ok ? reinterpret_cast<const char*>( g + (int64_t)s * ld + c )
		                     : reinterpret_cast<const char*>( g );
		const uint32_t dst = sbase + (uint32_t)( ( c * ( LS + PAD ) + s ) * (int)sizeof(T) );
		#pragma unroll
		for ( int h = 0; h < PER; ++h )
			cp_async<CPB>( dst + h * CPB, src + ( ok ? h * CPB : 0 ), ok ? CPB : 0 );
	}
}

template <typename T, int BP, int BQ, int BK, int TP, int TQ, int STAGES>
struct FfmaCfg
{
	static constexpr bool CPLX = Elem<T>::cplx;
	static constexpr int  VE   = 16 / (int)sizeof(T);   // elements per 16-byte vector
	static constexpr int  TY   = BP / TP, TX = BQ / TQ; // thread grid
	static constexpr int  NT   = TX * TY;
	static constexpr int  GP   = TP / VE, GQ = TQ / VE; // vector groups per thread
	static constexpr int  PAD  = VE;
	static constexpr int  XS_ELEMS = BK * ( BP + PAD );
	static constexpr int  YS_ELEMS = BK * ( BQ + PAD );
	static constexpr int  STAGE_BYTES = ( XS_ELEMS + YS_ELEMS ) * (int)sizeof(T);
	static constexpr int  SMEM_BYTES  = STAGE_BYTES * STAGES;
	static_assert( TP % VE == 0 && TQ % VE == 0, "microtile" );
};

template <typename T> struct Vec16;
template <> struct Vec16<float>  { using type = float4; };
template <> struct Vec16<float2> { using type = float4; };

template <typename T, int BP, int BQ, int BK, int TP, int TQ, int STAGES, bool XK, bool YK, bool AL>
__global__ void __launch_bounds__( ( BP / TP ) * ( BQ / TQ ), 1 )
gemm_ffma_kernel( const GemmArgs<T> g )
{
	using Cfg = FfmaCfg<T, BP, BQ, BK, TP, TQ, STAGES>;
	constexpr bool CPLX = Cfg::CPLX;
	constexpr int  NT = Cfg::NT, VE = Cfg::VE, GP = Cfg::GP, GQ = Cfg::GQ, PAD = Cfg::PAD;
	constexpr int  SXP = BP + PAD, SYQ = BQ + PAD;

	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* const smem = reinterpret_cast<T*>( smem_raw );

	const int tid = threadIdx.x;
	const int tx  = tid % Cfg::TX, ty = tid / Cfg::TX;

	const int64_t KT = ( g.K + BK - 1 ) / BK;
	const int num_tiles = g.tiles_p * g.tiles_q;
	const bool cjx = CPLX && g.conjx, cjy = CPLX && g.conjy;

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

		auto issue = [&]( int64_t kt, int stage )
		{
			const int k_lim = (int)min( (int64_t)BK, g.K - kt * BK );
			T* xs = smem + (size_t)stage * ( Cfg::XS_ELEMS + Cfg::YS_ELEMS );
			T* ys = xs + Cfg::XS_ELEMS;
			if constexpr ( XK ) load_tile_t<T, BP, BK, PAD, NT>( smem_u32( xs ), gx + kt * BK, g.ldx, p_lim, k_lim, tid );
			else                load_tile<T, BK, BP, PAD, NT, AL>( smem_u32( xs ), gx + kt * BK * g.ldx, g.ldx, k_lim, p_lim, tid );
			if constexpr ( YK ) load_tile_t<T, BQ, BK, PAD, NT>( smem_u32( ys ), gy + kt * BK, g.ldy, q_lim, k_lim, tid );
			else                load_tile<T, BK, BQ, PAD, NT, AL>( smem_u32( ys ), gy + kt * BK * g.ldy, g.ldy, k_lim, q_lim, tid );
		};

		T acc[TP][TQ];
		#pragma unroll
		for ( int i = 0; i < TP; ++i )
			#pragma unroll
			for ( int j = 0; j < TQ; ++j )
			{
				if constexpr ( CPLX ) acc[i][j] = make_float2( 0.f, 0.f );
				else                  acc[i][j] = 0.f;
			}

		#pragma unroll
		for ( int s = 0; s < STAGES - 1; ++s )
		{
			if ( s < KT ) issue( s, s );
			cp_async_commit();
		}

		for ( int64_t kt = 0; kt < KT; ++kt )
		{
			cp_async_wait<STAGES - 2>();
			__syncthreads();
			{
				const int64_t kn = kt + STAGES - 1;
				if ( kn < KT ) issue( kn, (int)( kn % STAGES ) );
				cp_async_commit();
			}
			const T* xs = smem + (size_t)( kt % STAGES ) * ( Cfg::XS_ELEMS + Cfg::YS_ELEMS );
			const T* ys = xs + Cfg::XS_ELEMS;

			#pragma unroll
			for ( int k = 0; k < BK; ++k )
			{
				T x[TP], y[TQ];
				#pragma unroll
				for ( int gp = 0; gp < GP; ++gp )
				{
					const float4 v = *reinterpret_cast<const float4*>( xs + k * SXP + gp * ( BP / GP ) + ty * VE );
					if constexpr ( CPLX ) { x[gp * VE] = make_float2( v.x, v.y ); x[gp * VE + 1] = make_float2( v.z, v.w ); }
					else { x[gp * VE] = v.x; x[gp * VE + 1] = v.y; x[gp * VE + 2] = v.z; x[gp * VE + 3] = v.w; }
				}
				#pragma unroll
				for ( int gq = 0; gq < GQ; ++gq )
				{
					const float4 v = *reinterpret_cast<const float4*>( ys + k * SYQ + gq * ( BQ / GQ ) + tx * VE );
					if constexpr ( CPLX ) { y[gq * VE] = make_float2( v.x, v.y ); y[gq * VE + 1] = make_float2( v.z, v.w ); }
					else { y[gq * VE] = v.x; y[gq * VE + 1] = v.y; y[gq * VE + 2] = v.z; y[gq * VE + 3] = v.w; }
				}
				if constexpr ( !CPLX )
				{
					#pragma unroll
					for ( int i = 0; i < TP; ++i )
						#pragma unroll
						for ( int j = 0; j < TQ; ++j )
							acc[i][j] = fmaf( x[i], y[j], acc[i][j] );
				}
				else
				{
					#pragma unroll
					for ( int i = 0; i < TP; ++i )
					{
						const float xr = x[i].x, xi = flip_sign( x[i].y, cjx );
						#pragma unroll
						for ( int j = 0; j < TQ; ++j )
						{
							const float yr = y[j].x, yi = flip_sign( y[j].y, cjy );
							acc[i][j].x = fmaf( xr, yr, acc[i][j].x );
							acc[i][j].x = fmaf( -xi, yi, acc[i][j].x );
							acc[i][j].y = fmaf( xr, yi, acc[i][j].y );
							acc[i][j].y = fmaf( xi, yr, acc[i][j].y );
						}
					}
				}
			}
		}
		cp_async_wait<0>();

		// ---- epilogue
		int dlo, dhi;
		tri_band( g, p0, q0, dlo, dhi );
		#pragma unroll
		for ( int i = 0; i < TP; ++i )
		{
			const int pl = ( i / VE ) * ( BP / GP ) + ty * VE + ( i % VE );
			if ( pl >= p_lim ) continue;
			T* drow = g.D + ( p0 + pl ) * g.ldd + q0;
			#pragma unroll
			for ( int gq = 0; gq < GQ; ++gq )
			{
				const int ql = gq * ( BQ / GQ ) + tx * VE;
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
		__syncthreads();
	}
}

} // namespace b200
