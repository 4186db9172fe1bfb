// gemm_dmma.cuh -- FP64 tensor-core (DMMA) gemm for d and z.
//
// Replaces, for datatypes d/z, the reference's five-loop gemm:
//   jc/pc/ic loops            frame/3/gemm/bli_gemm_blk_var{2,3,1}.c
//   packm of A and B panels   frame/1m/packm/bli_packm_blk_var1.c
//   jr/ir macrokernel         frame/3/gemm/bli_gemm_ker_var2.c:216-278
//   gemm microkernel          ref_kernels/3/bli_gemm_ref.c:163-317
//
// B200 mapping
//   * MC x NC cache blocks / thread ranges  -> persistent CTAs walking a
//     rasterised list of BP x BQ output tiles (one CTA per SM).
//   * packm (HBM repack into MR/NR micropanels) -> none.  Tiles are staged
//     straight from the caller's layout into padded shared memory by an
//     asynchronous multi-stage cp.async (LDGSTS) pipeline; edge panels are
//     zero-filled by the copy itself (the reference's zero padding,
//     bli_packm_cxk_ref.c:80-145) and conjugation is folded into the fragment
//     load.
//   * MR x NR register microtile -> 8x8 DMMA tiles, (WTP/8) x (WTQ/8) per warp.
//   * alpha/beta epilogue keeps the reference's rule that beta == 0 never
//     reads C (ref_kernels/3/bli_gemm_ref.c:250-314).
//
// The kernel computes a "row-major" product
//     D(p,q) = alpha * sum_k X(p,k) * Y(k,q) + beta * D(p,q),   D[p*ldd + q],
// and the host maps column-major C onto it as D = C^T (X = op(B)^T,
// Y = op(A)^T) so that the two accumulators a lane owns are adjacent in
// memory.  X and Y may each be k-contiguous or p/q-contiguous.
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T>
struct GemmArgs
{
	const T* X; const T* Y; T* D;
	int64_t  P, Q, K;
	int64_t  ldx, ldy, ldd;
	T        alpha, beta;
	int      conjx, conjy;      // complex only
	int      tiles_p, tiles_q;
	int      beta_is_zero;
	int      d_vec_ok;          // D rows 16B aligned -> vector epilogue
	// k-panel accumulation (the pc loop of bli_gemm_blk_var3 inside one launch):
	// D = beta*D + alpha * sum_s X_s * Y_s, every panel K wide with the same strides.
	// Panel 0 is (X, Y); panels 1..nseg-1 are (Xseg[s-1], Yseg[s-1]).  Warp-specialised kernels only.
	int*     tile_counter;      // {next tile, finished CTAs}: dynamic tile scheduling (nullptr = static stride)
	int      nseg;
	const T* Xseg[7];
	const T* Yseg[7];
	// TMA kernels, nseg > 1: the panels of an operand sit at base + slot*stride (one 3-D tensor map per operand, third
	// coordinate = slot); segx[s] / segy[s] is the slot of k panel s (launch_dmma_tma fills them in).
	int      segx[8], segy[8];
	// Triangular D (the gemmt family: frame/3/gemmt/bli_gemmt_{l,u}_ker_var2.c computes only the stored
	// triangle of C).  tri == 0: full;  tri == 1: only q - p <= tri_off;  tri == 2: only q - p >= tri_off.
	// Tiles wholly outside are skipped, rows of tiles crossing the diagonal are clipped in the epilogue.
	int      tri;
	int64_t  tri_off;
	// Triangular OPERAND (trmm, trmm3: frame/3/trmm/bli_trmm_{ll,lu,rl,ru}_ker_var2.c skip the k range where the
	// packed triangular panel is identically zero).  The operand itself holds explicit zeros on its unstored side
	// (as bli_packm_struc_cxk.c:185-196,262-273 packs them), so this only trims the k loop of a tile:
	//   ktri == 1: X(p,k) == 0 for k > p   -> k < p0 + p_lim      ktri == 2: X(p,k) == 0 for k < p -> k >= p0
	//   ktri == 3: Y(k,q) == 0 for k > q   -> k < q0 + q_lim      ktri == 4: Y(k,q) == 0 for k < q -> k >= q0
	int      ktri;
	int      raster;            // tile rows per raster group (tile_coords); 8 unless tuned
};

// k-tile range [kt0, kt1) a tile has to visit (all of [0, KT) unless an operand is triangular).
template <typename T>
__device__ __forceinline__ void tile_k_range( const GemmArgs<T>& g, int64_t p0, int p_lim, int64_t q0, int q_lim,
                                              int BK, int64_t KT, int64_t& kt0, int64_t& kt1 )
{
	kt0 = 0; kt1 = KT;
	switch ( g.ktri )
	{
		case 1: kt1 = min( KT, ( p0 + p_lim + BK - 1 ) / BK ); break;
		case 2: kt0 = p0 / BK; break;
		case 3: kt1 = min( KT, ( q0 + q_lim + BK - 1 ) / BK ); break;
		case 4: kt0 = q0 / BK; break;
		default: break;
	}
}

template <typename T>
__device__ __forceinline__ bool tri_skip_tile( const GemmArgs<T>& g, int64_t p0, int64_t q0, int p_lim, int q_lim )
{
	if ( g.tri == 0 ) return false;
	if ( g.tri == 1 ) return ( q0 - ( p0 + p_lim - 1 ) ) > g.tri_off;     // smallest q - p of the tile
	return ( q0 + q_lim - 1 - p0 ) < g.tri_off;                           // largest q - p of the tile
}

template <typename T>
__device__ __forceinline__ bool tri_tile_interior( const GemmArgs<T>& g, int64_t p0, int64_t q0, int p_lim, int q_lim )
{
	if ( g.tri == 0 ) return true;
	if ( g.tri == 1 ) return ( q0 + q_lim - 1 - p0 ) <= g.tri_off;
	return ( q0 - ( p0 + p_lim - 1 ) ) >= g.tri_off;
}

// Element (pl, ql) of the tile at (p0, q0) belongs to the stored triangle iff dlo <= ql - pl <= dhi.
template <typename T>
__device__ __forceinline__ void tri_band( const GemmArgs<T>& g, int64_t p0, int64_t q0, int& dlo, int& dhi )
{
	constexpr int BIG = 1 << 30;
	const int64_t d = max( (int64_t)-BIG, min( (int64_t)BIG, p0 - q0 + g.tri_off ) );
	dlo = ( g.tri == 2 ) ? (int)d : -BIG;
	dhi = ( g.tri == 1 ) ? (int)d :  BIG;
}
__device__ __forceinline__ bool in_band( int d, int dlo, int dhi ) { return d >= dlo && d <= dhi; }

// Tile -> (tp,tq) with a grouped raster so that the ~148 concurrently running
// tiles form a compact block and share X/Y panels in L2.
__device__ __forceinline__ void tile_coords( int tile, int tiles_p, int tiles_q, int GROUP, int& tp, int& tq )
{
	const int per_group = GROUP * tiles_q;
	const int grp   = tile / per_group;
	const int first = grp * GROUP;
	const int gsize = min( GROUP, tiles_p - first );
	const int r     = tile - grp * per_group;
	tp = first + r % gsize;
	tq = r / gsize;
}

// Copy one LS x LC tile (LC contiguous in global memory, leading dimension ld)
// into shared memory rows of LC+PAD elements.  Rows >= s_lim and columns
// >= c_lim are zero-filled.
template <typename T, int LS, int LC, int PAD, int NT, bool AL>
__device__ __forceinline__ void load_tile( uint32_t sbase, const T* __restrict__ g, int64_t ld,
                                           int s_lim, int c_lim, int tid )
{
	using R = typename Elem<T>::real;
	constexpr int CPB   = AL ? 16 : ( sizeof(R) >= 8 ? 8 : 4 );
	constexpr int ROWB  = LC * (int)sizeof(T);
	constexpr int CPR   = ROWB / CPB;
	constexpr int TOTAL = LS * CPR;
	constexpr int ITERS = ( TOTAL + NT - 1 ) / NT;
	const int lim_bytes = c_lim * (int)sizeof(T);
	constexpr int UNROLL = ITERS <= 8 ? ITERS : 4;     // long (unaligned / producer-warp) copies: bound register use
	#pragma unroll UNROLL
	for ( int i = 0; i < ITERS; ++i )
	{
		const int id = tid + i * NT;
		if ( TOTAL % NT != 0 && id >= TOTAL ) break;
		const int s  = id / CPR;
		const int cb = ( id % CPR ) * CPB;
		int nbytes   = ( s < s_lim ) ? min( max( lim_bytes - cb, 0 ), CPB ) : 0;
		const char* src = ( nbytes > 0 )
		                ? reinterpret_cast<const char*>( g + (int64_t)s * ld ) + cb
		                : reinterpret_cast<const char*>( g );
		cp_async<CPB>( sbase + (uint32_t)( s * ( LC + PAD ) * (int)sizeof(T) + cb ), src, nbytes );
	}
}

template <typename T, int BP, int BQ, int BK, int WP, int WQ, int STAGES, bool XK, bool YK, bool AL>
struct DmmaCfg
{
	static constexpr bool CPLX = Elem<T>::cplx;
	static constexpr int  NT   = WP * WQ * 32;
	static constexpr int  WTP  = BP / WP;           // warp tile
	static constexpr int  WTQ  = BQ / WQ;
	static constexpr int  MT   = WTP / 8;           // 8x8 DMMA tiles per warp
	static constexpr int  NTL  = WTQ / 8;
	// Row paddings that make every fragment load bank-conflict free:
	//  real:    row stride == 4 (mod 16) doubles for both orientations
	//  complex: k-contiguous rows == 4 (mod 8), p/q-contiguous rows == 2 (mod 8)
	static constexpr int  PADK = 4;
	static constexpr int  PADC = CPLX ? 2 : 4;
	static constexpr int  XS_ELEMS = XK ? BP * ( BK + PADK ) : BK * ( BP + PADC );
	static constexpr int  YS_ELEMS = YK ? BQ * ( BK + PADK ) : BK * ( BQ + PADC );
	static constexpr int  STAGE_BYTES = ( XS_ELEMS + YS_ELEMS ) * (int)sizeof(T);
	static constexpr int  SMEM_BYTES  = STAGE_BYTES * STAGES;
	static_assert( BK % 8 == 0 && BP % ( 8 * WP ) == 0 && BQ % ( 8 * WQ ) == 0, "tile shape" );
};

template <typename T, int BP, int BQ, int BK, int WP, int WQ, int STAGES, bool XK, bool YK, bool AL>
__global__ void __launch_bounds__( WP * WQ * 32, 1 )
gemm_dmma_kernel( const GemmArgs<T> g )
{
	using Cfg = DmmaCfg<T, BP, BQ, BK, WP, WQ, STAGES, XK, YK, AL>;
	constexpr bool CPLX = Cfg::CPLX;
	constexpr int  NT = Cfg::NT, MT = Cfg::MT, NTL = Cfg::NTL;
	constexpr int  SXK = BK + Cfg::PADK, SXP = BP + Cfg::PADC, SYQ = BQ + Cfg::PADC;

	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* const smem = reinterpret_cast<T*>( smem_raw );

	const int tid  = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	const int gq   = lane >> 2;      // "g": row of the A fragment / column of B
	const int t4   = lane & 3;       // "t": k index inside a k4 step
	const int wp0  = ( warp / WQ ) * Cfg::WTP;
	const int wq0  = ( warp % WQ ) * Cfg::WTQ;

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
			if constexpr ( XK ) load_tile<T, BP, BK, Cfg::PADK, NT, AL>( smem_u32( xs ), gx + kt * BK, g.ldx, p_lim, k_lim, tid );
			else                load_tile<T, BK, BP, Cfg::PADC, NT, AL>( smem_u32( xs ), gx + kt * BK * g.ldx, g.ldx, k_lim, p_lim, tid );
			if constexpr ( YK ) load_tile<T, BQ, BK, Cfg::PADK, NT, AL>( smem_u32( ys ), gy + kt * BK, g.ldy, q_lim, k_lim, tid );
			else                load_tile<T, BK, BQ, Cfg::PADC, NT, AL>( smem_u32( ys ), gy + kt * BK * g.ldy, g.ldy, k_lim, q_lim, tid );
		};

		// accumulators: real -> acc[MT][NTL][2]; complex -> re and im planes
		double acc[CPLX ? 2 : 1][MT][NTL][2];
		#pragma unroll
		for ( int c = 0; c < ( CPLX ? 2 : 1 ); ++c )
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
				#pragma unroll
				for ( int j = 0; j < NTL; ++j ) { acc[c][i][j][0] = 0.0; acc[c][i][j][1] = 0.0; }

		// ---- pipeline prologue
		#pragma unroll
		for ( int s = 0; s < STAGES - 1; ++s )
		{
			if ( s < KT ) issue( s, s );
			cp_async_commit();
		}

		// ---- main loop over k tiles
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
			for ( int kk = 0; kk < BK / 4; ++kk )
			{
				T xf[MT], yf[NTL];
				#pragma unroll
				for ( int i = 0; i < MT; ++i )
				{
					const int p = wp0 + i * 8 + gq, k = kk * 4 + t4;
					xf[i] = XK ? xs[p * SXK + k] : xs[k * SXP + p];
				}
				#pragma unroll
				for ( int j = 0; j < NTL; ++j )
				{
					const int q = wq0 + j * 8 + gq, k = kk * 4 + t4;
					yf[j] = YK ? ys[q * SXK + k] : ys[k * SYQ + q];
				}
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
					// (xr + i xi)(yr + i yi): re += xr*yr - xi*yi ; im += xr*yi + xi*yr
					double xr[MT], xi[MT], nxi[MT];
					#pragma unroll
					for ( int i = 0; i < MT; ++i )
					{
						xr[i]  = xf[i].x;
						xi[i]  = flip_sign( xf[i].y, cjx );
						nxi[i] = -xi[i];
					}
					#pragma unroll
					for ( int j = 0; j < NTL; ++j )
					{
						const double yr = yf[j].x, yi = flip_sign( yf[j].y, cjy );
						#pragma unroll
						for ( int i = 0; i < MT; ++i )
						{
							dmma884( acc[0][i][j][0], acc[0][i][j][1], xr[i],  yr );
							dmma884( acc[1][i][j][0], acc[1][i][j][1], xr[i],  yi );
							dmma884( acc[0][i][j][0], acc[0][i][j][1], nxi[i], yi );
							dmma884( acc[1][i][j][0], acc[1][i][j][1], xi[i],  yr );
						}
					}
				}
			}
		}
		cp_async_wait<0>();

		// ---- epilogue: D = alpha*acc + beta*D   (beta == 0: D is not read)
		int dlo, dhi;
		tri_band( g, p0, q0, dlo, dhi );
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
				const bool one = ( ql < q_lim && in_band( ql - pl, dlo, dhi ) );
				const bool two = ( ql + 1 < q_lim && in_band( ql + 1 - pl, dlo, dhi ) );
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
						// ab *= alpha (bli_tscals), then c := ab + beta*c (bli_txpbys)
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
		__syncthreads();   // all fragment reads done before the next tile's prologue overwrites smem
	}
}

} // namespace b200
