// gemm_dmma.cuh -- FP64 tensor-core (DMMA) gemm for d and z: the kernel contract (GemmArgs), the tile raster, the triangular
// schedules and the cp.async tile loader shared by the kernels in gemm_dmma_ws.cuh / gemm_dmma_tma.cuh / gemm_dmma_pp.cuh /
// gemm_zmma_tma.cuh (the single-role kernel that first lived here was retired in round 2: 31.3 against 36.4 TFLOP/s).
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
	// Split-k tail (gemm_dmma_tma_kernel<..., SK>; mid-size problems whose last wave would leave SMs idle): work units
	// 0 .. sk_full-1 are whole tiles; every later tile is cut into sk_split equal k chunks, one unit each.  A chunk parks
	// its accumulators in sk_ws (slot tail*sk_split + chunk, 128x128 doubles in lane order); per consumer warp (a warp owns
	// the same part of the tile in every chunk) the LAST arrival at sk_flags[8*tail + warp] adds the sk_split slots IN
	// CHUNK ORDER (so the result does not depend on who was last), re-arms the counter and runs the ordinary epilogue.
	// sk_split == 0: off.
	int      sk_full, sk_split;
	T*       sk_ws;
	int*     sk_flags;
};

// Work unit -> (tile, k chunk or -1, k-tile range) under the split-k tail schedule above.
template <typename T>
__device__ __forceinline__ void sk_unit( const GemmArgs<T>& g, int unit, int64_t KT, int& tile, int& chunk, int64_t& kt0, int64_t& kt1 )
{
	if ( unit < g.sk_full ) { tile = unit; chunk = -1; kt0 = 0; kt1 = KT; return; }
	const int v = unit - g.sk_full;
	tile  = g.sk_full + v / g.sk_split;
	chunk = v - ( tile - g.sk_full ) * g.sk_split;
	kt0 = KT * chunk / g.sk_split;
	kt1 = KT * ( chunk + 1 ) / g.sk_split;
}

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

} // namespace b200
