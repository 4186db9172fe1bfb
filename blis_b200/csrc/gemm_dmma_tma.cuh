// gemm_dmma_tma.cuh -- dgemm with TMA tensor-map staging (aligned operands).
//
// Same contract, same consumer math and same tile scheduler as gemm_dmma_ws.cuh;
// what changes is the PRODUCER: instead of 128 threads issuing cp.async, ONE
// thread issues cp.async.bulk.tensor (TMA, UTMALDG in SASS) per operand and
// k-slab, straight from the caller's matrix described by a CUtensorMap, into
// 128-byte-swizzled shared memory; completion is tracked by the stage's mbarrier
// transaction count.  This is the north-star form of "packm replaced by TMA
// tensor-map staging": the reference's packm (frame/1m/packm/bli_packm_blk_var1.c)
// copies an MC x KC block into MR-row micropanels with zero padding at the edges;
// here the tensor map describes the unpacked matrix and the TMA unit does the
// copy, the bounds check and the zero fill (out-of-bounds box elements read as 0).
//
// Shared-memory layout (per stage, 2 x 16 KiB, no padding):
//   k-contiguous operand  : ONE box {16 k, 128 rows}: row r = 128 bytes = 16 doubles of k
//   p/q-contiguous operand: EIGHT boxes {16 rows, 16 k}: box b holds rows 16b..16b+15,
//                           inside a box row k = 128 bytes = 16 consecutive p (or q)
// both with CU_TENSOR_MAP_SWIZZLE_128B: 16-byte chunk c of line L is stored at chunk c ^ (L % 8).
// An 8-byte fragment load of DMMA.8x8x4 touches, per half-warp, 4 lines x 4 k values.  To make
// that bank-conflict free under the 128B swizzle the k index a lane uses in k4-step s is PERMUTED:
//       k(s, t) = (t & 1) | ((t >> 1) << 3) | (((t + s) & 3) << 1)        t = lane % 4
// (A and B fragments use the same k, so the product is unchanged; over s = 0..3 every k in 0..15
// is used once).  For the k-contiguous layout the 16 lanes of a half-warp then hit 16 distinct
// 8-byte bank pairs because (k>>1)^g covers both 4-chunk cosets x both halves; for the other layout
// because the lines k(s,t) % 8 have pairwise distinct (k % 8) >> 1.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_ws.cuh"

namespace b200 {

__device__ __forceinline__ void mbar_arrive_expect_tx( uint32_t bar, uint32_t bytes )
{
	asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory" );
}
__device__ __forceinline__ void tma_load_2d( uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar )
{
	asm volatile( "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
	              :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory" );
}

__device__ __forceinline__ void tma_load_3d( uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar )
{
	asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
	              :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory" );
}

// Producer side: ask L2 for the BP x BQ tile of D at (p0, q0) now (one bulk prefetch per row); the consumers' epilogue
// reads it after the k loop.  Needs 16-byte aligned rows (d_vec_ok); nothing is requested when beta == 0.
template <typename T>
__device__ __forceinline__ void prefetch_d_tile_l2( const GemmArgs<T>& g, int p0, int q0, int BP, int BQ )
{
	if ( g.beta_is_zero || !g.d_vec_ok ) return;
	const int rows = (int)min( (int64_t)BP, g.P - p0 );
	const uint32_t bytes = (uint32_t)( min( (int64_t)BQ, g.Q - q0 ) * (int64_t)sizeof(T) ) & ~15u;
	const T* dt = g.D + (int64_t)p0 * g.ldd + q0;
	if ( bytes == 0 ) return;
	for ( int r = 0; r < rows; ++r )
		asm volatile( "cp.async.bulk.prefetch.L2.global [%0], %1;\n" :: "l"(dt + (int64_t)r * g.ldd), "r"(bytes) : "memory" );
}

template <int ST>
struct DmmaTmaCfgT
{
	static constexpr int BP = 128, BQ = 128, BK = 16, WP = 4, WQ = 2, STAGES = ST;
	static constexpr int WTP = BP / WP, WTQ = BQ / WQ, MT = WTP / 8, NTL = WTQ / 8;
	static constexpr int OPER_BYTES  = 128 * 128;                 // 16 KiB per operand per stage
	static constexpr int STAGE_BYTES = 2 * OPER_BYTES;
	static constexpr int NCONS = WP * WQ * 32, NPROD = 128, NT_ALL = NCONS + NPROD;
	static constexpr int BAR_BYTES  = 3 * STAGES * 8 + 4 * 8 + 16;               // full, empty, staged (CST) per stage + 4 scheduler barriers + tile slots
	static constexpr int SMEM_BYTES = STAGE_BYTES * STAGES + BAR_BYTES + 1024;   // + slack for 1 KiB alignment
};
using DmmaTmaCfg    = DmmaTmaCfgT<6>;
using DmmaTmaCfgCst = DmmaTmaCfgT<7>;      // CST holds the D stages until their TMA stores have drained: one more stage (224 KiB)

__device__ __forceinline__ void tma_store_2d( const CUtensorMap* map, int c0, int c1, uint32_t src )
{
	asm volatile( "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n"
	              :: "l"(map), "r"(src), "r"(c0), "r"(c1) : "memory" );
}
__device__ __forceinline__ void mbar_arrive_n( uint32_t bar, uint32_t n )
{
	asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(n) : "memory" );
}

// CST ("C staged"): for small k the epilogue's read-modify-write of D dominates (k = 64: 8.4 us of DMMA per tile against
// ~4 us of exposed global-load latency, ncu: 37 % of the stall samples behind the last DMMA).  With CST the producer also
// TMA-loads the D tile, as FOUR extra ring stages of 32 rows x 128 columns (eight 128B-swizzled {16, 32} boxes = one
// 32 KiB stage) issued right behind the tile's k stages, and the consumers take D from shared memory: warp row-group r
// (rows 32r..32r+31) reads exactly extra stage r.  The loads travel through the same full/empty ring, so they are in
// flight while the k loop runs and no register is spent on latency hiding.
// The WRITE half goes the same way back: a consumer overwrites the D values it read with beta*D + alpha*acc in place
// (STS.128) and moves on to the next tile's k loop at once; a store thread (second producer warp) waits until the two
// warps of a row group have staged their stage, hands it to the TMA unit (cp.async.bulk.tensor global <- shared, eight
// {8, 32} boxes), and releases the ring slot when the unit has read it.  The tensor pipe therefore never waits for the
// 128 KiB store burst of a tile to drain to HBM (ncu before: k = 64, beta = 0 -- no D read at all -- still only 81 % DMMA
// busy).  The D stages are NOT swizzled: sixteen {8 columns, 32 rows} boxes with 64-byte rows, so that the eight lanes of
// an LDS.128 / STS.128 quarter-warp (two rows x four chunks) cover 128 contiguous bytes (the 128B-swizzled {16, 32} boxes
// of the first version cost 28 % bank conflicts, ncu r01).
// beta == 0: the producer passes the four stages on empty (nothing is loaded).  Used when D rows are 16-byte aligned and
// K <= 1024 (gemm_d.cu); seven ring stages; the default instantiation is untouched.
// SK ("split k"): the tail schedule of GemmArgs::sk_* -- tiles of the last, partial wave are cut into k chunks so that
// every SM has work until the end (2048^3: 256 tiles on 148 SMs = 2 rounds for 1.73 rounds of work).  The reference never
// splits k inside one gemm (bli_gemm_blk_var3.c:110-112 runs the pc loop sequentially); here the chunks of a tile are added
// in chunk order by whichever CTA finishes last, so the result is reproducible, but it is rounded differently from the
// unsplit product (within the same error bound).  Plain products only (no TRI, no CST, one k panel).
template <bool XK, bool YK, bool TRI = false, bool CST = false, bool SK = false>
__global__ void __launch_bounds__( 384, 1 )
gemm_dmma_tma_kernel( const GemmArgs<double> g, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                      const __grid_constant__ CUtensorMap tmd )
{
	using Cfg = typename std::conditional<CST, DmmaTmaCfgCst, DmmaTmaCfg>::type;
	constexpr int BP = Cfg::BP, BQ = Cfg::BQ, BK = Cfg::BK, WQ = Cfg::WQ, STAGES = Cfg::STAGES;
	constexpr int MT = Cfg::MT, NTL = Cfg::NTL, KS = BK / 4;
	static_assert( !( SK && ( TRI || CST ) ), "split k: plain products only" );

	extern __shared__ unsigned char smem_unaligned[];
	const uint32_t raw = smem_u32( smem_unaligned );
	const uint32_t sbase = ( raw + 1023u ) & ~1023u;                         // swizzle atoms need 1 KiB alignment
	unsigned char* const smem = smem_unaligned + ( sbase - raw );
	const uint32_t bar_base = sbase + (uint32_t)Cfg::STAGE_BYTES * STAGES;
	auto full_bar    = [&]( int s ) { return bar_base + (uint32_t)s * 8u; };
	auto empty_bar   = [&]( int s ) { return bar_base + (uint32_t)( STAGES + s ) * 8u; };
	auto staged_bar  = [&]( int s ) { return bar_base + (uint32_t)( 2 * STAGES + s ) * 8u; };      // CST: a row group has written its results
	auto sched_full  = [&]( int s ) { return bar_base + (uint32_t)( 3 * STAGES + s ) * 8u; };
	auto sched_empty = [&]( int s ) { return bar_base + (uint32_t)( 3 * STAGES + 2 + s ) * 8u; };
	volatile int* const sched_tile = reinterpret_cast<volatile int*>( smem + (size_t)Cfg::STAGE_BYTES * STAGES + ( 3 * STAGES + 4 ) * 8 );

	const int tid = threadIdx.x;
	if ( tid == 0 )
	{
		#pragma unroll
		for ( int s = 0; s < STAGES; ++s )
		{
			mbar_init( full_bar( s ),  1 );                    // the producer's expect_tx arrival; bytes complete it
			mbar_init( empty_bar( s ), Cfg::NCONS / 32 );
			mbar_init( staged_bar( s ), WQ );                  // the two warps of a row group
		}
		#pragma unroll
		for ( int s = 0; s < 2; ++s )
		{
			mbar_init( sched_full( s ),  1 );
			mbar_init( sched_empty( s ), Cfg::NCONS / 32 + ( CST ? 1 : 0 ) );    // CST: the store thread follows the tiles too
		}
		asm volatile( "fence.mbarrier_init.release.cluster;\n" ::: "memory" );   // visible to the async proxy
	}
	__syncthreads();

	// k-panel accumulation (b200_gemm_kpanels: the pc loop of bli_gemm_blk_var3 inside one launch): nseg panels of K each,
	// described by 3-D tensor maps whose third coordinate selects the panel; the consumers see one long k loop.
	const int64_t KT_SEG = ( g.K + BK - 1 ) / BK;
	const int64_t KT = KT_SEG * g.nseg;
	const int num_tiles = SK ? g.sk_full + ( g.tiles_p * g.tiles_q - g.sk_full ) * g.sk_split : g.tiles_p * g.tiles_q;   // work units

	if ( tid >= Cfg::NCONS )
	{
		// ============ PRODUCER warpgroup: one thread drives the TMA unit ============
		setmaxnreg_dec<SK ? 56 : 40>();               // 256 x 224 + 128 x 56 = 64512 registers
		if constexpr ( CST )
		{
			if ( tid == Cfg::NCONS + 32 )
			{
				// ============ STORE thread: D stages -> global through the TMA unit, then the ring slot is free ============
				asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmd) : "memory" );
				int stage = 0; uint32_t phase = 0;
				uint32_t staged_parity = 0;                    // bit s: parity the next completion of staged_bar( s ) will have (a stage is a D stage only now and then)
				for ( int it = 0; ; ++it )
				{
					const int slot = it & 1;
					mbar_wait( sched_full( slot ), ( it >> 1 ) & 1 );
					const int tile = sched_tile[slot];
					mbar_arrive( sched_empty( slot ) );
					if ( tile >= num_tiles ) break;
					int tp, tq;
					tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
					const int p0 = tp * BP, q0 = tq * BQ;
					const bool fast = ( g.d_vec_ok && min( (int64_t)BQ, g.Q - q0 ) == BQ );
					for ( int64_t kt = 0; kt < KT; ++kt ) { if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; } }   // the k stages are not mine
					int st0 = stage;
					#pragma unroll 1
					for ( int qd = 0; qd < BP / 32; ++qd )
					{
						// The row group's two warps arrive here for EVERY tile, after the stage's load has landed (they waited for it) and,
						// on the fast path, after they have written their results.  This thread never waits on a full barrier itself: it
						// skips the k stages without looking at them, so it may be several ring wraps away from what a parity can tell apart.
						mbar_wait( staged_bar( stage ), ( staged_parity >> stage ) & 1u );
						staged_parity ^= 1u << stage;
						if ( fast )
						{
							const uint32_t cs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES;
							#pragma unroll
							for ( int b = 0; b < BQ / 8; ++b ) tma_store_2d( &tmd, q0 + b * 8, p0 + qd * 32, cs + b * 2048 );
							asm volatile( "cp.async.bulk.commit_group;\n" ::: "memory" );
						}
						if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
					}
					// release the four slots in order, each as soon as the unit has READ it
					#pragma unroll 1
					for ( int qd = 0; qd < BP / 32; ++qd )
					{
						if ( fast )
						{
							if      ( qd == 0 ) asm volatile( "cp.async.bulk.wait_group.read 3;\n" ::: "memory" );
							else if ( qd == 1 ) asm volatile( "cp.async.bulk.wait_group.read 2;\n" ::: "memory" );
							else if ( qd == 2 ) asm volatile( "cp.async.bulk.wait_group.read 1;\n" ::: "memory" );
							else                asm volatile( "cp.async.bulk.wait_group.read 0;\n" ::: "memory" );
						}
						mbar_arrive_n( empty_bar( st0 ), Cfg::NCONS / 32 );
						if ( ++st0 == STAGES ) st0 = 0;
					}
				}
				asm volatile( "cp.async.bulk.wait_group 0;\n" ::: "memory" );     // all stores complete before the CTA exits
				return;
			}
		}
		if ( tid != Cfg::NCONS ) return;
		asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmx) : "memory" );
		asm volatile( "prefetch.tensormap [%0];\n" :: "l"(&tmy) : "memory" );
		int stage = 0; uint32_t phase = 0;
		for ( int it = 0; ; ++it )
		{
			const int slot = it & 1;
			mbar_wait( sched_empty( slot ), ( ( it >> 1 ) & 1 ) ^ 1u );
			int tile = g.tile_counter ? atomicAdd( g.tile_counter, 1 ) : (int)( blockIdx.x + (unsigned)it * gridDim.x );
			sched_tile[slot] = tile;
			mbar_arrive( sched_full( slot ) );
			if ( tile >= num_tiles ) break;
			int64_t kt0 = 0, kt1 = KT;
			if constexpr ( SK ) { int chunk; sk_unit( g, tile, KT, tile, chunk, kt0, kt1 ); }
			int tp, tq;
			tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
			const int p0 = tp * BP, q0 = tq * BQ;
			if ( TRI && tri_skip_tile( g, p0, q0, (int)min( (int64_t)BP, g.P - p0 ), (int)min( (int64_t)BQ, g.Q - q0 ) ) ) continue;
			if constexpr ( TRI ) tile_k_range( g, p0, (int)min( (int64_t)BP, g.P - p0 ), q0, (int)min( (int64_t)BQ, g.Q - q0 ), BK, KT, kt0, kt1 );
			if constexpr ( !CST ) prefetch_d_tile_l2( g, p0, q0, BP, BQ );
			for ( int64_t kt = kt0; kt < kt1; ++kt )
			{
				mbar_wait( empty_bar( stage ), phase ^ 1u );
				const uint32_t xs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES, ys = xs + Cfg::OPER_BYTES;
				const uint32_t fb = full_bar( stage );
				mbar_arrive_expect_tx( fb, 2u * Cfg::OPER_BYTES );
				if ( g.nseg == 1 )
				{
					const int k0 = (int)( kt * BK );
					if constexpr ( XK ) tma_load_2d( xs, &tmx, k0, p0, fb );
					else
					{
						#pragma unroll
						for ( int b = 0; b < BP / 16; ++b ) tma_load_2d( xs + b * 2048, &tmx, p0 + b * 16, k0, fb );
					}
					if constexpr ( YK ) tma_load_2d( ys, &tmy, k0, q0, fb );
					else
					{
						#pragma unroll
						for ( int b = 0; b < BQ / 16; ++b ) tma_load_2d( ys + b * 2048, &tmy, q0 + b * 16, k0, fb );
					}
				}
				else
				{
					const int seg = (int)( kt / KT_SEG );
					const int k0 = (int)( ( kt - seg * KT_SEG ) * BK );
					const int sx = g.segx[seg], sy = g.segy[seg];
					if constexpr ( XK ) tma_load_3d( xs, &tmx, k0, p0, sx, fb );
					else
					{
						#pragma unroll
						for ( int b = 0; b < BP / 16; ++b ) tma_load_3d( xs + b * 2048, &tmx, p0 + b * 16, k0, sx, fb );
					}
					if constexpr ( YK ) tma_load_3d( ys, &tmy, k0, q0, sy, fb );
					else
					{
						#pragma unroll
						for ( int b = 0; b < BQ / 16; ++b ) tma_load_3d( ys + b * 2048, &tmy, q0 + b * 16, k0, sy, fb );
					}
				}
				if ( ++stage == STAGES ) { stage = 0; phase ^= 1u; }
			}
			if constexpr ( CST )
			{
				#pragma unroll 1
				for ( int qd = 0; qd < BP / 32; ++qd )
				{
					mbar_wait( empty_bar( stage ), phase ^ 1u );
					const uint32_t cs = sbase + (uint32_t)stage * Cfg::STAGE_BYTES;
					const uint32_t fb = full_bar( stage );
					if ( g.beta_is_zero ) mbar_arrive( fb );                       // D is not read: the stage is handed over empty
					else
					{
						mbar_arrive_expect_tx( fb, (uint32_t)Cfg::STAGE_BYTES );
						#pragma unroll
						for ( int b = 0; b < BQ / 8; ++b ) tma_load_2d( cs + b * 2048, &tmd, q0 + b * 8, p0 + qd * 32, fb );
					}
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
	const int gq = lane >> 2, t4 = lane & 3;
	const int wp0 = ( warp / WQ ) * Cfg::WTP;
	const int wq0 = ( warp % WQ ) * Cfg::WTQ;

	// byte offset of the fragment element of 8x8 tile `i` (rows w0 + 8i + g) in k4-step s
	auto frag_off = [&]( bool kmajor, int w0, int i, int s ) -> int
	{
		const int ts = ( t4 + s ) & 3;
		if ( kmajor )
		{
			const int chunk = ( ( t4 >> 1 ) << 2 ) | ts;                       // (k >> 1)
			return ( w0 + i * 8 + gq ) * 128 + ( ( chunk ^ gq ) << 4 ) + ( t4 & 1 ) * 8;
		}
		const int k  = ( t4 & 1 ) | ( ( t4 >> 1 ) << 3 ) | ( ts << 1 );
		const int k7 = ( t4 & 1 ) | ( ts << 1 );
		const int chunk = ( ( i & 1 ) << 2 ) | ( gq >> 1 );                    // ((row % 16) >> 1)
		return ( ( w0 >> 4 ) + ( i >> 1 ) ) * 2048 + k * 128 + ( ( chunk ^ k7 ) << 4 ) + ( gq & 1 ) * 8;
	};

	int stage = 0; uint32_t phase = 0;
	auto load_frags = [&]( double ( &xf )[MT], double ( &yf )[NTL], int st, int s )
	{
		const unsigned char* xs = smem + (size_t)st * Cfg::STAGE_BYTES;
		const unsigned char* ys = xs + Cfg::OPER_BYTES;
		#pragma unroll
		for ( int i = 0; i < MT; ++i ) xf[i] = *reinterpret_cast<const double*>( xs + frag_off( XK, wp0, i, s ) );
		#pragma unroll
		for ( int j = 0; j < NTL; ++j ) yf[j] = *reinterpret_cast<const double*>( ys + frag_off( YK, wq0, j, s ) );
	};

	for ( int it = 0; ; ++it )
	{
		const int slot = it & 1;
		mbar_wait( sched_full( slot ), ( it >> 1 ) & 1 );
		int tile = sched_tile[slot];
		__syncwarp();
		if ( lane == 0 ) mbar_arrive( sched_empty( slot ) );
		if ( tile >= num_tiles ) break;
		int64_t kt0 = 0, kt1 = KT;
		int chunk = -1;
		if constexpr ( SK ) sk_unit( g, tile, KT, tile, chunk, kt0, kt1 );
		int tp, tq;
		tile_coords( tile, g.tiles_p, g.tiles_q, g.raster, tp, tq );
		const int64_t p0 = (int64_t)tp * BP, q0 = (int64_t)tq * BQ;
		const int p_lim = (int)min( (int64_t)BP, g.P - p0 );
		const int q_lim = (int)min( (int64_t)BQ, g.Q - q0 );
		if ( TRI && tri_skip_tile( g, p0, q0, p_lim, q_lim ) ) continue;

		double acc[MT][NTL][2];
		#pragma unroll
		for ( int i = 0; i < MT; ++i )
			#pragma unroll
			for ( int j = 0; j < NTL; ++j ) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

		double xa[MT], ya[NTL], xb[MT], yb[NTL];
		mbar_wait( full_bar( stage ), phase );
		load_frags( xa, ya, stage, 0 );

		auto mma_step = [&]( double ( &xf )[MT], double ( &yf )[NTL] )
		{
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
				#pragma unroll
				for ( int j = 0; j < NTL; ++j )
					dmma884( acc[i][j][0], acc[i][j][1], xf[i], yf[j] );
		};

		if constexpr ( TRI ) tile_k_range( g, p0, p_lim, q0, q_lim, BK, KT, kt0, kt1 );
		for ( int64_t kt = kt0; kt < kt1; ++kt )
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

		if constexpr ( SK )
		{
			if ( chunk >= 0 )
			{
				// A k chunk of a tail tile.  Every warp owns the same 32 x 64 part of the tile in every chunk, so the fix-up is
				// a per-WARP affair (no CTA barrier): park the accumulators (coalesced 16-byte stores, L2), count in at the
				// warp's own counter, and go on to the next unit -- unless this warp is the last of the tile's sk_split chunks
				// to arrive: then it adds all slots IN CHUNK ORDER (so the rounding never depends on who was last), re-arms
				// the counter and runs the ordinary epilogue.  Nobody ever waits for another CTA.
				const int tail = tile - g.sk_full;
				constexpr int SLOT = MT * NTL * Cfg::NCONS;                      // double2 per accumulator slot (128 KiB)
				double2* const slot0 = reinterpret_cast<double2*>( g.sk_ws ) + (int64_t)tail * g.sk_split * SLOT + tid;
				double2* const mine  = slot0 + (int64_t)chunk * SLOT;
				int* const flag = g.sk_flags + tail * ( Cfg::NCONS / 32 ) + warp;
				#pragma unroll
				for ( int i = 0; i < MT; ++i )
					#pragma unroll
					for ( int j = 0; j < NTL; ++j ) __stcg( mine + ( i * NTL + j ) * Cfg::NCONS, make_double2( acc[i][j][0], acc[i][j][1] ) );
				__syncwarp();
				int seen = 0;
				if ( lane == 0 ) { __threadfence(); seen = atomicAdd( flag, 1 ); }
				seen = __shfl_sync( 0xffffffffu, seen, 0 );
				if ( seen != g.sk_split - 1 ) continue;
				if ( lane == 0 ) { __threadfence(); *flag = 0; }
				__syncwarp();
				#pragma unroll 1
				for ( int c = 0; c < g.sk_split; ++c )
				{
					const double2* part = slot0 + (int64_t)c * SLOT;
					#pragma unroll
					for ( int i = 0; i < MT; ++i )
						#pragma unroll
						for ( int j = 0; j < NTL; ++j )
						{
							const double2 w = __ldcg( part + ( i * NTL + j ) * Cfg::NCONS );
							if ( c == 0 ) { acc[i][j][0] = w.x; acc[i][j][1] = w.y; }
							else          { acc[i][j][0] += w.x; acc[i][j][1] += w.y; }
						}
				}
			}
		}

		// ---- epilogue: D = alpha*acc + beta*D   (beta == 0: D is not read)
		const bool interior = ( !TRI || tri_tile_interior( g, p0, q0, p_lim, q_lim ) );
		int dlo = 0, dhi = 0;
		if constexpr ( TRI ) tri_band( g, p0, q0, dlo, dhi );
		auto keep = [&]( int d ) { if constexpr ( TRI ) return in_band( d, dlo, dhi ); else return true; };
		if constexpr ( CST )
		{
			// the four D stages of this tile: row-group r = warp / WQ owns stage r (rows 32r..32r+31); nobody else touches it.
			// The store thread releases all four, so a consumer only steps over them.
			const bool fast = ( g.d_vec_ok && q_lim == BQ );
			int st = stage + warp / WQ; uint32_t ph = phase;
			if ( st >= STAGES ) { st -= STAGES; ph ^= 1u; }
			mbar_wait( full_bar( st ), ph );                        // also on edge tiles: the store thread releases the slot on my word
			if ( fast )
			{
				// stage layout: sixteen un-swizzled {8 columns, 32 rows} boxes of 2 KiB (one per 8x8 tile column): a row is 64 bytes,
				// so the eight lanes of an LDS.128 / STS.128 quarter-warp (two rows x four 16-byte chunks) cover 128 contiguous bytes
				unsigned char* cs = smem + (size_t)st * Cfg::STAGE_BYTES + ( wq0 >> 3 ) * 2048 + gq * 64 + t4 * 16;
				#pragma unroll
				for ( int i = 0; i < MT; ++i )
				{
					double2 o[NTL];
					if ( !g.beta_is_zero )
					{
						#pragma unroll
						for ( int j = 0; j < NTL; ++j ) o[j] = *reinterpret_cast<const double2*>( cs + j * 2048 + i * 512 );
					}
					#pragma unroll
					for ( int j = 0; j < NTL; ++j )
					{
						double r0 = g.alpha * acc[i][j][0], r1 = g.alpha * acc[i][j][1];
						if ( !g.beta_is_zero ) { r0 = fma( g.beta, o[j].x, r0 ); r1 = fma( g.beta, o[j].y, r1 ); }
						*reinterpret_cast<double2*>( cs + j * 2048 + i * 512 ) = make_double2( r0, r1 );
					}
				}
				asm volatile( "fence.proxy.async.shared::cta;\n" ::: "memory" );    // my STS before the TMA unit's reads
			}
			__syncwarp();
			if ( lane == 0 ) mbar_arrive( staged_bar( st ) );
			stage += BP / 32;
			if ( stage >= STAGES ) { stage -= STAGES; phase ^= 1u; }
			if ( fast ) continue;
		}
		if ( g.d_vec_ok && q_lim == BQ && interior )
		{
			// Interior tile: all loads of a row are issued before its first store (NTL 16-byte loads per lane in flight;
			// the producer has already asked L2 for the tile).  Keeping two rows in flight would double the read rate of
			// small-k problems but needs 64 more registers than the 232 a consumer can have (ptxas spills the loaded
			// values; measured slower), see DESIGN.md section 10.
			#pragma unroll
			for ( int i = 0; i < MT; ++i )
			{
				if ( wp0 + i * 8 + gq >= p_lim ) continue;
				double2* __restrict__ dp = reinterpret_cast<double2*>( g.D + ( p0 + wp0 + i * 8 + gq ) * g.ldd + q0 + wq0 + 2 * t4 );
				double2 o[NTL];
				if ( !g.beta_is_zero )
				{
					#pragma unroll
					for ( int j = 0; j < NTL; ++j ) o[j] = __ldcs( dp + j * 4 );
				}
				#pragma unroll
				for ( int j = 0; j < NTL; ++j )
				{
					double r0 = g.alpha * acc[i][j][0], r1 = g.alpha * acc[i][j][1];
					if ( !g.beta_is_zero ) { r0 = fma( g.beta, o[j].x, r0 ); r1 = fma( g.beta, o[j].y, r1 ); }
					__stcs( dp + j * 4, make_double2( r0, r1 ) );
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
				const int ql = wq0 + j * 8 + 2 * t4;
				double r0 = g.alpha * acc[i][j][0], r1 = g.alpha * acc[i][j][1];
				if ( ql < q_lim && keep( ql - pl ) )
				{
					if ( !g.beta_is_zero ) r0 = fma( g.beta, drow[ql], r0 );
					drow[ql] = r0;
				}
				if ( ql + 1 < q_lim && keep( ql + 1 - pl ) )
				{
					if ( !g.beta_is_zero ) r1 = fma( g.beta, drow[ql + 1], r1 );
					drow[ql + 1] = r1;
				}
			}
		}
	}
}

} // namespace b200
