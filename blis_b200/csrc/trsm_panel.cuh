// trsm_panel.cuh -- fused diagonal-PANEL triangular solve for d (the "gemmtrsm" path on the FP64 tensor pipe).
//
// Replaces, for one PB x PB diagonal panel of A (PB = 256) and all n right-hand sides, in ONE launch:
//   bli_trsm_blk_var1 over the panel        frame/3/trsm/bli_trsm_blk_var1.c:40-188
//   bli_trsm_ll_ker_var2 / _lu_             frame/3/trsm/bli_trsm_l{l,u}_ker_var2.c:38-335
//   gemmtrsm microkernel                    ref_kernels/3/bli_gemmtrsm_ref.c:43-196  ( b11 := alpha*b11 - a1x*bx1, then trsm )
//   trsm microkernel                        ref_kernels/3/bli_trsm_ref.c:44-128,140-224
//   triangular packm of the panel           frame/1m/packm/bli_packm_struc_cxk.c:155-301, bli_packm_cxc_diag_ref.c:161-236
//
// Round 1 solved 64-row diagonal blocks one launch at a time with one thread per right-hand side and ran every update,
// down to k = 64, as a separate gemm launch (~1000 launches for m = 32768, 13 ms of latency-bound work).  Here one CTA
// owns CN = 64 columns of B for the whole panel and keeps them in shared memory; the panel is processed LEFT-LOOKING in
// 64-row blocks, the reference's gemmtrsm structure at tile scale:
//
//   for jb = 0 .. PB/64-1:
//       acc      := alpha * B_jb                                      (8 warps x [64 rows x 8 columns] in registers)
//       acc      -= A(jb,kb) * X_kb   for kb < jb                     DMMA.8x8x4, A blocks streamed through a 2-stage
//                                                                     cp.async ring, X_kb read from shared memory
//       X_jb     := inv(A(jb,jb)) * acc                               diagonal block: eight 8 x 8 micro-blocks
//
// and the diagonal 64 x 64 block itself is the same recursion one level down: micro-block r is solved by forward
// substitution inside the warp -- row l's solution is broadcast with warp shuffles from the four lanes that own it and
// subtracted from the rows below, x_i = ( alpha*b_i - sum_{l<i} a_il x_l ) * inv(a_ii) with the diagonal PRE-INVERTED
// as the reference does (BLIS_ENABLE_TRSM_PREINVERSION; bli_trsm_ref.c:130-134) -- and then applied to the micro-blocks
// below it with two DMMAs each, the X_r fragment re-laid out for the tensor pipe with shuffles (no shared-memory round
// trip on the critical path).  All eight warps work all the time; a warp never needs another warp's columns, so the
// only CTA-wide barrier is the one that hands an A stage over (one per 64 x 64 block of A).
//
// What is kept from the reference: only the stored triangle of A influences the result (the other triangle may hold
// NaN), a unit diagonal is never read, alpha is applied to b11 before the first update, a ragged panel is extended
// with identity rows (zero right-hand sides), upper-triangular panels run the mirror image (last block first, rows
// descending) through index reflection; integer-valued systems are reproduced exactly (every intermediate is exact).
#pragma once
#include "common.cuh"

namespace b200 {

struct TrsmPanelArgs
{
	const double* A;  int64_t rs_a, cs_a;     // diagonal panel of the (effective) triangular matrix, pb x pb
	double*       B;  int64_t rs_b, cs_b;     // pb x n right-hand sides, overwritten with the solution
	int64_t       n;
	int           pb;                          // rows of this panel (1 .. PB)
	int           upper, unit;
	double        alpha;
};

struct TrsmPanelCfg
{
	static constexpr int PB = 256, CN = 64, NB = 64, NT = 256;
	static constexpr int SA  = NB + 4;         // A stage row stride (== 4 mod 16 doubles: conflict-free fragment loads)
	static constexpr int SBK = PB + 4;         // X tile, k-contiguous layout   Xs[n*SBK + k]
	static constexpr int SBN = CN + 4;         // X tile, n-contiguous layout   Xs[k*SBN + n]
	static constexpr int XS_ELEMS = ( CN * SBK > PB * SBN ) ? CN * SBK : PB * SBN;
	static constexpr int AS_ELEMS = NB * SA;
	static constexpr int SMEM_BYTES = ( XS_ELEMS + 2 * AS_ELEMS + PB ) * 8;
};

__device__ __forceinline__ double shfl_f64( double v, int src )
{
	return __shfl_sync( 0xffffffffu, v, src );
}

__global__ void __launch_bounds__( TrsmPanelCfg::NT, 1 )
trsm_panel_kernel( const TrsmPanelArgs a )
{
	using Cfg = TrsmPanelCfg;
	constexpr int PB = Cfg::PB, CN = Cfg::CN, NB = Cfg::NB, NT = Cfg::NT, SA = Cfg::SA;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	double* const Xs   = reinterpret_cast<double*>( smem_raw );
	double* const As   = Xs + Cfg::XS_ELEMS;
	double* const dinv = As + 2 * Cfg::AS_ELEMS;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, t = lane & 3;
	const int pb = a.pb;
	const int nblk = ( pb + NB - 1 ) / NB;
	const int64_t j0 = (int64_t)blockIdx.x * CN;
	const int nc = (int)min( (int64_t)CN, a.n - j0 );
	const bool upper = a.upper != 0;
	const bool a_ifast = ( a.rs_a <= a.cs_a );           // rows of A contiguous (column-major): stage as [l][i]
	const bool b_kfast = ( a.rs_b <= a.cs_b );           // rows of B contiguous (column-major): X tile as [n][k]
	const int xk_s = b_kfast ? 1 : Cfg::SBN, xn_s = b_kfast ? Cfg::SBK : 1;
	const int ai_s = a_ifast ? 1 : SA, al_s = a_ifast ? SA : 1;

	// in-block reflection: logical row r of a 64-block is physical row 63 - r of an upper-triangular panel
	auto refl = [&]( int r ) { return upper ? NB - 1 - r : r; };
	auto pblk = [&]( int jb ) { return upper ? nblk - 1 - jb : jb; };

	// ---- loaders (8-byte cp.async: any strides, any alignment; zero fill outside the panel) ----
	auto load_a_block = [&]( int stage, int pj, int pk )
	{
		const uint32_t sb = smem_u32( As + stage * Cfg::AS_ELEMS );
		#pragma unroll 4
		for ( int it = 0; it < NB * NB / NT; ++it )
		{
			const int e = tid + it * NT;
			int pi, pl;
			if ( a_ifast ) { pi = e % NB; pl = e / NB; } else { pl = e % NB; pi = e / NB; }
			const int gi = pj * NB + pi, gl = pk * NB + pl;
			const bool ok = ( gi < pb && gl < pb );
			const double* src = ok ? a.A + gi * a.rs_a + gl * a.cs_a : a.A;
			cp_async<8>( sb + (uint32_t)( pi * ai_s + pl * al_s ) * 8u, src, ok ? 8 : 0 );
		}
	};
	auto load_b_block = [&]( int pj )
	{
		const uint32_t sb = smem_u32( Xs );
		#pragma unroll 4
		for ( int it = 0; it < NB * CN / NT; ++it )
		{
			const int e = tid + it * NT;
			int k, n;
			if ( b_kfast ) { k = e % NB; n = e / NB; } else { n = e % CN; k = e / CN; }
			const int gk = pj * NB + k;
			const bool ok = ( gk < pb && n < nc );
			const double* src = ok ? a.B + gk * a.rs_b + ( j0 + n ) * a.cs_b : a.B;
			cp_async<8>( sb + (uint32_t)( gk * xk_s + n * xn_s ) * 8u, src, ok ? 8 : 0 );
		}
	};

	// ---- pre-inverted diagonal (unit diagonal: never read) ----
	for ( int k = tid; k < PB; k += NT )
	{
		double d = 1.0;
		if ( k < pb && !a.unit ) d = 1.0 / a.A[k * ( a.rs_a + a.cs_a )];
		dinv[k] = d;
	}

	// work items: the 64 x 64 blocks (jb, kb <= jb) of the panel in the order they are used
	const int n_items = nblk * ( nblk + 1 ) / 2;
	{
		load_b_block( pblk( 0 ) );
		load_a_block( 0, pblk( 0 ), pblk( 0 ) );
		cp_async_commit();
	}

	double acc[8][2];
	const int wn = warp * 8;                                 // this warp's columns of the tile
	int jb = 0, kb = 0;
	for ( int w = 0; w < n_items; ++w )
	{
		cp_async_wait<0>();
		__syncthreads();                                      // item w (and its B block) landed; stage (w+1)&1 is free
		{
			int njb = jb, nkb = kb + 1;
			if ( nkb > njb ) { ++njb; nkb = 0; }
			if ( w + 1 < n_items )
			{
				if ( nkb == 0 ) load_b_block( pblk( njb ) );
				load_a_block( ( w + 1 ) & 1, pblk( njb ), pblk( nkb ) );
			}
			cp_async_commit();
		}
		const double* as = As + ( w & 1 ) * Cfg::AS_ELEMS;
		const int pj = pblk( jb );

		if ( kb == 0 )
		{
			// acc := alpha * B_jb   (C-fragment layout: lane (g,t) holds rows 8*mt + g, columns wn + 2t, wn + 2t + 1)
			#pragma unroll
			for ( int mt = 0; mt < 8; ++mt )
			{
				const int k = pj * NB + refl( mt * 8 + g );
				acc[mt][0] = a.alpha * Xs[k * xk_s + ( wn + 2 * t ) * xn_s];
				acc[mt][1] = a.alpha * Xs[k * xk_s + ( wn + 2 * t + 1 ) * xn_s];
			}
		}

		if ( kb < jb )
		{
			// acc -= A(jb,kb) * X_kb : 16 k4-steps x 8 row tiles of DMMA.8x8x4
			const int pk = pblk( kb );
			const double* xb = Xs + ( wn + g ) * xn_s;
			#pragma unroll 4
			for ( int s = 0; s < NB / 4; ++s )
			{
				const int l = refl( s * 4 + t );
				const double bf = xb[( pk * NB + l ) * xk_s];
				double af[8];
				#pragma unroll
				for ( int mt = 0; mt < 8; ++mt ) af[mt] = -as[refl( mt * 8 + g ) * ai_s + l * al_s];
				#pragma unroll
				for ( int mt = 0; mt < 8; ++mt ) dmma884( acc[mt][0], acc[mt][1], af[mt], bf );
			}
		}
		else
		{
			// diagonal block: micro-block mt is solved by substitution inside the warp, then applied to the micro-blocks below
			#pragma unroll
			for ( int mt = 0; mt < 8; ++mt )
			{
				const int r = mt * 8 + g;                        // this lane's logical row in the block
				const int pr = refl( r );
				const double dg = dinv[pj * NB + pr];
				double arow[8];
				#pragma unroll
				for ( int l = 0; l < 8; ++l ) arow[l] = ( l < g ) ? as[pr * ai_s + refl( mt * 8 + l ) * al_s] : 0.0;
				#pragma unroll
				for ( int l = 0; l < 8; ++l )
				{
					const double x0 = acc[mt][0] * dg, x1 = acc[mt][1] * dg;
					const double b0 = shfl_f64( x0, 4 * l + t ), b1 = shfl_f64( x1, 4 * l + t );
					if ( g == l ) { acc[mt][0] = x0; acc[mt][1] = x1; }
					if ( g > l )  { acc[mt][0] = fma( -arow[l], b0, acc[mt][0] ); acc[mt][1] = fma( -arow[l], b1, acc[mt][1] ); }
				}
				// X_mt (C layout) -> shared memory (later blocks and the final store read it from there)
				{
					const int k = pj * NB + pr;
					Xs[k * xk_s + ( wn + 2 * t ) * xn_s]     = acc[mt][0];
					Xs[k * xk_s + ( wn + 2 * t + 1 ) * xn_s] = acc[mt][1];
				}
				if ( mt < 7 )
				{
					// X_mt as B-operand fragments: lane (g,t) needs X[row 4s + t][column g], held by lane (4s + t, g >> 1), element g & 1
					double bf[2];
					#pragma unroll
					for ( int s = 0; s < 2; ++s )
					{
						const int src = ( ( 4 * s + t ) << 2 ) | ( g >> 1 );
						const double v0 = shfl_f64( acc[mt][0], src ), v1 = shfl_f64( acc[mt][1], src );
						bf[s] = ( g & 1 ) ? v1 : v0;
					}
					#pragma unroll
					for ( int m2 = mt + 1; m2 < 8; ++m2 )
					{
						const int pr2 = refl( m2 * 8 + g );
						#pragma unroll
						for ( int s = 0; s < 2; ++s )
						{
							const double af = -as[pr2 * ai_s + refl( mt * 8 + 4 * s + t ) * al_s];
							dmma884( acc[m2][0], acc[m2][1], af, bf[s] );
						}
					}
				}
			}
			__syncwarp();                                        // X_jb visible to the whole warp before later blocks read it
		}
		if ( ++kb > jb ) { ++jb; kb = 0; }
	}
	cp_async_wait<0>();
	__syncthreads();

	// ---- X tile -> B (coalesced along B's contiguous dimension) ----
	#pragma unroll 4
	for ( int it = 0; it < PB * CN / NT; ++it )
	{
		const int e = tid + it * NT;
		int k, n;
		if ( b_kfast ) { k = e % PB; n = e / PB; } else { n = e % CN; k = e / CN; }
		if ( k < pb && n < nc ) a.B[k * a.rs_b + ( j0 + n ) * a.cs_b] = Xs[k * xk_s + n * xn_s];
	}
}

} // namespace b200
