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
// owns CN = 64 columns of B for the whole panel; the panel is processed LEFT-LOOKING in 64-row blocks, the reference's
// gemmtrsm structure at tile scale:
//
//   for jb = 0 .. PB/64-1:
//       acc      := alpha * B_jb                                      (8 warps x [64 rows x 8 columns] in registers,
//                                                                      fetched from global memory one block ahead)
//       acc      -= A(jb,kb) * X_kb   for kb < jb                     DMMA.8x8x4, A blocks streamed through a 2-stage
//                                                                     cp.async ring, X_kb read from shared memory
//       X_jb     := inv(A(jb,jb)) * acc                               diagonal block: eight 8 x 8 micro-blocks
//
// and the diagonal 64 x 64 block itself is the same recursion one level down: micro-block r is solved by forward
// substitution inside the warp -- row l's solution is broadcast with warp shuffles from the four lanes that own it and
// subtracted from the rows below, x_i = ( alpha*b_i - sum_{l<i} a_il x_l ) * inv(a_ii) with the diagonal PRE-INVERTED
// as the reference does (BLIS_ENABLE_TRSM_PREINVERSION; bli_trsm_ref.c:130-134) -- and then applied to the micro-blocks
// below it with two DMMAs each, the X_r fragment re-laid out for the tensor pipe with shuffles (no shared-memory round
// trip on the critical path).  Solved rows go to global memory straight from the registers and to shared memory for
// the blocks below.  All eight warps work all the time; a warp never needs another warp's columns, so the only
// CTA-wide barrier is the one that hands an A stage over (one per 64 x 64 block of A).
//
// What is kept from the reference: only the stored triangle of A influences the result (the other triangle may hold
// NaN), a unit diagonal is never read, alpha is applied to b11 before the first update, a ragged panel is extended
// with identity rows (zero right-hand sides), upper-triangular panels run the mirror image (last block first, rows
// descending) through index reflection; integer-valued systems are reproduced exactly (every intermediate is exact).
//
// Template parameters fix the three layout decisions at compile time so that every fragment address in the k loops is
// "lane base + constant": UPPER (index reflection), AI (rows of A contiguous: A stage kept as [l][i], else [i][l]),
// BK (rows of B contiguous: X tile kept as [n][k], else [k][n]).  Both shared-memory tiles use row strides == 4 (mod 16)
// doubles, which makes every 8-byte fragment load of a warp bank-conflict free in either orientation.
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
	int           a_vec;                       // A can be staged with 16-byte copies (unit stride along the fast index, aligned)
	double        alpha;
};

struct TrsmPanelCfg
{
	static constexpr int PB = 256, CN = 64, NB = 64, NT = 256;
	static constexpr int SA  = NB + 4;         // A stage row stride
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

template <bool UPPER, bool AI, bool BK>
__global__ void __launch_bounds__( TrsmPanelCfg::NT, 1 )
trsm_panel_kernel( const TrsmPanelArgs a )
{
	using Cfg = TrsmPanelCfg;
	constexpr int PB = Cfg::PB, CN = Cfg::CN, NB = Cfg::NB, NT = Cfg::NT, SA = Cfg::SA;
	constexpr int XK = BK ? 1 : Cfg::SBN, XN = BK ? Cfg::SBK : 1;       // X tile strides along k (panel row) and n (column)
	constexpr int AIS = AI ? 1 : SA, ALS = AI ? SA : 1;                 // A stage strides along i (row) and l (column)
	constexpr int SG = UPPER ? -1 : 1;                                  // direction of a logical step in physical indices
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
	const int wn = warp * 8;                                 // this warp's columns of the tile

	// logical row r of a 64-block is physical row 63 - r of an upper-triangular panel; logical block jb is physical block nblk-1-jb
	auto refl = [&]( int r ) { return UPPER ? NB - 1 - r : r; };
	auto pblk = [&]( int jb ) { return UPPER ? nblk - 1 - jb : jb; };

	// ---- A stage loader (cp.async; zero fill outside the panel; 16-byte copies when the layout allows) ----
	auto load_a_block = [&]( int stage, int pj, int pk )
	{
		const uint32_t sb = smem_u32( As + stage * Cfg::AS_ELEMS );
		const int i_lim = pb - pj * NB, l_lim = pb - pk * NB;              // rows / columns of this block inside the panel
		const double* src0 = a.A + (int64_t)( pj * NB ) * a.rs_a + (int64_t)( pk * NB ) * a.cs_a;
		if ( a.a_vec )
		{
			// fast index f (i for AI, l otherwise) in pairs, slow index s: NB*NB/2 chunks of 16 bytes
			const int64_t fs = AI ? a.cs_a : a.rs_a;                       // stride of the slow index (the fast one has stride 1)
			const int f_lim = AI ? i_lim : l_lim, s_lim = AI ? l_lim : i_lim;
			#pragma unroll
			for ( int it = 0; it < NB * NB / 2 / NT; ++it )
			{
				const int e = tid + it * NT;
				const int f = ( e % ( NB / 2 ) ) * 2, s = e / ( NB / 2 );
				const int nb_ = ( s < s_lim ) ? min( max( f_lim - f, 0 ), 2 ) * 8 : 0;
				cp_async<16>( sb + (uint32_t)( s * SA + f ) * 8u, nb_ ? (const void*)( src0 + s * fs + f ) : (const void*)a.A, nb_ );
			}
		}
		else
		{
			#pragma unroll 4
			for ( int it = 0; it < NB * NB / NT; ++it )
			{
				const int e = tid + it * NT;
				int pi, pl;
				if ( AI ) { pi = e % NB; pl = e / NB; } else { pl = e % NB; pi = e / NB; }
				const bool ok = ( pi < i_lim && pl < l_lim );
				cp_async<8>( sb + (uint32_t)( pi * AIS + pl * ALS ) * 8u, ok ? (const void*)( src0 + pi * a.rs_a + pl * a.cs_a ) : (const void*)a.A, ok ? 8 : 0 );
			}
		}
	};

	// ---- right-hand sides: this lane's C-fragment elements of block jb straight from global memory ----
	// lane (g,t) holds rows 8*mt + g (logical), columns wn + 2t, wn + 2t + 1
	const bool c0_ok = ( wn + 2 * t < nc ), c1_ok = ( wn + 2 * t + 1 < nc );
	double* const bcol = a.B + ( j0 + wn + 2 * t ) * a.cs_b;
	auto load_b_frag = [&]( double ( &bf )[8][2], int jb )
	{
		const int pj = pblk( jb );
		#pragma unroll
		for ( int mt = 0; mt < 8; ++mt )
		{
			const int k = pj * NB + refl( mt * 8 + g );
			const bool rok = ( k < pb );
			bf[mt][0] = ( rok && c0_ok ) ? __ldcs( bcol + k * a.rs_b ) : 0.0;
			bf[mt][1] = ( rok && c1_ok ) ? __ldcs( bcol + k * a.rs_b + a.cs_b ) : 0.0;
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
	load_a_block( 0, pblk( 0 ), pblk( 0 ) );
	cp_async_commit();

	double acc[8][2], bnext[8][2];
	load_b_frag( bnext, 0 );

	// per-lane fragment bases (logical row g / logical k index t of a block; + constant offsets in the loops)
	const int a_lane = refl( g ) * AIS + refl( t ) * ALS;               // A operand: A(8mt + g, 4s + t)
	const int x_lane = ( wn + g ) * XN + refl( t ) * XK;                // B operand: X(4s + t, wn + g)
	const int c_lane = refl( g ) * XK + ( wn + 2 * t ) * XN;            // C layout:  X(8mt + g, wn + 2t)

	int jb = 0, kb = 0;
	for ( int w = 0; w < n_items; ++w )
	{
		cp_async_wait<0>();
		__syncthreads();                                      // item w landed; every warp is done with stage (w+1)&1
		{
			int njb = jb, nkb = kb + 1;
			if ( nkb > njb ) { ++njb; nkb = 0; }
			if ( w + 1 < n_items ) load_a_block( ( w + 1 ) & 1, pblk( njb ), pblk( nkb ) );
			cp_async_commit();
		}
		const double* as = As + ( w & 1 ) * Cfg::AS_ELEMS;
		const int pj = pblk( jb );

		if ( kb == 0 )
		{
			#pragma unroll
			for ( int mt = 0; mt < 8; ++mt ) { acc[mt][0] = a.alpha * bnext[mt][0]; acc[mt][1] = a.alpha * bnext[mt][1]; }
			if ( jb + 1 < nblk ) load_b_frag( bnext, jb + 1 );        // in flight under this block's work
		}

		if ( kb < jb )
		{
			// acc -= A(jb,kb) * X_kb : 16 k4-steps x 8 row tiles of DMMA.8x8x4 (the sign rides on the X fragment)
			const double* ap = as + a_lane;
			const double* xp = Xs + x_lane + pblk( kb ) * NB * XK;
			#pragma unroll
			for ( int s = 0; s < NB / 4; ++s )
			{
				const double bf = -xp[SG * 4 * s * XK];
				double af[8];
				#pragma unroll
				for ( int mt = 0; mt < 8; ++mt ) af[mt] = ap[SG * ( 8 * mt * AIS + 4 * s * ALS )];
				#pragma unroll
				for ( int mt = 0; mt < 8; ++mt ) dmma884( acc[mt][0], acc[mt][1], af[mt], bf );
			}
		}
		else
		{
			// diagonal block: micro-block mt is solved by substitution inside the warp, then applied to the micro-blocks below
			double* const xc = Xs + c_lane + pj * NB * XK;
			#pragma unroll
			for ( int mt = 0; mt < 8; ++mt )
			{
				const int pr = refl( mt * 8 + g );                // this lane's physical row in the block
				const double dg = dinv[pj * NB + pr];
				double arow[8];
				#pragma unroll
				for ( int l = 0; l < 7; ++l ) arow[l] = ( l < g ) ? as[pr * AIS + refl( mt * 8 + l ) * ALS] : 0.0;
				#pragma unroll
				for ( int l = 0; l < 8; ++l )
				{
					const double x0 = acc[mt][0] * dg, x1 = acc[mt][1] * dg;
					if ( g == l ) { acc[mt][0] = x0; acc[mt][1] = x1; }
					if ( l < 7 )
					{
						const double b0 = shfl_f64( x0, 4 * l + t ), b1 = shfl_f64( x1, 4 * l + t );
						if ( g > l ) { acc[mt][0] = fma( -arow[l], b0, acc[mt][0] ); acc[mt][1] = fma( -arow[l], b1, acc[mt][1] ); }
					}
				}
				// X_mt: to shared memory for the blocks below, to global memory because it is final
				xc[SG * 8 * mt * XK]      = acc[mt][0];
				xc[SG * 8 * mt * XK + XN] = acc[mt][1];
				{
					const int k = pj * NB + pr;
					if ( k < pb )
					{
						if ( c0_ok ) bcol[k * a.rs_b] = acc[mt][0];
						if ( c1_ok ) bcol[k * a.rs_b + a.cs_b] = acc[mt][1];
					}
				}
				if ( mt < 7 )
				{
					// -X_mt as B-operand fragments: lane (g,t) needs X[row 4s + t][column g], held by lane (4s + t, g >> 1), element g & 1
					double bf[2];
					#pragma unroll
					for ( int s = 0; s < 2; ++s )
					{
						const int src = ( ( 4 * s + t ) << 2 ) | ( g >> 1 );
						const double v0 = shfl_f64( acc[mt][0], src ), v1 = shfl_f64( acc[mt][1], src );
						bf[s] = -( ( g & 1 ) ? v1 : v0 );
					}
					const double* ap = as + a_lane + SG * 8 * mt * ALS;
					#pragma unroll
					for ( int m2 = mt + 1; m2 < 8; ++m2 )
						#pragma unroll
						for ( int s = 0; s < 2; ++s )
							dmma884( acc[m2][0], acc[m2][1], ap[SG * ( 8 * m2 * AIS + 4 * s * ALS )], bf[s] );
				}
			}
			__syncwarp();                                        // X_jb visible to the whole warp before later blocks read it
		}
		if ( ++kb > jb ) { ++jb; kb = 0; }
	}
	cp_async_wait<0>();
}

} // namespace b200
