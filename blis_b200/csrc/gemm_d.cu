// gemm_d.cu -- kernel selection and launch for datatype double (see gemm_launch.cuh).
#define B200_GEMM_LAUNCHERS
#include "gemm_launch.cuh"

namespace b200 {

// Split-k tail schedule (GemmArgs::sk_*): `tiles` 128x128 tiles on `grid` persistent CTAs.  The r = tiles % grid tiles of the
// last wave are cut into S chunks each; with S chosen so that r*S units fill whole rounds of the grid, the tail costs
// ceil(r*S/grid)/S of a tile time instead of 1.  Returns S (0: leave the schedule alone).
static int splitk_plan( int64_t tiles, int grid, int64_t kt, int* full )
{
	if ( grid <= 0 || tiles <= grid || kt <= 0 ) return 0;        // less than one wave: measured slower with chunks only (1280^3 forced: 22.9 -> 18.0)
	const int64_t waves = ( tiles + grid - 1 ) / grid;
	const int64_t r = tiles % grid;
	if ( r == 0 || waves > 8 ) return 0;                          // nothing idle / the tail is under ~1.5 % of the run
	// Cost of the tail in tile times: ceil(r*S/grid) rounds of 1/S each, plus the fix-up of a round -- ~11 us on a B200
	// (148 CTAs park 128 KiB each and the last arrivals read the slots back: [B200] ncu, 2048^3), i.e. 5/kt of a tile time
	// (a 16-wide k step of a 128x128 tile takes 2.1 us).
	double best = 1.0; int S = 0;
	for ( int s : { 2, 3, 4, 5, 6, 8 } )
	{
		if ( kt / s < 16 ) break;                                   // >= 256 k per chunk
		const double rounds = (double)( ( r * s + grid - 1 ) / grid );
		const double t = rounds * ( 1.0 / s + 5.0 / (double)kt );
		if ( t < best - 1e-9 ) { best = t; S = s; }
	}
	if ( S == 0 || ( 1.0 - best ) / (double)waves < 0.03 ) return 0;   // under 3 % of the whole launch
	*full = (int)( tiles - r );
	return S;
}
// the launch itself: the accumulator slots (r*S x 128 KiB) come from the stream-ordered pool, the counters are the stream's
// own (self re-arming, context.cu)
static int launch_dmma_tma_splitk( GemmArgs<double>& g, bool xk, bool yk, int grid, int S, int full, cudaStream_t st )
{
	const int64_t r = (int64_t)g.tiles_p * g.tiles_q - full;
	int* flags = sk_flags_of( g.tile_counter );
	if ( !flags || r > Context::kSkTail ) return launch_dmma_tma( g, xk, yk, grid, st );
	void* ws = nullptr;
	if ( dev_alloc( &ws, (size_t)r * S * 128 * 128 * 8, st ) != kSuccess ) return kFailure;
	g.sk_full = full; g.sk_split = S; g.sk_ws = (double*)ws; g.sk_flags = flags;
	const int64_t units = full + r * S;
	const int rc = launch_dmma_tma<false, false, true>( g, xk, yk, (int)std::min<int64_t>( units, grid ), st );
	dev_free( ws, st );
	g.sk_full = 0; g.sk_split = 0; g.sk_ws = nullptr; g.sk_flags = nullptr;
	return rc;
}

template <>
int launch_gemm_kernel<double>( GemmArgs<double>& g, bool xk, bool yk, bool al, cudaStream_t st )
{
	Context& c = ctx();
	// grid = min( tiles, SMs x resident CTAs per SM ); `per_sm` = 2 for the kernels with one consumer warpgroup
	auto tiles = [&]( int bp, int bq, int per_sm = 1 ) {
		g.tiles_p = (int)( ( g.P + bp - 1 ) / bp ); g.tiles_q = (int)( ( g.Q + bq - 1 ) / bq );
		return (int)std::min<int64_t>( (int64_t)g.tiles_p * g.tiles_q, (int64_t)c.num_sms * c.grid_mult * per_sm );
	};
	int cfg = c.dgemm_cfg;
	if ( cfg >= 0 && cfg < 4 ) cfg = -1;                // 0-3 were the retired single-role kernel
	if ( g.tri || g.ktri )
	{
		// triangular D (gemmt family): the TRI instantiations of the default kernels
		const int64_t t128 = ( ( g.P + 127 ) / 128 ) * ( ( g.Q + 127 ) / 128 );
		if ( t128 < 2 * c.num_sms )
		{
			return launch_dmma_ws<double, 128, 64, 16, 4, 1, 3, true>( g, xk, yk, al, tiles( 128, 64, 2 ), st );
		}
		if ( tma_eligible( g, xk, yk, al ) ) return launch_dmma_tma<true>( g, xk, yk, tiles( 128, 128 ), st );
		return launch_dmma_ws<double, 128, 128, 16, 4, 2, 5, true>( g, xk, yk, al, tiles( 128, 128 ), st );
	}
	if ( cfg < 0 )
	{
		// auto: the cooperative 128x128 tile unless it cannot fill the SMs once; then 128x64 tiles, two CTAs per SM
		const int64_t t128 = ( ( g.P + 127 ) / 128 ) * ( ( g.Q + 127 ) / 128 );
		// ... and 64x64 tiles when even 128x64 tiles leave SMs idle (1024^3: 64 -> 256 tiles)
		// [B200] 1536^3 (144 tiles): 128x128 wins; 1024^3 (64 tiles): 13 -> 21.4; 768^3: 11.1 (128x64) -> 17.0 (64x64); 512^3: 4.4 -> 6.4
		cfg = ( 4 * t128 < 3 * c.num_sms ) ? ( ( 2 * t128 < c.num_sms && g.nseg == 1 ) ? 10 : 7 ) : ( tma_eligible( g, xk, yk, al ) ? 9 : 6 );
	}
	switch ( cfg )
	{
		default:
		case 4: return launch_dmma_ws<double, 128, 128, 16, 2, 4, 5>( g, xk, yk, al, tiles( 128, 128 ), st );
		case 5: return launch_dmma_ws<double, 128, 128, 32, 2, 4, 3>( g, xk, yk, al, tiles( 128, 128 ), st );
		case 6: return launch_dmma_ws<double, 128, 128, 16, 4, 2, 5>( g, xk, yk, al, tiles( 128, 128 ), st );
		case 9:
			if ( tma_eligible( g, xk, yk, al ) )
			{
				// small k, opt-in: two consumer groups take turns on the tensor pipe (gemm_dmma_pp.cuh; measured slower)
				if ( c.dmma_pp && g.nseg == 1 && g.K <= c.dmma_pp && g.d_vec_ok )
					return launch_dmma_pp( g, xk, yk, tiles( 128, 128 ), st );
				// small k: D travels through the TMA ring in both directions (gemm_dmma_tma.cuh, CST)
				if ( c.dmma_cst && g.d_vec_ok && g.K * g.nseg <= c.dmma_cst && g.ldd >= g.Q && g.ldd * 8 < ( 1ll << 40 ) )
					return launch_dmma_tma<false, true>( g, xk, yk, tiles( 128, 128 ), st );
				// mid-size: a last wave that would leave SMs idle is cut into k chunks (split-k tail, gemm_dmma_tma.cuh SK)
				if ( c.dgemm_splitk && g.nseg == 1 )
				{
					const int grid = tiles( 128, 128 );
					int full = 0;
					const int S = splitk_plan( (int64_t)g.tiles_p * g.tiles_q, c.num_sms * c.grid_mult, ( g.K + 15 ) / 16, &full );
					if ( S ) return launch_dmma_tma_splitk( g, xk, yk, grid, S, full, st );
				}
				return launch_dmma_tma( g, xk, yk, tiles( 128, 128 ), st );
			}
			return launch_dmma_ws<double, 128, 128, 16, 4, 2, 5>( g, xk, yk, al, tiles( 128, 128 ), st );
		case 7:  return launch_dmma_ws<double, 128, 64, 16, 4, 1, 3>( g, xk, yk, al, tiles( 128, 64, 2 ), st );
		case 10: return launch_dmma_ws<double, 64, 64, 16, 2, 2, 4>( g, xk, yk, al, tiles( 64, 64, 2 ), st );
		case 8:  return launch_dmma_ws<double, 64, 128, 16, 1, 4, 3>( g, xk, yk, al, tiles( 64, 128, 2 ), st );
	}
}

} // namespace b200

extern "C" int b200_splitk_plan( int64_t tiles, int grid, int64_t kt, int* full )
{
	int f = 0;
	const int S = b200::splitk_plan( tiles, grid, kt, &f );
	if ( full ) *full = S ? f : (int)tiles;
	return S;
}
