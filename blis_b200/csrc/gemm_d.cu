// gemm_d.cu -- kernel selection and launch for datatype double (see gemm_launch.cuh).
#define B200_GEMM_LAUNCHERS
#include "gemm_launch.cuh"

namespace b200 {

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
				return launch_dmma_tma( g, xk, yk, tiles( 128, 128 ), st );
			}
			return launch_dmma_ws<double, 128, 128, 16, 4, 2, 5>( g, xk, yk, al, tiles( 128, 128 ), st );
		case 7:  return launch_dmma_ws<double, 128, 64, 16, 4, 1, 3>( g, xk, yk, al, tiles( 128, 64, 2 ), st );
		case 10: return launch_dmma_ws<double, 64, 64, 16, 2, 2, 4>( g, xk, yk, al, tiles( 64, 64, 2 ), st );
		case 8:  return launch_dmma_ws<double, 64, 128, 16, 1, 4, 3>( g, xk, yk, al, tiles( 64, 128, 2 ), st );
	}
}

} // namespace b200
