// gemm_s.cu -- kernel selection and launch for datatype float (see gemm_launch.cuh).
#define B200_GEMM_LAUNCHERS
#include "gemm_launch.cuh"

namespace b200 {

template <>
int launch_gemm_kernel<float>( GemmArgs<float>& g, bool xk, bool yk, bool al, cudaStream_t st )
{
	Context& c = ctx();
	if ( g.nseg > 1 ) return fail( "b200_gemm_kpanels: only d and z are supported" );
	g.tiles_p = (int)( ( g.P + 127 ) / 128 ); g.tiles_q = (int)( ( g.Q + 127 ) / 128 );
	const int grid = (int)std::min<int64_t>( (int64_t)g.tiles_p * g.tiles_q, (int64_t)c.num_sms * c.grid_mult );
	// default (sgemm_cfg < 0 or 3): TMA + packed-FFMA2 kernel when the operands are 16-byte aligned
	if ( g.tri || g.ktri )
	{
		if ( tma_eligible( g, xk, yk, al ) ) return launch_ffma_tma<true>( g, xk, yk, grid, st );
		return launch_ffma<float, 128, 128, 16, 8, 8, 4>( g, xk, yk, al, grid, st );       // run-time tri support
	}
	const int64_t t128 = (int64_t)g.tiles_p * g.tiles_q;
	// [B200] 1024^3 (64 tiles of 128x128): 18.7 -> 23.8 TFLOP/s, 768^3: 9.8 -> 16.8, 512^3: 3.8 -> 6.6; at 1536^3 (144 tiles) the
	// 128x128 TMA kernel is better again (45.8 vs 31.7), hence the 55 % threshold
	if ( ( c.sgemm_cfg < 0 && 20 * t128 < 11 * c.num_sms ) || c.sgemm_cfg == 4 || c.sgemm_cfg == 5 )
	{
		// problems that cannot fill the SMs with 128x128 tiles: 64x128 tiles, or 64x64 when those are still too few
		// (cp.async kernel, several CTAs per SM)
		const bool tiny = ( c.sgemm_cfg == 5 ) || ( c.sgemm_cfg < 0 && 2 * t128 < c.num_sms );
		g.tiles_p = (int)( ( g.P + 63 ) / 64 ); g.tiles_q = (int)( ( g.Q + ( tiny ? 63 : 127 ) ) / ( tiny ? 64 : 128 ) );
		const int small_grid = (int)std::min<int64_t>( (int64_t)g.tiles_p * g.tiles_q, (int64_t)c.num_sms * 4 );
		if ( tiny ) return launch_ffma<float, 64, 64, 16, 4, 4, 4>( g, xk, yk, al, small_grid, st );
		return launch_ffma<float, 64, 128, 16, 4, 8, 4>( g, xk, yk, al, small_grid, st );
	}
	if ( ( c.sgemm_cfg < 0 || c.sgemm_cfg == 3 ) && tma_eligible( g, xk, yk, al ) )
	{
		// small k: the read-modify-write of D is staged through the TMA ring as well (gemm_ffma_tma.cuh, CST)
		if ( c.dmma_cst && !yk && !g.beta_is_zero && g.d_vec_ok && g.K <= c.dmma_cst && g.ldd >= g.Q && g.ldd * 4 < ( 1ll << 40 ) )
			return launch_ffma_tma<false, true>( g, xk, yk, grid, st );
		return launch_ffma_tma( g, xk, yk, grid, st );
	}
	if ( c.sgemm_cfg == 1 ) return launch_ffma_ws<float, 128, 128, 16, 8, 8, 5>( g, xk, yk, al, grid, st );
	if ( c.sgemm_cfg == 2 ) return launch_ffma_ws<float, 128, 128, 32, 8, 8, 4>( g, xk, yk, al, grid, st );
	return launch_ffma<float, 128, 128, 16, 8, 8, 4>( g, xk, yk, al, grid, st );
}

} // namespace b200
