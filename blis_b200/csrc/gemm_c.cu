// gemm_c.cu -- kernel selection and launch for datatype float2 (see gemm_launch.cuh).
#define B200_GEMM_LAUNCHERS
#include "gemm_launch.cuh"

namespace b200 {

template <>
int launch_gemm_kernel<float2>( GemmArgs<float2>& g, bool xk, bool yk, bool al, cudaStream_t st )
{
	Context& c = ctx();
	if ( g.nseg > 1 ) return fail( "b200_gemm_kpanels: only d and z are supported" );
	g.tiles_p = (int)( ( g.P + 63 ) / 64 ); g.tiles_q = (int)( ( g.Q + 127 ) / 128 );
	const int grid = (int)std::min<int64_t>( (int64_t)g.tiles_p * g.tiles_q, (int64_t)c.num_sms * c.grid_mult );
	if ( g.tri || g.ktri )
	{
		if ( tma_eligible( g, xk, yk, al ) ) return launch_cfma_tma<true>( g, xk, yk, grid, st );
		return launch_ffma_ws<float2, 64, 128, 16, 4, 8, 5>( g, xk, yk, al, grid, st );     // run-time tri support
	}
	if ( ( c.cgemm_cfg < 0 || c.cgemm_cfg == 3 ) && tma_eligible( g, xk, yk, al ) )
	{
		// small k: the read-modify-write of D is staged through the TMA ring as well (gemm_cfma_tma.cuh, CST)
		if ( c.dmma_cst && !yk && !g.beta_is_zero && g.d_vec_ok && g.K <= c.dmma_cst && g.ldd >= g.Q && g.ldd * 8 < ( 1ll << 40 ) )
			return launch_cfma_tma<false, true>( g, xk, yk, grid, st );
		return launch_cfma_tma( g, xk, yk, grid, st );
	}
	if ( c.cgemm_cfg != 0 ) return launch_ffma_ws<float2, 64, 128, 16, 4, 8, 5>( g, xk, yk, al, grid, st );
	return launch_ffma<float2, 64, 128, 16, 4, 8, 4>( g, xk, yk, al, grid, st );
}

} // namespace b200
