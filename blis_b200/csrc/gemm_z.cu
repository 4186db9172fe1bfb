// gemm_z.cu -- kernel selection and launch for datatype double2 (see gemm_launch.cuh).
#define B200_GEMM_LAUNCHERS
#include "gemm_launch.cuh"

namespace b200 {

template <>
int launch_gemm_kernel<double2>( GemmArgs<double2>& g, bool xk, bool yk, bool al, cudaStream_t st )
{
	Context& c = ctx();
	g.tiles_p = (int)( ( g.P + 63 ) / 64 ); g.tiles_q = (int)( ( g.Q + 127 ) / 128 );
	const int grid = (int)std::min<int64_t>( (int64_t)g.tiles_p * g.tiles_q, (int64_t)c.num_sms * c.grid_mult );
	// zgemm_cfg: 1 (and 0) warp-specialised cp.async kernel, 2 (and < 0 = auto) TMA kernel when eligible
	const bool tma = ( c.zgemm_cfg == 2 || c.zgemm_cfg < 0 ) && tma_eligible_z( g, xk, yk );
	if ( g.tri || g.ktri )
	{
		if ( tma ) return launch_zmma_tma<true>( g, xk, yk, grid, st );
		return launch_dmma_ws<double2, 64, 128, 8, 2, 4, 5, true>( g, xk, yk, al, grid, st );
	}
	if ( tma ) return launch_zmma_tma( g, xk, yk, grid, st );
	return launch_dmma_ws<double2, 64, 128, 8, 2, 4, 5>( g, xk, yk, al, grid, st );
}

} // namespace b200
