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
	if ( ( c.sgemm_cfg < 0 || c.sgemm_cfg == 3 ) && tma_eligible( g, xk, yk, al ) ) return launch_ffma_tma( g, xk, yk, grid, st );
	if ( c.sgemm_cfg == 1 ) return launch_ffma_ws<float, 128, 128, 16, 8, 8, 5>( g, xk, yk, al, grid, st );
	if ( c.sgemm_cfg == 2 ) return launch_ffma_ws<float, 128, 128, 32, 8, 8, 4>( g, xk, yk, al, grid, st );
	return launch_ffma<float, 128, 128, 16, 8, 8, 4>( g, xk, yk, al, grid, st );
}

} // namespace b200
