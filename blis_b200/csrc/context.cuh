// context.cuh -- host-side runtime of the engine: device context, per-thread
// stream, error reporting, pointer classification and pinned staging.
//
// Reference counterparts (what this stands in for on the device side):
//   bli_init_once / bli_finalize          frame/base/bli_init.c:87-99
//   pack-buffer allocator (pba) + pools   frame/base/bli_pba.c:93-188
//   bli_check_error_code -> abort         frame/base/bli_error.c:126-139
// BLIS is re-entrant and may be called from many application threads at once
// (SURVEY.md 8b "Threading"), so global state is guarded and the stream
// selection is thread-local.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <atomic>
#include "common.cuh"

namespace b200 {

// ---- errors -------------------------------------------------------------------
extern thread_local char g_err[512];
int fail( const char* fmt, ... );

#define B200_CUDA( call ) \
	do { cudaError_t e__ = ( call ); if ( e__ != cudaSuccess ) \
	     return ::b200::fail( "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString( e__ ) ); } while ( 0 )

// ---- context --------------------------------------------------------------------
struct Context
{
	std::atomic<bool> ready{false};        // set last by do_init (release), read by every entry point (acquire)
	int          device  = -1;
	int          num_sms = kNumSMs;
	cudaStream_t stream  = nullptr;       // engine-owned default stream
	cudaStream_t copy_stream = nullptr;   // host->device staging of pipelined calls
	cudaStream_t d2h_stream  = nullptr;   // device->host of pipelined calls
	static constexpr int kBatchStreams = 8;
	cudaStream_t batch_streams[kBatchStreams] = {};   // b200_gemm_batch: independent small problems run concurrently
	cudaEvent_t  batch_fork = nullptr, batch_join[kBatchStreams] = {};
	std::mutex   batch_mu;
	void*        batch_desc = nullptr;    // pinned upload buffer of the grouped kernel's problem records (guarded by batch_mu)
	size_t       batch_desc_bytes = 0;
	cudaEvent_t  batch_desc_done = nullptr;
	// pinned staging ring for pageable host operands
	static constexpr int    kStageBufs  = 2;
	static constexpr size_t kStageBytes = (size_t)64 << 20;
	void*        stage[kStageBufs]      = { nullptr, nullptr };
	cudaEvent_t  stage_free[kStageBufs] = { nullptr, nullptr };
	std::mutex   stage_mu;
	// tuning knobs (b200_set_option)
	int          dgemm_cfg = -1;         // auto: TMA 128x128x16 (cfg 9) when aligned, cp.async ws (cfg 6) otherwise, 128x64 2 CTAs/SM (cfg 7) for small problems
	int          zgemm_cfg = 1;          // warp-specialised 64x128x8, 5 stages
	int          sgemm_cfg = -1;         // auto: TMA + FFMA2 kernel when aligned, cp.async kernel otherwise
	int          cgemm_cfg = -1;         // auto: TMA + FFMA2 kernel when aligned, cp.async ws kernel otherwise
	int          trsm_nb   = 0;           // 0 = default
	int          trsm_fused = 1;          // dtrsm: fused 256-row diagonal-panel kernel (trsm_panel.cuh); 0 = 64-row block solves + gemm updates
	int          grid_mult = 1;           // persistent CTAs per SM
	// dynamic tile scheduling: self re-arming {tile, done} pairs, ONE PAIR PER STREAM (kernels of one stream run one
	// after the other, so a stream's pair is never shared by two running kernels; see sched_slot)
	static constexpr int kSchedSlots = 2048;
	static constexpr int kSkTail = 160;  // split-k tail (gemm_dmma_tma.cuh SK): at most this many tail tiles, one self re-arming counter per consumer warp (8) each, per stream
	int*         sched_counters = nullptr;   // kSchedSlots pairs, then kSchedSlots x 8*kSkTail split-k counters
	int          dynamic_tiles = 1;
	int          raster_group = 8;      // tile rows per raster group: the ~148 running tiles form a raster_group x 148/raster_group block
	int          tma_l2_promotion = 2;  // CUtensorMapL2promotion: 0 none, 1 64 B, 2 128 B, 3 256 B
	int          dmma_pp  = 0;          // dgemm TMA kernel: ping-pong the two q-halves of a tile for K <= this (0 = never)
	int          dmma_cst = 256;        // dgemm TMA kernel: stage D through the ring (TMA load + TMA store) for K <= this (0 = never); [B200] wins up to k = 256
	int          dgemm_splitk = 1;      // dgemm TMA kernel: cut the tiles of a partial last wave into k chunks (mid-size problems; 0 = never split k)
	int          batch_grouped = 1;     // b200_gemm_batch: small device-resident problems share ONE launch (gemm_grouped.cuh); 0 = a launch each on the stream pool
	long long    batch_grouped_max = 128ll * 128 * 128;   // ... "small" = m*n*k at most this
	int          trsm_host_pipe = 1;    // trsm with pinned host operands: transfers run under the solve (row-block pipeline, host_trsm.cuh); 0 = sequential transfers
	long long    trsm_host_rb = 0;      // its rows per block (0 = the engine's choice: m / trsm_host_rb_div in whole 256 rows for m >= trsm_host_rb_min_m; -1 = sequential transfers)
	long long    trsm_host_rb_div = 32;
	long long    trsm_host_rb_min_m = 4096;
	int          host_trace = 0;        // print the event timeline of pipelined host-operand calls to stderr (diagnostic)
	int          host_kpipe = 1;        // host operands with long k: pipeline over k panels instead of column blocks
	int          ktri_skip = 1;         // trmm/trmm3: tiles skip the k range in which the triangular operand is zero
	int          transpose_y = 1;       // s/c: transpose a k-contiguous Y panel once instead of re-pairing registers in the k loop
	std::atomic<unsigned long long> launches{0};   // kernels launched by this engine (b200_launch_count)
};

Context& ctx();
int ensure_init();                         // lazy init + cudaSetDevice(engine device) for the calling thread; kSuccess/kFailure
// Counts one kernel launch and records its name (per-thread "last kernel" and a per-name histogram: b200_last_kernel,
// b200_kernel_stats).  `name` must have static storage duration.
void note_launch( const char* name );
// The {tile, done} counter pair dynamic tile scheduling uses on stream `st` (nullptr: none left -> static schedule).
int* sched_slot( cudaStream_t st );
// The split-k counters that belong to a stream's pair (8*Context::kSkTail ints, all zero between launches).
int* sk_flags_of( int* sched_pair );
cudaStream_t cur_stream();                 // thread's selected stream or the engine's

// ---- pointer classification ------------------------------------------------------
enum class MemKind { Device, HostPinned, HostPageable };
MemKind classify( const void* p );

// Temporary device memory, stream ordered (cudaMallocAsync pool).
int dev_alloc( void** p, size_t bytes, cudaStream_t st );
void dev_free( void* p, cudaStream_t st );

// Copy an m x n matrix of `es`-byte elements between host (rs,cs strides, any
// kind of host memory) and a dense column-major device buffer (ld = m).
int stage_to_device( void* dst, const void* src, int64_t m, int64_t n, int64_t rs, int64_t cs,
                     size_t es, cudaStream_t st );
int stage_to_host( void* dst, int64_t rs, int64_t cs, const void* src, int64_t m, int64_t n,
                   size_t es, cudaStream_t st );
// One rectangular piece of a larger column-major device image (leading dimension ldd elements) <- host (rs, cs).
int stage_block_to_device( void* dst, int64_t ldd, const void* src, int64_t m, int64_t n, int64_t rs, int64_t cs, size_t es, cudaStream_t st );
int stage_block_to_host( void* dst, int64_t rs, int64_t cs, const void* src, int64_t ldd, int64_t m, int64_t n, size_t es, cudaStream_t st );
// m x m triangular host matrix -> dense column-major device image (ld = m); only the stored triangle is transferred.
int stage_tri_to_device( void* dst, const void* src, int64_t m, int64_t rs, int64_t cs, bool upper,
                         size_t es, cudaStream_t st );

} // namespace b200
