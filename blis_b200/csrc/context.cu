// context.cu -- engine runtime: init, streams, errors, pinned staging.
#include "context.cuh"
#include <stdlib.h>
#include "../../include/blis_b200.h"
#include <algorithm>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace b200 {

thread_local char g_err[512] = "";
static thread_local cudaStream_t t_stream = nullptr;
static thread_local bool t_stream_set = false;

int fail( const char* fmt, ... )
{
	va_list ap; va_start( ap, fmt );
	vsnprintf( g_err, sizeof( g_err ), fmt, ap );
	va_end( ap );
	return kFailure;
}

Context& ctx() { static Context c; return c; }
static std::mutex g_init_mu;

// ---- launch bookkeeping ---------------------------------------------------------------------
static thread_local const char* t_last_kernel = "";
static std::mutex g_stats_mu;
static std::map<std::string, unsigned long long> g_kernel_stats;

void note_launch( const char* name )
{
	ctx().launches++;
	t_last_kernel = name;
	std::lock_guard<std::mutex> lk( g_stats_mu );
	g_kernel_stats[name]++;
}

// ---- tile-scheduler counters ------------------------------------------------------------------
// Every launch of a persistent gemm kernel draws its tiles from a {next tile, finished CTAs} pair that the last CTA
// re-arms (gemm_dmma_ws.cuh).  Two kernels running at the same time must never share a pair, so pairs are owned by
// STREAMS: kernels of one stream execute one after the other (the engine does not use programmatic dependent launch),
// and different streams -- the caller's threads (b200_set_stream), the 8-stream batch pool -- get different pairs.
// A stream handle that is destroyed and reused simply inherits the (re-armed) pair.
static std::mutex g_sched_mu;
static std::unordered_map<cudaStream_t, int> g_sched_of_stream;

int* sched_slot( cudaStream_t st )
{
	Context& c = ctx();
	if ( !c.dynamic_tiles || !c.sched_counters ) return nullptr;
	std::lock_guard<std::mutex> lk( g_sched_mu );
	auto it = g_sched_of_stream.find( st );
	if ( it == g_sched_of_stream.end() )
	{
		if ( (int)g_sched_of_stream.size() >= Context::kSchedSlots ) return nullptr;      // static schedule from here on
		it = g_sched_of_stream.emplace( st, (int)g_sched_of_stream.size() ).first;
	}
	return c.sched_counters + 2 * it->second;
}
int* sk_flags_of( int* sched_pair )
{
	Context& c = ctx();
	if ( !sched_pair || !c.sched_counters ) return nullptr;
	const int64_t slot = ( sched_pair - c.sched_counters ) / 2;
	return c.sched_counters + 2 * Context::kSchedSlots + slot * 8 * Context::kSkTail;
}

// Environment convention of the reference (frame/base/bli_env.c:68: BLIS_NUM_THREADS, BLIS_JC_NT, ... read once at
// initialisation): every tuning knob of b200_set_option can be preset as BLIS_B200_<KEY> (upper case), e.g.
// BLIS_B200_DGEMM_CFG=6, BLIS_B200_TRSM_FUSED=0, BLIS_B200_DMMA_CST=0; BLIS_B200_DEVICE picks the GPU when the caller did
// not (b200_init( -1 ) with no device made current).  Unparsable values are ignored, as bli_env_get_var does.
extern "C" int b200_set_option( const char* key, long long value );
static void apply_env_options()
{
	static const char* keys[] = { "dgemm_cfg", "zgemm_cfg", "sgemm_cfg", "cgemm_cfg", "grid_mult", "dynamic_tiles", "transpose_y", "ktri_skip",
	                              "host_kpipe", "host_trace", "dmma_cst", "dmma_pp", "dgemm_splitk", "trsm_host_pipe", "trsm_host_rb", "trsm_host_rb_min_m", "trsm_host_rb_div", "batch_grouped", "batch_grouped_max", "trsm_fused", "dist_ab_static", "tma_l2_promotion", "raster_group", "reserve_sms" };
	for ( const char* k : keys )
	{
		char name[64] = "BLIS_B200_"; size_t n = strlen( name );
		for ( const char* p = k; *p && n + 1 < sizeof( name ); ++p ) name[n++] = (char)( *p >= 'a' && *p <= 'z' ? *p - 32 : *p );
		name[n] = 0;
		const char* v = getenv( name );
		if ( !v || !*v ) continue;
		char* end = nullptr; const long long val = strtoll( v, &end, 10 );
		if ( end && *end == 0 ) b200_set_option( k, val );
	}
}

static int do_init( int device )
{
	Context& c = ctx();
	std::lock_guard<std::mutex> lk( g_init_mu );
	if ( c.ready ) return kSuccess;
	int ndev = 0;
	if ( cudaGetDeviceCount( &ndev ) != cudaSuccess || ndev == 0 )
		return fail( "b200_init: no CUDA device visible; this engine has no CPU fallback" );
	if ( device < 0 )
	{
		const char* ev = getenv( "BLIS_B200_DEVICE" );
		if ( ev && *ev >= '0' && *ev <= '9' && atoi( ev ) < ndev ) device = atoi( ev );
		else B200_CUDA( cudaGetDevice( &device ) );
	}
	B200_CUDA( cudaSetDevice( device ) );
	cudaDeviceProp prop;
	B200_CUDA( cudaGetDeviceProperties( &prop, device ) );
	if ( prop.major != 10 )
		return fail( "b200_init: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor );
	c.device  = device;
	c.num_sms = prop.multiProcessorCount;
	B200_CUDA( cudaStreamCreateWithFlags( &c.stream, cudaStreamNonBlocking ) );
	B200_CUDA( cudaStreamCreateWithFlags( &c.copy_stream, cudaStreamNonBlocking ) );
	B200_CUDA( cudaStreamCreateWithFlags( &c.d2h_stream, cudaStreamNonBlocking ) );
	for ( int i = 0; i < Context::kBatchStreams; ++i )
	{
		B200_CUDA( cudaStreamCreateWithFlags( &c.batch_streams[i], cudaStreamNonBlocking ) );
		B200_CUDA( cudaEventCreateWithFlags( &c.batch_join[i], cudaEventDisableTiming ) );
	}
	B200_CUDA( cudaEventCreateWithFlags( &c.batch_fork, cudaEventDisableTiming ) );
	// keep freed workspace cached in the pool instead of returning it to the OS
	cudaMemPool_t pool;
	if ( cudaDeviceGetDefaultMemPool( &pool, device ) == cudaSuccess )
	{
		uint64_t thresh = UINT64_MAX;
		cudaMemPoolSetAttribute( pool, cudaMemPoolAttrReleaseThreshold, &thresh );
	}
	const size_t sched_bytes = (size_t)( 2 + 8 * Context::kSkTail ) * Context::kSchedSlots * sizeof(int);
	B200_CUDA( cudaMalloc( (void**)&c.sched_counters, sched_bytes ) );
	B200_CUDA( cudaMemset( c.sched_counters, 0, sched_bytes ) );
	{ std::lock_guard<std::mutex> lk2( g_sched_mu ); g_sched_of_stream.clear(); }
	c.ready = true;
	apply_env_options();
	return kSuccess;
}

int ensure_init()
{
	if ( !ctx().ready && do_init( -1 ) != kSuccess ) return kFailure;
	// every calling thread must have the engine's device current: streams, events and the workspace pool belong to it
	// (a fresh host thread starts on device 0)
	static thread_local int t_device = -1;
	const int dev = ctx().device;
	if ( t_device != dev )
	{
		int cur = -1;
		if ( cudaGetDevice( &cur ) != cudaSuccess || cur != dev ) B200_CUDA( cudaSetDevice( dev ) );
		t_device = dev;
	}
	return kSuccess;
}

cudaStream_t cur_stream()
{
	return t_stream_set ? t_stream : ctx().stream;
}

MemKind classify( const void* p )
{
	cudaPointerAttributes at;
	if ( cudaPointerGetAttributes( &at, p ) != cudaSuccess ) { cudaGetLastError(); return MemKind::HostPageable; }
	if ( at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged ) return MemKind::Device;
	if ( at.type == cudaMemoryTypeHost ) return MemKind::HostPinned;
	return MemKind::HostPageable;
}

int dev_alloc( void** p, size_t bytes, cudaStream_t st )
{
	*p = nullptr;
	if ( bytes == 0 ) bytes = 16;
	B200_CUDA( cudaMallocAsync( p, bytes, st ) );
	return kSuccess;
}
void dev_free( void* p, cudaStream_t st ) { if ( p ) cudaFreeAsync( p, st ); }

// ---- pinned staging --------------------------------------------------------------
static int ensure_stage_bufs()
{
	Context& c = ctx();
	if ( c.stage[0] ) return kSuccess;
	for ( int i = 0; i < Context::kStageBufs; ++i )
	{
		B200_CUDA( cudaMallocHost( &c.stage[i], Context::kStageBytes ) );
		B200_CUDA( cudaEventCreateWithFlags( &c.stage_free[i], cudaEventDisableTiming ) );
	}
	return kSuccess;
}

// memcpy of `lines` lines of `len` bytes with different pitches, split over a
// few host threads when the block is large.
static void copy_lines( char* dst, size_t dpitch, const char* src, size_t spitch, size_t len, size_t lines )
{
	const size_t total = len * lines;
	unsigned nthr = 1;
	if ( total >= ( (size_t)8 << 20 ) )
		nthr = std::min<unsigned>( 12, std::max<unsigned>( 1, std::thread::hardware_concurrency() ) );
	auto work = [=]( size_t l0, size_t l1 )
	{
		if ( dpitch == len && spitch == len ) { memcpy( dst + l0 * len, src + l0 * len, ( l1 - l0 ) * len ); return; }
		for ( size_t l = l0; l < l1; ++l ) memcpy( dst + l * dpitch, src + l * spitch, len );
	};
	if ( nthr <= 1 || lines < nthr ) { work( 0, lines ); return; }
	std::vector<std::thread> th;
	for ( unsigned t = 0; t < nthr; ++t )
		th.emplace_back( work, lines * t / nthr, lines * ( t + 1 ) / nthr );
	for ( auto& x : th ) x.join();
}

// Generic element gather for general-stride host matrices (rs != 1 && cs != 1).
static void gather_elems( char* dst, const char* src, int64_t m, int64_t j0, int64_t j1, int64_t rs, int64_t cs, size_t es, bool to_host )
{
	// dense side: column-major m x (j1-j0) block starting at dst/src
	for ( int64_t j = j0; j < j1; ++j )
		for ( int64_t i = 0; i < m; ++i )
		{
			const int64_t so = ( i * rs + j * cs ) * (int64_t)es;
			const int64_t d  = ( i + ( j - j0 ) * m ) * (int64_t)es;
			if ( !to_host ) memcpy( dst + d, src + so, es );
			else            memcpy( dst + so, src + d, es );
		}
}

// column-major device block (leading dimension ldd >= m elements)  <->  host (rs, cs)
static int stage_xfer( void* dev, void* host, int64_t m, int64_t n, int64_t rs, int64_t cs, size_t es,
                       cudaStream_t st, bool to_host, int64_t ldd = 0 )
{
	if ( m <= 0 || n <= 0 ) return kSuccess;
	if ( ldd <= 0 ) ldd = m;
	Context& c = ctx();
	const MemKind kind = classify( host );
	// Dense device image is column-major m x n.  If the host matrix is
	// row-stored (cs == 1) we view the transposed problem: lines are rows.
	const bool col_lines = ( rs == 1 );
	const bool row_lines = ( !col_lines && cs == 1 );
	// the direct 2-D copy needs a pitch that covers the line (BLIS also allows m x 1 operands with cs < m and negative
	// strides: those take the ring / gather path below)
	if ( kind == MemKind::HostPinned && col_lines && ( cs >= m || n == 1 ) )
	{
		if ( n == 1 )
		{
			if ( !to_host ) B200_CUDA( cudaMemcpyAsync( dev, host, m * es, cudaMemcpyHostToDevice, st ) );
			else            B200_CUDA( cudaMemcpyAsync( host, dev, m * es, cudaMemcpyDeviceToHost, st ) );
			return kSuccess;
		}
		if ( !to_host ) B200_CUDA( cudaMemcpy2DAsync( dev, ldd * es, host, cs * es, m * es, n, cudaMemcpyHostToDevice, st ) );
		else            B200_CUDA( cudaMemcpy2DAsync( host, cs * es, dev, ldd * es, m * es, n, cudaMemcpyDeviceToHost, st ) );
		return kSuccess;
	}
	(void)row_lines;
	// pageable (or pinned but not column-stored): go through the pinned ring,
	// packing to / unpacking from dense column-major blocks of columns.
	std::lock_guard<std::mutex> lk( c.stage_mu );
	if ( ensure_stage_bufs() != kSuccess ) return kFailure;
	const size_t colb = (size_t)m * es;
	if ( colb > Context::kStageBytes )
		return fail( "stage: one column (%zu bytes) exceeds the staging buffer", colb );
	const int64_t cols_per = std::max<int64_t>( 1, (int64_t)( Context::kStageBytes / colb ) );
	int buf = 0;
	if ( !to_host )
	{
		for ( int64_t j0 = 0; j0 < n; j0 += cols_per, buf ^= 1 )
		{
			const int64_t j1 = std::min( n, j0 + cols_per );
			char* pin = (char*)c.stage[buf];
			char* d   = (char*)dev + (size_t)j0 * (size_t)ldd * es;
			const size_t dpitch = (size_t)ldd * es;        // == colb for a dense image
			B200_CUDA( cudaEventSynchronize( c.stage_free[buf] ) );
			if ( col_lines ) copy_lines( pin, colb, (const char*)host + (size_t)j0 * cs * es, (size_t)cs * es, colb, (size_t)( j1 - j0 ) );
			else             gather_elems( pin, (const char*)host, m, j0, j1, rs, cs, es, false );
			if ( dpitch == colb ) B200_CUDA( cudaMemcpyAsync( d, pin, colb * ( j1 - j0 ), cudaMemcpyHostToDevice, st ) );
			else                  B200_CUDA( cudaMemcpy2DAsync( d, dpitch, pin, colb, colb, (size_t)( j1 - j0 ), cudaMemcpyHostToDevice, st ) );
			B200_CUDA( cudaEventRecord( c.stage_free[buf], st ) );
		}
		return kSuccess;
	}
	// device -> host, double buffered: the DMA of chunk i runs while the host threads unpack chunk i-1 (the first version
	// waited for every chunk before unpacking it: DMA and unpacking took turns)
	auto unpack = [&]( int b, int64_t j0, int64_t j1 ) -> int
	{
		B200_CUDA( cudaEventSynchronize( c.stage_free[b] ) );
		const char* pin = (const char*)c.stage[b];
		if ( col_lines ) copy_lines( (char*)host + (size_t)j0 * cs * es, (size_t)cs * es, pin, colb, colb, (size_t)( j1 - j0 ) );
		else             gather_elems( (char*)host, pin, m, j0, j1, rs, cs, es, true );
		return kSuccess;
	};
	int64_t prev0 = -1, prev1 = -1;
	for ( int64_t j0 = 0; j0 < n; j0 += cols_per, buf ^= 1 )
	{
		const int64_t j1 = std::min( n, j0 + cols_per );
		char* pin = (char*)c.stage[buf];
		char* d   = (char*)dev + (size_t)j0 * (size_t)ldd * es;
		const size_t dpitch = (size_t)ldd * es;
		B200_CUDA( cudaEventSynchronize( c.stage_free[buf] ) );          // (an earlier upload from this buffer)
		if ( dpitch == colb ) B200_CUDA( cudaMemcpyAsync( pin, d, colb * ( j1 - j0 ), cudaMemcpyDeviceToHost, st ) );
		else                  B200_CUDA( cudaMemcpy2DAsync( pin, colb, d, dpitch, colb, (size_t)( j1 - j0 ), cudaMemcpyDeviceToHost, st ) );
		B200_CUDA( cudaEventRecord( c.stage_free[buf], st ) );
		if ( prev0 >= 0 && unpack( buf ^ 1, prev0, prev1 ) != kSuccess ) return kFailure;
		prev0 = j0; prev1 = j1;
	}
	if ( prev0 >= 0 && unpack( buf ^ 1, prev0, prev1 ) != kSuccess ) return kFailure;
	return kSuccess;
}

int stage_to_device( void* dst, const void* src, int64_t m, int64_t n, int64_t rs, int64_t cs, size_t es, cudaStream_t st )
{
	return stage_xfer( dst, const_cast<void*>( src ), m, n, rs, cs, es, st, false );
}
int stage_to_host( void* dst, int64_t rs, int64_t cs, const void* src, int64_t m, int64_t n, size_t es, cudaStream_t st )
{
	return stage_xfer( const_cast<void*>( src ), dst, m, n, rs, cs, es, st, true );
}

int stage_block_to_device( void* dst, int64_t ldd, const void* src, int64_t m, int64_t n, int64_t rs, int64_t cs, size_t es, cudaStream_t st )
{
	return stage_xfer( dst, const_cast<void*>( src ), m, n, rs, cs, es, st, false, ldd );
}
int stage_block_to_host( void* dst, int64_t rs, int64_t cs, const void* src, int64_t ldd, int64_t m, int64_t n, size_t es, cudaStream_t st )
{
	return stage_xfer( const_cast<void*>( src ), dst, m, n, rs, cs, es, st, true, ldd );
}

// Only the stored triangle of a host-resident triangular matrix travels (the reference's packm never reads the other one
// either: bli_packm_struc_cxk.c:155-301).  The matrix is cut into column panels and each panel into the rows the
// triangle reaches there -- [p0, m) for a lower, [0, p1) for an upper triangle --, so (np + 1) / (2 np) of the square
// moves: 53 % with 16 panels.  The rest of the device image stays uninitialised; the kernels never read it.
int stage_tri_to_device( void* dst, const void* src, int64_t m, int64_t rs, int64_t cs, bool upper, size_t es, cudaStream_t st )
{
	if ( m < 2048 ) return stage_xfer( dst, const_cast<void*>( src ), m, m, rs, cs, es, st, false );
	const int64_t pw = ( ( m + 15 ) / 16 + 63 ) / 64 * 64;
	for ( int64_t p0 = 0; p0 < m; p0 += pw )
	{
		const int64_t p1 = std::min( m, p0 + pw );
		const int64_t r0 = upper ? 0 : p0, r1 = upper ? p1 : m;
		const char* h = (const char*)src + ( r0 * rs + p0 * cs ) * (int64_t)es;
		char*       d = (char*)dst + ( r0 + p0 * m ) * (int64_t)es;
		if ( stage_xfer( d, const_cast<char*>( h ), r1 - r0, p1 - p0, rs, cs, es, st, false, m ) != kSuccess ) return kFailure;
	}
	return kSuccess;
}

} // namespace b200

// ---- C ABI: lifetime -----------------------------------------------------------
using namespace b200;

extern "C" b200_err_t b200_init( int device )
{
	if ( ctx().ready ) return kSuccess;
	return do_init( device );
}

extern "C" void b200_finalize( void )
{
	Context& c = ctx();
	std::lock_guard<std::mutex> lk( g_init_mu );
	if ( !c.ready ) return;
	cudaStreamSynchronize( c.stream );
	for ( int i = 0; i < Context::kStageBufs; ++i )
	{
		if ( c.stage[i] ) { cudaFreeHost( c.stage[i] ); c.stage[i] = nullptr; }
		if ( c.stage_free[i] ) { cudaEventDestroy( c.stage_free[i] ); c.stage_free[i] = nullptr; }
	}
	cudaStreamDestroy( c.stream ); cudaStreamDestroy( c.copy_stream ); cudaStreamDestroy( c.d2h_stream );
	c.stream = c.copy_stream = c.d2h_stream = nullptr;
	for ( int i = 0; i < Context::kBatchStreams; ++i )
	{
		if ( c.batch_streams[i] ) { cudaStreamDestroy( c.batch_streams[i] ); c.batch_streams[i] = nullptr; }
		if ( c.batch_join[i] )    { cudaEventDestroy( c.batch_join[i] );     c.batch_join[i] = nullptr; }
	}
	if ( c.batch_fork ) { cudaEventDestroy( c.batch_fork ); c.batch_fork = nullptr; }
	if ( c.batch_desc_done ) { cudaEventDestroy( c.batch_desc_done ); c.batch_desc_done = nullptr; }
	if ( c.batch_desc ) { cudaFreeHost( c.batch_desc ); c.batch_desc = nullptr; c.batch_desc_bytes = 0; }
	if ( c.sched_counters ) { cudaFree( c.sched_counters ); c.sched_counters = nullptr; }
	c.ready = false;
}

extern "C" const char* b200_last_error( void ) { return g_err; }

extern "C" const char* b200_last_kernel( void ) { return t_last_kernel; }

extern "C" size_t b200_kernel_stats( char* buf, size_t len, int reset )
{
	std::lock_guard<std::mutex> lk( g_stats_mu );
	std::string out;
	for ( const auto& kv : g_kernel_stats ) { out += kv.first; out += '\t'; out += std::to_string( kv.second ); out += '\n'; }
	if ( buf && len ) { const size_t nb = std::min( len - 1, out.size() ); memcpy( buf, out.data(), nb ); buf[nb] = 0; }
	if ( reset ) g_kernel_stats.clear();
	return out.size();
}

// BLIS_MALLOC_USER / BLIS_FREE_USER hooks of config/b200 (bli_family_b200.h):
// page-locked host memory for matrices created by bli_obj_create().
extern "C" void* b200_malloc_pinned( size_t size )
{
	void* p = nullptr;
	if ( ensure_init() != kSuccess ) return nullptr;
	if ( cudaMallocHost( &p, size ? size : 1 ) != cudaSuccess ) { cudaGetLastError(); return nullptr; }
	return p;
}
extern "C" void b200_free_pinned( void* p ) { if ( p ) cudaFreeHost( p ); }

extern "C" int b200_pointer_kind( const void* p )
{
	if ( ensure_init() != kSuccess ) return -1;
	switch ( classify( p ) )
	{
		case MemKind::Device:     return 0;
		case MemKind::HostPinned: return 1;
		default:                  return 2;
	}
}

extern "C" int b200_device_count( void )
{
	int n = 0;
	if ( cudaGetDeviceCount( &n ) != cudaSuccess ) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" const char* b200_info( void )
{
	return "blis_b200 0.1 (sm_100a; d/z: DMMA.8x8x4 tiles, s/c: FFMA2 register tiles; "
	       "TMA (cp.async for unaligned views) multi-stage staging; persistent dynamic tile scheduler)";
}

extern "C" void b200_set_stream( void* stream )
{
	t_stream = (cudaStream_t)stream;
	t_stream_set = ( stream != nullptr );
}
extern "C" void* b200_get_stream( void )
{
	if ( ensure_init() != kSuccess ) return nullptr;
	return (void*)cur_stream();
}
extern "C" b200_err_t b200_sync( void )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	B200_CUDA( cudaStreamSynchronize( cur_stream() ) );
	return kSuccess;
}
