// host_dist.cuh -- multi-GPU gemm/trsm behind the C ABI: one process per GPU, NCCL over NVLink/NVSwitch.
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
//
// The reference has no distributed layer; what it has is the jc x ic partitioning of C among
// threads -- bli_rntm_factorize -> bli_thread_partition_2x2 (frame/thread/bli_thread.c:194-320) picks the ic x jc
// factor pair, bli_thread_range_sub (frame/thread/bli_thread_range.c:38-184) hands every thread a contiguous range
// with the ragged edge on the last one -- and the rule that k is NEVER split across threads
// (frame/3/gemm/bli_gemm_blk_var3.c:110-112), so there is no reduction and a result does not depend on the thread
// count.  Exactly that arithmetic is applied across GPUs here:
//
//   gemm   C split into a Pr x Pc grid of blocks, rank (i, j) owns C_ij.  A's row panel i lives block-cyclically along
//          k (panel width kb) on the Pc ranks of grid row i, B's column panel j on the Pr ranks of grid column j.
//          Step s all-gathers the next L = lcm(Pr, Pc) k panels in the row group (A) and the column group (B) on the
//          engine's communication stream into one of two receive buffers, while ONE k-panel launch
//          (gemm_dev with nseg = L: 3-D tensor maps over the receive buffer, gemm_dmma_tma.cuh) accumulates step s-1.
//          The gather of step s+2 is ordered only behind the kernels that read its buffer (an event), never behind the
//          caller's stream as a whole, so consecutive products keep the pipeline full (B200_DIST_AB_STATIC).
//   gemm, skinny (sup shapes: m, n >> k)   1-D split of C's columns (or rows) over all ranks; the operand that spans the
//          split dimension is local, the small one is broadcast once (ncclBroadcast) -- SURVEY.md 8e row 3.
//   trsm   column blocks of B (left side; rows for the right side) -- the reference's only parallel trsm loops run over
//          them (frame/3/trsm/bli_trsm_cntl.c:446-451) -- with the triangular A replicated, optionally broadcast from
//          one rank first; no collective during the solve.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded by the host application -- e.g. torch's --
// or the system one), so libblis_b200.so keeps no link-time dependency on it and single-GPU users never load it.
// The communicator is bootstrapped by the CALLER's launcher: rank 0 obtains 128 bytes from b200_dist_unique_id(), the
// launcher (MPI_Bcast, torch.distributed, a file) hands them to every rank, every rank calls b200_dist_init().
#pragma once
#include <dlfcn.h>
#include <numeric>
#include <vector>
#include "host_gemm.cuh"
#include "host_trsm.cuh"

namespace b200 {

// ---- the reference's partitioning arithmetic (bit-exact; tests/test_partition.py checks the C entry points against
// ---- golden vectors produced by the real library) -----------------------------------------------------------------
// bli_thread_range_sub, frame/thread/bli_thread_range.c:38-184
static void range_sub( int64_t work_id, int64_t n_way, int64_t n, int64_t bf, bool handle_edge_low, int64_t* start, int64_t* end )
{
	if ( n_way == 1 ) { *start = 0; *end = n; return; }
	const int64_t n_bf_whole = n / bf, n_bf_left = n % bf;
	int64_t n_bf_lo = n_bf_whole / n_way, n_bf_hi = n_bf_whole / n_way;
	if ( !handle_edge_low )
	{
		const int64_t n_th_lo = n_bf_whole % n_way;
		if ( n_th_lo != 0 ) n_bf_lo += 1;
		const int64_t size_lo = n_bf_lo * bf, size_hi = n_bf_hi * bf;
		const int64_t lo_start = 0, hi_start = n_th_lo * size_lo;
		if ( work_id < n_th_lo ) { *start = lo_start + work_id * size_lo; *end = lo_start + ( work_id + 1 ) * size_lo; }
		else
		{
			*start = hi_start + ( work_id - n_th_lo ) * size_hi;
			*end   = hi_start + ( work_id - n_th_lo + 1 ) * size_hi;
			if ( work_id == n_way - 1 ) *end += n_bf_left;
		}
		return;
	}
	const int64_t n_th_hi = n_bf_whole % n_way, n_th_lo = n_way - n_th_hi;
	if ( n_th_hi != 0 ) n_bf_hi += 1;
	const int64_t size_lo = n_bf_lo * bf, size_hi = n_bf_hi * bf;
	const int64_t hi_start = n_th_lo * size_lo + n_bf_left;
	if ( work_id < n_th_lo )
	{
		*start = work_id * size_lo; *end = ( work_id + 1 ) * size_lo;
		if ( work_id == 0 ) *end += n_bf_left; else { *start += n_bf_left; *end += n_bf_left; }
	}
	else { *start = hi_start + ( work_id - n_th_lo ) * size_hi; *end = hi_start + ( work_id - n_th_lo + 1 ) * size_hi; }
}

// bli_thread_partition_2x2 (fast heuristic), frame/thread/bli_thread.c:194-320
static void partition_2x2( int64_t n_thread, int64_t work1, int64_t work2, int64_t* nt1, int64_t* nt2 )
{
	if ( n_thread < 4 ) { *nt1 = ( work1 >= work2 ? n_thread : 1 ); *nt2 = ( work1 < work2 ? n_thread : 1 ); return; }
	int64_t tn1 = 1, tn2 = 1, rem = n_thread, f = 2;
	while ( rem > 1 )
	{
		while ( rem % f ) ++f;
		rem /= f;
		if ( work1 > work2 ) { work1 /= f; tn1 *= f; } else { work2 /= f; tn2 *= f; }
	}
	auto iabs = []( int64_t v ) { return v < 0 ? -v : v; };
	if ( work1 > work2 )      { if ( tn2 % 2 == 0 && iabs( work1 / 2 - work2 * 2 ) < work1 - work2 ) { tn1 *= 2; tn2 /= 2; } }
	else if ( work1 < work2 ) { if ( tn1 % 2 == 0 && iabs( work2 / 2 - work1 * 2 ) < work2 - work1 ) { tn1 /= 2; tn2 *= 2; } }
	*nt1 = tn1; *nt2 = tn2;
}

static int dist_plan( int world, int rank, int64_t m, int64_t n, int64_t k, int64_t kb, b200_dist_plan_t* p )
{
	if ( world < 1 || rank < 0 || rank >= world ) return fail( "b200_dist_plan: bad world/rank %d/%d", world, rank );
	if ( m < 0 || n < 0 || k < 0 || kb < 1 ) return fail( "b200_dist_plan: bad dimensions" );
	int64_t pr, pc;
	partition_2x2( world, m, n, &pr, &pc );
	if ( pr * pc != world ) return fail( "b200_dist_plan: %d ranks do not factor into a grid", world );
	p->world = world; p->rank = rank; p->pr = (int)pr; p->pc = (int)pc;
	p->i = rank / (int)pc; p->j = rank % (int)pc;
	p->L = (int)std::lcm( pr, pc );
	p->kb = kb; p->T = (int)( ( k + kb - 1 ) / kb );
	if ( k % kb != 0 || p->T % p->L != 0 )
		return fail( "b200_dist_plan: k = %lld must be a multiple of kb * lcm(Pr, Pc) = %lld", (long long)k, (long long)( kb * p->L ) );
	p->steps = p->T / p->L;
	range_sub( p->i, pr, m, 1, false, &p->m0, &p->m1 );
	range_sub( p->j, pc, n, 1, false, &p->n0, &p->n1 );
	p->na = p->T / p->pc; p->nb = p->T / p->pr;                   // T is a multiple of both
	return kSuccess;
}

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------
// Minimal declarations (nccl.h 2.18+: ncclUniqueId is 128 bytes; ncclDataType_t ncclUint8 = 1; ncclResult_t 0 = success).
struct NcclUid { char b[128]; };
struct NcclApi
{
	void* lib = nullptr;
	int ( *GetUniqueId )( void* ) = nullptr;
	int ( *CommInitRank )( void**, int, NcclUid /* ncclUniqueId, by value */, int ) = nullptr;
	int ( *CommSplit )( void*, int, int, void**, void* ) = nullptr;
	int ( *CommDestroy )( void* ) = nullptr;
	int ( *AllGather )( const void*, void*, size_t, int, void*, cudaStream_t ) = nullptr;
	int ( *Broadcast )( const void*, void*, size_t, int, int, void*, cudaStream_t ) = nullptr;
	int ( *GroupStart )() = nullptr;
	int ( *GroupEnd )() = nullptr;
	int ( *GetVersion )( int* ) = nullptr;
	const char* ( *GetErrorString )( int ) = nullptr;
};

struct DistState
{
	std::mutex mu;
	NcclApi    nccl;
	bool       up = false;
	int        world = 1, rank = 0, nccl_version = 0;
	void*      comm = nullptr;
	// row / column communicators of the grid in use (re-split when a product picks another grid)
	int        pr = 0, pc = 0;
	void      *row = nullptr, *col = nullptr;
	cudaStream_t comm_stream = nullptr;
	// double-buffered receive areas of the k-panel gathers, grown on demand and kept across calls
	void*      abuf[2] = { nullptr, nullptr }; size_t abytes = 0;
	void*      bbuf[2] = { nullptr, nullptr }; size_t bbytes = 0;
	cudaEvent_t gathered[2] = { nullptr, nullptr };            // recorded on comm_stream after the gathers into buffer b
	cudaEvent_t buf_free[2] = { nullptr, nullptr };            // recorded on the compute stream after the kernels reading buffer b
	cudaEvent_t inputs = nullptr, done = nullptr;
	bool       buf_used[2] = { false, false };
	int        ab_static = 0;                                  // b200_set_option("dist_ab_static", 1)
	// one-sided transport: shards registered with b200_dist_register are mapped into every rank (CUDA IPC) and the k
	// panels are PULLED by the copy engines over NVLink (cudaMemcpyAsync peer copies): no SM is taken from the DMMA kernel
	cudaStream_t comm_stream2 = nullptr;                       // B panels (A panels travel on comm_stream)
	cudaEvent_t  pulled2 = nullptr;
	struct Reg { const void *a, *b; std::vector<const char*> peer_a, peer_b; };
	std::vector<Reg> regs;
	struct Opened { int rank; cudaIpcMemHandle_t h; char* base; };
	std::vector<Opened> opened;                                // every peer allocation mapped so far (an allocation is opened once)
	int        last_transport = 0;                             // 0 none, 1 NCCL all-gather, 2 copy-engine gets
	// timing of the last b200_dist_gemm (events on the compute stream; b200_dist_last_wait_ms)
	std::vector<cudaEvent_t> ev_wait0, ev_wait1;
	int        last_steps = 0;
};
static DistState& dist() { static DistState d; return d; }

#define B200_NCCL( call ) \
	do { int r__ = ( call ); if ( r__ != 0 ) \
	     return ::b200::fail( "%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, r__, \
	                          dist().nccl.GetErrorString ? dist().nccl.GetErrorString( r__ ) : "?" ); } while ( 0 )

static int nccl_load()
{
	NcclApi& n = dist().nccl;
	if ( n.lib ) return kSuccess;
	const char* names[] = { "libnccl.so.2", "libnccl.so" };
	for ( const char* nm : names ) { n.lib = dlopen( nm, RTLD_NOW | RTLD_GLOBAL ); if ( n.lib ) break; }
	if ( !n.lib ) return fail( "b200_dist: libnccl.so.2 not found (%s); the multi-GPU path has no other transport", dlerror() );
	auto sym = [&]( const char* s ) { return dlsym( n.lib, s ); };
	*(void**)&n.GetUniqueId    = sym( "ncclGetUniqueId" );
	*(void**)&n.CommInitRank   = sym( "ncclCommInitRank" );
	*(void**)&n.CommSplit      = sym( "ncclCommSplit" );
	*(void**)&n.CommDestroy    = sym( "ncclCommDestroy" );
	*(void**)&n.AllGather      = sym( "ncclAllGather" );
	*(void**)&n.Broadcast      = sym( "ncclBroadcast" );
	*(void**)&n.GroupStart     = sym( "ncclGroupStart" );
	*(void**)&n.GroupEnd       = sym( "ncclGroupEnd" );
	*(void**)&n.GetVersion     = sym( "ncclGetVersion" );
	*(void**)&n.GetErrorString = sym( "ncclGetErrorString" );
	if ( !n.GetUniqueId || !n.CommInitRank || !n.CommSplit || !n.CommDestroy || !n.AllGather || !n.Broadcast || !n.GroupStart || !n.GroupEnd )
	{
		dlclose( n.lib ); n.lib = nullptr;
		return fail( "b200_dist: libnccl.so.2 lacks a required symbol (need NCCL >= 2.18 for ncclCommSplit)" );
	}
	return kSuccess;
}

static int dist_unique_id( void* id128 )
{
	if ( nccl_load() != kSuccess ) return kFailure;
	memset( id128, 0, B200_DIST_ID_BYTES );
	B200_NCCL( dist().nccl.GetUniqueId( id128 ) );
	return kSuccess;
}

static void dist_free_buffers( DistState& d )
{
	for ( int b = 0; b < 2; ++b )
	{
		if ( d.abuf[b] ) cudaFree( d.abuf[b] );
		if ( d.bbuf[b] ) cudaFree( d.bbuf[b] );
		d.abuf[b] = d.bbuf[b] = nullptr; d.buf_used[b] = false;
	}
	d.abytes = d.bbytes = 0;
}

static int dist_finalize()
{
	DistState& d = dist();
	std::lock_guard<std::mutex> lk( d.mu );
	if ( !d.up ) return kSuccess;
	cudaDeviceSynchronize();
	if ( d.row ) d.nccl.CommDestroy( d.row );
	if ( d.col ) d.nccl.CommDestroy( d.col );
	if ( d.comm ) d.nccl.CommDestroy( d.comm );
	d.row = d.col = d.comm = nullptr; d.pr = d.pc = 0;
	dist_free_buffers( d );
	for ( int b = 0; b < 2; ++b ) { cudaEventDestroy( d.gathered[b] ); cudaEventDestroy( d.buf_free[b] ); d.gathered[b] = d.buf_free[b] = nullptr; }
	cudaEventDestroy( d.inputs ); cudaEventDestroy( d.done ); d.inputs = d.done = nullptr;
	for ( auto e : d.ev_wait0 ) cudaEventDestroy( e );
	for ( auto e : d.ev_wait1 ) cudaEventDestroy( e );
	d.ev_wait0.clear(); d.ev_wait1.clear();
	for ( auto& o : d.opened ) cudaIpcCloseMemHandle( o.base );
	d.opened.clear(); d.regs.clear();
	if ( d.comm_stream2 ) { cudaStreamDestroy( d.comm_stream2 ); d.comm_stream2 = nullptr; }
	if ( d.pulled2 ) { cudaEventDestroy( d.pulled2 ); d.pulled2 = nullptr; }
	cudaStreamDestroy( d.comm_stream ); d.comm_stream = nullptr;
	d.up = false; d.world = 1; d.rank = 0;
	return kSuccess;
}

static int dist_init( int world, int rank, const void* id128 )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( world < 1 || rank < 0 || rank >= world || !id128 ) return fail( "b200_dist_init: bad arguments" );
	if ( nccl_load() != kSuccess ) return kFailure;
	DistState& d = dist();
	if ( d.up ) { if ( dist_finalize() != kSuccess ) return kFailure; }
	std::lock_guard<std::mutex> lk( d.mu );
	NcclUid uid; memcpy( uid.b, id128, 128 );
	B200_NCCL( d.nccl.CommInitRank( &d.comm, world, uid, rank ) );
	if ( d.nccl.GetVersion ) d.nccl.GetVersion( &d.nccl_version );
	int lo = 0, hi = 0;
	B200_CUDA( cudaDeviceGetStreamPriorityRange( &lo, &hi ) );
	// the gathers are short and on the critical path of the NEXT step only: give them priority over the persistent gemm CTAs
	B200_CUDA( cudaStreamCreateWithPriority( &d.comm_stream, cudaStreamNonBlocking, hi ) );
	B200_CUDA( cudaStreamCreateWithPriority( &d.comm_stream2, cudaStreamNonBlocking, hi ) );
	B200_CUDA( cudaEventCreateWithFlags( &d.pulled2, cudaEventDisableTiming ) );
	for ( int b = 0; b < 2; ++b )
	{
		B200_CUDA( cudaEventCreateWithFlags( &d.gathered[b], cudaEventDisableTiming ) );
		B200_CUDA( cudaEventCreateWithFlags( &d.buf_free[b], cudaEventDisableTiming ) );
	}
	B200_CUDA( cudaEventCreateWithFlags( &d.inputs, cudaEventDisableTiming ) );
	B200_CUDA( cudaEventCreateWithFlags( &d.done, cudaEventDisableTiming ) );
	d.world = world; d.rank = rank; d.up = true;
	return kSuccess;
}

// Row / column communicators of a Pr x Pc grid (collective: every rank calls it with the same grid).
static int dist_grid_comms( DistState& d, int pr, int pc )
{
	if ( d.pr == pr && d.pc == pc ) return kSuccess;
	if ( d.row ) { d.nccl.CommDestroy( d.row ); d.row = nullptr; }
	if ( d.col ) { d.nccl.CommDestroy( d.col ); d.col = nullptr; }
	const int i = d.rank / pc, j = d.rank % pc;
	B200_NCCL( d.nccl.CommSplit( d.comm, i, j, &d.row, nullptr ) );       // grid row i: ranks ordered by column
	B200_NCCL( d.nccl.CommSplit( d.comm, pr + j, i, &d.col, nullptr ) );  // grid column j: ranks ordered by row
	d.pr = pr; d.pc = pc;
	return kSuccess;
}

static int dist_grow( DistState& d, size_t abytes, size_t bbytes )
{
	if ( abytes <= d.abytes && bbytes <= d.bbytes ) return kSuccess;
	B200_CUDA( cudaDeviceSynchronize() );
	const size_t na = std::max( abytes, d.abytes ), nb = std::max( bbytes, d.bbytes );
	dist_free_buffers( d );
	for ( int b = 0; b < 2; ++b )
	{
		B200_CUDA( cudaMalloc( &d.abuf[b], na ) );
		B200_CUDA( cudaMalloc( &d.bbuf[b], nb ) );
	}
	d.abytes = na; d.bbytes = nb;
	return kSuccess;
}

// ---- registration of static shards for one-sided gets (collective over all ranks) ------------------------------------
// Every rank passes its A and B shards (device memory from cudaMalloc, e.g. a torch tensor); the CUDA IPC handles of the
// allocations behind them travel through the communicator, every rank maps every other rank's shards, and from then on
// b200_dist_gemm( ..., B200_DIST_AB_STATIC ) on exactly these pointers moves the k panels with peer copies issued by the
// RECEIVER (the copy engines pull over NVLink; nothing runs on an SM and nothing is asked of the owner).  The contract of
// B200_DIST_AB_STATIC is what makes a one-sided get legal: the shards are complete before the call and are not written
// while products that use them are in flight.  Fails (on every rank alike) when the memory cannot be shared between the
// processes; the caller then simply keeps the NCCL transport.
typedef CUresult ( *MemGetAddressRangeFn )( CUdeviceptr*, size_t*, CUdeviceptr );
static int dist_register( const void* a_loc, const void* b_loc )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	DistState& d = dist();
	if ( !d.up ) return fail( "b200_dist_register: call b200_dist_init first" );
	std::lock_guard<std::mutex> lk( d.mu );
	for ( auto& r : d.regs ) if ( r.a == a_loc && r.b == b_loc ) return kSuccess;
	struct Rec { cudaIpcMemHandle_t h[2]; unsigned long long off[2]; int ok; int pad; };
	static_assert( sizeof( Rec ) % 8 == 0, "record size" );
	Rec mine; memset( &mine, 0, sizeof( mine ) ); mine.ok = 1;
	static MemGetAddressRangeFn range_fn = nullptr;
	if ( !range_fn )
	{
		void* fp = nullptr; cudaDriverEntryPointQueryResult q;
		if ( cudaGetDriverEntryPoint( "cuMemGetAddressRange", &fp, cudaEnableDefault, &q ) == cudaSuccess && q == cudaDriverEntryPointSuccess )
			range_fn = (MemGetAddressRangeFn)fp;
	}
	const void* ptrs[2] = { a_loc, b_loc };
	for ( int x = 0; x < 2 && mine.ok; ++x )
	{
		CUdeviceptr base = 0; size_t size = 0;
		if ( !range_fn || classify( ptrs[x] ) != MemKind::Device || range_fn( &base, &size, (CUdeviceptr)ptrs[x] ) != CUDA_SUCCESS ) { mine.ok = 0; break; }
		if ( cudaIpcGetMemHandle( &mine.h[x], (void*)base ) != cudaSuccess ) { cudaGetLastError(); mine.ok = 0; break; }
		mine.off[x] = (unsigned long long)( (CUdeviceptr)ptrs[x] - base );
	}
	// all records to all ranks (through the communicator; this is a set-up call, so it may synchronise)
	Rec* dev = nullptr;
	std::vector<Rec> all( d.world );
	B200_CUDA( cudaMalloc( &dev, sizeof( Rec ) * ( d.world + 1 ) ) );
	B200_CUDA( cudaMemcpyAsync( dev + d.world, &mine, sizeof( Rec ), cudaMemcpyHostToDevice, d.comm_stream ) );
	B200_NCCL( d.nccl.AllGather( dev + d.world, dev, sizeof( Rec ), 1, d.comm, d.comm_stream ) );
	B200_CUDA( cudaMemcpyAsync( all.data(), dev, sizeof( Rec ) * d.world, cudaMemcpyDeviceToHost, d.comm_stream ) );
	B200_CUDA( cudaStreamSynchronize( d.comm_stream ) );
	int ok = 1;
	for ( auto& r : all ) ok &= r.ok;
	DistState::Reg reg; reg.a = a_loc; reg.b = b_loc;
	reg.peer_a.assign( d.world, nullptr ); reg.peer_b.assign( d.world, nullptr );
	for ( int r = 0; r < d.world && ok; ++r )
	{
		if ( r == d.rank ) { reg.peer_a[r] = (const char*)a_loc; reg.peer_b[r] = (const char*)b_loc; continue; }
		for ( int x = 0; x < 2 && ok; ++x )
		{
			char* base = nullptr;
			for ( auto& o : d.opened ) if ( o.rank == r && !memcmp( &o.h, &all[r].h[x], sizeof( cudaIpcMemHandle_t ) ) ) base = o.base;
			if ( !base )
			{
				void* vp = nullptr;
				if ( cudaIpcOpenMemHandle( &vp, all[r].h[x], cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess ) { cudaGetLastError(); ok = 0; break; }
				base = (char*)vp;
				d.opened.push_back( { r, all[r].h[x], base } );
			}
			( x == 0 ? reg.peer_a : reg.peer_b )[r] = base + all[r].off[x];
		}
	}
	// agree on the outcome: one rank that could not map a peer makes everybody stay on NCCL
	mine.ok = ok;
	B200_CUDA( cudaMemcpyAsync( dev + d.world, &mine, sizeof( Rec ), cudaMemcpyHostToDevice, d.comm_stream ) );
	B200_NCCL( d.nccl.AllGather( dev + d.world, dev, sizeof( Rec ), 1, d.comm, d.comm_stream ) );
	B200_CUDA( cudaMemcpyAsync( all.data(), dev, sizeof( Rec ) * d.world, cudaMemcpyDeviceToHost, d.comm_stream ) );
	B200_CUDA( cudaStreamSynchronize( d.comm_stream ) );
	cudaFree( dev );
	for ( auto& r : all ) ok &= r.ok;
	if ( !ok ) return fail( "b200_dist_register: the shards cannot be shared between the ranks (CUDA IPC); NCCL transport stays in use" );
	d.regs.push_back( std::move( reg ) );
	return kSuccess;
}
static int dist_unregister( const void* a_loc, const void* b_loc )
{
	DistState& d = dist();
	std::lock_guard<std::mutex> lk( d.mu );
	if ( d.up ) cudaDeviceSynchronize();
	for ( size_t x = 0; x < d.regs.size(); ++x )
		if ( d.regs[x].a == a_loc && d.regs[x].b == b_loc ) { d.regs.erase( d.regs.begin() + x ); return kSuccess; }
	return kSuccess;
}

// ---- gemm on a Pr x Pc grid -----------------------------------------------------------------------------------------
// Shard layout (all column-major, device resident):
//   a_loc  my k panels of A's row block i, one after the other: panel q (global panel t = q*Pc + j) is the m_loc x kb
//          matrix at a_loc + q*kb*m_loc, leading dimension m_loc
//   b_loc  my k panels of B's column block j: panel q (t = q*Pr + i) is the kb x n_loc matrix at b_loc + q*kb*n_loc,
//          leading dimension kb
//   c_loc  my block C_ij, m_loc x n_loc with strides (rs_c, cs_c)
template <typename T>
static int dist_gemm( int64_t m, int64_t n, int64_t k, int64_t kb, const T* alpha, const T* a_loc, const T* b_loc,
                      const T* beta, T* c_loc, int64_t rs_c, int64_t cs_c, int flags )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	DistState& d = dist();
	if ( !d.up ) return fail( "b200_dist_gemm: call b200_dist_init first" );
	std::lock_guard<std::mutex> lk( d.mu );
	b200_dist_plan_t p;
	if ( dist_plan( d.world, d.rank, m, n, k, kb, &p ) != kSuccess ) return kFailure;
	if ( p.L > 8 ) return fail( "b200_dist_gemm: lcm(Pr, Pc) = %d panels per step (at most 8)", p.L );
	const int64_t m_loc = p.m1 - p.m0, n_loc = p.n1 - p.n0;
	if ( dist_grid_comms( d, p.pr, p.pc ) != kSuccess ) return kFailure;
	if ( m_loc == 0 || n_loc == 0 )
		return fail( "b200_dist_gemm: empty block of C on rank %d (fewer rows/columns than grid rows/columns)", d.rank );
	for ( const void* q : { (const void*)a_loc, (const void*)b_loc, (const void*)c_loc } )
		if ( classify( q ) != MemKind::Device ) return fail( "b200_dist_gemm: shards must be device resident" );
	cudaStream_t st = cur_stream();
	const int qa = p.L / p.pc, qb = p.L / p.pr;                               // panels I contribute per step
	const size_t a_panel = (size_t)kb * m_loc, b_panel = (size_t)kb * n_loc;   // elements
	// every rank of my grid row has the same m_loc (same i), every rank of my grid column the same n_loc (same j)
	// a group of ONE rank needs no gather at all: the kernels read that operand's panels straight from the caller's shard
	const bool gather_a = ( p.pc > 1 ), gather_b = ( p.pr > 1 );
	if ( dist_grow( d, gather_a ? (size_t)p.L * a_panel * sizeof(T) : 0, gather_b ? (size_t)p.L * b_panel * sizeof(T) : 0 ) != kSuccess ) return kFailure;

	const bool ab_static = ( flags & B200_DIST_AB_STATIC ) || d.ab_static;
	if ( !ab_static )
	{
		// the shards may be the output of work already queued on the caller's stream
		B200_CUDA( cudaEventRecord( d.inputs, st ) );
		B200_CUDA( cudaStreamWaitEvent( d.comm_stream, d.inputs, 0 ) );
	}
	const DistState::Reg* reg = nullptr;
	if ( ab_static ) for ( auto& r : d.regs ) if ( r.a == a_loc && r.b == b_loc ) reg = &r;
	d.last_transport = ( gather_a || gather_b ) ? ( reg ? 2 : 1 ) : 0;
	auto start = [&]( int s ) -> int
	{
		const int bf = s & 1;
		if ( !gather_a && !gather_b ) return kSuccess;            // one rank: nothing to move
		if ( d.buf_used[bf] ) B200_CUDA( cudaStreamWaitEvent( d.comm_stream, d.buf_free[bf], 0 ) );
		if ( reg )
		{
			// one-sided: I pull every group member's panels of this step into my receive buffer (same slot layout as the
			// all-gather); A panels on comm_stream, B panels on comm_stream2, so two copy engines work at once
			if ( gather_b )
			{
				if ( d.buf_used[bf] ) B200_CUDA( cudaStreamWaitEvent( d.comm_stream2, d.buf_free[bf], 0 ) );
				for ( int ii = 0; ii < p.pr; ++ii )
					B200_CUDA( cudaMemcpyAsync( (char*)d.bbuf[bf] + (size_t)ii * qb * b_panel * sizeof(T),
					                            reg->peer_b[ii * p.pc + p.j] + (size_t)s * qb * b_panel * sizeof(T),
					                            (size_t)qb * b_panel * sizeof(T), cudaMemcpyDeviceToDevice, d.comm_stream2 ) );
				B200_CUDA( cudaEventRecord( d.pulled2, d.comm_stream2 ) );
			}
			if ( gather_a )
				for ( int jj = 0; jj < p.pc; ++jj )
					B200_CUDA( cudaMemcpyAsync( (char*)d.abuf[bf] + (size_t)jj * qa * a_panel * sizeof(T),
					                            reg->peer_a[p.i * p.pc + jj] + (size_t)s * qa * a_panel * sizeof(T),
					                            (size_t)qa * a_panel * sizeof(T), cudaMemcpyDeviceToDevice, d.comm_stream ) );
			if ( gather_b ) B200_CUDA( cudaStreamWaitEvent( d.comm_stream, d.pulled2, 0 ) );
			B200_CUDA( cudaEventRecord( d.gathered[bf], d.comm_stream ) );
			d.buf_used[bf] = true;
			return kSuccess;
		}
		// receive layout: [source rank in group][its q-th panel of this step] -> slot = src*q_per_rank + q
		B200_NCCL( d.nccl.GroupStart() );
		if ( gather_a ) B200_NCCL( d.nccl.AllGather( a_loc + (size_t)s * qa * a_panel, d.abuf[bf], (size_t)qa * a_panel * sizeof(T), /*ncclUint8*/ 1, d.row, d.comm_stream ) );
		if ( gather_b ) B200_NCCL( d.nccl.AllGather( b_loc + (size_t)s * qb * b_panel, d.bbuf[bf], (size_t)qb * b_panel * sizeof(T), 1, d.col, d.comm_stream ) );
		B200_NCCL( d.nccl.GroupEnd() );
		B200_CUDA( cudaEventRecord( d.gathered[bf], d.comm_stream ) );
		d.buf_used[bf] = true;
		return kSuccess;
	};
	while ( (int)d.ev_wait0.size() < p.steps )
	{
		cudaEvent_t e0, e1;
		B200_CUDA( cudaEventCreate( &e0 ) ); B200_CUDA( cudaEventCreate( &e1 ) );
		d.ev_wait0.push_back( e0 ); d.ev_wait1.push_back( e1 );
	}
	d.last_steps = p.steps;
	const bool trace = ( flags & B200_DIST_TRACE ) != 0;
	if ( start( 0 ) != kSuccess ) return kFailure;
	if ( p.steps > 1 && start( 1 ) != kSuccess ) return kFailure;
	const T one = Scalar<T>::make( 1.0, 0.0 );
	for ( int s = 0; s < p.steps; ++s )
	{
		const int bf = s & 1;
		if ( trace ) B200_CUDA( cudaEventRecord( d.ev_wait0[s], st ) );
		if ( gather_a || gather_b ) B200_CUDA( cudaStreamWaitEvent( st, d.gathered[bf], 0 ) );
		if ( trace ) B200_CUDA( cudaEventRecord( d.ev_wait1[s], st ) );
		// global panel t = s*L + l: A from grid column t % Pc (its (l / Pc)-th panel of the step), B from grid row t % Pr
		const T* ap[8]; const T* bp[8];
		for ( int l = 0; l < p.L; ++l )
		{
			const int t = s * p.L + l;
			ap[l] = gather_a ? (const T*)d.abuf[bf] + ( (size_t)( t % p.pc ) * qa + (size_t)( l / p.pc ) ) * a_panel : a_loc + (size_t)t * a_panel;
			bp[l] = gather_b ? (const T*)d.bbuf[bf] + ( (size_t)( t % p.pr ) * qb + (size_t)( l / p.pr ) ) * b_panel : b_loc + (size_t)t * b_panel;
		}
		if ( gemm_dev<T>( false, false, m_loc, n_loc, kb, *alpha, ap[0], 1, m_loc, bp[0], 1, kb, s == 0 ? *beta : one,
		                  c_loc, rs_c, cs_c, st, p.L, ap + 1, bp + 1 ) != kSuccess ) return kFailure;
		if ( gather_a || gather_b ) B200_CUDA( cudaEventRecord( d.buf_free[bf], st ) );
		if ( s + 2 < p.steps && start( s + 2 ) != kSuccess ) return kFailure;
	}
	return kSuccess;
}

// Time the compute stream spent waiting for gathers in the last traced b200_dist_gemm (B200_DIST_TRACE), summed over
// its steps; synchronises.  < 0: nothing traced.
static double dist_last_wait_ms()
{
	DistState& d = dist();
	std::lock_guard<std::mutex> lk( d.mu );
	if ( !d.up || d.last_steps == 0 ) return -1.0;
	if ( cudaDeviceSynchronize() != cudaSuccess ) return -1.0;
	double tot = 0.0;
	for ( int s = 0; s < d.last_steps; ++s )
	{
		float ms = 0.f;
		if ( cudaEventElapsedTime( &ms, d.ev_wait0[s], d.ev_wait1[s] ) != cudaSuccess ) { cudaGetLastError(); return -1.0; }
		tot += ms;
	}
	return tot;
}

// ---- skinny gemm: 1-D split (the shapes bli_gemmsup serves, frame/3/bli_l3_sup.c:37-135) ------------
//   split = B200_DIST_COLS: rank r owns columns [n0, n1) = range_sub(r, world, n, bf=128) of C and of B; A (m x k, the
//           small operand: column-major, leading dimension lda) is broadcast from `root` into every rank's `a`
//   split = B200_DIST_ROWS: rank r owns rows [m0, m1) of C and of A; B (k x n, column-major, ldb) is broadcast
// root < 0: the small operand is already replicated (no collective at all).  c_loc/the local operand: any strides.
template <typename T>
static int dist_gemm_1d( int split, int root, int64_t m, int64_t n, int64_t k, const T* alpha,
                         T* a, int64_t rs_a, int64_t cs_a, T* b, int64_t rs_b, int64_t cs_b,
                         const T* beta, T* c_loc, int64_t rs_c, int64_t cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	DistState& d = dist();
	if ( !d.up ) return fail( "b200_dist_gemm_1d: call b200_dist_init first" );
	std::lock_guard<std::mutex> lk( d.mu );
	if ( split != B200_DIST_COLS && split != B200_DIST_ROWS ) return fail( "b200_dist_gemm_1d: split must be B200_DIST_COLS or B200_DIST_ROWS" );
	if ( root >= d.world ) return fail( "b200_dist_gemm_1d: root %d of %d ranks", root, d.world );
	int64_t lo, hi;
	range_sub( d.rank, d.world, split == B200_DIST_COLS ? n : m, 128, false, &lo, &hi );
	cudaStream_t st = cur_stream();
	T* small = ( split == B200_DIST_COLS ) ? a : b;
	if ( root >= 0 )
	{
		// dense footprint of the small operand (column-major with leading dimension cs)
		const int64_t rows = ( split == B200_DIST_COLS ) ? m : k, cols = ( split == B200_DIST_COLS ) ? k : n;
		const int64_t rs = ( split == B200_DIST_COLS ) ? rs_a : rs_b, cs = ( split == B200_DIST_COLS ) ? cs_a : cs_b;
		if ( rs != 1 || cs < rows ) return fail( "b200_dist_gemm_1d: the broadcast operand must be column-major" );
		if ( classify( small ) != MemKind::Device ) return fail( "b200_dist_gemm_1d: the broadcast operand must be device resident" );
		const size_t bytes = ( (size_t)cs * ( cols - 1 ) + rows ) * sizeof(T);
		// the broadcast runs on the caller's stream: it is the first thing the product needs (8 MB at n = 16384, k = 64)
		B200_NCCL( d.nccl.Broadcast( small, small, bytes, 1, root, d.comm, st ) );
	}
	if ( hi <= lo ) return kSuccess;
	if ( split == B200_DIST_COLS )
		return gemm_dev<T>( false, false, m, hi - lo, k, *alpha, a, rs_a, cs_a, b, rs_b, cs_b, *beta, c_loc, rs_c, cs_c, st );
	return gemm_dev<T>( false, false, hi - lo, n, k, *alpha, a, rs_a, cs_a, b, rs_b, cs_b, *beta, c_loc, rs_c, cs_c, st );
}

// ---- trsm: column blocks of B (left side) / row blocks (right side), A replicated --------------------------------------
// b_loc is this rank's block: columns range_sub(rank, world, n, 128) of B for side = left (m x n_loc), rows
// range_sub(rank, world, m, 128) for side = right (m_loc x n).  root >= 0: A (column-major, dense leading dimension
// cs_a, device resident on every rank) is first broadcast from that rank; root < 0: A is already replicated.
template <typename T>
static int dist_trsm( int side, int uplo, int transa, int diag, int root, int64_t m, int64_t n, const T* alpha,
                      T* a, int64_t rs_a, int64_t cs_a, T* b_loc, int64_t rs_b, int64_t cs_b )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	DistState& d = dist();
	if ( !d.up ) return fail( "b200_dist_trsm: call b200_dist_init first" );
	int64_t lo, hi;
	{
		std::lock_guard<std::mutex> lk( d.mu );
		if ( root >= d.world ) return fail( "b200_dist_trsm: root %d of %d ranks", root, d.world );
		range_sub( d.rank, d.world, side == B200_LEFT ? n : m, 128, false, &lo, &hi );
		if ( root >= 0 )
		{
			const int64_t ma = ( side == B200_LEFT ) ? m : n;
			if ( rs_a != 1 || cs_a < ma ) return fail( "b200_dist_trsm: a broadcast A must be column-major" );
			if ( classify( a ) != MemKind::Device ) return fail( "b200_dist_trsm: a broadcast A must be device resident" );
			B200_NCCL( d.nccl.Broadcast( a, a, ( (size_t)cs_a * ( ma - 1 ) + ma ) * sizeof(T), 1, root, d.comm, cur_stream() ) );
		}
	}
	if ( hi <= lo ) return kSuccess;
	if ( side == B200_LEFT ) return trsm_front<T>( side, uplo, transa, diag, m, hi - lo, alpha, a, rs_a, cs_a, b_loc, rs_b, cs_b );
	return trsm_front<T>( side, uplo, transa, diag, hi - lo, n, alpha, a, rs_a, cs_a, b_loc, rs_b, cs_b );
}

} // namespace b200
