// gemm_launch.cuh -- host-side launchers of the gemm kernels (shared by the per-datatype translation units
// gemm_{d,z,s,c}.cu, which are compiled in parallel) and the kernel-selection entry point used by capi.cu.
#pragma once
#include "context.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_ws.cuh"
#include "gemm_dmma_tma.cuh"
#include "gemm_dmma_pp.cuh"
#include "gemm_ffma_tma.cuh"
#include "gemm_cfma_tma.cuh"
#include "gemm_zmma_tma.cuh"
#include "gemm_ffma.cuh"
#include "gemm_ffma_ws.cuh"
#include "../../include/blis_b200.h"
#include <algorithm>
#include <numeric>
#include <string>
#include <type_traits>
#include <utility>

namespace b200 {

// Picks the kernel for one normalised problem (gemm_dmma.cuh: GemmArgs) and launches it on `st`.
// One explicit specialisation per datatype, each in its own translation unit.
template <typename T>
int launch_gemm_kernel( GemmArgs<T>& g, bool xk, bool yk, bool al, cudaStream_t st );
template <> int launch_gemm_kernel<double> ( GemmArgs<double>&  g, bool xk, bool yk, bool al, cudaStream_t st );
template <> int launch_gemm_kernel<double2>( GemmArgs<double2>& g, bool xk, bool yk, bool al, cudaStream_t st );
template <> int launch_gemm_kernel<float>  ( GemmArgs<float>&   g, bool xk, bool yk, bool al, cudaStream_t st );
template <> int launch_gemm_kernel<float2> ( GemmArgs<float2>&  g, bool xk, bool yk, bool al, cudaStream_t st );

// printf-style kernel name with static lifetime (one per template instantiation of the calling lambda)
static std::string kfmt( const char* fmt, ... )
{
	char buf[160]; va_list ap; va_start( ap, fmt ); vsnprintf( buf, sizeof( buf ), fmt, ap ); va_end( ap );
	return std::string( buf );
}
template <typename T> static const char* tname()
{
	return std::is_same<T, double>::value ? "double" : std::is_same<T, float>::value ? "float" : std::is_same<T, float2>::value ? "float2" : "double2";
}

template <typename KernT>
static int set_smem( KernT kern, int bytes )
{
	B200_CUDA( cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes ) );
	return kSuccess;
}

#ifdef B200_GEMM_LAUNCHERS      // defined by gemm_{d,z,s,c}.cu only: capi.cu does not instantiate any gemm kernel

// ---- kernel launchers -------------------------------------------------------------

template <typename T, int BP, int BQ, int BK, int WP, int WQ, int ST, bool TRI = false>
static int launch_dmma_ws( const GemmArgs<T>& g, bool xk, bool yk, bool al, int grid, cudaStream_t st )
{
	auto go = [&]( auto XKc, auto YKc, auto ALc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value, AL = decltype( ALc )::value;
		using Cfg = DmmaWsCfg<T, BP, BQ, BK, WP, WQ, ST, XK, YK, AL>;
		auto kern = gemm_dmma_ws_kernel<T, BP, BQ, BK, WP, WQ, ST, XK, YK, AL, TRI>;
		static const std::string kname = kfmt( "gemm_dmma_ws_kernel<%s,%dx%dx%d,%dst,XK=%d,YK=%d,AL=%d,TRI=%d>", tname<T>(), BP, BQ, BK, ST, XK, YK, AL, TRI );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, Cfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, Cfg::NT_ALL, Cfg::SMEM_BYTES, st>>>( g );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	const int sel = ( xk ? 4 : 0 ) | ( yk ? 2 : 0 ) | ( al ? 1 : 0 );
	switch ( sel )
	{
		case 0: return go( Ff{}, Ff{}, Ff{} );  case 1: return go( Ff{}, Ff{}, Tt{} );
		case 2: return go( Ff{}, Tt{}, Ff{} );  case 3: return go( Ff{}, Tt{}, Tt{} );
		case 4: return go( Tt{}, Ff{}, Ff{} );  case 5: return go( Tt{}, Ff{}, Tt{} );
		case 6: return go( Tt{}, Tt{}, Ff{} );  default: return go( Tt{}, Tt{}, Tt{} );
	}
}

// ---- TMA path (dgemm, 16-byte aligned operands) ---------------------------------------------
typedef CUresult ( *EncodeTiledFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
static EncodeTiledFn encode_tiled_fn()
{
	static EncodeTiledFn fn = nullptr;
	if ( !fn )
	{
		void* p = nullptr; cudaDriverEntryPointQueryResult q;
		if ( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q ) == cudaSuccess && q == cudaDriverEntryPointSuccess )
			fn = (EncodeTiledFn)p;
	}
	return fn;
}
// Tensor map of an operand with `rows` rows: k-contiguous (element (r,k) at base[r*ld + k]) -> dims {K, rows}, box {16, 128};
// row-contiguous (element (r,k) at base[k*ld + r]) -> dims {rows, K}, box {16, 16}.  128-byte swizzle, zero fill out of bounds.
// (float: the same with 32-element = 128-byte box rows: box {32, 128} resp. {32, 32}.)
static int make_tmap( CUtensorMap* tm, const void* base, size_t es, bool kmajor, int64_t rows, int64_t K, int64_t ld, int box_rows = 128 )
{
	EncodeTiledFn enc = encode_tiled_fn();
	if ( !enc ) return fail( "cuTensorMapEncodeTiled not available" );
	const cuuint32_t inner = (cuuint32_t)( 128 / es );
	cuuint64_t dims[2]    = { (cuuint64_t)( kmajor ? K : rows ), (cuuint64_t)( kmajor ? rows : K ) };
	cuuint64_t strides[1] = { (cuuint64_t)ld * es };
	cuuint32_t box[2]     = { inner, (cuuint32_t)( kmajor ? box_rows : inner ) };
	cuuint32_t estr[2]    = { 1, 1 };
	const CUresult r = enc( tm, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
	                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)ctx().tma_l2_promotion,
	                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
	if ( r != CUDA_SUCCESS ) return fail( "cuTensorMapEncodeTiled failed (%d)", (int)r );
	return kSuccess;
}
// k panels of one operand as slots of ONE 3-D tensor map: all panel pointers must sit at base + slot*stride with a
// 16-byte aligned stride (true for the gather buffers of the multi-GPU path and for panels cut out of one matrix).
template <typename T>
static bool seg_slots( const T* p0, const T* const* more, int nseg, const T** base, int64_t* stride_bytes, int* slots )
{
	const T* lo = p0; int64_t S = 0;
	for ( int s = 1; s < nseg; ++s ) lo = std::min( lo, more[s - 1] );
	for ( int s = 0; s < nseg; ++s )
	{
		const int64_t d = (const char*)( s ? more[s - 1] : p0 ) - (const char*)lo;
		S = std::gcd( S, d );                                      // gcd( 0, d ) = d
	}
	if ( S == 0 ) S = 16;                                        // all panels identical
	if ( S % 16 != 0 || S >= ( 1ll << 40 ) ) return false;
	for ( int s = 0; s < nseg; ++s )
	{
		const int64_t d = (const char*)( s ? more[s - 1] : p0 ) - (const char*)lo;
		if ( d % S != 0 || d / S >= ( 1 << 20 ) ) return false;
		slots[s] = (int)( d / S );
	}
	*base = lo; *stride_bytes = S;
	return true;
}
// 3-D variant of make_tmap: dims {.., .., nslots}, box {.., .., 1}; the shared-memory image of a box equals the 2-D one.
static int make_tmap3( CUtensorMap* tm, const void* base, bool kmajor, int64_t rows, int64_t K, int64_t ld, int64_t slot_bytes, int nslots )
{
	EncodeTiledFn enc = encode_tiled_fn();
	if ( !enc ) return fail( "cuTensorMapEncodeTiled not available" );
	cuuint64_t dims[3]    = { (cuuint64_t)( kmajor ? K : rows ), (cuuint64_t)( kmajor ? rows : K ), (cuuint64_t)nslots };
	cuuint64_t strides[2] = { (cuuint64_t)ld * 8, (cuuint64_t)slot_bytes };
	cuuint32_t box[3]     = { 16, (cuuint32_t)( kmajor ? 128 : 16 ), 1 };
	cuuint32_t estr[3]    = { 1, 1, 1 };
	const CUresult r = enc( tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr,
	                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)ctx().tma_l2_promotion,
	                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
	if ( r != CUDA_SUCCESS ) return fail( "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r );
	return kSuccess;
}
// D tile stages of the dgemm CST path: dims {Q, P}, box {8 columns, 32 rows}, no swizzle (64-byte rows in shared memory).
static int make_tmap_d8( CUtensorMap* tm, const void* base, int64_t P, int64_t Q, int64_t ldd )
{
	EncodeTiledFn enc = encode_tiled_fn();
	if ( !enc ) return fail( "cuTensorMapEncodeTiled not available" );
	cuuint64_t dims[2]    = { (cuuint64_t)Q, (cuuint64_t)P };
	cuuint64_t strides[1] = { (cuuint64_t)ldd * 8 };
	cuuint32_t box[2]     = { 8, 32 };
	cuuint32_t estr[2]    = { 1, 1 };
	const CUresult r = enc( tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
	                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)ctx().tma_l2_promotion,
	                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
	if ( r != CUDA_SUCCESS ) return fail( "cuTensorMapEncodeTiled (D) failed (%d)", (int)r );
	return kSuccess;
}
template <typename T>
static bool tma_eligible( const GemmArgs<T>& g, bool xk, bool yk, bool al )
{
	// TMA needs 16-byte aligned bases and strides (== al), a leading dimension that covers the row, 32-bit box coordinates
	if ( g.nseg > 1 )
	{
		if ( !std::is_same<T, double>::value || g.tri || g.ktri ) return false;
		const T* b; int64_t S; int sl[8];
		if ( !seg_slots( g.X, g.Xseg, g.nseg, &b, &S, sl ) || !seg_slots( g.Y, g.Yseg, g.nseg, &b, &S, sl ) ) return false;
	}
	return al && g.P < ( 1ll << 31 ) && g.Q < ( 1ll << 31 ) && g.K < ( 1ll << 31 ) &&
	       g.ldx >= ( xk ? g.K : g.P ) && g.ldy >= ( yk ? g.K : g.Q ) && g.ldx * 8 < ( 1ll << 40 ) && g.ldy * 8 < ( 1ll << 40 );
}
template <bool TRI = false, bool CST = false, bool SK = false>
static int launch_dmma_tma( const GemmArgs<double>& g_in, bool xk, bool yk, int grid, cudaStream_t st )
{
	GemmArgs<double> g = g_in;
	CUtensorMap tmx, tmy, tmd;
	if ( g.nseg > 1 )
	{
		const double *bx, *by; int64_t sx, sy;
		if ( !seg_slots( g.X, g.Xseg, g.nseg, &bx, &sx, g.segx ) || !seg_slots( g.Y, g.Yseg, g.nseg, &by, &sy, g.segy ) )
			return fail( "k panels are not slots of one strided buffer" );
		int nx = 0, ny = 0;
		for ( int s = 0; s < g.nseg; ++s ) { nx = std::max( nx, g.segx[s] + 1 ); ny = std::max( ny, g.segy[s] + 1 ); }
		if ( make_tmap3( &tmx, bx, xk, g.P, g.K, g.ldx, sx, nx ) != kSuccess ) return kFailure;
		if ( make_tmap3( &tmy, by, yk, g.Q, g.K, g.ldy, sy, ny ) != kSuccess ) return kFailure;
	}
	else
	{
		if ( make_tmap( &tmx, g.X, 8, xk, g.P, g.K, g.ldx ) != kSuccess ) return kFailure;
		if ( make_tmap( &tmy, g.Y, 8, yk, g.Q, g.K, g.ldy ) != kSuccess ) return kFailure;
	}
	// D as un-swizzled {8 columns, 32 rows} boxes (CST: the epilogue works on D in shared memory); a copy of tmx when unused
	if ( CST ) { if ( make_tmap_d8( &tmd, g.D, g.P, g.Q, g.ldd ) != kSuccess ) return kFailure; }
	else tmd = tmx;
	auto go = [&]( auto XKc, auto YKc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value;
		auto kern = gemm_dmma_tma_kernel<XK, YK, TRI, CST, SK>;
		using KCfg = typename std::conditional<CST, DmmaTmaCfgCst, DmmaTmaCfg>::type;
		static const std::string kname = SK ? kfmt( "gemm_dmma_tma_kernel<XK=%d,YK=%d,TRI=%d,CST=%d,SK=1>", XK, YK, TRI, CST )
		                                    : kfmt( "gemm_dmma_tma_kernel<XK=%d,YK=%d,TRI=%d,CST=%d>", XK, YK, TRI, CST );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, KCfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, KCfg::NT_ALL, KCfg::SMEM_BYTES, st>>>( g, tmx, tmy, tmd );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	if ( xk ) return yk ? go( Tt{}, Tt{} ) : go( Tt{}, Ff{} );
	return yk ? go( Ff{}, Tt{} ) : go( Ff{}, Ff{} );
}

// small k: two consumer groups taking turns on the tensor pipe (gemm_dmma_pp.cuh); nseg == 1, full D
static int launch_dmma_pp( const GemmArgs<double>& g, bool xk, bool yk, int grid, cudaStream_t st )
{
	CUtensorMap tmx, tmy;
	if ( make_tmap( &tmx, g.X, 8, xk, g.P, g.K, g.ldx ) != kSuccess ) return kFailure;
	if ( make_tmap( &tmy, g.Y, 8, yk, g.Q, g.K, g.ldy, DmmaPpCfg::BQH ) != kSuccess ) return kFailure;
	auto go = [&]( auto XKc, auto YKc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value;
		auto kern = gemm_dmma_pp_kernel<XK, YK>;
		static const std::string kname = kfmt( "gemm_dmma_pp_kernel<XK=%d,YK=%d>", XK, YK );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, DmmaPpCfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, DmmaPpCfg::NT_ALL, DmmaPpCfg::SMEM_BYTES, st>>>( g, tmx, tmy );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	if ( xk ) return yk ? go( Tt{}, Tt{} ) : go( Tt{}, Ff{} );
	return yk ? go( Ff{}, Tt{} ) : go( Ff{}, Ff{} );
}

template <bool TRI = false, bool CST = false>
static int launch_ffma_tma( const GemmArgs<float>& g, bool xk, bool yk, int grid, cudaStream_t st )
{
	CUtensorMap tmx, tmy, tmd;
	if ( make_tmap( &tmx, g.X, 4, xk, g.P, g.K, g.ldx ) != kSuccess ) return kFailure;
	if ( make_tmap( &tmy, g.Y, 4, yk, g.Q, g.K, g.ldy ) != kSuccess ) return kFailure;
	// D as {32 columns, 64 rows} boxes (CST: the epilogue reads D from shared memory); a copy of tmx when unused
	if ( CST ) { if ( make_tmap( &tmd, g.D, 4, true, g.P, g.Q, g.ldd, 64 ) != kSuccess ) return kFailure; }
	else tmd = tmx;
	auto go = [&]( auto XKc, auto YKc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value;
		auto kern = gemm_ffma_tma_kernel<XK, YK, TRI, CST>;
		static const std::string kname = kfmt( "gemm_ffma_tma_kernel<XK=%d,YK=%d,TRI=%d,CST=%d>", XK, YK, TRI, CST );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, FfmaTmaCfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, FfmaTmaCfg::NT_ALL, FfmaTmaCfg::SMEM_BYTES, st>>>( g, tmx, tmy, tmd );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	if ( xk ) return yk ? go( Tt{}, Tt{} ) : go( Tt{}, Ff{} );
	return yk ? go( Ff{}, Tt{} ) : go( Ff{}, Ff{} );
}

template <bool TRI = false, bool CST = false>
static int launch_cfma_tma( const GemmArgs<float2>& g, bool xk, bool yk, int grid, cudaStream_t st )
{
	// float2 elements are moved as opaque 8-byte elements (FLOAT64-typed map; zero fill out of bounds)
	CUtensorMap tmx, tmy, tmd;
	if ( CST ) { if ( make_tmap( &tmd, g.D, 8, true, g.P, g.Q, g.ldd, 16 ) != kSuccess ) return kFailure; }
	if ( make_tmap( &tmx, g.X, 8, xk, g.P, g.K, g.ldx, CfmaTmaCfg::BP ) != kSuccess ) return kFailure;
	if ( make_tmap( &tmy, g.Y, 8, yk, g.Q, g.K, g.ldy, CfmaTmaCfg::BQ ) != kSuccess ) return kFailure;
	auto go = [&]( auto XKc, auto YKc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value;
		auto kern = gemm_cfma_tma_kernel<XK, YK, TRI, CST>;
		static const std::string kname = kfmt( "gemm_cfma_tma_kernel<XK=%d,YK=%d,TRI=%d,CST=%d>", XK, YK, TRI, CST );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, CfmaTmaCfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		if ( !CST ) tmd = tmx;
		kern<<<grid, CfmaTmaCfg::NT_ALL, CfmaTmaCfg::SMEM_BYTES, st>>>( g, tmx, tmy, tmd );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	if ( xk ) return yk ? go( Tt{}, Tt{} ) : go( Tt{}, Ff{} );
	return yk ? go( Ff{}, Tt{} ) : go( Ff{}, Ff{} );
}

// zgemm: FLOAT64-typed maps with two elements per complex number; k-contiguous: dims {2K, rows}, box {16, box_rows};
// row-contiguous: dims {2 rows, K}, box {16, 8} (eight complex rows x eight k lines).  128-byte swizzle, zero fill.
static int make_tmap_z( CUtensorMap* tm, const void* base, bool kmajor, int64_t rows, int64_t K, int64_t ld, int box_rows )
{
	EncodeTiledFn enc = encode_tiled_fn();
	if ( !enc ) return fail( "cuTensorMapEncodeTiled not available" );
	cuuint64_t dims[2]    = { (cuuint64_t)( kmajor ? 2 * K : 2 * rows ), (cuuint64_t)( kmajor ? rows : K ) };
	cuuint64_t strides[1] = { (cuuint64_t)ld * 16 };
	cuuint32_t box[2]     = { 16, (cuuint32_t)( kmajor ? box_rows : 8 ) };
	cuuint32_t estr[2]    = { 1, 1 };
	const CUresult r = enc( tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
	                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)ctx().tma_l2_promotion,
	                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
	if ( r != CUDA_SUCCESS ) return fail( "cuTensorMapEncodeTiled failed (%d)", (int)r );
	return kSuccess;
}
static bool tma_eligible_z( const GemmArgs<double2>& g, bool xk, bool yk )
{
	return g.nseg == 1 && ( (uintptr_t)g.X % 16 == 0 ) && ( (uintptr_t)g.Y % 16 == 0 ) &&
	       g.P < ( 1ll << 30 ) && g.Q < ( 1ll << 30 ) && g.K < ( 1ll << 30 ) &&
	       g.ldx >= ( xk ? g.K : g.P ) && g.ldy >= ( yk ? g.K : g.Q ) && g.ldx * 16 < ( 1ll << 40 ) && g.ldy * 16 < ( 1ll << 40 );
}
template <bool TRI = false>
static int launch_zmma_tma( const GemmArgs<double2>& g, bool xk, bool yk, int grid, cudaStream_t st )
{
	CUtensorMap tmx, tmy;
	if ( make_tmap_z( &tmx, g.X, xk, g.P, g.K, g.ldx, ZmmaTmaCfg::BP ) != kSuccess ) return kFailure;
	if ( make_tmap_z( &tmy, g.Y, yk, g.Q, g.K, g.ldy, ZmmaTmaCfg::BQ ) != kSuccess ) return kFailure;
	auto go = [&]( auto XKc, auto YKc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value;
		auto kern = gemm_zmma_tma_kernel<XK, YK, TRI>;
		static const std::string kname = kfmt( "gemm_zmma_tma_kernel<XK=%d,YK=%d,TRI=%d>", XK, YK, TRI );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, ZmmaTmaCfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, ZmmaTmaCfg::NT_ALL, ZmmaTmaCfg::SMEM_BYTES, st>>>( g, tmx, tmy );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	if ( xk ) return yk ? go( Tt{}, Tt{} ) : go( Tt{}, Ff{} );
	return yk ? go( Ff{}, Tt{} ) : go( Ff{}, Ff{} );
}

template <typename T, int BP, int BQ, int BK, int TP, int TQ, int ST>
static int launch_ffma( const GemmArgs<T>& g, bool xk, bool yk, bool al, int grid, cudaStream_t st )
{
	auto go = [&]( auto XKc, auto YKc, auto ALc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value, AL = decltype( ALc )::value;
		using Cfg = FfmaCfg<T, BP, BQ, BK, TP, TQ, ST>;
		auto kern = gemm_ffma_kernel<T, BP, BQ, BK, TP, TQ, ST, XK, YK, AL>;
		static const std::string kname = kfmt( "gemm_ffma_kernel<%s,%dx%dx%d,XK=%d,YK=%d,AL=%d>", tname<T>(), BP, BQ, BK, XK, YK, AL );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, Cfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, Cfg::NT, Cfg::SMEM_BYTES, st>>>( g );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	const int sel = ( xk ? 4 : 0 ) | ( yk ? 2 : 0 ) | ( al ? 1 : 0 );
	switch ( sel )
	{
		case 0: return go( Ff{}, Ff{}, Ff{} );  case 1: return go( Ff{}, Ff{}, Tt{} );
		case 2: return go( Ff{}, Tt{}, Ff{} );  case 3: return go( Ff{}, Tt{}, Tt{} );
		case 4: return go( Tt{}, Ff{}, Ff{} );  case 5: return go( Tt{}, Ff{}, Tt{} );
		case 6: return go( Tt{}, Tt{}, Ff{} );  default: return go( Tt{}, Tt{}, Tt{} );
	}
}

template <typename T, int BP, int BQ, int BK, int TP, int TQ, int ST>
static int launch_ffma_ws( const GemmArgs<T>& g, bool xk, bool yk, bool al, int grid, cudaStream_t st )
{
	auto go = [&]( auto XKc, auto YKc, auto ALc ) -> int
	{
		constexpr bool XK = decltype( XKc )::value, YK = decltype( YKc )::value, AL = decltype( ALc )::value;
		using Cfg = FfmaWsCfg<T, BP, BQ, BK, TP, TQ, ST>;
		auto kern = gemm_ffma_ws_kernel<T, BP, BQ, BK, TP, TQ, ST, XK, YK, AL>;
		static const std::string kname = kfmt( "gemm_ffma_ws_kernel<%s,%dx%dx%d,XK=%d,YK=%d,AL=%d>", tname<T>(), BP, BQ, BK, XK, YK, AL );
		static std::atomic<bool> attr{ false };
		if ( !attr ) { if ( set_smem( kern, Cfg::SMEM_BYTES ) != kSuccess ) return kFailure; attr = true; }
		kern<<<grid, Cfg::NT_ALL, Cfg::SMEM_BYTES, st>>>( g );
		B200_CUDA( cudaGetLastError() );
		note_launch( kname.c_str() );
		return kSuccess;
	};
	using Tt = std::true_type; using Ff = std::false_type;
	const int sel = ( xk ? 4 : 0 ) | ( yk ? 2 : 0 ) | ( al ? 1 : 0 );
	switch ( sel )
	{
		case 0: return go( Ff{}, Ff{}, Ff{} );  case 1: return go( Ff{}, Ff{}, Tt{} );
		case 2: return go( Ff{}, Tt{}, Ff{} );  case 3: return go( Ff{}, Tt{}, Tt{} );
		case 4: return go( Tt{}, Ff{}, Ff{} );  case 5: return go( Tt{}, Ff{}, Tt{} );
		case 6: return go( Tt{}, Tt{}, Ff{} );  default: return go( Tt{}, Tt{}, Tt{} );
	}
}

#endif // B200_GEMM_LAUNCHERS

} // namespace b200
