// capi.cu -- the extern "C" level-3 entry points and their host-side front ends.
//
// Host logic that mirrors the reference's object-API front ends:
//   bli_gemm_ex   frame/3/bli_l3_oapi_ex.c:48-148   (trivial cases, storage
//                 based operand swap as in bli_gemm_cntl.c:98-161)
//   bli_trsm_ex   frame/3/bli_l3_oapi_ex.c:692-801  (right side solved as the
//                 transposed left-side problem, :748-759)
//   bli_l3_return_early_if_trivial   frame/3/bli_l3_util.c:40-65
//   bli_trsm_blk_var1 (solve block, then rank-k update of the remaining rows)
//                 frame/3/trsm/bli_trsm_blk_var1.c:40-188
#include "gemm_launch.cuh"
#include "trsm.cuh"
#include <vector>

namespace b200 {

// ---- small helpers -------------------------------------------------------------
template <typename T> struct Scalar;
template <> struct Scalar<float>
{
	static float   make( double r, double ) { return (float)r; }
	static bool    is_zero( float a ) { return a == 0.0f; }
	static bool    is_one( float a )  { return a == 1.0f; }
};
template <> struct Scalar<double>
{
	static double  make( double r, double ) { return r; }
	static bool    is_zero( double a ) { return a == 0.0; }
	static bool    is_one( double a )  { return a == 1.0; }
};
template <> struct Scalar<float2>
{
	static float2  make( double r, double i ) { return make_float2( (float)r, (float)i ); }
	static bool    is_zero( float2 a ) { return a.x == 0.0f && a.y == 0.0f; }
	static bool    is_one( float2 a )  { return a.x == 1.0f && a.y == 0.0f; }
};
template <> struct Scalar<double2>
{
	static double2 make( double r, double i ) { return make_double2( r, i ); }
	static bool    is_zero( double2 a ) { return a.x == 0.0 && a.y == 0.0; }
	static bool    is_one( double2 a )  { return a.x == 1.0 && a.y == 0.0; }
};

// ---- strided copy / scale kernels (component-wise, so complex data only
//      needs the alignment of its real type) -----------------------------------
template <typename R, int NC>
__global__ void copy2d_kernel( R* __restrict__ dst, int64_t rsd, int64_t csd,
                               const R* __restrict__ src, int64_t rss, int64_t css,
                               int64_t m, int64_t n, int inner_is_row, int tri )
{
	const int64_t total = m * n;
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_is_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		if ( ( tri == 1 && i < j ) || ( tri == 2 && i > j ) ) continue;      // only the stored triangle (1: lower, 2: upper)
		const R* s = src + ( i * rss + j * css ) * NC;
		R*       d = dst + ( i * rsd + j * csd ) * NC;
		#pragma unroll
		for ( int c = 0; c < NC; ++c ) d[c] = s[c];
	}
}

// C := beta * C  (beta == 0 stores zeros without reading C: bli_scalm / bli_setm)
template <typename R, int NC>
__global__ void scal2d_kernel( R* __restrict__ c, int64_t rs, int64_t cs, int64_t m, int64_t n,
                               R br, R bi, int beta_is_zero, int inner_is_row, int tri )
{
	const int64_t total = m * n;
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_is_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		if ( ( tri == 1 && i < j ) || ( tri == 2 && i > j ) ) continue;      // only the stored triangle (1: lower, 2: upper)
		R* p = c + ( i * rs + j * cs ) * NC;
		if ( beta_is_zero ) { for ( int q = 0; q < NC; ++q ) p[q] = (R)0; }
		else if ( NC == 1 ) p[0] = br * p[0];
		else { const R xr = p[0], xi = p[1]; p[0] = br * xr - bi * xi; p[1] = br * xi + bi * xr; }
	}
}

static inline int64_t iabs64( int64_t x ) { return x < 0 ? -x : x; }

// dst[c*ldd + r] = src[r*lds + c]: 32 x 32 tiles through shared memory, both sides coalesced.
template <typename T>
__global__ void __launch_bounds__( 256 ) transpose2d_kernel( T* __restrict__ dst, int64_t ldd, const T* __restrict__ src, int64_t lds,
                                                             int64_t R, int64_t Cn, int tiles_c )
{
	__shared__ T tile[32][33];
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const int64_t ntiles = ( ( R + 31 ) / 32 ) * tiles_c;
	for ( int64_t t = blockIdx.x; t < ntiles; t += gridDim.x )
	{
		const int64_t r0 = ( t / tiles_c ) * 32, c0 = ( t % tiles_c ) * 32;
		#pragma unroll
		for ( int i = 0; i < 4; ++i )
		{
			const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
			if ( r < R && c < Cn ) tile[ty + 8 * i][tx] = src[r * lds + c];
		}
		__syncthreads();
		#pragma unroll
		for ( int i = 0; i < 4; ++i )
		{
			const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
			if ( r < R && c < Cn ) dst[c * ldd + r] = tile[tx][ty + 8 * i];
		}
		__syncthreads();
	}
}

template <typename T>
static int transpose2d( T* dst, int64_t ldd, const T* src, int64_t lds, int64_t R, int64_t Cn, cudaStream_t st )
{
	const int64_t tiles_c = ( Cn + 31 ) / 32, ntiles = ( ( R + 31 ) / 32 ) * tiles_c;
	if ( ntiles <= 0 ) return kSuccess;
	if ( tiles_c >= ( 1ll << 31 ) ) return fail( "transpose2d: matrix too wide" );
	const int blocks = (int)std::min<int64_t>( ntiles, (int64_t)ctx().num_sms * 32 );
	transpose2d_kernel<T><<<blocks, 256, 0, st>>>( dst, ldd, src, lds, R, Cn, (int)tiles_c );
	B200_CUDA( cudaGetLastError() );
	ctx().launches++;
	return kSuccess;
}

template <typename T>
static int copy2d( T* dst, int64_t rsd, int64_t csd, const T* src, int64_t rss, int64_t css,
                   int64_t m, int64_t n, cudaStream_t st, int uplo = 0 )
{
	if ( m <= 0 || n <= 0 ) return kSuccess;
	const int tri = ( uplo == B200_LOWER ) ? 1 : ( uplo == B200_UPPER ) ? 2 : 0;
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	const int inner_is_row = ( iabs64( rss ) + iabs64( rsd ) <= iabs64( css ) + iabs64( csd ) );
	const int64_t total = m * n;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	copy2d_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)dst, rsd, csd, (const R*)src, rss, css, m, n, inner_is_row, tri );
	B200_CUDA( cudaGetLastError() );
	ctx().launches++;
	return kSuccess;
}

template <typename T>
static int scal2d( T* c, int64_t rs, int64_t cs, int64_t m, int64_t n, T beta, cudaStream_t st, int uplo = 0 )
{
	if ( m <= 0 || n <= 0 || Scalar<T>::is_one( beta ) ) return kSuccess;
	const int tri = ( uplo == B200_LOWER ) ? 1 : ( uplo == B200_UPPER ) ? 2 : 0;
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	R br, bi;
	if constexpr ( Elem<T>::cplx ) { br = beta.x; bi = beta.y; } else { br = beta; bi = 0; }
	const int inner_is_row = ( iabs64( rs ) <= iabs64( cs ) );
	const int64_t total = m * n;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	scal2d_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)c, rs, cs, m, n, br, bi, Scalar<T>::is_zero( beta ) ? 1 : 0, inner_is_row, tri );
	B200_CUDA( cudaGetLastError() );
	ctx().launches++;
	return kSuccess;
}

// Tile shapes per datatype = the "blocksizes" this engine registers
// (MR/NR become the warp tile, MC/NC the CTA tile, KC the staged k slab).
template <typename T> struct Tiles;
template <> struct Tiles<double>  { static constexpr int BP = 128, BQ = 128, BK = 16, MR = 32, NR = 64; };
template <> struct Tiles<double2> { static constexpr int BP = 64,  BQ = 128, BK = 8,  MR = 32, NR = 32; };
template <> struct Tiles<float>   { static constexpr int BP = 128, BQ = 128, BK = 16, MR = 8,  NR = 8;  };
template <> struct Tiles<float2>  { static constexpr int BP = 64,  BQ = 128, BK = 16, MR = 4,  NR = 8;  };

enum { kTriA = 1, kTriB = 2, kTriLower = 4, kTriUpper = 8 };

// ---- gemm on device-resident strided views ------------------------------------
// C(m x n) := beta*C + alpha * A(m x k) * B(k x n); A/B views already carry any
// transposition in their strides; conja/conjb request conjugation.
template <typename T>
static int gemm_dev( bool conja, bool conjb, int64_t m, int64_t n, int64_t k, T alpha,
                     const T* a, int64_t rs_a, int64_t cs_a,
                     const T* b, int64_t rs_b, int64_t cs_b,
                     T beta, T* c, int64_t rs_c, int64_t cs_c, cudaStream_t st,
                     int nseg = 1, const T* const* a_more = nullptr, const T* const* b_more = nullptr,
                     int uplo_c = 0,      // 0: all of C; B200_LOWER / B200_UPPER: only that triangle of C is computed and stored
                     int tri_operand = 0 ) // 0: none; kTriA/kTriB | kTriLower/kTriUpper: that operand is (effectively) triangular with
                                           // explicit zeros on the other side -> tiles skip the k range that only multiplies zeros
{
	if ( m <= 0 || n <= 0 ) return kSuccess;
	// bli_l3_return_early_if_trivial: alpha == 0 or k == 0  ->  C := beta*C
	if ( k <= 0 || Scalar<T>::is_zero( alpha ) ) return scal2d( c, rs_c, cs_c, m, n, beta, st, uplo_c );

	constexpr size_t ES = sizeof(T);
	void *tmp_c = nullptr, *tmp_x = nullptr, *tmp_y = nullptr;
	int rc = kSuccess;

	// Complex element accesses in the kernels need natural alignment of T.
	const bool c_misaligned = ( (uintptr_t)c % ( Elem<T>::cplx ? ES : sizeof( typename Elem<T>::real ) ) ) != 0;

	// -- output: make it "q-contiguous" (D = C or D = C^T)
	T* cd = c; int64_t rs_cd = rs_c, cs_cd = cs_c;
	const bool c_general = !( ( rs_c == 1 && ( cs_c >= m || n == 1 ) ) || ( cs_c == 1 && ( rs_c >= n || m == 1 ) ) ) || c_misaligned;
	if ( c_general )
	{
		if ( dev_alloc( &tmp_c, (size_t)m * n * ES, st ) != kSuccess ) return kFailure;
		cd = (T*)tmp_c; rs_cd = 1; cs_cd = m;
		if ( !Scalar<T>::is_zero( beta ) ) rc = copy2d( cd, rs_cd, cs_cd, c, rs_c, cs_c, m, n, st, uplo_c );
	}

	GemmArgs<T> g;
	int64_t xs_p, xs_k, ys_k, ys_q;
	bool swapped = false;                              // column-stored C: X panels come from B, Y panels from A
	if ( rs_cd == 1 && !( cs_cd == 1 && m > 1 ) )
	{
		swapped = true;
		// column-stored C: D = C^T,  X = B^T (P = n),  Y = A^T (Q = m)
		g.P = n; g.Q = m; g.ldd = ( n == 1 ? m : cs_cd );
		g.X = b; xs_p = cs_b; xs_k = rs_b; g.conjx = conjb;
		g.Y = a; ys_k = cs_a; ys_q = rs_a; g.conjy = conja;
	}
	else
	{
		// row-stored C: D = C,  X = A (P = m),  Y = B (Q = n)
		g.P = m; g.Q = n; g.ldd = ( m == 1 ? n : rs_cd );
		g.X = a; xs_p = rs_a; xs_k = cs_a; g.conjx = conja;
		g.Y = b; ys_k = rs_b; ys_q = cs_b; g.conjy = conjb;
	}
	g.D = cd; g.K = k; g.alpha = alpha; g.beta = beta;
	g.beta_is_zero = Scalar<T>::is_zero( beta ) ? 1 : 0;
	g.nseg = nseg;
	// stored triangle in D coordinates: C lower = {i >= j}.  D = C: p = i, q = j -> q - p <= 0 (tri 1);
	// D = C^T: p = j, q = i -> q - p >= 0 (tri 2); upper is the mirror image.
	g.tri = 0; g.tri_off = 0;
	if ( uplo_c == B200_LOWER ) g.tri = swapped ? 2 : 1;
	if ( uplo_c == B200_UPPER ) g.tri = swapped ? 1 : 2;
	g.raster = ctx().raster_group;
	g.ktri = 0;
	if ( tri_operand && ctx().ktri_skip )
	{
		const bool on_a = ( tri_operand & kTriA ) != 0, lower = ( tri_operand & kTriLower ) != 0;
		// a(i,l) lower: zero for l > i.  b(l,j) lower: zero for l < j.  X(p,k)/Y(k,q) as mapped above.
		if ( on_a ) g.ktri = swapped ? ( lower ? 3 : 4 ) : ( lower ? 1 : 2 );
		else        g.ktri = swapped ? ( lower ? 2 : 1 ) : ( lower ? 4 : 3 );
	}
	g.tile_counter = ctx().dynamic_tiles ? ctx().sched_counters + 2 * ( ctx().sched_next++ % 64 ) : nullptr;
	for ( int sgm = 1; sgm < nseg; ++sgm )
	{
		g.Xseg[sgm - 1] = swapped ? b_more[sgm - 1] : a_more[sgm - 1];
		g.Yseg[sgm - 1] = swapped ? a_more[sgm - 1] : b_more[sgm - 1];
	}

	// -- X: k-contiguous, p-contiguous, or packed
	bool xk = false, yk = false;
	auto misaligned = [&]( const T* p ) { return Elem<T>::cplx && ( (uintptr_t)p % ES ) != 0 && ES == 8; };
	if      ( !misaligned( g.X ) && ( xs_k == 1 || k == 1 ) && ( xs_p >= k || g.P == 1 ) && xs_k >= 0 ) { xk = true;  g.ldx = ( g.P == 1 ? k : xs_p ); }
	else if ( !misaligned( g.X ) && ( xs_p == 1 || g.P == 1 ) && ( xs_k >= g.P || k == 1 ) )            { xk = false; g.ldx = ( k == 1 ? g.P : xs_k ); }
	else if ( nseg > 1 ) rc = fail( "b200_gemm_kpanels: panels must be row- or column-stored" );
	else if ( rc == kSuccess )
	{
		if ( dev_alloc( &tmp_x, (size_t)g.P * k * ES, st ) != kSuccess ) rc = kFailure;
		else { rc = copy2d( (T*)tmp_x, k, (int64_t)1, g.X, xs_p, xs_k, g.P, k, st ); g.X = (const T*)tmp_x; xk = true; g.ldx = k; }
	}
	if      ( !misaligned( g.Y ) && ( ys_k == 1 || k == 1 ) && ( ys_q >= k || g.Q == 1 ) && ys_k >= 0 ) { yk = true;  g.ldy = ( g.Q == 1 ? k : ys_q ); }
	else if ( !misaligned( g.Y ) && ( ys_q == 1 || g.Q == 1 ) && ( ys_k >= g.Q || k == 1 ) )            { yk = false; g.ldy = ( k == 1 ? g.Q : ys_k ); }
	else if ( nseg > 1 ) rc = fail( "b200_gemm_kpanels: panels must be row- or column-stored" );
	else if ( rc == kSuccess )
	{
		if ( dev_alloc( &tmp_y, (size_t)g.Q * k * ES, st ) != kSuccess ) rc = kFailure;
		else { rc = copy2d( (T*)tmp_y, (int64_t)1, k, g.Y, ys_k, ys_q, k, g.Q, st ); g.Y = (const T*)tmp_y; yk = true; g.ldy = k; }
	}

	// FP32 kernels pair accumulators along q (packed FFMA2), which a k-contiguous Y can only feed through two
	// register moves per pair (ncu/SASS: +1000 MOV/IMAD per 1024 FFMA2, 33 instead of 58 TFLOP/s for sgemm "TN").
	// For problems large enough to notice, Y is transposed once into a q-contiguous temporary instead:
	// O(K*Q) traffic against O(P*Q*K) flops (0.3 % of the run time at 16384^3).
	if ( rc == kSuccess && yk && !tmp_y && nseg == 1 && ( std::is_same<T, float>::value || std::is_same<T, float2>::value ) &&
	     ctx().transpose_y && g.P >= 512 && (double)g.P * (double)g.Q * (double)k >= 1e9 && ( g.Q * ES ) % 16 == 0 )
	{
		if ( dev_alloc( &tmp_y, (size_t)g.Q * k * ES, st ) != kSuccess ) rc = kFailure;
		else
		{
			rc = transpose2d( (T*)tmp_y, g.Q, g.Y, g.ldy, g.Q, k, st );
			g.Y = (const T*)tmp_y; yk = false; g.ldy = g.Q;
		}
	}

	if ( rc == kSuccess )
	{
		bool al = ( (uintptr_t)g.X % 16 == 0 ) && ( (uintptr_t)g.Y % 16 == 0 ) &&
		          ( ( g.ldx * ES ) % 16 == 0 ) && ( ( g.ldy * ES ) % 16 == 0 );
		for ( int sgm = 1; sgm < nseg; ++sgm )
			al = al && ( (uintptr_t)g.Xseg[sgm - 1] % 16 == 0 ) && ( (uintptr_t)g.Yseg[sgm - 1] % 16 == 0 );
		g.d_vec_ok = ( (uintptr_t)g.D % 16 == 0 ) && ( ( g.ldd * ES ) % 16 == 0 );
		rc = launch_gemm_kernel<T>( g, xk, yk, al, st );
	}
	if ( rc == kSuccess && c_general ) rc = copy2d( c, rs_c, cs_c, cd, rs_cd, cs_cd, m, n, st, uplo_c );
	dev_free( tmp_x, st ); dev_free( tmp_y, st ); dev_free( tmp_c, st );
	return rc;
}

// C(i,j) += beta * S(i,j) on dense column-major device blocks (the host C of a k-panel pipelined call is added once,
// after its alpha*A*B part has been accumulated)
template <typename R, int NC>
__global__ void add_scaled_kernel( R* __restrict__ c, const R* __restrict__ s, int64_t total, R br, R bi )
{
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		if ( NC == 1 ) c[e] = fma( br, s[e], c[e] );
		else
		{
			const R sr = s[2 * e], si = s[2 * e + 1];
			c[2 * e]     += br * sr - bi * si;
			c[2 * e + 1] += br * si + bi * sr;
		}
	}
}

// ---- host operands, long k: pipeline over k PANELS -------------------------------------------------------------
// C := beta*C + alpha*A*B with every operand in host memory costs 8(mk + kn + 2mn) bytes of PCIe traffic.  The column
// block pipeline of gemm_front cannot start the second block before ALL of A has arrived.  Here the product is
// accumulated panel by panel over k (the pc loop of bli_gemm_blk_var3): round p needs only A(:, panel p) and
// B(panel p, :), 1/np of the traffic, and is one full-size launch; the host C is staged meanwhile into a separate buffer
// and merged (beta) during the last round, which runs per column block so that each finished block of C leaves on the
// D2H stream under the kernels of the next one.  Exposed transfer: the first pair of panels and the last block of C.
template <typename T>
static int gemm_host_kpipe( bool conja, bool conjb, int64_t m, int64_t n, int64_t k, T al, const T* a, int64_t rs_a, int64_t cs_a,
                            const T* b, int64_t rs_b, int64_t cs_b, T be, T* c, int64_t rs_c, int64_t cs_c, cudaStream_t st )
{
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	constexpr size_t ES = sizeof(T);
	Context& cx = ctx();
	cudaStream_t s_in = cx.copy_stream, s_out = cx.d2h_stream;
	const bool a_host = classify( a ) != MemKind::Device, b_host = classify( b ) != MemKind::Device;
	const bool load_c = !Scalar<T>::is_zero( be );
	// panel / block boundaries: uniform eighths, except that the LAST column block is quartered so that the exposed tail
	// (last block of C going home) is short.  (A short FIRST k panel was measured slower: a row panel of a column-major B
	// is a 2-D copy whose chunks are kw*8 bytes, and 4 KiB chunks move at a fraction of the PCIe rate.)
	const int64_t kb = std::max<int64_t>( 512, ( ( k + 7 ) / 8 + 127 ) / 128 * 128 );
	const int64_t nb = std::max<int64_t>( 512, ( ( n + 7 ) / 8 + 127 ) / 128 * 128 );
	std::vector<int64_t> pk{ 0 }, pn{ 0 };
	{
		while ( pk.back() < k ) pk.push_back( std::min( k, pk.back() + kb ) );
		while ( pn.back() < n ) pn.push_back( std::min( n, pn.back() + nb ) );
		const int64_t n0 = std::max<int64_t>( 512, ( nb / 4 + 127 ) / 128 * 128 );
		if ( pn.size() > 2 && n - pn[pn.size() - 2] > n0 ) pn.insert( pn.end() - 1, n - n0 );
	}
	const int np = (int)pk.size() - 1, nblk = (int)pn.size() - 1;
	void *da = nullptr, *db = nullptr, *dc = nullptr, *ds = nullptr;
	int rc = kSuccess;
	if ( ( a_host && dev_alloc( &da, (size_t)m * k * ES, st ) != kSuccess ) || ( b_host && dev_alloc( &db, (size_t)k * n * ES, st ) != kSuccess ) ||
	     dev_alloc( &dc, (size_t)m * n * ES, st ) != kSuccess || ( load_c && dev_alloc( &ds, (size_t)m * n * ES, st ) != kSuccess ) ) rc = kFailure;
	std::vector<cudaEvent_t> ev( np + 2 * nblk + 1 );
	for ( auto& e : ev ) cudaEventCreateWithFlags( &e, cudaEventDisableTiming );
	cudaEvent_t* ev_p = ev.data(); cudaEvent_t* ev_c = ev.data() + np; cudaEvent_t* ev_done = ev.data() + np + nblk; cudaEvent_t ev_alloc = ev.back();
	cudaEventRecord( ev_alloc, st );
	cudaStreamWaitEvent( s_in, ev_alloc, 0 ); cudaStreamWaitEvent( s_out, ev_alloc, 0 );
	const T one = Scalar<T>::make( 1.0, 0.0 ), zero = Scalar<T>::make( 0.0, 0.0 );
	int c_sent = 0;                                  // column blocks of the host C already queued for staging
	auto send_c = [&]( int upto ) -> int
	{
		int r = kSuccess;
		for ( ; c_sent < upto && c_sent < nblk && r == kSuccess; ++c_sent )
		{
			const int64_t j0 = pn[c_sent], w = pn[c_sent + 1] - j0;
			if ( load_c ) r = stage_to_device( (T*)ds + j0 * m, c + j0 * cs_c, m, w, rs_c, cs_c, ES, s_in );
			cudaEventRecord( ev_c[c_sent], s_in );
		}
		return r;
	};
	for ( int p = 0; p < np && rc == kSuccess; ++p )
	{
		const int64_t p0 = pk[p], kw = pk[p + 1] - p0;
		// panel p of A (m x kw, stored densely at da + p0*m) and of B (kw x n, stored densely at db + p0*n)
		if ( a_host ) rc = stage_to_device( (T*)da + p0 * m, a + p0 * cs_a, m, kw, rs_a, cs_a, ES, s_in );
		if ( rc == kSuccess && b_host ) rc = stage_to_device( (T*)db + p0 * n, b + p0 * rs_b, kw, n, rs_b, cs_b, ES, s_in );
		cudaEventRecord( ev_p[p], s_in );
		// the host C trickles in behind the panels, starting behind the second pair so that round 1 is never kept waiting
		if ( rc == kSuccess && p >= 1 ) rc = send_c( np > 1 ? ( p * nblk ) / ( np - 1 ) : nblk );
		const T* ap = a_host ? (const T*)da + p0 * m : a + p0 * cs_a;  const int64_t rs_ap = a_host ? 1 : rs_a, cs_ap = a_host ? m : cs_a;
		const T* bp = b_host ? (const T*)db + p0 * n : b + p0 * rs_b;  const int64_t rs_bp = b_host ? 1 : rs_b, cs_bp = b_host ? kw : cs_b;
		cudaStreamWaitEvent( st, ev_p[p], 0 );
		if ( rc != kSuccess ) break;
		if ( p + 1 < np )
			rc = gemm_dev<T>( conja, conjb, m, n, kw, al, ap, rs_ap, cs_ap, bp, rs_bp, cs_bp, p == 0 ? zero : one, (T*)dc, 1, m, st );
		else
		{
			rc = send_c( nblk );
			for ( int j = 0; j < nblk && rc == kSuccess; ++j )
			{
				const int64_t j0 = pn[j], w = pn[j + 1] - j0;
				rc = gemm_dev<T>( conja, conjb, m, w, kw, al, ap, rs_ap, cs_ap, bp + j0 * cs_bp, rs_bp, cs_bp, p == 0 ? zero : one, (T*)dc + j0 * m, 1, m, st );
				if ( rc == kSuccess && load_c )
				{
					cudaStreamWaitEvent( st, ev_c[j], 0 );
					const int64_t total = m * w;
					const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)cx.num_sms * 16 );
					R br, bi; if constexpr ( Elem<T>::cplx ) { br = be.x; bi = be.y; } else { br = be; bi = 0; }
					add_scaled_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)( (T*)dc + j0 * m ), (const R*)( (const T*)ds + j0 * m ), total, br, bi );
					if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm: launch failed" );
					cx.launches++;
				}
				cudaEventRecord( ev_done[j], st );
				cudaStreamWaitEvent( s_out, ev_done[j], 0 );
				if ( rc == kSuccess ) rc = stage_to_host( c + j0 * cs_c, rs_c, cs_c, (T*)dc + j0 * m, m, w, ES, s_out );
			}
		}
	}
	if ( cudaStreamSynchronize( s_out ) != cudaSuccess || cudaStreamSynchronize( s_in ) != cudaSuccess || cudaStreamSynchronize( st ) != cudaSuccess )
		rc = fail( "b200_gemm: stream sync failed: %s", cudaGetErrorString( cudaGetLastError() ) );
	for ( auto& e : ev ) cudaEventDestroy( e );
	dev_free( da, st ); dev_free( db, st ); dev_free( dc, st ); dev_free( ds, st );
	return rc;
}

// ---- gemm front end: transposition bits + host operand staging ---------------------
template <typename T>
static int gemm_front( int transa, int transb, int64_t m, int64_t n, int64_t k,
                       const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                       const T* b, int64_t rs_b, int64_t cs_b,
                       const T* beta, T* c, int64_t rs_c, int64_t cs_c, int tri_operand = 0 )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || n < 0 || k < 0 ) return fail( "b200_gemm: negative dimension" );
	if ( !alpha || !beta ) return fail( "b200_gemm: alpha/beta must be non-NULL host pointers" );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	constexpr size_t ES = sizeof(T);

	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool conja = Elem<T>::cplx && ( transa & B200_CONJ_NO_TRANSPOSE );
	const bool conjb = Elem<T>::cplx && ( transb & B200_CONJ_NO_TRANSPOSE );
	const T al = *alpha, be = *beta;
	const bool need_ab = ( k > 0 && !Scalar<T>::is_zero( al ) );

	void *da = nullptr, *db = nullptr, *dc = nullptr;
	int rc = kSuccess;
	const bool c_host = ( classify( c ) != MemKind::Device );
	// long k, everything large: accumulate over k panels (gemm_host_kpipe above)
	if ( c_host && need_ab && tri_operand == 0 && ctx().host_kpipe && k >= 4096 && n >= 2048 && m >= 512 &&
	     (double)m * (double)n * (double)k >= 6e10 && ( classify( a ) != MemKind::Device || classify( b ) != MemKind::Device ) )
		return gemm_host_kpipe<T>( conja, conjb, m, n, k, al, a, rs_a, cs_a, b, rs_b, cs_b, be, c, rs_c, cs_c, st );
	// Host C of a large problem: pipeline over column blocks of C (and of B when it is a host
	// operand) so that H2D of block j+1 and D2H of block j-1 run under the kernels of block j.
	const bool pipelined = c_host && need_ab && n >= 1024 && (double)m * (double)n * (double)k >= 2e9 && tri_operand == 0;
	// A host-resident A is needed by every column block.  Pipelined calls move it in k panels behind the first B/C block
	// and start computing that block panel by panel (k-panel accumulation) instead of waiting for all of A.
	const T* a_host = nullptr; int64_t rs_ah = 0, cs_ah = 0;
	if ( need_ab && classify( a ) != MemKind::Device )
	{
		if ( dev_alloc( &da, (size_t)m * k * ES, st ) != kSuccess ) return kFailure;
		if ( pipelined && k >= 2048 ) { a_host = a; rs_ah = rs_a; cs_ah = cs_a; }
		else rc = stage_to_device( da, a, m, k, rs_a, cs_a, ES, st );
		a = (const T*)da; rs_a = 1; cs_a = m;
	}
	const T* b_host = nullptr; int64_t rs_bh = 0, cs_bh = 0;      // set when B moves block-wise
	if ( rc == kSuccess && need_ab && classify( b ) != MemKind::Device )
	{
		if ( dev_alloc( &db, (size_t)k * n * ES, st ) != kSuccess ) rc = kFailure;
		else if ( pipelined ) { b_host = b; rs_bh = rs_b; cs_bh = cs_b; }
		else rc = stage_to_device( db, b, k, n, rs_b, cs_b, ES, st );
		b = (const T*)db; rs_b = 1; cs_b = k;
	}
	T* cdev = c; int64_t rs_cd = rs_c, cs_cd = cs_c;
	if ( rc == kSuccess && c_host && !pipelined )
	{
		if ( dev_alloc( &dc, (size_t)m * n * ES, st ) != kSuccess ) rc = kFailure;
		else if ( !Scalar<T>::is_zero( be ) ) rc = stage_to_device( dc, c, m, n, rs_c, cs_c, ES, st );
		cdev = (T*)dc; rs_cd = 1; cs_cd = m;
	}
	if ( rc == kSuccess && !pipelined )
		rc = gemm_dev<T>( conja, conjb, m, n, k, al, a, rs_a, cs_a, b, rs_b, cs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, 0, tri_operand );
	if ( rc == kSuccess && c_host && !pipelined )
	{
		rc = stage_to_host( c, rs_c, cs_c, dc, m, n, ES, st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "b200_gemm: stream sync failed: %s", cudaGetErrorString( cudaGetLastError() ) );
	}
	if ( rc == kSuccess && pipelined )
	{
		// B and C move block-wise; a host-resident A moves in k panels under the first block (see a_host above).
		Context& cx = ctx();
		cudaStream_t s_in = cx.copy_stream, s_out = cx.d2h_stream;
		const int64_t nb = std::max<int64_t>( 512, ( ( n + 7 ) / 8 + 127 ) / 128 * 128 );
		const int nblk = (int)( ( n + nb - 1 ) / nb );
		std::vector<cudaEvent_t> ev_in( nblk ), ev_done( nblk );
		for ( int j = 0; j < nblk; ++j )
		{
			cudaEventCreateWithFlags( &ev_in[j], cudaEventDisableTiming );
			cudaEventCreateWithFlags( &ev_done[j], cudaEventDisableTiming );
		}
		cudaEvent_t ev_alloc; cudaEventCreateWithFlags( &ev_alloc, cudaEventDisableTiming );
		if ( dev_alloc( &dc, (size_t)m * n * ES, st ) != kSuccess ) rc = kFailure;
		cudaEventRecord( ev_alloc, st );                 // dc usable on the other streams after this
		cudaStreamWaitEvent( s_in, ev_alloc, 0 );
		cudaStreamWaitEvent( s_out, ev_alloc, 0 );
		const bool load_c = !Scalar<T>::is_zero( be );
		auto h2d_block = [&]( int j ) -> int
		{
			const int64_t j0 = (int64_t)j * nb, w = std::min( nb, n - j0 );
			int r = kSuccess;
			if ( b_host ) r = stage_to_device( (T*)db + j0 * k, b_host + j0 * cs_bh, k, w, rs_bh, cs_bh, ES, s_in );
			if ( r == kSuccess && load_c ) r = stage_to_device( (T*)dc + j0 * m, c + j0 * cs_c, m, w, rs_c, cs_c, ES, s_in );
			cudaEventRecord( ev_in[j], s_in );
			return r;
		};
		if ( rc == kSuccess ) rc = h2d_block( 0 );
		for ( int j = 0; j < nblk && rc == kSuccess; ++j )
		{
			const int64_t j0 = (int64_t)j * nb, w = std::min( nb, n - j0 );
			cudaStreamWaitEvent( st, ev_in[j], 0 );
			if ( j == 0 && a_host )
			{
				// first block: C_0 := beta*C_0 + alpha * sum_p A(:, panel p) * B_0(panel p, :), each step waiting only for its panel of A
				const int64_t kb = std::max<int64_t>( 512, ( ( k + 7 ) / 8 + 127 ) / 128 * 128 );
				const T one = Scalar<T>::make( 1.0, 0.0 );
				for ( int64_t p0 = 0; p0 < k && rc == kSuccess; p0 += kb )
				{
					const int64_t kw = std::min( kb, k - p0 );
					rc = stage_to_device( (T*)da + p0 * m, a_host + p0 * cs_ah, m, kw, rs_ah, cs_ah, ES, s_in );
					cudaEvent_t ev_a; cudaEventCreateWithFlags( &ev_a, cudaEventDisableTiming );
					cudaEventRecord( ev_a, s_in );
					cudaStreamWaitEvent( st, ev_a, 0 );
					cudaEventDestroy( ev_a );
					if ( rc == kSuccess )
						rc = gemm_dev<T>( conja, conjb, m, w, kw, al, a + p0 * cs_a, rs_a, cs_a, b + p0 * rs_b, rs_b, cs_b,
						                  p0 == 0 ? be : one, (T*)dc, 1, m, st );
				}
			}
			else
			rc = gemm_dev<T>( conja, conjb, m, w, k, al, a, rs_a, cs_a, b + j0 * cs_b, rs_b, cs_b, be,
			                  (T*)dc + j0 * m, 1, m, st );
			cudaEventRecord( ev_done[j], st );
			if ( rc == kSuccess && j + 1 < nblk ) rc = h2d_block( j + 1 );
			cudaStreamWaitEvent( s_out, ev_done[j], 0 );
			if ( rc == kSuccess ) rc = stage_to_host( c + j0 * cs_c, rs_c, cs_c, (T*)dc + j0 * m, m, w, ES, s_out );
		}
		if ( cudaStreamSynchronize( s_out ) != cudaSuccess || cudaStreamSynchronize( s_in ) != cudaSuccess ||
		     cudaStreamSynchronize( st ) != cudaSuccess )
			rc = fail( "b200_gemm: stream sync failed: %s", cudaGetErrorString( cudaGetLastError() ) );
		for ( int j = 0; j < nblk; ++j ) { cudaEventDestroy( ev_in[j] ); cudaEventDestroy( ev_done[j] ); }
		cudaEventDestroy( ev_alloc );
	}
	dev_free( da, st ); dev_free( db, st ); dev_free( dc, st );
	return rc;
}

// ---- trsm -----------------------------------------------------------------------------
template <typename T> struct TrsmBlk;
template <> struct TrsmBlk<float>   { static constexpr int NB = 64, CN = 64; };
template <> struct TrsmBlk<double>  { static constexpr int NB = 64, CN = 64; };
template <> struct TrsmBlk<float2>  { static constexpr int NB = 32, CN = 64; };
template <> struct TrsmBlk<double2> { static constexpr int NB = 32, CN = 64; };

template <typename T>
struct TrsmPlan
{
	const T* A; int64_t rs_a, cs_a;      // effective triangular matrix (trans folded into strides)
	T*       B; int64_t rs_b, cs_b;
	int64_t  n;
	bool     upper, unit, conj;
	cudaStream_t st;
};

template <typename T>
static int trsm_base( const TrsmPlan<T>& p, int64_t i0, int mb, T alpha )
{
	constexpr int NB = TrsmBlk<T>::NB, CN = TrsmBlk<T>::CN;
	TrsmBaseArgs<T> a;
	a.A = p.A + i0 * ( p.rs_a + p.cs_a ); a.rs_a = p.rs_a; a.cs_a = p.cs_a;
	a.B = p.B + i0 * p.rs_b;              a.rs_b = p.rs_b; a.cs_b = p.cs_b;
	a.n = p.n; a.mb = mb; a.upper = p.upper; a.unit = p.unit; a.conj = p.conj; a.alpha = alpha;
	constexpr int NT = 256;
	auto kern = trsm_base_kernel<T, NB, CN, NT>;
	constexpr int smem = trsm_base_smem<T, NB, CN>();
	static bool attr = false;
	if ( !attr ) { if ( set_smem( kern, smem ) != kSuccess ) return kFailure; attr = true; }
	const int64_t grid = ( p.n + CN - 1 ) / CN;
	kern<<<(unsigned)grid, NT, smem, p.st>>>( a );
	B200_CUDA( cudaGetLastError() );
	ctx().launches++;
	return kSuccess;
}

// Recursive blocked solve of rows [i0, i0+mb): solve one half, rank-k update of
// the other half with the gemm kernel, solve the other half.  alpha is applied
// exactly once to every row (either by the base kernel or as the update's beta,
// as bli_trsm_ex passes alpha as beta: bli_l3_oapi_ex.c:778-789).
template <typename T>
static int trsm_rec( const TrsmPlan<T>& p, int64_t i0, int64_t mb, T alpha )
{
	constexpr int NB = TrsmBlk<T>::NB;
	if ( mb <= NB ) return trsm_base( p, i0, (int)mb, alpha );
	const int64_t nblk = ( mb + NB - 1 ) / NB;
	const int64_t m1 = ( ( nblk + 1 ) / 2 ) * NB, m2 = mb - m1;
	const T one = Scalar<T>::make( 1.0, 0.0 ), mone = Scalar<T>::make( -1.0, 0.0 );
	if ( !p.upper )
	{
		if ( trsm_rec( p, i0, m1, alpha ) != kSuccess ) return kFailure;
		// B2 := alpha*B2 - A21 * X1
		if ( gemm_dev<T>( p.conj, false, m2, p.n, m1, mone,
		                  p.A + ( i0 + m1 ) * p.rs_a + i0 * p.cs_a, p.rs_a, p.cs_a,
		                  p.B + i0 * p.rs_b, p.rs_b, p.cs_b,
		                  alpha, p.B + ( i0 + m1 ) * p.rs_b, p.rs_b, p.cs_b, p.st ) != kSuccess ) return kFailure;
		return trsm_rec( p, i0 + m1, m2, one );
	}
	else
	{
		// upper: the trailing block is solved first; split so the LAST block is the ragged one's partner
		if ( trsm_rec( p, i0 + m2, m1, alpha ) != kSuccess ) return kFailure;
		// B1 := alpha*B1 - A12 * X2
		if ( gemm_dev<T>( p.conj, false, m2, p.n, m1, mone,
		                  p.A + i0 * p.rs_a + ( i0 + m2 ) * p.cs_a, p.rs_a, p.cs_a,
		                  p.B + ( i0 + m2 ) * p.rs_b, p.rs_b, p.cs_b,
		                  alpha, p.B + i0 * p.rs_b, p.rs_b, p.cs_b, p.st ) != kSuccess ) return kFailure;
		return trsm_rec( p, i0, m2, one );
	}
}

template <typename T>
static int trsm_front( int side, int uplo, int transa, int diag, int64_t m, int64_t n,
                       const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                       T* b, int64_t rs_b, int64_t cs_b )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || n < 0 ) return fail( "b200_trsm: negative dimension" );
	if ( !alpha ) return fail( "b200_trsm: alpha must be a non-NULL host pointer" );
	if ( uplo != B200_LOWER && uplo != B200_UPPER ) return fail( "b200_trsm: uplo must be BLIS_LOWER or BLIS_UPPER" );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	constexpr size_t ES = sizeof(T);
	const T al = *alpha;

	// right side: X * op(A) = alpha*B  <=>  op(A)^T * X^T = alpha * B^T   (bli_l3_oapi_ex.c:748-759)
	if ( side == B200_RIGHT )
	{
		std::swap( m, n ); std::swap( rs_b, cs_b );
		transa ^= B200_TRANSPOSE;
	}
	bool upper = ( uplo == B200_UPPER );
	if ( transa & B200_TRANSPOSE ) { std::swap( rs_a, cs_a ); upper = !upper; }
	const bool conj = Elem<T>::cplx && ( transa & B200_CONJ_NO_TRANSPOSE );
	// now: A is m x m (effective uplo `upper`), B is m x n

	void *da = nullptr, *db = nullptr;
	int rc = kSuccess;
	const bool b_host = ( classify( b ) != MemKind::Device );
	const bool zero_alpha = Scalar<T>::is_zero( al );
	T* bdev = b; int64_t rs_bd = rs_b, cs_bd = cs_b;
	if ( b_host )
	{
		if ( dev_alloc( &db, (size_t)m * n * ES, st ) != kSuccess ) return kFailure;
		if ( !zero_alpha ) rc = stage_to_device( db, b, m, n, rs_b, cs_b, ES, st );
		bdev = (T*)db; rs_bd = 1; cs_bd = m;
	}
	if ( zero_alpha )
	{
		// bli_l3_return_early_if_trivial( alpha, a, b, &BLIS_ZERO, b ):  B := 0
		if ( rc == kSuccess ) rc = scal2d( bdev, rs_bd, cs_bd, m, n, al, st );
	}
	else
	{
		if ( rc == kSuccess && classify( a ) != MemKind::Device )
		{
			if ( dev_alloc( &da, (size_t)m * m * ES, st ) != kSuccess ) rc = kFailure;
			else rc = stage_to_device( da, a, m, m, rs_a, cs_a, ES, st );
			a = (const T*)da; rs_a = 1; cs_a = m;
		}
		if ( rc == kSuccess )
		{
			TrsmPlan<T> p{ a, rs_a, cs_a, bdev, rs_bd, cs_bd, n, upper, diag == B200_UNIT_DIAG, conj, st };
			rc = trsm_rec( p, 0, m, al );
		}
	}
	if ( rc == kSuccess && b_host )
	{
		rc = stage_to_host( b, rs_b, cs_b, db, m, n, ES, st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "b200_trsm: stream sync failed" );
	}
	dev_free( da, st ); dev_free( db, st );
	return rc;
}

// ---- gemmt family: gemmt, syrk, herk, syr2k, her2k ---------------------------------------
// bli_gemmt_ex / bli_syrk_ex / bli_herk_ex / bli_syr2k_ex / bli_her2k_ex (frame/3/bli_l3_oapi_ex.c:151-346):
// every one of them is one or two gemmt's, C := beta*C + alpha*A*B restricted to the stored triangle of the
// m x m matrix C (macrokernels frame/3/gemmt/bli_gemmt_{l,u}_ker_var2.c); herk/her2k then zero the imaginary
// part of the diagonal (bli_setid).  Here a gemmt is the gemm kernel with a triangular tile schedule.
enum { kOpGemmt = 0, kOpSyrk = 1, kOpHerk = 2, kOpSyr2k = 3, kOpHer2k = 4 };

template <typename R>
__global__ void zero_diag_imag_kernel( R* c, int64_t inc, int64_t m )
{
	for ( int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x )
		c[2 * i * inc + 1] = (R)0;
}

// Device view of a host or device operand: host data is staged into a dense column-major temporary.
template <typename T>
static int operand_to_device( const T*& p, int64_t& rs, int64_t& cs, int64_t m, int64_t n, void** tmp, cudaStream_t st )
{
	*tmp = nullptr;
	if ( m <= 0 || n <= 0 || classify( p ) == MemKind::Device ) return kSuccess;
	if ( dev_alloc( tmp, (size_t)m * n * sizeof(T), st ) != kSuccess ) return kFailure;
	if ( stage_to_device( *tmp, p, m, n, rs, cs, sizeof(T), st ) != kSuccess ) return kFailure;
	p = (const T*)*tmp; rs = 1; cs = m;
	return kSuccess;
}

template <typename T>
static int gemmt_family_front( int op, int uploc, int transa, int transb, int64_t m, int64_t k,
                               const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                               const T* b, int64_t rs_b, int64_t cs_b,
                               const T* beta, T* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || k < 0 ) return fail( "%s: negative dimension", name );
	if ( !alpha || !beta ) return fail( "%s: alpha/beta must be non-NULL host pointers", name );
	if ( uploc != B200_LOWER && uploc != B200_UPPER ) return fail( "%s: uplo must be BLIS_LOWER or BLIS_UPPER", name );
	if ( m == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	constexpr bool CPLX = Elem<T>::cplx;
	const bool two_operands = ( op == kOpGemmt || op == kOpSyr2k || op == kOpHer2k );
	const bool hermitian    = ( op == kOpHerk || op == kOpHer2k ); (void)hermitian;
	if ( !two_operands ) { b = a; rs_b = rs_a; cs_b = cs_a; transb = transa; }

	// op(A): m x k.  op(B): k x m for gemmt, m x k for the rank-2k operations (bli_l3_tapi_ex.c:251-252).
	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool ca = CPLX && ( transa & B200_CONJ_NO_TRANSPOSE );
	const bool cb = CPLX && ( transb & B200_CONJ_NO_TRANSPOSE );
	const T al = *alpha, be = *beta;
	const bool need_ab = ( k > 0 && !Scalar<T>::is_zero( al ) );

	void *da = nullptr, *db = nullptr, *dc = nullptr;
	int rc = kSuccess;
	const bool shared_ab = ( op != kOpGemmt && a == b && rs_a == rs_b && cs_a == cs_b );
	if ( need_ab )
	{
		const T* a0 = a;
		rc = operand_to_device( a, rs_a, cs_a, m, k, &da, st );
		if ( rc == kSuccess )
		{
			if ( !two_operands || ( shared_ab && a0 != a ) ) { b = a; rs_b = rs_a; cs_b = cs_a; }
			else if ( op == kOpGemmt ) rc = operand_to_device( b, rs_b, cs_b, k, m, &db, st );
			else                       rc = operand_to_device( b, rs_b, cs_b, m, k, &db, st );
		}
	}
	const bool c_host = ( classify( c ) != MemKind::Device );
	T* cdev = c; int64_t rs_cd = rs_c, cs_cd = cs_c;
	if ( rc == kSuccess && c_host )
	{
		// the whole array travels both ways, so the triangle that is not stored returns unchanged
		const T* cc = c;
		rc = operand_to_device( cc, rs_cd, cs_cd, m, m, &dc, st );
		cdev = (T*)dc;
	}

	const T one = Scalar<T>::make( 1.0, 0.0 );
	if ( rc == kSuccess )
	{
		switch ( op )
		{
			case kOpGemmt:          // C := beta*C + alpha * op(A) * op(B)
				rc = gemm_dev<T>( ca, cb, m, m, k, al, a, rs_a, cs_a, b, rs_b, cs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpSyrk:           // C := beta*C + alpha * op(A) * op(A)^T
				rc = gemm_dev<T>( ca, ca, m, m, k, al, a, rs_a, cs_a, a, cs_a, rs_a, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpHerk:           // C := beta*C + alpha * op(A) * op(A)^H   (alpha, beta real)
				rc = gemm_dev<T>( ca, !ca && CPLX, m, m, k, al, a, rs_a, cs_a, a, cs_a, rs_a, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpSyr2k:          // C := beta*C + alpha * op(A) * op(B)^T + alpha * op(B) * op(A)^T
				rc = gemm_dev<T>( ca, cb, m, m, k, al, a, rs_a, cs_a, b, cs_b, rs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				if ( rc == kSuccess )
				rc = gemm_dev<T>( cb, ca, m, m, k, al, b, rs_b, cs_b, a, cs_a, rs_a, one, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			case kOpHer2k:          // C := beta*C + alpha * op(A) * op(B)^H + conj(alpha) * op(B) * op(A)^H   (beta real)
			{
				T alh = al;
				if constexpr ( CPLX ) alh.y = -alh.y;
				rc = gemm_dev<T>( ca, !cb && CPLX, m, m, k, al, a, rs_a, cs_a, b, cs_b, rs_b, be, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				if ( rc == kSuccess )
				rc = gemm_dev<T>( cb, !ca && CPLX, m, m, k, alh, b, rs_b, cs_b, a, cs_a, rs_a, one, cdev, rs_cd, cs_cd, st, 1, nullptr, nullptr, uploc );
				break;
			}
			default: rc = fail( "%s: unknown operation", name );
		}
	}
	if constexpr ( CPLX )
	{
		if ( rc == kSuccess && hermitian )
		{
			using R = typename Elem<T>::real;
			const int blocks = (int)std::min<int64_t>( ( m + 255 ) / 256, (int64_t)ctx().num_sms * 4 );
			zero_diag_imag_kernel<R><<<blocks, 256, 0, st>>>( (R*)cdev, rs_cd + cs_cd, m );
			if ( cudaGetLastError() != cudaSuccess ) rc = fail( "%s: launch failed", name );
			ctx().launches++;
		}
	}
	if ( rc == kSuccess && c_host )
	{
		rc = stage_to_host( c, rs_c, cs_c, dc, m, m, sizeof(T), st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "%s: stream sync failed", name );
	}
	dev_free( da, st ); dev_free( db, st ); dev_free( dc, st );
	return rc;
}

} // namespace b200

// ---- C ABI --------------------------------------------------------------------------
using namespace b200;

#define B200_DEF_GEMM( ch, ctype, T ) \
extern "C" b200_err_t b200_##ch##gemm( int transa, int transb, b200_dim_t m, b200_dim_t n, b200_dim_t k, \
	const ctype* alpha, const ctype* a, b200_inc_t rs_a, b200_inc_t cs_a, \
	const ctype* b, b200_inc_t rs_b, b200_inc_t cs_b, const ctype* beta, \
	ctype* c, b200_inc_t rs_c, b200_inc_t cs_c ) \
{ \
	return gemm_front<T>( transa, transb, m, n, k, (const T*)alpha, (const T*)a, rs_a, cs_a, \
	                      (const T*)b, rs_b, cs_b, (const T*)beta, (T*)c, rs_c, cs_c ); \
}
B200_DEF_GEMM( s, float, float )
B200_DEF_GEMM( d, double, double )
B200_DEF_GEMM( c, b200_scomplex, float2 )
B200_DEF_GEMM( z, b200_dcomplex, double2 )

extern "C" b200_err_t b200_gemm( int dt, int transa, int transb, b200_dim_t m, b200_dim_t n, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
	const void* b, b200_inc_t rs_b, b200_inc_t cs_b, const void* beta,
	void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{
	switch ( dt )
	{
		case B200_FLOAT:    return gemm_front<float>  ( transa, transb, m, n, k, (const float*)alpha,   (const float*)a,   rs_a, cs_a, (const float*)b,   rs_b, cs_b, (const float*)beta,   (float*)c,   rs_c, cs_c );
		case B200_DOUBLE:   return gemm_front<double> ( transa, transb, m, n, k, (const double*)alpha,  (const double*)a,  rs_a, cs_a, (const double*)b,  rs_b, cs_b, (const double*)beta,  (double*)c,  rs_c, cs_c );
		case B200_SCOMPLEX: return gemm_front<float2> ( transa, transb, m, n, k, (const float2*)alpha,  (const float2*)a,  rs_a, cs_a, (const float2*)b,  rs_b, cs_b, (const float2*)beta,  (float2*)c,  rs_c, cs_c );
		case B200_DCOMPLEX: return gemm_front<double2>( transa, transb, m, n, k, (const double2*)alpha, (const double2*)a, rs_a, cs_a, (const double2*)b, rs_b, cs_b, (const double2*)beta, (double2*)c, rs_c, cs_c );
	}
	return fail( "b200_gemm: unsupported datatype %d (mixed-datatype gemm is out of scope)", dt );
}

#define B200_DEF_TRSM( ch, ctype, T ) \
extern "C" b200_err_t b200_##ch##trsm( int side, int uploa, int transa, int diaga, b200_dim_t m, b200_dim_t n, \
	const ctype* alpha, const ctype* a, b200_inc_t rs_a, b200_inc_t cs_a, \
	ctype* b, b200_inc_t rs_b, b200_inc_t cs_b ) \
{ \
	return trsm_front<T>( side, uploa, transa, diaga, m, n, (const T*)alpha, (const T*)a, rs_a, cs_a, (T*)b, rs_b, cs_b ); \
}
B200_DEF_TRSM( s, float, float )
B200_DEF_TRSM( d, double, double )
B200_DEF_TRSM( c, b200_scomplex, float2 )
B200_DEF_TRSM( z, b200_dcomplex, double2 )

extern "C" b200_err_t b200_trsm( int dt, int side, int uploa, int transa, int diaga, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
	void* b, b200_inc_t rs_b, b200_inc_t cs_b )
{
	switch ( dt )
	{
		case B200_FLOAT:    return trsm_front<float>  ( side, uploa, transa, diaga, m, n, (const float*)alpha,   (const float*)a,   rs_a, cs_a, (float*)b,   rs_b, cs_b );
		case B200_DOUBLE:   return trsm_front<double> ( side, uploa, transa, diaga, m, n, (const double*)alpha,  (const double*)a,  rs_a, cs_a, (double*)b,  rs_b, cs_b );
		case B200_SCOMPLEX: return trsm_front<float2> ( side, uploa, transa, diaga, m, n, (const float2*)alpha,  (const float2*)a,  rs_a, cs_a, (float2*)b,  rs_b, cs_b );
		case B200_DCOMPLEX: return trsm_front<double2>( side, uploa, transa, diaga, m, n, (const double2*)alpha, (const double2*)a, rs_a, cs_a, (double2*)b, rs_b, cs_b );
	}
	return fail( "b200_trsm: unsupported datatype %d", dt );
}

// k-panel accumulation: C := beta*C + alpha * sum_{s<npanels} op(A_s) * op(B_s), every panel k wide, all
// A panels (resp. B panels) with the same strides.  Device-resident operands, d and z only.
// ---- hemm, symm, trmm, trmm3 -----------------------------------------------------------------
// bli_hemm_ex / bli_symm_ex / bli_trmm3_ex / bli_trmm_ex (frame/3/bli_l3_oapi_ex.c:349-689): the gemm control tree
// with a structured A.  The reference resolves the structure while PACKING (bli_packm_struc_cxk.c:146-301: the
// unstored side of a Hermitian/symmetric matrix is read from its mirror image, conjugated for Hermitian; the unstored
// side of a triangular matrix is packed as explicit zeros; ref_kernels/1m/bli_packm_cxc_diag_ref.c:36-98: a unit
// diagonal is packed as one, a Hermitian diagonal loses its imaginary part) and then runs gemm-shaped macrokernels
// (trmm ones skip the zero k range).  Here the structure is resolved ONCE into a dense m x m device matrix
// (O(m^2) traffic against O(m^2 n) flops) and the product is the gemm kernel, trimmed in k for trmm.
namespace b200 {

enum { kStrucTri = 0, kStrucSym = 1, kStrucHerm = 2 };

template <typename R, int NC>
__global__ void densify_kernel( R* dst, int64_t ldd, const R* src, int64_t rs, int64_t cs,
                                int64_t m, int struc, int lower, int unit )
{
	const int64_t total = m * m;
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		const int64_t i = e % m, j = e / m;
		const bool stored = lower ? ( i >= j ) : ( i <= j );
		R re = (R)0, im = (R)0;
		if ( i == j )
		{
			if ( struc == kStrucTri && unit ) re = (R)1;
			else
			{
				const R* p = src + ( i * rs + j * cs ) * NC;
				re = p[0];
				if ( NC == 2 && struc != kStrucHerm ) im = p[NC - 1];
			}
		}
		else if ( stored )
		{
			const R* p = src + ( i * rs + j * cs ) * NC;
			re = p[0]; if ( NC == 2 ) im = p[NC - 1];
		}
		else if ( struc != kStrucTri )
		{
			const R* p = src + ( j * rs + i * cs ) * NC;         // mirror image
			re = p[0]; if ( NC == 2 ) im = ( struc == kStrucHerm ) ? -p[NC - 1] : p[NC - 1];
		}
		R* d = dst + ( i + j * ldd ) * NC;
		if ( src == dst && stored && i != j ) continue;       // in place (staged copy): stored elements stay
		d[0] = re; if ( NC == 2 ) d[NC - 1] = im;
	}
}

// Dense, structure-resolved copy of the ma x ma matrix A (host or device) in `*da` (column-major, ld = ma).
template <typename T>
static int densify_operand( void** da, const T* a, int64_t rs_a, int64_t cs_a, int64_t ma, int struc, int uplo, bool unit, cudaStream_t st )
{
	using R = typename Elem<T>::real;
	constexpr int NC = Elem<T>::cplx ? 2 : 1;
	*da = nullptr;
	if ( dev_alloc( da, (size_t)ma * ma * sizeof(T), st ) != kSuccess ) return kFailure;
	const T* src = a; int64_t rs = rs_a, cs = cs_a;
	if ( classify( a ) != MemKind::Device )
	{
		// the whole array is staged (the unstored triangle may hold anything; it is never used) and resolved in place
		if ( stage_to_device( *da, a, ma, ma, rs_a, cs_a, sizeof(T), st ) != kSuccess ) return kFailure;
		src = (const T*)*da; rs = 1; cs = ma;
	}
	const int64_t total = ma * ma;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	densify_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)*da, ma, (const R*)src, rs, cs, ma, struc, uplo == B200_LOWER ? 1 : 0, unit ? 1 : 0 );
	B200_CUDA( cudaGetLastError() );
	ctx().launches++;
	return kSuccess;
}

// op: 0 hemm, 1 symm, 2 trmm3, 3 trmm (C == B, beta ignored)
template <typename T>
static int struc_mm_front( int op, int side, int uploa, int transa, int diaga, int transb, int64_t m, int64_t n,
                           const T* alpha, const T* a, int64_t rs_a, int64_t cs_a,
                           const T* b, int64_t rs_b, int64_t cs_b,
                           const T* beta, T* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( m < 0 || n < 0 ) return fail( "%s: negative dimension", name );
	if ( !alpha || ( op != 3 && !beta ) ) return fail( "%s: alpha/beta must be non-NULL host pointers", name );
	if ( uploa != B200_LOWER && uploa != B200_UPPER ) return fail( "%s: uplo must be BLIS_LOWER or BLIS_UPPER", name );
	if ( side != B200_LEFT && side != B200_RIGHT ) return fail( "%s: side must be BLIS_LEFT or BLIS_RIGHT", name );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	const int64_t ma = ( side == B200_LEFT ) ? m : n;
	const T zero = Scalar<T>::make( 0.0, 0.0 );
	const T al = *alpha, be = ( op == 3 ) ? zero : *beta;
	const int struc = ( op == 0 ) ? kStrucHerm : ( op == 1 ) ? kStrucSym : kStrucTri;

	void *da = nullptr, *dt = nullptr;
	int rc = kSuccess;
	if ( !Scalar<T>::is_zero( al ) )
		rc = densify_operand<T>( &da, a, rs_a, cs_a, ma, struc, uploa, diaga == B200_UNIT_DIAG, st );
	const T* ad = (const T*)da;

	const T* bsrc = b; int64_t rs_bs = rs_b, cs_bs = cs_b; int transb_use = transb;
	if ( op == 3 )
	{
		// trmm is in place: B := alpha * transa(A) * B.  The product reads a copy of B (bli_trmm_ex aliases C = B and
		// relies on the macrokernel's loop order; a copy costs O(mn) against O(m^2 n)).
		transb_use = B200_NO_TRANSPOSE;
		if ( rc == kSuccess && !Scalar<T>::is_zero( al ) )
		{
			if ( dev_alloc( &dt, (size_t)m * n * sizeof(T), st ) != kSuccess ) rc = kFailure;
			else if ( classify( b ) != MemKind::Device ) rc = stage_to_device( dt, b, m, n, rs_b, cs_b, sizeof(T), st );
			else rc = copy2d( (T*)dt, (int64_t)1, m, b, rs_b, cs_b, m, n, st );
			bsrc = (const T*)dt; rs_bs = 1; cs_bs = m;
		}
	}
	if ( rc == kSuccess )
	{
		// effective triangle of transa(A): transposition mirrors it
		int tri_operand = 0;
		if ( struc == kStrucTri )
		{
			const bool lower_eff = ( uploa == B200_LOWER ) != ( ( transa & B200_TRANSPOSE ) != 0 );
			tri_operand = ( side == B200_LEFT ? kTriA : kTriB ) | ( lower_eff ? kTriLower : kTriUpper );
		}
		const int ta = ( struc == kStrucTri ) ? transa : ( transa & B200_CONJ_NO_TRANSPOSE );   // hemm/symm: conja only
		if ( side == B200_LEFT )
			rc = gemm_front<T>( ta, transb_use, m, n, m, &al, ad, 1, ma, bsrc, rs_bs, cs_bs, &be, c, rs_c, cs_c, tri_operand );
		else
			rc = gemm_front<T>( transb_use, ta, m, n, n, &al, bsrc, rs_bs, cs_bs, ad, 1, ma, &be, c, rs_c, cs_c, tri_operand );
	}
	dev_free( da, st ); dev_free( dt, st );
	return rc;
}

} // namespace b200

static int struc_mm_dt( int dt, int op, int side, int uploa, int transa, int diaga, int transb, int64_t m, int64_t n,
                        const void* alpha, const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                        const void* beta, void* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	switch ( dt )
	{
		case B200_FLOAT:    return struc_mm_front<float>  ( op, side, uploa, transa, diaga, transb, m, n, (const float*)alpha,   (const float*)a,   rs_a, cs_a, (const float*)b,   rs_b, cs_b, (const float*)beta,   (float*)c,   rs_c, cs_c, name );
		case B200_DOUBLE:   return struc_mm_front<double> ( op, side, uploa, transa, diaga, transb, m, n, (const double*)alpha,  (const double*)a,  rs_a, cs_a, (const double*)b,  rs_b, cs_b, (const double*)beta,  (double*)c,  rs_c, cs_c, name );
		case B200_SCOMPLEX: return struc_mm_front<float2> ( op, side, uploa, transa, diaga, transb, m, n, (const float2*)alpha,  (const float2*)a,  rs_a, cs_a, (const float2*)b,  rs_b, cs_b, (const float2*)beta,  (float2*)c,  rs_c, cs_c, name );
		case B200_DCOMPLEX: return struc_mm_front<double2>( op, side, uploa, transa, diaga, transb, m, n, (const double2*)alpha, (const double2*)a, rs_a, cs_a, (const double2*)b, rs_b, cs_b, (const double2*)beta, (double2*)c, rs_c, cs_c, name );
	}
	return fail( "%s: unsupported datatype %d", name, dt );
}

extern "C" b200_err_t b200_hemm( int dt, int side, int uploa, int conja, int transb, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return struc_mm_dt( dt, 0, side, uploa, conja, B200_NONUNIT_DIAG, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_hemm" ); }

extern "C" b200_err_t b200_symm( int dt, int side, int uploa, int conja, int transb, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return struc_mm_dt( dt, 1, side, uploa, conja, B200_NONUNIT_DIAG, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_symm" ); }

extern "C" b200_err_t b200_trmm3( int dt, int side, int uploa, int transa, int diaga, int transb, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return struc_mm_dt( dt, 2, side, uploa, transa, diaga, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_trmm3" ); }

extern "C" b200_err_t b200_trmm( int dt, int side, int uploa, int transa, int diaga, b200_dim_t m, b200_dim_t n,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, void* b, b200_inc_t rs_b, b200_inc_t cs_b )
{ return struc_mm_dt( dt, 3, side, uploa, transa, diaga, B200_NO_TRANSPOSE, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, nullptr, b, rs_b, cs_b, "b200_trmm" ); }

// ---- gemmt family C ABI ------------------------------------------------------------------
// Real-typed scalars of herk (alpha, beta) and her2k (beta) are widened to the matrix datatype (imaginary part 0),
// as bli_obj_init_finish_1x1( dt_r, ... ) + typecast does in bli_l3_tapi_ex.c:184-185,259-260.
template <typename T> static T widen_real( const void* p )
{
	using R = typename Elem<T>::real;
	return Scalar<T>::make( (double)*(const R*)p, 0.0 );
}

template <typename T>
static int gemmt_family_any( int op, int uploc, int transa, int transb, int64_t m, int64_t k, const void* alpha,
                             const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                             const void* beta, void* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	if ( !alpha || !beta ) return fail( "%s: alpha/beta must be non-NULL host pointers", name );
	const T al = ( op == kOpHerk )                    ? widen_real<T>( alpha ) : *(const T*)alpha;
	const T be = ( op == kOpHerk || op == kOpHer2k ) ? widen_real<T>( beta )  : *(const T*)beta;
	return gemmt_family_front<T>( op, uploc, transa, transb, m, k, &al, (const T*)a, rs_a, cs_a, (const T*)b, rs_b, cs_b,
	                              &be, (T*)c, rs_c, cs_c, name );
}

static int gemmt_family_dt( int dt, int op, int uploc, int transa, int transb, int64_t m, int64_t k, const void* alpha,
                            const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                            const void* beta, void* c, int64_t rs_c, int64_t cs_c, const char* name )
{
	switch ( dt )
	{
		case B200_FLOAT:    return gemmt_family_any<float>  ( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
		case B200_DOUBLE:   return gemmt_family_any<double> ( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
		case B200_SCOMPLEX: return gemmt_family_any<float2> ( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
		case B200_DCOMPLEX: return gemmt_family_any<double2>( op, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, name );
	}
	return fail( "%s: unsupported datatype %d", name, dt );
}

extern "C" b200_err_t b200_gemmt( int dt, int uploc, int transa, int transb, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpGemmt, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_gemmt" ); }

extern "C" b200_err_t b200_syr2k( int dt, int uploc, int transa, int transb, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpSyr2k, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c, "b200_syr2k" ); }

extern "C" b200_err_t b200_her2k( int dt, int uploc, int transa, int transb, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
	const void* beta_real, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpHer2k, uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta_real, c, rs_c, cs_c, "b200_her2k" ); }

extern "C" b200_err_t b200_syrk( int dt, int uploc, int transa, b200_dim_t m, b200_dim_t k,
	const void* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpSyrk, uploc, transa, transa, m, k, alpha, a, rs_a, cs_a, a, rs_a, cs_a, beta, c, rs_c, cs_c, "b200_syrk" ); }

extern "C" b200_err_t b200_herk( int dt, int uploc, int transa, b200_dim_t m, b200_dim_t k,
	const void* alpha_real, const void* a, b200_inc_t rs_a, b200_inc_t cs_a, const void* beta_real, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{ return gemmt_family_dt( dt, kOpHerk, uploc, transa, transa, m, k, alpha_real, a, rs_a, cs_a, a, rs_a, cs_a, beta_real, c, rs_c, cs_c, "b200_herk" ); }

// ---- mixed-datatype gemm ---------------------------------------------
// bli_gemm_ex with operands of different domain and/or precision (docs/MixedDatatypes.md; frame/3/gemm/bli_gemm_cntl.c:
// 87-392): A and B are typecast to the computation precision while they are packed, the product runs in the computation
// precision in the smallest domain that holds it (table of MixedDatatypes.md: "R += C*C" keeps only the real part and
// costs 4mnk, "C += R*C" treats the complex operand as a real matrix with twice the rows, ...), and the result is
// typecast and accumulated into C with beta in C's own datatype (ref_kernels/3/bli_gemm_ref.c:318-385,
// ref_kernels/ind/bli_gemm_{ccr,crr,rcc}_ref.c).  Here: one conversion pass per operand (typecast, transposition,
// conjugation, alpha where it has to act before a projection), ONE homogeneous real or complex gemm of the computation
// precision with the kernels above, one combine pass into C.
namespace b200 {

struct MdElem { double r, i; };

__device__ __forceinline__ MdElem md_load( const void* base, int dt, int64_t off )
{
	MdElem e; e.i = 0.0;
	switch ( dt )
	{
		case B200_FLOAT:    e.r = ( (const float*)base )[off]; break;
		case B200_DOUBLE:   e.r = ( (const double*)base )[off]; break;
		case B200_SCOMPLEX: { const float2 v = ( (const float2*)base )[off]; e.r = v.x; e.i = v.y; break; }
		default:            { const double2 v = ( (const double2*)base )[off]; e.r = v.x; e.i = v.y; break; }
	}
	return e;
}
__device__ __forceinline__ void md_store( void* base, int dt, int64_t off, MdElem e )
{
	switch ( dt )
	{
		case B200_FLOAT:    ( (float*)base )[off] = (float)e.r; break;
		case B200_DOUBLE:   ( (double*)base )[off] = e.r; break;
		case B200_SCOMPLEX: ( (float2*)base )[off] = make_float2( (float)e.r, (float)e.i ); break;
		default:            ( (double2*)base )[off] = make_double2( e.r, e.i ); break;
	}
}

// dst(i,j) [dense, strides rs_d/cs_d, datatype dt_d] := f( src(i,j) ) for an m x n view of src:
// typecast to the precision of dt_d, optional conjugation, optional multiplication by kappa (after the cast, as
// packm's scal2s does), real projection when dt_d is real, optional negation of the imaginary part afterwards.
__global__ void md_convert_kernel( void* dst, int dt_d, int64_t rs_d, int64_t cs_d, const void* src, int dt_s, int64_t rs_s, int64_t cs_s,
                                   int64_t m, int64_t n, int conj, int use_kappa, double kr, double ki, int neg_imag, int single_prec )
{
	const int64_t total = m * n;
	const bool inner_row = ( rs_d <= cs_d );
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		MdElem v = md_load( src, dt_s, i * rs_s + j * cs_s );
		if ( single_prec ) { v.r = (double)(float)v.r; v.i = (double)(float)v.i; }
		if ( conj ) v.i = -v.i;
		if ( use_kappa )
		{
			MdElem w;
			if ( single_prec ) { w.r = (double)( (float)kr * (float)v.r - (float)ki * (float)v.i ); w.i = (double)( (float)kr * (float)v.i + (float)ki * (float)v.r ); }
			else               { w.r = kr * v.r - ki * v.i; w.i = kr * v.i + ki * v.r; }
			v = w;
		}
		if ( neg_imag ) v.i = -v.i;
		md_store( dst, dt_d, i * rs_d + j * cs_d, v );
	}
}

// C(i,j) := beta * C(i,j) + alpha * T(i,j), evaluated in C's precision on the typecast T (bli_txpbys / bli_taxpbys
// with the C datatype as computation type); beta == 0 does not read C; a real C keeps the real part.
__global__ void md_combine_kernel( void* c, int dt_c, int64_t rs_c, int64_t cs_c, const void* t, int dt_t, int64_t rs_t, int64_t cs_t,
                                   int64_t m, int64_t n, double ar, double ai, double br, double bi, int beta_is_zero )
{
	const int64_t total = m * n;
	const bool inner_row = ( ( rs_c < 0 ? -rs_c : rs_c ) <= ( cs_c < 0 ? -cs_c : cs_c ) );
	const bool c_single = ( dt_c == B200_FLOAT || dt_c == B200_SCOMPLEX );
	const bool c_real = ( dt_c == B200_FLOAT || dt_c == B200_DOUBLE );
	for ( int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x )
	{
		int64_t i, j;
		if ( inner_row ) { i = e % m; j = e / m; } else { j = e % n; i = e / n; }
		MdElem x; x.r = 0.0; x.i = 0.0;
		if ( ar != 0.0 || ai != 0.0 ) x = md_load( t, dt_t, i * rs_t + j * cs_t );       // alpha == 0: T is not read
		MdElem y; y.r = 0.0; y.i = 0.0;
		if ( !beta_is_zero ) y = md_load( c, dt_c, i * rs_c + j * cs_c );
		MdElem o;
		if ( c_single )
		{
			const float xr = (float)x.r, xi = (float)x.i, far_ = (float)ar, fai = (float)ai, fbr = (float)br, fbi = (float)bi;
			float pr = far_ * xr - fai * xi, pi = far_ * xi + fai * xr;
			if ( !beta_is_zero ) { pr += fbr * (float)y.r - fbi * (float)y.i; pi += fbr * (float)y.i + fbi * (float)y.r; }
			o.r = pr; o.i = pi;
		}
		else
		{
			double pr = ar * x.r - ai * x.i, pi = ar * x.i + ai * x.r;
			if ( !beta_is_zero ) { pr += br * y.r - bi * y.i; pi += br * y.i + bi * y.r; }
			o.r = pr; o.i = pi;
		}
		if ( c_real ) o.i = 0.0;
		md_store( c, dt_c, i * rs_c + j * cs_c, o );
	}
}

static inline bool dt_is_real( int dt ) { return dt == B200_FLOAT || dt == B200_DOUBLE; }
static inline size_t dt_size( int dt ) { return dt == B200_FLOAT ? 4 : dt == B200_DCOMPLEX ? 16 : 8; }
static inline int dt_make( bool real, bool single ) { return real ? ( single ? B200_FLOAT : B200_DOUBLE ) : ( single ? B200_SCOMPLEX : B200_DCOMPLEX ); }

static int md_convert( void* dst, int dt_d, int64_t rs_d, int64_t cs_d, const void* src, int dt_s, int64_t rs_s, int64_t cs_s,
                       int64_t m, int64_t n, bool conj, bool use_kappa, double kr, double ki, bool neg_imag, cudaStream_t st )
{
	const int64_t total = m * n;
	if ( total <= 0 ) return kSuccess;
	const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
	const int single = ( dt_d == B200_FLOAT || dt_d == B200_SCOMPLEX ) ? 1 : 0;
	md_convert_kernel<<<blocks, 256, 0, st>>>( dst, dt_d, rs_d, cs_d, src, dt_s, rs_s, cs_s, m, n, conj ? 1 : 0, use_kappa ? 1 : 0, kr, ki, neg_imag ? 1 : 0, single );
	B200_CUDA( cudaGetLastError() );
	ctx().launches++;
	return kSuccess;
}

// Real or complex homogeneous product T := X * Y (beta = 0) of the computation precision on dense device operands.
static int md_gemm( bool real, bool single, int64_t m, int64_t n, int64_t k, const void* x, int64_t rs_x, int64_t cs_x,
                    const void* y, int64_t rs_y, int64_t cs_y, void* t, int64_t rs_t, int64_t cs_t, cudaStream_t st )
{
	if ( real && single )   return gemm_dev<float>  ( false, false, m, n, k, 1.f, (const float*)x, rs_x, cs_x, (const float*)y, rs_y, cs_y, 0.f, (float*)t, rs_t, cs_t, st );
	if ( real )             return gemm_dev<double> ( false, false, m, n, k, 1.0, (const double*)x, rs_x, cs_x, (const double*)y, rs_y, cs_y, 0.0, (double*)t, rs_t, cs_t, st );
	if ( single )           return gemm_dev<float2> ( false, false, m, n, k, make_float2( 1.f, 0.f ), (const float2*)x, rs_x, cs_x, (const float2*)y, rs_y, cs_y, make_float2( 0.f, 0.f ), (float2*)t, rs_t, cs_t, st );
	return gemm_dev<double2>( false, false, m, n, k, make_double2( 1.0, 0.0 ), (const double2*)x, rs_x, cs_x, (const double2*)y, rs_y, cs_y, make_double2( 0.0, 0.0 ), (double2*)t, rs_t, cs_t, st );
}

static int gemm_md_front( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb, int64_t m, int64_t n, int64_t k,
                          const double* alpha, const void* a, int64_t rs_a, int64_t cs_a, const void* b, int64_t rs_b, int64_t cs_b,
                          const double* beta, void* c, int64_t rs_c, int64_t cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	for ( int dt : { dt_a, dt_b, dt_c } ) if ( dt < 0 || dt > 3 ) return fail( "b200_gemm_md: unsupported datatype %d", dt );
	if ( comp_prec != 0 && comp_prec != 2 ) return fail( "b200_gemm_md: computation precision must be BLIS_SINGLE_PREC (0) or BLIS_DOUBLE_PREC (2)" );
	if ( m < 0 || n < 0 || k < 0 ) return fail( "b200_gemm_md: negative dimension" );
	if ( !alpha || !beta ) return fail( "b200_gemm_md: alpha/beta must be non-NULL host pointers (dcomplex)" );
	if ( m == 0 || n == 0 ) return kSuccess;
	cudaStream_t st = cur_stream();
	const bool a_real = dt_is_real( dt_a ), b_real = dt_is_real( dt_b ), c_real = dt_is_real( dt_c ), single = ( comp_prec == 0 );
	// alpha lives in the computation precision, complex if any operand is; beta in C's datatype (bli_gemm_cntl.c:174-189)
	double ar = alpha[0], ai = ( a_real && b_real && c_real ) ? 0.0 : alpha[1];
	double br = beta[0],  bi = c_real ? 0.0 : beta[1];
	if ( single ) { ar = (double)(float)ar; ai = (double)(float)ai; }
	if ( dt_c == B200_FLOAT || dt_c == B200_SCOMPLEX ) { br = (double)(float)br; bi = (double)(float)bi; }
	const bool beta_zero = ( br == 0.0 && bi == 0.0 );

	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool conja = !a_real && ( transa & B200_CONJ_NO_TRANSPOSE ), conjb = !b_real && ( transb & B200_CONJ_NO_TRANSPOSE );

	void *da = nullptr, *db = nullptr, *dc = nullptr, *pa = nullptr, *pb = nullptr, *pt = nullptr;
	int rc = kSuccess;
	// host operands: raw bytes to the device first
	const bool c_host = ( classify( c ) != MemKind::Device );
	const bool trivial = ( k == 0 || ( ar == 0.0 && ai == 0.0 ) );
	void* cdev = c; int64_t rs_cd = rs_c, cs_cd = cs_c;
	if ( c_host )
	{
		if ( dev_alloc( &dc, (size_t)m * n * dt_size( dt_c ), st ) != kSuccess ) return kFailure;
		if ( !beta_zero ) rc = stage_to_device( dc, c, m, n, rs_c, cs_c, dt_size( dt_c ), st );
		cdev = dc; rs_cd = 1; cs_cd = m;
	}
	if ( trivial )
	{
		// bli_l3_return_early_if_trivial: C := beta * C
		if ( rc == kSuccess )
		{
			const int blocks = (int)std::min<int64_t>( ( m * n + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
			md_combine_kernel<<<blocks, 256, 0, st>>>( cdev, dt_c, rs_cd, cs_cd, cdev, dt_c, rs_cd, cs_cd, m, n, 0.0, 0.0, br, bi, beta_zero ? 1 : 0 );
			if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm_md: launch failed" );
			ctx().launches++;
		}
	}
	else
	{
		if ( rc == kSuccess && classify( a ) != MemKind::Device )
		{
			if ( dev_alloc( &da, (size_t)m * k * dt_size( dt_a ), st ) != kSuccess ) rc = kFailure;
			else rc = stage_to_device( da, a, m, k, rs_a, cs_a, dt_size( dt_a ), st );
			a = da; rs_a = 1; cs_a = m;
		}
		if ( rc == kSuccess && classify( b ) != MemKind::Device )
		{
			if ( dev_alloc( &db, (size_t)k * n * dt_size( dt_b ), st ) != kSuccess ) rc = kFailure;
			else rc = stage_to_device( db, b, k, n, rs_b, cs_b, dt_size( dt_b ), st );
			b = db; rs_b = 1; cs_b = k;
		}
		const size_t es_r = single ? 4 : 8, es_z = 2 * es_r;
		const int dt_r = dt_make( true, single ), dt_z = dt_make( false, single );
		int dt_t = dt_r; int64_t rs_t = 1, cs_t = m; double car = ar, cai = ai;      // T layout and the alpha left for the combine step
		if ( rc == kSuccess && ( dev_alloc( &pa, (size_t)m * k * es_z, st ) != kSuccess || dev_alloc( &pb, (size_t)k * n * es_z, st ) != kSuccess ||
		                         dev_alloc( &pt, (size_t)m * n * es_z, st ) != kSuccess ) ) rc = kFailure;
		if ( rc == kSuccess )
		{
			if ( a_real && b_real )
			{
				// R*R (C real or complex): real product, alpha (complex when C is) applied by the combine step
				rc = md_convert( pa, dt_r, 1, m, a, dt_a, rs_a, cs_a, m, k, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_r, 1, k, b, dt_b, rs_b, cs_b, k, n, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, m, n, k, pa, 1, m, pb, 1, k, pt, 1, m, st );
			}
			else if ( !a_real && !b_real && !c_real )
			{
				rc = md_convert( pa, dt_z, 1, m, a, dt_a, rs_a, cs_a, m, k, conja, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_z, 1, k, b, dt_b, rs_b, cs_b, k, n, conjb, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( false, single, m, n, k, pa, 1, m, pb, 1, k, pt, 1, m, st );
				dt_t = dt_z;
			}
			else if ( !c_real && !a_real && b_real )
			{
				// C += C*R: the complex A is a real matrix with 2m rows (interleaved re/im), T likewise: 4mnk flops
				rc = md_convert( pa, dt_z, 1, m, a, dt_a, rs_a, cs_a, m, k, conja, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_r, 1, k, b, dt_b, rs_b, cs_b, k, n, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, 2 * m, n, k, pa, 1, 2 * m, pb, 1, k, pt, 1, 2 * m, st );
				dt_t = dt_z;
			}
			else if ( !c_real && a_real && !b_real )
			{
				// C += R*C: transposed, T^T = B^T A^T with B^T a real matrix with 2n rows; T comes out row-major
				rc = md_convert( pb, dt_z, 1, n, b, dt_b, cs_b, rs_b, n, k, conjb, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_convert( pa, dt_r, 1, k, a, dt_a, cs_a, rs_a, k, m, false, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, 2 * n, m, k, pb, 1, 2 * n, pa, 1, k, pt, 1, 2 * n, st );
				dt_t = dt_z; rs_t = n; cs_t = 1;
			}
			else if ( c_real && !a_real && !b_real )
			{
				// R += C*C: T = Re( alpha*A * B ) = [ Re | -Im ]( alpha*A ) * [ Re ; Im ]( B ): a real product with 2k inner
				// dimension (the reference's 1r packing with one operand conjugated, bli_gemm_cntl.c:349-366): 4mnk flops
				rc = md_convert( pa, dt_z, k, 1, a, dt_a, rs_a, cs_a, m, k, conja, true, ar, ai, true, st );      // row-major m x k
				if ( rc == kSuccess ) rc = md_convert( pb, dt_z, 1, k, b, dt_b, rs_b, cs_b, k, n, conjb, false, 0, 0, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, m, n, 2 * k, pa, 2 * k, 1, pb, 1, 2 * k, pt, 1, m, st );
				car = 1.0; cai = 0.0;
			}
			else
			{
				// R += C*R or R += R*C: only the real part of ( alpha * the complex operand ) takes part
				// (BLIS_PACKED_PANELS_RO, bli_gemm_cntl.c:374-389)
				rc = md_convert( pa, dt_r, 1, m, a, dt_a, rs_a, cs_a, m, k, conja, !a_real, ar, ai, false, st );
				if ( rc == kSuccess ) rc = md_convert( pb, dt_r, 1, k, b, dt_b, rs_b, cs_b, k, n, conjb, !b_real, ar, ai, false, st );
				if ( rc == kSuccess ) rc = md_gemm( true, single, m, n, k, pa, 1, m, pb, 1, k, pt, 1, m, st );
				car = 1.0; cai = 0.0;
			}
		}
		if ( rc == kSuccess )
		{
			const int blocks = (int)std::min<int64_t>( ( m * n + 255 ) / 256, (int64_t)ctx().num_sms * 16 );
			md_combine_kernel<<<blocks, 256, 0, st>>>( cdev, dt_c, rs_cd, cs_cd, pt, dt_t, rs_t, cs_t, m, n, car, cai, br, bi, beta_zero ? 1 : 0 );
			if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm_md: launch failed" );
			ctx().launches++;
		}
	}
	if ( rc == kSuccess && c_host )
	{
		rc = stage_to_host( c, rs_c, cs_c, dc, m, n, dt_size( dt_c ), st );
		if ( rc == kSuccess && cudaStreamSynchronize( st ) != cudaSuccess ) rc = fail( "b200_gemm_md: stream sync failed" );
	}
	dev_free( da, st ); dev_free( db, st ); dev_free( dc, st ); dev_free( pa, st ); dev_free( pb, st ); dev_free( pt, st );
	return rc;
}

} // namespace b200

extern "C" b200_err_t b200_gemm_md( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb,
	b200_dim_t m, b200_dim_t n, b200_dim_t k, const double* alpha, const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
	const void* b, b200_inc_t rs_b, b200_inc_t cs_b, const double* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{
	return gemm_md_front( dt_a, dt_b, dt_c, comp_prec, transa, transb, m, n, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c );
}

// ---- batched gemm --------------------------------------------------------
// ?gemm_batch_ / cblas_?gemm_batch (frame/compat/extra/bla_gemm_batch.c:44-131): group i holds group_size[i]
// independent problems with the same shape, transposition and scalars; the reference loops over them calling
// bli_?gemm_ex one after the other.  Here the problems of a batch whose operands are device resident (or pinned)
// are issued round-robin on a pool of streams, so that small problems, which cannot fill 148 SMs one at a time,
// run side by side; problems with pageable host operands take the ordinary (synchronous, staged) path.
namespace b200 {

template <typename T>
static int gemm_batch_front( int group_count, const int* group_size, const int* transa, const int* transb,
                             const int64_t* m, const int64_t* n, const int64_t* k, const T* alpha,
                             const T* const* a, const int64_t* rs_a, const int64_t* cs_a,
                             const T* const* b, const int64_t* rs_b, const int64_t* cs_b,
                             const T* beta, T* const* c, const int64_t* rs_c, const int64_t* cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( group_count < 0 ) return fail( "b200_gemm_batch: negative group count" );
	if ( group_count == 0 ) return kSuccess;
	if ( !group_size || !transa || !transb || !m || !n || !k || !alpha || !beta || !a || !b || !c ||
	     !rs_a || !cs_a || !rs_b || !cs_b || !rs_c || !cs_c ) return fail( "b200_gemm_batch: NULL argument array" );
	Context& cx = ctx();
	cudaStream_t st = cur_stream();
	std::lock_guard<std::mutex> lock( cx.batch_mu );          // one batch at a time owns the stream pool
	int rc = kSuccess;
	B200_CUDA( cudaEventRecord( cx.batch_fork, st ) );
	for ( int s = 0; s < Context::kBatchStreams; ++s ) B200_CUDA( cudaStreamWaitEvent( cx.batch_streams[s], cx.batch_fork, 0 ) );
	int64_t idx = 0; int next = 0;
	for ( int g = 0; g < group_count && rc == kSuccess; ++g )
	{
		if ( group_size[g] < 0 || m[g] < 0 || n[g] < 0 || k[g] < 0 ) { rc = fail( "b200_gemm_batch: negative size in group %d", g ); break; }
		int64_t ra = rs_a[g], ca = cs_a[g], rb = rs_b[g], cb = cs_b[g];
		if ( transa[g] & B200_TRANSPOSE ) std::swap( ra, ca );
		if ( transb[g] & B200_TRANSPOSE ) std::swap( rb, cb );
		const bool conja = Elem<T>::cplx && ( transa[g] & B200_CONJ_NO_TRANSPOSE ), conjb = Elem<T>::cplx && ( transb[g] & B200_CONJ_NO_TRANSPOSE );
		const bool need_ab = ( k[g] > 0 && !Scalar<T>::is_zero( alpha[g] ) );
		for ( int j = 0; j < group_size[g] && rc == kSuccess; ++j, ++idx )
		{
			if ( m[g] == 0 || n[g] == 0 ) continue;
			const bool on_device = classify( c[idx] ) == MemKind::Device &&
			                       ( !need_ab || ( classify( a[idx] ) == MemKind::Device && classify( b[idx] ) == MemKind::Device ) );
			if ( on_device )
			{
				cudaStream_t bs = cx.batch_streams[next]; next = ( next + 1 ) % Context::kBatchStreams;
				rc = gemm_dev<T>( conja, conjb, m[g], n[g], k[g], alpha[g], a[idx], ra, ca, b[idx], rb, cb, beta[g], c[idx], rs_c[g], cs_c[g], bs );
			}
			else
				rc = gemm_front<T>( transa[g], transb[g], m[g], n[g], k[g], alpha + g, a[idx], rs_a[g], cs_a[g], b[idx], rs_b[g], cs_b[g],
				                    beta + g, c[idx], rs_c[g], cs_c[g] );
		}
	}
	// join: the caller's stream continues after every pool stream has drained
	for ( int s = 0; s < Context::kBatchStreams; ++s )
	{
		cudaEventRecord( cx.batch_join[s], cx.batch_streams[s] );
		cudaStreamWaitEvent( st, cx.batch_join[s], 0 );
	}
	return rc;
}

} // namespace b200

extern "C" b200_err_t b200_gemm_batch( int dt, int group_count, const int* group_size, const int* transa, const int* transb,
	const b200_dim_t* m, const b200_dim_t* n, const b200_dim_t* k, const void* alpha,
	const void* const* a, const b200_inc_t* rs_a, const b200_inc_t* cs_a,
	const void* const* b, const b200_inc_t* rs_b, const b200_inc_t* cs_b,
	const void* beta, void* const* c, const b200_inc_t* rs_c, const b200_inc_t* cs_c )
{
	switch ( dt )
	{
		case B200_FLOAT:    return gemm_batch_front<float>  ( group_count, group_size, transa, transb, m, n, k, (const float*)alpha,   (const float* const*)a,   rs_a, cs_a, (const float* const*)b,   rs_b, cs_b, (const float*)beta,   (float* const*)c,   rs_c, cs_c );
		case B200_DOUBLE:   return gemm_batch_front<double> ( group_count, group_size, transa, transb, m, n, k, (const double*)alpha,  (const double* const*)a,  rs_a, cs_a, (const double* const*)b,  rs_b, cs_b, (const double*)beta,  (double* const*)c,  rs_c, cs_c );
		case B200_SCOMPLEX: return gemm_batch_front<float2> ( group_count, group_size, transa, transb, m, n, k, (const float2*)alpha,  (const float2* const*)a,  rs_a, cs_a, (const float2* const*)b,  rs_b, cs_b, (const float2*)beta,  (float2* const*)c,  rs_c, cs_c );
		case B200_DCOMPLEX: return gemm_batch_front<double2>( group_count, group_size, transa, transb, m, n, k, (const double2*)alpha, (const double2* const*)a, rs_a, cs_a, (const double2* const*)b, rs_b, cs_b, (const double2*)beta, (double2* const*)c, rs_c, cs_c );
	}
	return fail( "b200_gemm_batch: unsupported datatype %d", dt );
}

template <typename T>
static int kpanels_front( int transa, int transb, int64_t m, int64_t n, int64_t k, int npanels, const T* alpha,
                          const T* const* a, int64_t rs_a, int64_t cs_a, const T* const* b, int64_t rs_b, int64_t cs_b,
                          const T* beta, T* c, int64_t rs_c, int64_t cs_c )
{
	if ( ensure_init() != kSuccess ) return kFailure;
	if ( npanels < 1 || npanels > 8 ) return fail( "b200_gemm_kpanels: 1..8 panels per call" );
	if ( m < 0 || n < 0 || k < 0 ) return fail( "b200_gemm_kpanels: negative dimension" );
	if ( m == 0 || n == 0 ) return kSuccess;
	if ( transa & B200_TRANSPOSE ) std::swap( rs_a, cs_a );
	if ( transb & B200_TRANSPOSE ) std::swap( rs_b, cs_b );
	const bool conja = Elem<T>::cplx && ( transa & B200_CONJ_NO_TRANSPOSE );
	const bool conjb = Elem<T>::cplx && ( transb & B200_CONJ_NO_TRANSPOSE );
	for ( int sgm = 0; sgm < npanels; ++sgm )
		if ( classify( a[sgm] ) != MemKind::Device || classify( b[sgm] ) != MemKind::Device )
			return fail( "b200_gemm_kpanels: panels must be device resident" );
	if ( classify( c ) != MemKind::Device ) return fail( "b200_gemm_kpanels: C must be device resident" );
	return gemm_dev<T>( conja, conjb, m, n, k, *alpha, a[0], rs_a, cs_a, b[0], rs_b, cs_b, *beta, c, rs_c, cs_c,
	                    cur_stream(), npanels, a + 1, b + 1 );
}

extern "C" b200_err_t b200_gemm_kpanels( int dt, int transa, int transb, b200_dim_t m, b200_dim_t n, b200_dim_t k, int npanels,
	const void* alpha, const void* const* a, b200_inc_t rs_a, b200_inc_t cs_a,
	const void* const* b, b200_inc_t rs_b, b200_inc_t cs_b, const void* beta, void* c, b200_inc_t rs_c, b200_inc_t cs_c )
{
	if ( dt == B200_DOUBLE )
		return kpanels_front<double>( transa, transb, m, n, k, npanels, (const double*)alpha, (const double* const*)a, rs_a, cs_a,
		                              (const double* const*)b, rs_b, cs_b, (const double*)beta, (double*)c, rs_c, cs_c );
	if ( dt == B200_DCOMPLEX )
		return kpanels_front<double2>( transa, transb, m, n, k, npanels, (const double2*)alpha, (const double2* const*)a, rs_a, cs_a,
		                               (const double2* const*)b, rs_b, cs_b, (const double2*)beta, (double2*)c, rs_c, cs_c );
	return fail( "b200_gemm_kpanels: only d and z are supported" );
}

extern "C" b200_dim_t b200_blksz( int dt, int bs )
{
	auto pick = [&]( int mr, int nr, int mc, int kc, int nc ) -> b200_dim_t
	{
		switch ( bs ) { case B200_BS_MR: return mr; case B200_BS_NR: return nr; case B200_BS_MC: return mc;
		                case B200_BS_KC: return kc; case B200_BS_NC: return nc; }
		return -1;
	};
	switch ( dt )
	{
		case B200_FLOAT:    return pick( Tiles<float>::MR,   Tiles<float>::NR,   Tiles<float>::BQ,   Tiles<float>::BK,   Tiles<float>::BP );
		case B200_DOUBLE:   return pick( Tiles<double>::MR,  Tiles<double>::NR,  Tiles<double>::BQ,  Tiles<double>::BK,  Tiles<double>::BP );
		case B200_SCOMPLEX: return pick( Tiles<float2>::MR,  Tiles<float2>::NR,  Tiles<float2>::BQ,  Tiles<float2>::BK,  Tiles<float2>::BP );
		case B200_DCOMPLEX: return pick( Tiles<double2>::MR, Tiles<double2>::NR, Tiles<double2>::BQ, Tiles<double2>::BK, Tiles<double2>::BP );
	}
	return -1;
}

extern "C" unsigned long long b200_launch_count( void ) { return ctx().launches.load(); }

// Tuning knobs for sweeps (not part of the reference surface).
extern "C" b200_err_t b200_set_option( const char* key, long long value )
{
	Context& c = ctx();
	if      ( !strcmp( key, "dgemm_cfg" ) ) c.dgemm_cfg = (int)value;
	else if ( !strcmp( key, "zgemm_cfg" ) ) c.zgemm_cfg = (int)value;
	else if ( !strcmp( key, "sgemm_cfg" ) ) c.sgemm_cfg = (int)value;
	else if ( !strcmp( key, "cgemm_cfg" ) ) c.cgemm_cfg = (int)value;
	else if ( !strcmp( key, "grid_mult" ) ) c.grid_mult = (int)std::max<long long>( 1, value );
	else if ( !strcmp( key, "dynamic_tiles" ) ) c.dynamic_tiles = (int)value;
	else if ( !strcmp( key, "transpose_y" ) ) c.transpose_y = (int)value;
	else if ( !strcmp( key, "ktri_skip" ) ) c.ktri_skip = (int)value;
	else if ( !strcmp( key, "host_kpipe" ) ) c.host_kpipe = (int)value;
	else if ( !strcmp( key, "dmma_cst" ) ) c.dmma_cst = (int)value;
	else if ( !strcmp( key, "tma_l2_promotion" ) ) c.tma_l2_promotion = (int)std::min<long long>( 3, std::max<long long>( 0, value ) );
	else if ( !strcmp( key, "raster_group" ) ) c.raster_group = (int)std::max<long long>( 1, value );
	else if ( !strcmp( key, "reserve_sms" ) )
	{
		// leave SMs free for concurrently running communication kernels (multi-GPU overlap)
		cudaDeviceProp prop; int dev = 0; cudaGetDevice( &dev ); cudaGetDeviceProperties( &prop, dev );
		c.num_sms = std::max( 1, prop.multiProcessorCount - (int)std::max<long long>( 0, value ) );
	}
	else return fail( "b200_set_option: unknown key %s", key );
	return kSuccess;
}
