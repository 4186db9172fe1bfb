// host_gemm.cuh -- gemm on device views (gemm_dev), host-operand pipelines, gemm front end
// (host side of the engine; included by capi.cu, which holds the extern "C" entry points)
#pragma once
#include "host_util.cuh"
namespace b200 {

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
	g.tile_counter = sched_slot( st );
	g.sk_full = 0; g.sk_split = 0; g.sk_ws = nullptr; g.sk_flags = nullptr;
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
// and merged (beta) during the final phase, which runs per column block over the LAST TWO panels (one k-panel-accumulation
// launch per block) so that each finished block of C leaves on the D2H stream under the kernels of the next one and the
// downloads keep up.  Exposed transfer: the first, short pair of panels (2.5 ms) and the last, short block of C (2.4 ms).
// [B200] 16384^3 through dgemm_: 266.8 -> ~253 ms per call (kernels alone 241.6 ms); b200_set_option("host_trace", 1) prints
// the timeline.
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
	// (last block of C going home) is short, and the FIRST k panel is cut once more so that the first kernel starts early.
	const int64_t kb = std::max<int64_t>( 512, ( ( k + 7 ) / 8 + 127 ) / 128 * 128 );
	const int64_t nb = std::max<int64_t>( 512, ( ( n + 7 ) / 8 + 127 ) / 128 * 128 );
	std::vector<int64_t> pk{ 0 }, pn{ 0 };
	{
		// the first eighth is cut once more at a quarter (host_kpipe = divisor, 1 = 4): the first kernel starts after
		// 2.5 ms instead of 9.7 ms ([B200] 16384^3, host_trace timeline; 4 KiB lines of B still move at 54 GB/s)
		const int64_t first = kb / ( cx.host_kpipe >= 2 ? cx.host_kpipe : 4 ) / 128 * 128;
		if ( kb >= 1024 && first >= 256 ) pk.push_back( first );
		while ( pk.back() < k ) pk.push_back( std::min( k, pk.back() + kb - ( pk.size() == 2 && pk.back() < kb ? pk.back() : 0 ) ) );
		while ( pn.back() < n ) pn.push_back( std::min( n, pn.back() + nb ) );
		const int64_t n0 = std::max<int64_t>( 512, ( nb / 4 + 127 ) / 128 * 128 );
		if ( pn.size() > 2 && n - pn[pn.size() - 2] > n0 ) pn.insert( pn.end() - 1, n - n0 );
	}
	const int np = (int)pk.size() - 1, nblk = (int)pn.size() - 1;
	void *da = nullptr, *db = nullptr, *dc = nullptr, *ds = nullptr;
	int rc = kSuccess;
	if ( ( a_host && dev_alloc( &da, (size_t)m * k * ES, st ) != kSuccess ) || ( b_host && dev_alloc( &db, (size_t)k * n * ES, st ) != kSuccess ) ||
	     dev_alloc( &dc, (size_t)m * n * ES, st ) != kSuccess || ( load_c && dev_alloc( &ds, (size_t)m * n * ES, st ) != kSuccess ) ) rc = kFailure;
	const bool trace = cx.host_trace != 0;                       // host_trace: print the timeline of this call (ms since the first enqueue) to stderr
	std::vector<cudaEvent_t> ev( np + 2 * nblk + 1 );
	for ( auto& e : ev ) cudaEventCreateWithFlags( &e, trace ? cudaEventDefault : cudaEventDisableTiming );
	std::vector<cudaEvent_t> ev_k( trace ? np + 1 : 0 ), ev_out( trace ? nblk : 0 );
	for ( auto& e : ev_k ) cudaEventCreate( &e );
	for ( auto& e : ev_out ) cudaEventCreate( &e );
	cudaEvent_t* ev_p = ev.data(); cudaEvent_t* ev_c = ev.data() + np; cudaEvent_t* ev_done = ev.data() + np + nblk; cudaEvent_t ev_alloc = ev.back();
	cudaEventRecord( ev_alloc, st );
	cudaStreamWaitEvent( s_in, ev_alloc, 0 ); cudaStreamWaitEvent( s_out, ev_alloc, 0 );
	const T one = Scalar<T>::make( 1.0, 0.0 ), zero = Scalar<T>::make( 0.0, 0.0 );
	// the final phase folds the last two k panels into its per-block launches when they have the same width (d and z: the
	// k-panel accumulation kernels)
	constexpr bool kPanelsOk = std::is_same<T, double>::value || std::is_same<T, double2>::value;
	const int tail_p = ( kPanelsOk && np >= 3 && pk[np] - pk[np - 1] == pk[np - 1] - pk[np - 2] ) ? 2 : 1;
	const T *tail_a = nullptr, *tail_b = nullptr;
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
		if ( p < np - tail_p )
		{
			rc = gemm_dev<T>( conja, conjb, m, n, kw, al, ap, rs_ap, cs_ap, bp, rs_bp, cs_bp, p == 0 ? zero : one, (T*)dc, 1, m, st );
			if ( trace ) cudaEventRecord( ev_k[p], st );
		}
		else if ( p + 1 < np ) { tail_a = ap; tail_b = bp; }          // first panel of the final phase: used together with the last one
		else
		{
			// final phase, per column block: the last tail_p panels in ONE launch (k-panel accumulation), the merge of the
			// host C, and the block leaves.  With two panels a block computes for longer than it travels ([B200] 16384^3:
			// 8.1 ms against 4.7 ms), so the D2H stream keeps up and only the last, short block is exposed; with one panel
			// the downloads fell 9 ms behind (timeline: host_trace).
			rc = send_c( nblk );
			for ( int j = 0; j < nblk && rc == kSuccess; ++j )
			{
				const int64_t j0 = pn[j], w = pn[j + 1] - j0;
				const T beta_j = ( np == tail_p ) ? zero : one;
				if ( tail_p == 2 )
				{
					const T* a_more[1] = { ap };  const T* b_more[1] = { bp + j0 * cs_bp };
					rc = gemm_dev<T>( conja, conjb, m, w, kw, al, tail_a, rs_ap, cs_ap, tail_b + j0 * cs_bp, rs_bp, cs_bp, beta_j, (T*)dc + j0 * m, 1, m, st, 2, a_more, b_more );
				}
				else
					rc = gemm_dev<T>( conja, conjb, m, w, kw, al, ap, rs_ap, cs_ap, bp + j0 * cs_bp, rs_bp, cs_bp, beta_j, (T*)dc + j0 * m, 1, m, st );
				if ( rc == kSuccess && load_c )
				{
					cudaStreamWaitEvent( st, ev_c[j], 0 );
					const int64_t total = m * w;
					const int blocks = (int)std::min<int64_t>( ( total + 255 ) / 256, (int64_t)cx.num_sms * 16 );
					R br, bi; if constexpr ( Elem<T>::cplx ) { br = be.x; bi = be.y; } else { br = be; bi = 0; }
					add_scaled_kernel<R, NC><<<blocks, 256, 0, st>>>( (R*)( (T*)dc + j0 * m ), (const R*)( (const T*)ds + j0 * m ), total, br, bi );
					if ( cudaGetLastError() != cudaSuccess ) rc = fail( "b200_gemm: launch failed" );
					note_launch( "add_scaled_kernel" );
				}
				cudaEventRecord( ev_done[j], st );
			}
			// the blocks go home in a second pass: for a pageable C stage_to_host blocks this thread (it unpacks the pinned
			// ring), and every kernel of the final phase has to be queued before that starts
			for ( int j = 0; j < nblk && rc == kSuccess; ++j )
			{
				const int64_t j0 = pn[j], w = pn[j + 1] - j0;
				cudaStreamWaitEvent( s_out, ev_done[j], 0 );
				rc = stage_to_host( c + j0 * cs_c, rs_c, cs_c, (T*)dc + j0 * m, m, w, ES, s_out );
				if ( trace ) cudaEventRecord( ev_out[j], s_out );
			}
		}
	}
	if ( cudaStreamSynchronize( s_out ) != cudaSuccess || cudaStreamSynchronize( s_in ) != cudaSuccess || cudaStreamSynchronize( st ) != cudaSuccess )
		rc = fail( "b200_gemm: stream sync failed: %s", cudaGetErrorString( cudaGetLastError() ) );
	if ( trace && rc == kSuccess )
	{
		auto ms = [&]( cudaEvent_t e ) { float t = 0.f; cudaEventElapsedTime( &t, ev_alloc, e ); return t; };
		fprintf( stderr, "gemm_host_kpipe %lld x %lld x %lld: panels in at", (long long)m, (long long)n, (long long)k );
		for ( int p = 0; p < np; ++p ) fprintf( stderr, " %.1f", ms( ev_p[p] ) );
		fprintf( stderr, " | C staged at" );
		for ( int j = 0; j < nblk; ++j ) fprintf( stderr, " %.1f", ms( ev_c[j] ) );
		fprintf( stderr, " | rounds done at" );
		for ( int p = 0; p < np - tail_p; ++p ) fprintf( stderr, " %.1f", ms( ev_k[p] ) );
		fprintf( stderr, " | last round blocks done at" );
		for ( int j = 0; j < nblk; ++j ) fprintf( stderr, " %.1f", ms( ev_done[j] ) );
		fprintf( stderr, " | blocks home at" );
		for ( int j = 0; j < nblk; ++j ) fprintf( stderr, " %.1f", ms( ev_out[j] ) );
		fprintf( stderr, " ms\n" );
	}
	for ( auto& e : ev ) cudaEventDestroy( e );
	for ( auto& e : ev_k ) cudaEventDestroy( e );
	for ( auto& e : ev_out ) cudaEventDestroy( e );
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
	bool conja = Elem<T>::cplx && ( transa & B200_CONJ_NO_TRANSPOSE );
	bool conjb = Elem<T>::cplx && ( transb & B200_CONJ_NO_TRANSPOSE );
	// A ROW-stored host C (a row-major caller): everything below -- staging, column-block and k-panel pipelines -- moves
	// columns, and a row-stored host matrix would crawl through the element-wise gather.  Solve the transposed problem
	// instead, C^T = op(B)^T op(A)^T: its C is column-stored, and for a row-major caller so are its A and B.
	if ( tri_operand == 0 && cs_c == 1 && rs_c != 1 && m > 1 && classify( c ) != MemKind::Device )
	{
		std::swap( m, n ); std::swap( a, b ); std::swap( conja, conjb );
		const int64_t ra = rs_a, ca = cs_a;
		rs_a = cs_b; cs_a = rs_b; rs_b = ca; cs_b = ra;
		std::swap( rs_c, cs_c );
	}
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

} // namespace b200
