/* blis_oracle.c -- CPU restatement of the reference BLIS gemm/trsm path.
 * TEST INFRASTRUCTURE ONLY: see blis_oracle.h for scope, pinning and rules. */
#include <stdlib.h>
#include <string.h>
#include "blis_oracle.h"

/* ---- blocksizes: reference-context defaults (ref_kernels/bli_cntx_ref.c:377-384,
   frame/include/bli_kernel_macro_defs.h:260-289); the reference gemm ukr has
   no row preference (bli_cntx_ref.c, BLIS_GEMM_UKR_ROW_PREF = FALSE). */
static orc_blksz_t g_bs[4] = {
	/* s */ { 4, 16, 256, 256, 4096, 0 },
	/* c */ { 4,  8, 128, 256, 4096, 0 },
	/* d */ { 4,  8, 128, 256, 4096, 0 },
	/* z */ { 4,  4,  64, 256, 4096, 0 },
};
void orc_set_blksz( int dt, dim_t mr, dim_t nr, dim_t mc, dim_t kc, dim_t nc, int row_pref )
{
	orc_blksz_t b = { mr, nr, mc, kc, nc, row_pref };
	g_bs[ dt & 3 ] = b;
}
void orc_get_blksz( int dt, orc_blksz_t* out ) { *out = g_bs[ dt & 3 ]; }

/* frame/base/bli_blksz.c:236-282 */
dim_t orc_determine_blocksize( int backward, dim_t i, dim_t dim, dim_t b_alg, dim_t b_max )
{
	const dim_t dim_left_now = dim - i;
	if ( backward )
	{
		const dim_t dim_at_edge = dim_left_now % b_alg;
		if ( b_alg + dim_at_edge <= b_max ) return b_alg + dim_at_edge;
		return dim_at_edge;
	}
	if ( dim_left_now <= b_max ) return dim_left_now;
	return b_alg;
}

/* frame/thread/bli_thread_range.c:38-184 */
void orc_thread_range_sub( dim_t work_id, dim_t n_way, dim_t n, dim_t bf, int handle_edge_low,
                           dim_t* start, dim_t* end )
{
	if ( n_way == 1 ) { *start = 0; *end = n; return; }
	const dim_t n_bf_whole = n / bf, n_bf_left = n % bf;
	dim_t n_bf_lo = n_bf_whole / n_way, n_bf_hi = n_bf_whole / n_way;
	if ( !handle_edge_low )
	{
		const dim_t n_th_lo = n_bf_whole % n_way;
		if ( n_th_lo != 0 ) n_bf_lo += 1;
		const dim_t size_lo = n_bf_lo * bf, size_hi = n_bf_hi * bf;
		const dim_t hi_start = n_th_lo * size_lo;
		if ( work_id < n_th_lo ) { *start = work_id * size_lo; *end = ( work_id + 1 ) * size_lo; }
		else
		{
			*start = hi_start + ( work_id - n_th_lo ) * size_hi;
			*end   = hi_start + ( work_id - n_th_lo + 1 ) * size_hi;
			if ( work_id == n_way - 1 ) *end += n_bf_left;
		}
	}
	else
	{
		const dim_t n_th_hi = n_bf_whole % n_way, n_th_lo = n_way - n_th_hi;
		if ( n_th_hi != 0 ) n_bf_hi += 1;
		const dim_t size_lo = n_bf_lo * bf, size_hi = n_bf_hi * bf;
		const dim_t hi_start = n_th_lo * size_lo + n_bf_left;
		if ( work_id < n_th_lo )
		{
			*start = work_id * size_lo; *end = ( work_id + 1 ) * size_lo;
			if ( work_id == 0 ) *end += n_bf_left;
			else { *start += n_bf_left; *end += n_bf_left; }
		}
		else
		{
			*start = hi_start + ( work_id - n_th_lo ) * size_hi;
			*end   = hi_start + ( work_id - n_th_lo + 1 ) * size_hi;
		}
	}
}

/* frame/thread/bli_thread.c:194-320 (fast heuristic; prime factors ascending) */
void orc_thread_partition_2x2( dim_t n_thread, dim_t work1, dim_t work2, dim_t* nt1, dim_t* nt2 )
{
	if ( n_thread < 4 )
	{
		*nt1 = ( work1 >= work2 ? n_thread : 1 );
		*nt2 = ( work1 <  work2 ? n_thread : 1 );
		return;
	}
	dim_t tn1 = 1, tn2 = 1, rem = n_thread, f = 2;
	while ( rem > 1 )
	{
		while ( rem % f != 0 ) ++f;
		rem /= f;
		if ( work1 > work2 ) { work1 /= f; tn1 *= f; }
		else                 { work2 /= f; tn2 *= f; }
	}
	if ( work1 > work2 )
	{
		if ( tn2 % 2 == 0 )
		{
			const dim_t diff = work1 - work2;
			dim_t diff_mod = work1 / 2 - work2 * 2; if ( diff_mod < 0 ) diff_mod = -diff_mod;
			if ( diff_mod < diff ) { tn1 *= 2; tn2 /= 2; }
		}
	}
	else if ( work1 < work2 )
	{
		if ( tn1 % 2 == 0 )
		{
			const dim_t diff = work2 - work1;
			dim_t diff_mod = work2 / 2 - work1 * 2; if ( diff_mod < 0 ) diff_mod = -diff_mod;
			if ( diff_mod < diff ) { tn1 /= 2; tn2 *= 2; }
		}
	}
	*nt1 = tn1; *nt2 = tn2;
}

/* bli_align_dim_to_mult( dim, mult, round up ) (frame/base/bli_blksz.c) */
dim_t orc_align_dim_to_mult( dim_t dim, dim_t mult )
{
	if ( mult <= 0 ) return dim;
	return ( ( dim + mult - 1 ) / mult ) * mult;
}

/* panel stride: ldp * padded length, bumped to even (bli_packm_init.c:147-158,
   bli_packm_blk_var1.c:253-254, bli_trsm_ll_ker_var2.c:250-252) */
dim_t orc_packm_panel_stride( dim_t ldp, dim_t panel_len_max )
{
	dim_t ps = ldp * panel_len_max;
	if ( ps % 2 != 0 ) ps += 1;
	return ps;
}

/* ---- four instantiations of the generic body ---- */
#define ORC_CAT_( a, b ) a##b
#define ORC_CAT( a, b ) ORC_CAT_( a, b )
#define ELT    FN(elt_t)
#define E_ld   FN(E_ld)
#define E_st   FN(E_st)
#define E_mk   FN(E_mk)
#define E_conj FN(E_conj)
#define E_is0  FN(E_is0)
#define E_mul  FN(E_mul)
#define E_axpy FN(E_axpy)
#define E_xpby FN(E_xpby)
#define E_sub  FN(E_sub)
#define E_inv  FN(E_inv)

#define R float
#define CPLX 0
#define ORC_DT 0
#define FN( name ) ORC_CAT( orc_s, name )
#include "blis_oracle_t.inc"
#undef R
#undef CPLX
#undef ORC_DT
#undef FN

#define R double
#define CPLX 0
#define ORC_DT 2
#define FN( name ) ORC_CAT( orc_d, name )
#include "blis_oracle_t.inc"
#undef R
#undef CPLX
#undef ORC_DT
#undef FN

#define R float
#define CPLX 1
#define ORC_DT 1
#define FN( name ) ORC_CAT( orc_c, name )
#include "blis_oracle_t.inc"
#undef R
#undef CPLX
#undef ORC_DT
#undef FN

#define R double
#define CPLX 1
#define ORC_DT 3
#define FN( name ) ORC_CAT( orc_z, name )
#include "blis_oracle_t.inc"
#undef R
#undef CPLX
#undef ORC_DT
#undef FN

/* ------------------------------------------------------------------ mixed-datatype gemm */

/* Restatement of bli_gemm_ex for operands of different domain/precision
   (docs/MixedDatatypes.md; frame/3/gemm/bli_gemm_cntl.c:87-392):
   - dt_comp = ( all domains equal ? domain(C) : real ) | comp_prec (:99);  alpha is cast to ( complex if any operand is |
     comp_prec ), beta to dt_c (:174-189);
   - alpha is attached to B, or to A when alpha's domain is complex, A is complex and B is real (:194-206); an attached
     scalar with a non-zero imaginary part is applied while packing a complex (or real-only packed) operand
     (frame/1m/packm/bli_packm_scalar.c:46-68), otherwise by the microkernel;
   - A and B are typecast to comp_prec by the packing kernels (packm_struc_cxk[dt][dt_p], :253-254);
   - domain cases (:297-389): C+=C*R / C+=R*C treat the complex operand as a real matrix (no cross terms);
     R+=C*C packs both operands 1r with one of them conjugated and keeps Re(A*B) with a doubled k (KC halved);
     R+=C*R / R+=R*C pack only the real part (BLIS_PACKED_PANELS_RO) of ( scalar * complex operand );
     C+=R*R computes the real product and the crr wrapper applies the complex alpha
     (ref_kernels/ind/bli_gemm_crr_ref.c:98-117);
   - per KC block: ct = (alpha) * A_blk * B_blk in comp_prec from zero, then C := beta*C + cast(ct) evaluated in C's
     precision (ref_kernels/3/bli_gemm_ref.c:358-383), beta becoming one after the first block.
   The arithmetic is carried in float or double variables according to comp_prec / the precision of C. */
#define ORC_MD_CORE( NAME, RT ) \
static void NAME( int a_real, int b_real, int c_real, dim_t m, dim_t n, dim_t kb, const double* ap, const double* bp, \
                  double ukr_ar, double ukr_ai, double* ct ) \
{ \
	/* ap: m x kb, bp: kb x n, both dense column-major (re,im) pairs already in the computation precision */ \
	for ( dim_t j = 0; j < n; ++j ) for ( dim_t i = 0; i < m; ++i ) \
	{ \
		RT sr = 0, si = 0; \
		for ( dim_t l = 0; l < kb; ++l ) \
		{ \
			const RT xr = (RT)ap[ 2 * ( i + l * m ) ], xi = (RT)ap[ 2 * ( i + l * m ) + 1 ]; \
			const RT yr = (RT)bp[ 2 * ( l + j * kb ) ], yi = (RT)bp[ 2 * ( l + j * kb ) + 1 ]; \
			if ( a_real && b_real )        { sr += xr * yr; } \
			else if ( !a_real && b_real )  { sr += xr * yr; si += xi * yr; }            /* complex operand as a real matrix */ \
			else if ( a_real && !b_real )  { sr += xr * yr; si += xr * yi; } \
			else if ( c_real )             { sr += xr * yr; sr += ( -xi ) * yi; }         /* 1r packing, one operand conjugated */ \
			else                           { sr += xr * yr - xi * yi; si += xi * yr + xr * yi; } \
		} \
		const RT ar = (RT)ukr_ar, ai = (RT)ukr_ai; \
		RT pr, pi; \
		if ( ai == 0 ) { pr = ar * sr; pi = ar * si; } else { pr = ar * sr - ai * si; pi = ar * si + ai * sr; } \
		ct[ 2 * ( i + j * m ) ] = (double)pr; ct[ 2 * ( i + j * m ) + 1 ] = (double)pi; \
	} \
}
ORC_MD_CORE( orc_md_core_s, float )
ORC_MD_CORE( orc_md_core_d, double )

static void orc_md_ld( const void* p, int dt, inc_t off, double* r, double* i )
{
	*i = 0.0;
	if      ( dt == 0 ) *r = ( (const float*)p )[off];
	else if ( dt == 2 ) *r = ( (const double*)p )[off];
	else if ( dt == 1 ) { *r = ( (const float*)p )[2 * off]; *i = ( (const float*)p )[2 * off + 1]; }
	else                { *r = ( (const double*)p )[2 * off]; *i = ( (const double*)p )[2 * off + 1]; }
}
static void orc_md_st( void* p, int dt, inc_t off, double r, double i )
{
	if      ( dt == 0 ) ( (float*)p )[off] = (float)r;
	else if ( dt == 2 ) ( (double*)p )[off] = r;
	else if ( dt == 1 ) { ( (float*)p )[2 * off] = (float)r; ( (float*)p )[2 * off + 1] = (float)i; }
	else                { ( (double*)p )[2 * off] = r; ( (double*)p )[2 * off + 1] = i; }
}

/* C := beta*C + x in C's precision; beta == 0 overwrites (bli_txpbys_mxn) */
static void orc_md_update( void* c, int dt_c, inc_t off, double xr, double xi, double br, double bi )
{
	const int single = ( dt_c == 0 || dt_c == 1 ), real = ( dt_c == 0 || dt_c == 2 );
	double yr = 0, yi = 0, orr, oi;
	const int beta0 = ( br == 0.0 && bi == 0.0 );
	if ( !beta0 ) orc_md_ld( c, dt_c, off, &yr, &yi );
	if ( single )
	{
		float fr = (float)xr, fi = (float)xi;
		if ( !beta0 ) { fr = fr + ( (float)br * (float)yr - (float)bi * (float)yi ); fi = fi + ( (float)bi * (float)yr + (float)br * (float)yi ); }
		orr = fr; oi = fi;
	}
	else
	{
		orr = xr; oi = xi;
		if ( !beta0 ) { orr = xr + ( br * yr - bi * yi ); oi = xi + ( bi * yr + br * yi ); }
	}
	orc_md_st( c, dt_c, off, orr, real ? 0.0 : oi );
}

void orc_gemm_md( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb, dim_t m, dim_t n, dim_t k,
                  const double* alpha, const void* a, inc_t rs_a, inc_t cs_a, const void* b, inc_t rs_b, inc_t cs_b,
                  const double* beta, void* c, inc_t rs_c, inc_t cs_c )
{
	const int a_real = ( dt_a == 0 || dt_a == 2 ), b_real = ( dt_b == 0 || dt_b == 2 ), c_real = ( dt_c == 0 || dt_c == 2 );
	const int single = ( comp_prec == 0 ), c_single = ( dt_c == 0 || dt_c == 1 );
	if ( m == 0 || n == 0 ) return;
	/* beta in C's datatype */
	double br = beta[0], bi = c_real ? 0.0 : beta[1];
	if ( c_single ) { br = (float)br; bi = (float)bi; }
	/* bli_l3_return_early_if_trivial: alpha == 0 (as given) or k == 0: C := beta*C */
	if ( k == 0 || ( alpha[0] == 0.0 && alpha[1] == 0.0 ) )
	{
		for ( dim_t j = 0; j < n; ++j ) for ( dim_t i = 0; i < m; ++i ) orc_md_update( c, dt_c, i * rs_c + j * cs_c, 0.0, 0.0, br, bi );
		return;
	}
	/* alpha in the computation precision, complex if any operand is */
	const int any_cplx = !( a_real && b_real && c_real );
	double ar = alpha[0], ai = any_cplx ? alpha[1] : 0.0;
	if ( single ) { ar = (float)ar; ai = (float)ai; }
	if ( transa & ORC_TRANS_BIT ) { inc_t t = rs_a; rs_a = cs_a; cs_a = t; }
	if ( transb & ORC_TRANS_BIT ) { inc_t t = rs_b; rs_b = cs_b; cs_b = t; }
	const int conja = !a_real && ( transa & ORC_CONJ_BIT ), conjb = !b_real && ( transb & ORC_CONJ_BIT );
	/* which operand carries alpha, and is it applied while packing? */
	const int on_a = ( any_cplx && !a_real && b_real );
	const int ro_case = c_real && ( a_real != b_real );                 /* R += C*R or R += R*C */
	const int carrier_is_cplx = on_a ? !a_real : !b_real;
	const int at_pack = ( ai != 0.0 ) && ( carrier_is_cplx || ro_case );
	const double ukr_ar = at_pack ? 1.0 : ar, ukr_ai = at_pack ? 0.0 : ai;
	/* typecast (and scale) the operands: "packing" */
	double* ap = (double*)malloc( sizeof(double) * 2 * (size_t)( m * k ) );
	double* bp = (double*)malloc( sizeof(double) * 2 * (size_t)( k * n ) );
	for ( int which = 0; which < 2; ++which )
	{
		const void* src = which ? b : a; const int dt = which ? dt_b : dt_a;
		const inc_t rs = which ? rs_b : rs_a, cs = which ? cs_b : cs_a;
		const dim_t rows = which ? k : m, cols = which ? n : k;
		const int cj = which ? conjb : conja, scale = at_pack && ( which ? !on_a : on_a );
		const int keep_real_only = ro_case && !( which ? b_real : a_real );
		double* dst = which ? bp : ap;
		for ( dim_t j = 0; j < cols; ++j ) for ( dim_t i = 0; i < rows; ++i )
		{
			double r, im; orc_md_ld( src, dt, i * rs + j * cs, &r, &im );
			if ( single ) { r = (float)r; im = (float)im; }
			if ( cj ) im = -im;
			if ( scale )
			{
				double sr, si;
				if ( single ) { sr = (float)( (float)ar * (float)r - (float)ai * (float)im ); si = (float)( (float)ai * (float)r + (float)ar * (float)im ); }
				else          { sr = ar * r - ai * im; si = ai * r + ar * im; }
				r = sr; im = si;
			}
			if ( keep_real_only ) im = 0.0;
			dst[ 2 * ( i + j * rows ) ] = r; dst[ 2 * ( i + j * rows ) + 1 ] = im;
		}
	}
	/* after real-only packing the complex operand is real for the product */
	const int a_real_p = a_real || ro_case, b_real_p = b_real || ro_case;
	/* KC of the computation datatype; halved for R += C*C (kc_scale = 2, :363) */
	const int induced = !( a_real == b_real && b_real == c_real );
	orc_blksz_t bs; orc_get_blksz( ( ( induced || c_real ) ? 0 : 1 ) | ( single ? 0 : 2 ), &bs );
	dim_t kc = bs.kc; if ( c_real && !a_real && !b_real ) kc = kc / 2 > 0 ? kc / 2 : 1;
	double* ct = (double*)malloc( sizeof(double) * 2 * (size_t)( m * n ) );
	double* ablk = (double*)malloc( sizeof(double) * 2 * (size_t)( m * kc ) );
	double* bblk = (double*)malloc( sizeof(double) * 2 * (size_t)( kc * n ) );
	double cur_br = br, cur_bi = bi;
	for ( dim_t pc = 0; pc < k; )
	{
		const dim_t kb = orc_determine_blocksize( 0, pc, k, kc, kc );
		for ( dim_t l = 0; l < kb; ++l ) for ( dim_t i = 0; i < m; ++i )
		{ ablk[ 2 * ( i + l * m ) ] = ap[ 2 * ( i + ( pc + l ) * m ) ]; ablk[ 2 * ( i + l * m ) + 1 ] = ap[ 2 * ( i + ( pc + l ) * m ) + 1 ]; }
		for ( dim_t j = 0; j < n; ++j ) for ( dim_t l = 0; l < kb; ++l )
		{ bblk[ 2 * ( l + j * kb ) ] = bp[ 2 * ( pc + l + j * k ) ]; bblk[ 2 * ( l + j * kb ) + 1 ] = bp[ 2 * ( pc + l + j * k ) + 1 ]; }
		if ( single ) orc_md_core_s( a_real_p, b_real_p, c_real, m, n, kb, ablk, bblk, ukr_ar, ukr_ai, ct );
		else          orc_md_core_d( a_real_p, b_real_p, c_real, m, n, kb, ablk, bblk, ukr_ar, ukr_ai, ct );
		for ( dim_t j = 0; j < n; ++j ) for ( dim_t i = 0; i < m; ++i )
			orc_md_update( c, dt_c, i * rs_c + j * cs_c, ct[ 2 * ( i + j * m ) ], ct[ 2 * ( i + j * m ) + 1 ], cur_br, cur_bi );
		cur_br = 1.0; cur_bi = 0.0;
		pc += kb;
	}
	free( ap ); free( bp ); free( ct ); free( ablk ); free( bblk );
}
