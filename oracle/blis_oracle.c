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
