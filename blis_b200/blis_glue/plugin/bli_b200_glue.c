/*
   bli_b200_glue.c -- BLIS-side binding of the B200 engine (C99, includes blis.h).

   This is the reference-facing half of the drop-in: everything here speaks
   BLIS types (obj_t, cntx_t, rntm_t, err_t) and forwards to the C ABI in
   include/blis_b200.h.  It is shared by the two integration routes described
   in INTEGRATION.md:

     (1) plugin route  -- bli_plugin_register_b200() installs the whole-operation
         gemm hook on the active context of ANY stock libblis at run time
         (the mechanism of build/plugin/bli_plugin_register.c:37-81 and
         docs/PluginHowTo.md), no rebuild of BLIS needed;
     (2) config route  -- config/b200/bli_cntx_init_b200.c calls
         bli_b200_install( cntx ) while the gks registers the sub-configuration
         (frame/base/bli_gks.c:58-93).

   gemm: the hook has the gemmsup_oft signature (frame/3/bli_l3_sup_oft.h:46-59)
   and is reached from bli_gemm_ex before any host packing
   (frame/3/bli_l3_oapi_ex.c:76-77 -> frame/3/bli_l3_sup.c:37-135) for every
   homogeneous-datatype problem because the MT/NT/KT thresholds are set huge.
   trsm: BLIS has no whole-operation slot for trsm (bli_l3_oapi_ex.c:692-801
   goes straight to the control tree), so this file provides bli_trsm_ex_b200
   with bli_trsm_ex's exact signature; INTEGRATION.md shows the two-line guard
   that lets the b200 configuration resolve bli_trsm_ex to it (the sandbox
   trick of bli_l3_oapi_ex.c:45-49).

   Error convention: the engine has no CPU fallback, so a failure inside it is
   fatal exactly like every other BLIS error: message + bli_abort()
   (frame/base/bli_error.c:126-139).
*/
#include <stdio.h>
#include "blis.h"
#include "blis_b200.h"

static void bli_b200_die( const char* op )
{
	fprintf( stderr, "libblis (b200): %s failed: %s\n", op, b200_last_error() );
	fflush( stderr );
	bli_abort();
}

/* -- gemm: whole-operation handler (gemmsup_oft) ------------------------------ */

err_t bli_gemmsup_b200
     (
       const obj_t*  alpha,
       const obj_t*  a,
       const obj_t*  b,
       const obj_t*  beta,
       const obj_t*  c,
       const cntx_t* cntx,
             rntm_t* rntm
     )
{
	( void )cntx; ( void )rntm;

	/* Guaranteed by the caller (bli_gemmsup): one datatype, comp precision ==
	   storage precision, alpha/beta already cast to dt, non-trivial dims. */
	const num_t dt = bli_obj_dt( c );

	/* Scalars attached to A/B by earlier API layers must be the identity at
	   this point of the call stack (they are only set inside the control
	   tree, bli_gemm_cntl.c:168-207). */

	const err_t r = b200_gemm
	(
	  ( int )dt,
	  ( int )bli_obj_conjtrans_status( a ),
	  ( int )bli_obj_conjtrans_status( b ),
	  bli_obj_length( c ),
	  bli_obj_width( c ),
	  bli_obj_width_after_trans( a ),
	  bli_obj_buffer_for_1x1( dt, alpha ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
	  bli_obj_buffer_for_1x1( dt, beta ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c )
	);
	if ( r != BLIS_SUCCESS ) bli_b200_die( "gemm" );

	/* BLIS_SUCCESS ends bli_gemm_ex; returning BLIS_FAILURE would send the
	   problem to the CPU control tree, which the b200 configuration forbids. */
	return BLIS_SUCCESS;
}

/* -- trsm: same parameter list as bli_trsm_ex --------------------------------- */

void bli_trsm_ex_b200
     (
             side_t  side,
       const obj_t*  alpha,
       const obj_t*  a,
       const obj_t*  b,
       const cntx_t* cntx,
       const rntm_t* rntm
     )
{
	( void )rntm;
	bli_init_once();

	if ( bli_error_checking_is_enabled() )
		bli_trsm_check( side, alpha, a, b, cntx );

	const num_t dt = bli_obj_dt( b );
	if ( bli_obj_dt( a ) != dt )
	{
		fprintf( stderr, "libblis (b200): mixed-datatype trsm is not supported by the b200 engine.\n" );
		bli_abort();
	}
	if ( bli_obj_has_zero_dim( b ) ) return;

	obj_t alpha_cast;
	bli_obj_scalar_init_detached_copy_of( dt, BLIS_NO_CONJUGATE, alpha, &alpha_cast );

	/* alpha == 0 (B := 0) is handled by the engine so that it also works for
	   device-resident B; the reference's bli_l3_return_early_if_trivial would
	   run bli_scalm on the host pointer (frame/3/bli_l3_util.c:57-60). */
	const err_t r = b200_trsm
	(
	  ( int )dt,
	  ( int )side,
	  ( int )bli_obj_uplo( a ),
	  ( int )bli_obj_conjtrans_status( a ),
	  ( int )bli_obj_diag( a ),
	  bli_obj_length( b ),
	  bli_obj_width( b ),
	  bli_obj_buffer_for_1x1( dt, &alpha_cast ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b )
	);
	if ( r != BLIS_SUCCESS ) bli_b200_die( "trsm" );
}

#ifdef BLIS_B200_OVERRIDE_TRSM_EX
/* Build-time switch of the config route / LD_PRELOAD demo: bli_trsm_ex itself
   resolves to the engine ("trsm"). */
void bli_trsm_ex( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b,
                  const cntx_t* cntx, const rntm_t* rntm )
{
	bli_trsm_ex_b200( side, alpha, a, b, cntx, rntm );
}
#endif

/* -- gemmt family: same parameter lists as bli_gemmt_ex / bli_syrk_ex / bli_herk_ex /
      bli_syr2k_ex / bli_her2k_ex (frame/3/bli_l3_oapi_ex.c:151-346) ------------------

   bli_gemmt_ex never consults the BLIS_GEMMT handler slot in this snapshot (the slot exists,
   ref_kernels/bli_cntx_ref.c:563, but bli_l3_oapi_ex.c:151-228 goes straight to the control
   tree), so these operations are bound like trsm: functions with the expert-API signatures
   that the b200 configuration resolves the public names to ("gemmt family").
   The rank-k/2k operations are bound individually rather than through gemmt because the
   reference calls gemmt from inside the same translation unit. */

static void bli_b200_scalar( num_t dt, const obj_t* s, obj_t* out )
{
	bli_obj_scalar_init_detached_copy_of( dt, BLIS_NO_CONJUGATE, s, out );
}

static bool bli_b200_same_dt( const obj_t* a, const obj_t* b, const obj_t* c, const char* op )
{
	const num_t dt = bli_obj_dt( c );
	if ( bli_obj_dt( a ) == dt && ( b == NULL || bli_obj_dt( b ) == dt ) ) return TRUE;
	fprintf( stderr, "libblis (b200): mixed-datatype %s is not supported by the b200 engine.\n", op );
	bli_abort();
	return FALSE;
}

void bli_gemmt_ex_b200( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                        const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_gemmt_check( alpha, a, b, beta, c, cntx );
	bli_b200_same_dt( a, b, c, "gemmt" );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt = bli_obj_dt( c );
	obj_t al, be; bli_b200_scalar( dt, alpha, &al ); bli_b200_scalar( dt, beta, &be );
	const err_t r = b200_gemmt( ( int )dt, ( int )bli_obj_uplo( c ),
	  ( int )bli_obj_conjtrans_status( a ), ( int )bli_obj_conjtrans_status( b ),
	  bli_obj_length( c ), bli_obj_width_after_trans( a ),
	  bli_obj_buffer_for_1x1( dt, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
	  bli_obj_buffer_for_1x1( dt, &be ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "gemmt" );
}

void bli_syrk_ex_b200( const obj_t* alpha, const obj_t* a, const obj_t* beta, const obj_t* c,
                       const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_syrk_check( alpha, a, beta, c, cntx );
	bli_b200_same_dt( a, NULL, c, "syrk" );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt = bli_obj_dt( c );
	obj_t al, be; bli_b200_scalar( dt, alpha, &al ); bli_b200_scalar( dt, beta, &be );
	const err_t r = b200_syrk( ( int )dt, ( int )bli_obj_uplo( c ), ( int )bli_obj_conjtrans_status( a ),
	  bli_obj_length( c ), bli_obj_width_after_trans( a ),
	  bli_obj_buffer_for_1x1( dt, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_for_1x1( dt, &be ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "syrk" );
}

void bli_herk_ex_b200( const obj_t* alpha, const obj_t* a, const obj_t* beta, const obj_t* c,
                       const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_herk_check( alpha, a, beta, c, cntx );
	bli_b200_same_dt( a, NULL, c, "herk" );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt = bli_obj_dt( c ), dt_r = bli_dt_proj_to_real( dt );
	obj_t al, be; bli_b200_scalar( dt_r, alpha, &al ); bli_b200_scalar( dt_r, beta, &be );   /* real scalars */
	const err_t r = b200_herk( ( int )dt, ( int )bli_obj_uplo( c ), ( int )bli_obj_conjtrans_status( a ),
	  bli_obj_length( c ), bli_obj_width_after_trans( a ),
	  bli_obj_buffer_for_1x1( dt_r, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_for_1x1( dt_r, &be ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "herk" );
}

void bli_syr2k_ex_b200( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                        const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_syr2k_check( alpha, a, b, beta, c, cntx );
	bli_b200_same_dt( a, b, c, "syr2k" );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt = bli_obj_dt( c );
	obj_t al, be; bli_b200_scalar( dt, alpha, &al ); bli_b200_scalar( dt, beta, &be );
	const err_t r = b200_syr2k( ( int )dt, ( int )bli_obj_uplo( c ),
	  ( int )bli_obj_conjtrans_status( a ), ( int )bli_obj_conjtrans_status( b ),
	  bli_obj_length( c ), bli_obj_width_after_trans( a ),
	  bli_obj_buffer_for_1x1( dt, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
	  bli_obj_buffer_for_1x1( dt, &be ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "syr2k" );
}

void bli_her2k_ex_b200( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                        const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_her2k_check( alpha, a, b, beta, c, cntx );
	bli_b200_same_dt( a, b, c, "her2k" );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt = bli_obj_dt( c ), dt_r = bli_dt_proj_to_real( dt );
	obj_t al, be; bli_b200_scalar( dt, alpha, &al ); bli_b200_scalar( dt_r, beta, &be );     /* beta is real */
	const err_t r = b200_her2k( ( int )dt, ( int )bli_obj_uplo( c ),
	  ( int )bli_obj_conjtrans_status( a ), ( int )bli_obj_conjtrans_status( b ),
	  bli_obj_length( c ), bli_obj_width_after_trans( a ),
	  bli_obj_buffer_for_1x1( dt, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
	  bli_obj_buffer_for_1x1( dt_r, &be ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "her2k" );
}

/* -- hemm, symm, trmm3, trmm: bli_hemm_ex / bli_symm_ex / bli_trmm3_ex / bli_trmm_ex
      (frame/3/bli_l3_oapi_ex.c:349-689); no handler slot exists for them either ------------ */

static void bli_b200_struc_mm( int op, side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b,
                               const obj_t* beta, const obj_t* c, const char* name )
{
	bli_b200_same_dt( a, b, c, name );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt = bli_obj_dt( c );
	obj_t al, be; bli_b200_scalar( dt, alpha, &al ); bli_b200_scalar( dt, beta, &be );
	err_t r;
	if ( op == 0 || op == 1 )
	{
		/* A transposition bit on the structured operand (object API only): A^T == conj(A) for a Hermitian A and
		   A^T == A for a symmetric one, which is how the reference's packm resolves it
		   (frame/1m/packm/bli_packm_struc_cxk.c:146-301 reads the stored triangle through the toggled strides). */
		conj_t conja = bli_obj_conj_status( a );
		if ( op == 0 && bli_obj_has_trans( a ) ) conja = ( conja == BLIS_CONJUGATE ? BLIS_NO_CONJUGATE : BLIS_CONJUGATE );
		r = ( op == 0 ? b200_hemm : b200_symm )( ( int )dt, ( int )side, ( int )bli_obj_uplo( a ),
		  ( int )conja, ( int )bli_obj_conjtrans_status( b ),
		  bli_obj_length( c ), bli_obj_width( c ), bli_obj_buffer_for_1x1( dt, &al ),
		  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
		  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
		  bli_obj_buffer_for_1x1( dt, &be ),
		  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	}
	else
		r = b200_trmm3( ( int )dt, ( int )side, ( int )bli_obj_uplo( a ), ( int )bli_obj_conjtrans_status( a ),
		  ( int )bli_obj_diag( a ), ( int )bli_obj_conjtrans_status( b ),
		  bli_obj_length( c ), bli_obj_width( c ), bli_obj_buffer_for_1x1( dt, &al ),
		  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
		  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
		  bli_obj_buffer_for_1x1( dt, &be ),
		  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( name );
}

void bli_hemm_ex_b200( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                       const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_hemm_check( side, alpha, a, b, beta, c, cntx );
	bli_b200_struc_mm( 0, side, alpha, a, b, beta, c, "hemm" );
}

void bli_symm_ex_b200( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                       const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_symm_check( side, alpha, a, b, beta, c, cntx );
	bli_b200_struc_mm( 1, side, alpha, a, b, beta, c, "symm" );
}

void bli_trmm3_ex_b200( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                        const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_trmm3_check( side, alpha, a, b, beta, c, cntx );
	bli_b200_struc_mm( 2, side, alpha, a, b, beta, c, "trmm3" );
}

void bli_trmm_ex_b200( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b,
                       const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_trmm_check( side, alpha, a, b, cntx );
	bli_b200_same_dt( a, NULL, b, "trmm" );
	if ( bli_obj_has_zero_dim( b ) ) return;
	const num_t dt = bli_obj_dt( b );
	obj_t al; bli_b200_scalar( dt, alpha, &al );
	const err_t r = b200_trmm( ( int )dt, ( int )side, ( int )bli_obj_uplo( a ), ( int )bli_obj_conjtrans_status( a ),
	  ( int )bli_obj_diag( a ), bli_obj_length( b ), bli_obj_width( b ), bli_obj_buffer_for_1x1( dt, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "trmm" );
}

#ifdef BLIS_B200_OVERRIDE_GEMMT_EX
void bli_hemm_ex( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_hemm_ex_b200( side, alpha, a, b, beta, c, cntx, rntm ); }
void bli_symm_ex( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_symm_ex_b200( side, alpha, a, b, beta, c, cntx, rntm ); }
void bli_trmm3_ex( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_trmm3_ex_b200( side, alpha, a, b, beta, c, cntx, rntm ); }
void bli_trmm_ex( side_t side, const obj_t* alpha, const obj_t* a, const obj_t* b, const cntx_t* cntx, const rntm_t* rntm )
{ bli_trmm_ex_b200( side, alpha, a, b, cntx, rntm ); }
#endif

#ifdef BLIS_B200_OVERRIDE_GEMMT_EX
/* Build-time switch of the config route / LD_PRELOAD demo, as for trsm. */
void bli_gemmt_ex( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_gemmt_ex_b200( alpha, a, b, beta, c, cntx, rntm ); }
void bli_syrk_ex( const obj_t* alpha, const obj_t* a, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_syrk_ex_b200( alpha, a, beta, c, cntx, rntm ); }
void bli_herk_ex( const obj_t* alpha, const obj_t* a, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_herk_ex_b200( alpha, a, beta, c, cntx, rntm ); }
void bli_syr2k_ex( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_syr2k_ex_b200( alpha, a, b, beta, c, cntx, rntm ); }
void bli_her2k_ex( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_her2k_ex_b200( alpha, a, b, beta, c, cntx, rntm ); }
#endif

/* -- mixed-datatype gemm: bli_gemm_ex itself ------------------------------------------------
   bli_gemmsup rejects operands of different datatype or a computation precision other than C's before it
   reaches the handler (frame/3/bli_l3_sup.c:60-75), so mixed-datatype problems would run on the CPU control tree
   (frame/3/gemm/bli_gemm_cntl.c:87-392).  bli_gemm_ex_b200 has bli_gemm_ex's parameter list; homogeneous problems
   go to b200_gemm, everything else to b200_gemm_md.  Bound like trsm (-DBLIS_B200_OVERRIDE_GEMM_EX). */
void bli_gemm_ex_b200( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c,
                       const cntx_t* cntx, const rntm_t* rntm )
{
	( void )rntm;
	bli_init_once();
	if ( bli_error_checking_is_enabled() ) bli_gemm_check( alpha, a, b, beta, c, cntx );
	if ( bli_obj_has_zero_dim( c ) ) return;
	const num_t dt_a = bli_obj_dt( a ), dt_b = bli_obj_dt( b ), dt_c = bli_obj_dt( c );
	const prec_t cp = bli_obj_comp_prec( c );
	obj_t al, be;
	if ( dt_a == dt_c && dt_b == dt_c && cp == bli_dt_prec( dt_c ) )
	{
		bli_b200_scalar( dt_c, alpha, &al ); bli_b200_scalar( dt_c, beta, &be );
		const err_t r = b200_gemm( ( int )dt_c, ( int )bli_obj_conjtrans_status( a ), ( int )bli_obj_conjtrans_status( b ),
		  bli_obj_length( c ), bli_obj_width( c ), bli_obj_width_after_trans( a ), bli_obj_buffer_for_1x1( dt_c, &al ),
		  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
		  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ), bli_obj_buffer_for_1x1( dt_c, &be ),
		  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
		if ( r != BLIS_SUCCESS ) bli_b200_die( "gemm" );
		return;
	}
	bli_b200_scalar( BLIS_DCOMPLEX, alpha, &al ); bli_b200_scalar( BLIS_DCOMPLEX, beta, &be );
	const err_t r = b200_gemm_md( ( int )dt_a, ( int )dt_b, ( int )dt_c, ( int )cp,
	  ( int )bli_obj_conjtrans_status( a ), ( int )bli_obj_conjtrans_status( b ),
	  bli_obj_length( c ), bli_obj_width( c ), bli_obj_width_after_trans( a ),
	  ( const double* )bli_obj_buffer_for_1x1( BLIS_DCOMPLEX, &al ),
	  bli_obj_buffer_at_off( a ), bli_obj_row_stride( a ), bli_obj_col_stride( a ),
	  bli_obj_buffer_at_off( b ), bli_obj_row_stride( b ), bli_obj_col_stride( b ),
	  ( const double* )bli_obj_buffer_for_1x1( BLIS_DCOMPLEX, &be ),
	  bli_obj_buffer_at_off( c ), bli_obj_row_stride( c ), bli_obj_col_stride( c ) );
	if ( r != BLIS_SUCCESS ) bli_b200_die( "gemm (mixed datatype)" );
}

#ifdef BLIS_B200_OVERRIDE_GEMM_EX
void bli_gemm_ex( const obj_t* alpha, const obj_t* a, const obj_t* b, const obj_t* beta, const obj_t* c, const cntx_t* cntx, const rntm_t* rntm )
{ bli_gemm_ex_b200( alpha, a, b, beta, c, cntx, rntm ); }
#endif

/* -- batched gemm: ?gemm_batch_ (frame/compat/extra/bla_gemm_batch.c:44-131) ------------------
   The reference loops over the problems calling bli_?gemm_ex; these wrappers hand the whole batch to
   b200_gemm_batch (device-resident problems run concurrently; host problems are staged one by one). */
#ifdef BLIS_B200_OVERRIDE_GEMM_BATCH
#include <stdlib.h>
static void bli_b200_gemm_batch( num_t dt, const f77_char* ta, const f77_char* tb, const f77_int* m, const f77_int* n, const f77_int* k,
                                 const void* alpha, const void** a, const f77_int* lda, const void** b, const f77_int* ldb,
                                 const void* beta, void** c, const f77_int* ldc, const f77_int* group_count, const f77_int* group_size )
{
	const int ng = ( int )*group_count;
	bli_init_auto();
	if ( ng <= 0 ) return;
	int*        gs  = malloc( sizeof( int ) * 3 * ( size_t )ng );
	b200_dim_t* dim = malloc( sizeof( b200_dim_t ) * 9 * ( size_t )ng );
	int *tra = gs + ng, *trb = gs + 2 * ng;
	b200_dim_t *mm = dim, *nn = dim + ng, *kk = dim + 2 * ng, *rsa = dim + 3 * ng, *csa = dim + 4 * ng,
	           *rsb = dim + 5 * ng, *csb = dim + 6 * ng, *rsc = dim + 7 * ng, *csc = dim + 8 * ng;
	for ( int i = 0; i < ng; ++i )
	{
		trans_t t;
		gs[i] = ( int )group_size[i];
		bli_param_map_netlib_to_blis_trans( ta[i], &t ); tra[i] = ( int )t;
		bli_param_map_netlib_to_blis_trans( tb[i], &t ); trb[i] = ( int )t;
		mm[i] = m[i]; nn[i] = n[i]; kk[i] = k[i];
		rsa[i] = 1; csa[i] = lda[i]; rsb[i] = 1; csb[i] = ldb[i]; rsc[i] = 1; csc[i] = ldc[i];
	}
	const err_t r = b200_gemm_batch( ( int )dt, ng, gs, tra, trb, mm, nn, kk, alpha, ( const void* const* )a, rsa, csa,
	                                 ( const void* const* )b, rsb, csb, beta, ( void* const* )c, rsc, csc );
	free( gs ); free( dim );
	if ( r != BLIS_SUCCESS || b200_sync() != BLIS_SUCCESS ) bli_b200_die( "gemm_batch" );
}
#define BLI_B200_GEMM_BATCH( ch, ftype, dt ) \
void ch##gemm_batch_( const f77_char* ta, const f77_char* tb, const f77_int* m, const f77_int* n, const f77_int* k, \
                      const ftype* alpha, const ftype** a, const f77_int* lda, const ftype** b, const f77_int* ldb, \
                      const ftype* beta, ftype** c, const f77_int* ldc, const f77_int* group_count, const f77_int* group_size ) \
{ bli_b200_gemm_batch( dt, ta, tb, m, n, k, alpha, ( const void** )a, lda, ( const void** )b, ldb, beta, ( void** )c, ldc, group_count, group_size ); }
BLI_B200_GEMM_BATCH( s, float,    BLIS_FLOAT )
BLI_B200_GEMM_BATCH( d, double,   BLIS_DOUBLE )
BLI_B200_GEMM_BATCH( c, scomplex, BLIS_SCOMPLEX )
BLI_B200_GEMM_BATCH( z, dcomplex, BLIS_DCOMPLEX )
#endif

/* -- registration ----------------------------------------------------------- */

/* Install the engine into one context: tile shapes as blocksizes, thresholds
   that route every gemm to the handler, and the handler itself. */
void bli_b200_install( cntx_t* cntx )
{
	/* "every size": m < MT || n < NT || k < KT is the dispatch test
	   (frame/base/bli_cntx.h:186-192). */
	const dim_t all = ( dim_t )1 << 62;
	blksz_t mt, nt, kt;
	bli_blksz_init_easy( &mt, all, all, all, all );
	bli_blksz_init_easy( &nt, all, all, all, all );
	bli_blksz_init_easy( &kt, all, all, all, all );
	bli_cntx_set_blkszs
	(
	  cntx,
	  BLIS_MT, &mt, BLIS_MT,
	  BLIS_NT, &nt, BLIS_NT,
	  BLIS_KT, &kt, BLIS_KT,
	  BLIS_VA_END
	);
	bli_cntx_set_l3_sup_handlers
	(
	  cntx,
	  BLIS_GEMM, bli_gemmsup_b200,
	  BLIS_VA_END
	);
}

/* Plugin entry point: call once after bli_init() and before any computation
   ("registration must happen before any computations are performed with the
   plugin", docs/PluginHowTo.md:178). */
err_t bli_plugin_register_b200( void )
{
	bli_init();
	if ( b200_init( -1 ) != BLIS_SUCCESS ) bli_b200_die( "b200_init" );

	/* the context the library dispatches on (frame/base/bli_gks.c:276) */
	cntx_t* cntx = ( cntx_t* )bli_gks_lookup_id( bli_arch_query_id() );
	if ( cntx == NULL ) return BLIS_FAILURE;
	bli_b200_install( cntx );
	return BLIS_SUCCESS;
}

/* LD_PRELOAD route: with BLIS_B200_PLUGIN=1 in the environment the plugin
   registers itself as soon as the library is loaded, so an UNMODIFIED binary
   linked against libblis (e.g. the reference's own test_libblis.x) runs its
   gemm/trsm on the B200.  BLIS_B200_VERBOSE=1 reports the number of CUDA
   kernels the engine launched when the process exits. */
#include <stdlib.h>
unsigned long long b200_launch_count( void );
static void bli_b200_report( void )
{
	fprintf( stderr, "libblis (b200): %llu CUDA kernels launched by the engine\n", b200_launch_count() );
}
__attribute__((constructor)) static void bli_b200_ctor( void )
{
	const char* v = getenv( "BLIS_B200_VERBOSE" );
	if ( v != NULL && v[0] == '1' ) atexit( bli_b200_report );     /* also in a config/b200 build, where nothing is registered at run time */
	const char* e = getenv( "BLIS_B200_PLUGIN" );
	if ( e == NULL || e[0] != '1' ) return;
	bli_plugin_register_b200();
}
