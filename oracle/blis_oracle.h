/*
 * blis_oracle.h -- CPU restatement of the reference's gemm/trsm hot path (and the gemmt family next to it).
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA engine is compared
 * against; it is never linked into, called from, or shipped with the product
 * (blis_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the real reference library built from /root/reference by
 *       oracle/build_ref.py (index arithmetic and packing bit-exact, whole
 *       gemm/trsm bit-exact on power-of-two inputs and within the testsuite's
 *       tolerance otherwise), and
 *   (2) the golden fixtures under tests/golden/ generated from that library by
 *       tests/golden/make_golden.py (these travel to machines that have no
 *       /root/reference).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the BLIS tree).  Plain C99, single threaded, no dependencies.
 */
#ifndef BLIS_ORACLE_H
#define BLIS_ORACLE_H
#include <stdint.h>

typedef int64_t dim_t;
typedef int64_t inc_t;
typedef int64_t doff_t;

/* enum values of the reference (frame/include/bli_type_defs.h:278-455) */
#define ORC_TRANS_BIT  0x08
#define ORC_CONJ_BIT   0x10
#define ORC_UPPER      0x60
#define ORC_LOWER      0xC0
#define ORC_DENSE      0xE0
#define ORC_LEFT       0
#define ORC_RIGHT      1
#define ORC_UNIT_DIAG  0x100

/* cache / register blocksizes for one datatype (cntx blksz_t entries) */
typedef struct { dim_t mr, nr, mc, kc, nc; int row_pref; } orc_blksz_t;

/* dt: 0 = s, 1 = c, 2 = d, 3 = z (num_t).  Defaults are the reference
   context's values (ref_kernels/bli_cntx_ref.c:377-384). */
void orc_set_blksz( int dt, dim_t mr, dim_t nr, dim_t mc, dim_t kc, dim_t nc, int row_pref );
void orc_get_blksz( int dt, orc_blksz_t* out );

/* ---- index arithmetic (must be bit-exact) ---- */
dim_t orc_determine_blocksize( int backward, dim_t i, dim_t dim, dim_t b_alg, dim_t b_max );
void  orc_thread_range_sub( dim_t work_id, dim_t n_way, dim_t n, dim_t bf, int handle_edge_low,
                            dim_t* start, dim_t* end );
void  orc_thread_partition_2x2( dim_t n_thread, dim_t work1, dim_t work2, dim_t* nt1, dim_t* nt2 );
dim_t orc_align_dim_to_mult( dim_t dim, dim_t mult );
dim_t orc_packm_panel_stride( dim_t ldp, dim_t panel_len_max );

/* ---- typed entry points: X in {s,d,c,z}; complex data is interleaved (re,im) ---- */
#define ORC_DECL( ch, ctype ) \
void orc_##ch##packm_cxk( int conj, dim_t cdim, dim_t cdim_max, dim_t n, dim_t n_max, const ctype* kappa, \
                          const ctype* a, inc_t inca, inc_t lda, ctype* p, inc_t ldp ); \
void orc_##ch##packm_diag( int uplo, int unit, int conj, int invdiag, dim_t cdim, dim_t cdim_max, dim_t n_max, \
                          const ctype* kappa, const ctype* a, inc_t inca, inc_t lda, ctype* p, inc_t ldp ); \
void orc_##ch##packm_struc_cxk( int triangular, int uplo, int unit, int conj, int invdiag, \
                          dim_t panel_dim, dim_t panel_len, dim_t panel_dim_max, dim_t panel_len_max, \
                          dim_t panel_dim_off, dim_t panel_len_off, const ctype* kappa, \
                          const ctype* c, inc_t incc, inc_t ldc, ctype* p, inc_t ldp ); \
void orc_##ch##gemm_ukr( dim_t m, dim_t n, dim_t k, const ctype* alpha, const ctype* a, const ctype* b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c, dim_t mr, dim_t nr ); \
void orc_##ch##trsm_ukr( int upper, const ctype* a, ctype* b, ctype* c, inc_t rs_c, inc_t cs_c, dim_t mr, dim_t nr ); \
void orc_##ch##gemmtrsm_ukr( int upper, dim_t m, dim_t n, dim_t k, const ctype* alpha, const ctype* a1x, const ctype* a11, \
                          const ctype* bx1, ctype* b11, ctype* c11, inc_t rs_c, inc_t cs_c, dim_t mr, dim_t nr ); \
void orc_##ch##gemm( int transa, int transb, dim_t m, dim_t n, dim_t k, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##gemmt( int uploc, int transa, int transb, dim_t m, dim_t k, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##syrk( int uploc, int transa, dim_t m, dim_t k, const ctype* alpha, const ctype* a, inc_t rs_a, inc_t cs_a, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##herk( int uploc, int transa, dim_t m, dim_t k, const ctype* alpha_r, const ctype* a, inc_t rs_a, inc_t cs_a, \
                          const ctype* beta_r, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##syr2k( int uploc, int transa, int transb, dim_t m, dim_t k, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##her2k( int uploc, int transa, int transb, dim_t m, dim_t k, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta_r, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##hemm( int side, int uplo, int conja, int transb, dim_t m, dim_t n, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##symm( int side, int uplo, int conja, int transb, dim_t m, dim_t n, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##trmm3( int side, int uplo, int transa, int diag, int transb, dim_t m, dim_t n, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, const ctype* b, inc_t rs_b, inc_t cs_b, \
                          const ctype* beta, ctype* c, inc_t rs_c, inc_t cs_c ); \
void orc_##ch##trmm( int side, int uplo, int transa, int diag, dim_t m, dim_t n, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, ctype* b, inc_t rs_b, inc_t cs_b ); \
void orc_##ch##trsm( int side, int uplo, int transa, int diag, dim_t m, dim_t n, const ctype* alpha, \
                          const ctype* a, inc_t rs_a, inc_t cs_a, ctype* b, inc_t rs_b, inc_t cs_b );

ORC_DECL( s, float )
ORC_DECL( d, double )
ORC_DECL( c, float )    /* pointers to interleaved (re,im) pairs */
ORC_DECL( z, double )

/* mixed-datatype gemm (docs/MixedDatatypes.md; bli_gemm_cntl.c:87-392): dt_* are num_t values, comp_prec is 0 (single)
   or 2 (double), alpha and beta point to {real, imag} doubles */
void orc_gemm_md( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb, dim_t m, dim_t n, dim_t k,
                  const double* alpha, const void* a, inc_t rs_a, inc_t cs_a, const void* b, inc_t rs_b, inc_t cs_b,
                  const double* beta, void* c, inc_t rs_c, inc_t cs_c );

#endif
