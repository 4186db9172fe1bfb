/*
 * blis_b200.h -- C ABI of the B200-native level-3 engine for BLIS.
 *
 * This is the drop-in boundary: plain pointers and
 * 64-bit sizes, no C++/torch types.  Every entry point states the reference
 * interface it stands in for (paths relative to the BLIS tree).
 *
 * Conventions shared with the reference
 *   - dimensions and strides are 64-bit signed (BLIS dim_t/inc_t = gint_t,
 *     frame/include/bli_type_defs.h:78-116);
 *   - trans/conj/uplo/diag/side/datatype arguments use the numeric values of
 *     BLIS's own enums (frame/include/bli_type_defs.h:278-455), so the BLIS
 *     side binding passes its trans_t/uplo_t/... straight through;
 *   - matrices are described by (pointer, row stride, column stride) exactly
 *     like the typed API (frame/3/bli_l3_tapi.c:43-70); any of rs==1, cs==1
 *     or general stride is accepted;
 *   - return value is a BLIS err_t: B200_SUCCESS (-1) or B200_FAILURE (-2)
 *     (frame/include/bli_type_defs.h:1502-1607).  On failure
 *     b200_last_error() holds the message; the BLIS glue turns it into
 *     bli_abort() because this engine has no CPU fallback.
 *
 * Operand pointers may be device pointers (used in place), pinned host
 * pointers or pageable host pointers (staged through the engine's pinned
 * buffers); each pointer is classified per call.
 */
#ifndef BLIS_B200_H
#define BLIS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t b200_dim_t;   /* BLIS dim_t  */
typedef int64_t b200_inc_t;   /* BLIS inc_t  */
typedef int     b200_err_t;   /* BLIS err_t  */

#define B200_SUCCESS (-1)     /* BLIS_SUCCESS */
#define B200_FAILURE (-2)     /* BLIS_FAILURE */

/* num_t  (bli_type_defs.h:448-451) */
#define B200_FLOAT     0
#define B200_SCOMPLEX  1
#define B200_DOUBLE    2
#define B200_DCOMPLEX  3
/* trans_t (bli_type_defs.h:397-400) */
#define B200_NO_TRANSPOSE       0x00
#define B200_TRANSPOSE          0x08
#define B200_CONJ_NO_TRANSPOSE  0x10
#define B200_CONJ_TRANSPOSE     0x18
/* uplo_t (bli_type_defs.h:412-414) */
#define B200_UPPER  0x60
#define B200_LOWER  0xC0
/* side_t (bli_type_defs.h:419-420) */
#define B200_LEFT   0
#define B200_RIGHT  1
/* diag_t (bli_type_defs.h:425-426) */
#define B200_NONUNIT_DIAG 0x000
#define B200_UNIT_DIAG    0x100

typedef struct { float  real, imag; } b200_scomplex;  /* BLIS scomplex */
typedef struct { double real, imag; } b200_dcomplex;  /* BLIS dcomplex */

/* ---- lifetime ------------------------------------------------------------
 * Replaces bli_init()/bli_finalize() for the device side
 * (frame/base/bli_init.c:87-99).  b200_init is idempotent and thread safe;
 * device < 0 means "current CUDA device".  The engine also initialises
 * itself lazily on first use. */
b200_err_t  b200_init( int device );
void        b200_finalize( void );
const char* b200_last_error( void );
int         b200_device_count( void );
/* Version / build info string, like bli_info_get_version_str()
 * (frame/base/bli_info.c). */
const char* b200_info( void );

/* Name of the last kernel the calling thread launched (e.g. "gemm_dmma_tma_kernel<XK=1,YK=0,TRI=0,CST=0>"), and a
 * histogram "name\tcount\n..." of every kernel launched since the last reset, written to buf (NUL terminated, truncated
 * to len); returns the untruncated length.  The reference reports which microkernel a context holds through
 * bli_info_get_gemm_ukr_impl_string() (frame/base/bli_info.c:180-215); here tests and bench.py read which tile kernel
 * actually served a call. */
const char* b200_last_kernel( void );
size_t      b200_kernel_stats( char* buf, size_t len, int reset );

/* All work of the calling thread is issued on this CUDA stream
 * (cudaStream_t passed as void*); NULL selects the engine's own stream.
 * The torch harness passes torch's current stream so CUDA events see it. */
void        b200_set_stream( void* stream );
void*       b200_get_stream( void );
/* Block until everything the engine queued on its stream has finished. */
b200_err_t  b200_sync( void );

/* Page-locked host allocation: the BLIS_MALLOC_USER / BLIS_FREE_USER hooks of
 * config/b200 (docs/ConfigurationHowTo.md:195-207), signature void* f(size_t). */
void*       b200_malloc_pinned( size_t size );
void        b200_free_pinned( void* p );
/* Where the engine will find an operand: 0 device (or managed) memory, used in
 * place; 1 page-locked host memory (b200_malloc_pinned, bli_obj_create under
 * config/b200, cudaHostRegister), copied by the DMA engines directly; 2 pageable
 * host memory, copied through the engine's pinned staging ring.  This is the
 * per-call classification every entry point applies to its operand pointers;
 * -1 if no device is usable. */
int         b200_pointer_kind( const void* p );

/* ---- gemm ------------------------------------------------------------------
 * C := beta*C + alpha*transa(A)*transb(B),  C is m x n, k is the inner dim.
 *
 * b200_gemm is what the whole-operation gemm hook registered with
 *   bli_cntx_set_l3_sup_handler( BLIS_GEMM, ... )       frame/base/bli_cntx.h:335-343
 * (signature gemmsup_oft, frame/3/bli_l3_sup_oft.h:46-59) calls after
 * unpacking its obj_t arguments; the typed variants have the parameter list
 * of bli_?gemm (frame/3/bli_l3_tapi.c:43-70).
 * alpha/beta point to one element of type dt in HOST memory. */
b200_err_t b200_gemm( int dt, int transa, int transb,
                      b200_dim_t m, b200_dim_t n, b200_dim_t k,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );

#define B200_DECL_GEMM( ch, ctype ) \
b200_err_t b200_##ch##gemm( int transa, int transb, \
                      b200_dim_t m, b200_dim_t n, b200_dim_t k, \
                      const ctype* alpha, \
                      const ctype* a, b200_inc_t rs_a, b200_inc_t cs_a, \
                      const ctype* b, b200_inc_t rs_b, b200_inc_t cs_b, \
                      const ctype* beta, \
                      ctype*       c, b200_inc_t rs_c, b200_inc_t cs_c );
B200_DECL_GEMM( s, float )
B200_DECL_GEMM( d, double )
B200_DECL_GEMM( c, b200_scomplex )
B200_DECL_GEMM( z, b200_dcomplex )

/* k-panel accumulation in ONE launch:
 *   C := beta*C + alpha * sum_{s < npanels} transa(A_s) * transb(B_s)
 * i.e. the pc loop of bli_gemm_blk_var3 (frame/3/gemm/bli_gemm_blk_var3.c:37-114: one rank-KC update
 * per k block, beta reset to one after the first) folded into the kernel's k loop.  Every A_s has the same
 * shape/strides (m x k after transa) and every B_s likewise (k x n); 1 <= npanels <= 8; device-resident
 * operands; dt = d or z.  Used by the multi-GPU gemm to accumulate all-gathered k-panels (blis_b200/dist.py). */
b200_err_t b200_gemm_kpanels( int dt, int transa, int transb,
                      b200_dim_t m, b200_dim_t n, b200_dim_t k, int npanels,
                      const void* alpha,
                      const void* const* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* const* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );

/* ---- trsm ------------------------------------------------------------------
 * Solve  transa(A) * X = alpha * B  (side = left)  or
 *        X * transa(A) = alpha * B  (side = right), overwriting B with X.
 * A is triangular (uplo), unit or non-unit diagonal; B is m x n.
 *
 * Stands in for bli_trsm_ex (frame/3/bli_l3_oapi_ex.c:692-801); the typed
 * variants have the parameter list of bli_?trsm (frame/3/bli_l3_tapi.c,
 * GENTFUNC trsm). */
b200_err_t b200_trsm( int dt, int side, int uploa, int transa, int diaga,
                      b200_dim_t m, b200_dim_t n,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      void*       b, b200_inc_t rs_b, b200_inc_t cs_b );

#define B200_DECL_TRSM( ch, ctype ) \
b200_err_t b200_##ch##trsm( int side, int uploa, int transa, int diaga, \
                      b200_dim_t m, b200_dim_t n, \
                      const ctype* alpha, \
                      const ctype* a, b200_inc_t rs_a, b200_inc_t cs_a, \
                      ctype*       b, b200_inc_t rs_b, b200_inc_t cs_b );
B200_DECL_TRSM( s, float )
B200_DECL_TRSM( d, double )
B200_DECL_TRSM( c, b200_scomplex )
B200_DECL_TRSM( z, b200_dcomplex )

/* ---- gemmt family ------------------------------
 * Updates of ONE triangle (uploc) of the m x m matrix C; the other triangle is
 * never written.  transa(A) is m x k.
 *   gemmt:  C := beta*C + alpha * transa(A) * transb(B)            transb(B): k x m
 *   syrk:   C := beta*C + alpha * transa(A) * transa(A)^T
 *   herk:   C := beta*C + alpha * transa(A) * transa(A)^H          alpha, beta REAL (float/double)
 *   syr2k:  C := beta*C + alpha * transa(A) * transb(B)^T + alpha * transb(B) * transa(A)^T        transb(B): m x k
 *   her2k:  C := beta*C + alpha * transa(A) * transb(B)^H + conj(alpha) * transb(B) * transa(A)^H  beta REAL
 * herk/her2k set the imaginary parts of C's diagonal to zero afterwards.
 *
 * Stand in for bli_gemmt_ex / bli_syrk_ex / bli_herk_ex / bli_syr2k_ex /
 * bli_her2k_ex (frame/3/bli_l3_oapi_ex.c:151-346); argument order follows the
 * typed API bli_?gemmt / bli_?syrk / bli_?herk / bli_?syr2k / bli_?her2k
 * (frame/3/bli_l3_tapi.c:77-296) with the datatype as the first argument. */
b200_err_t b200_gemmt( int dt, int uploc, int transa, int transb,
                      b200_dim_t m, b200_dim_t k,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_syrk( int dt, int uploc, int transa,
                      b200_dim_t m, b200_dim_t k,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_herk( int dt, int uploc, int transa,
                      b200_dim_t m, b200_dim_t k,
                      const void* alpha_real,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* beta_real,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_syr2k( int dt, int uploc, int transa, int transb,
                      b200_dim_t m, b200_dim_t k,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_her2k( int dt, int uploc, int transa, int transb,
                      b200_dim_t m, b200_dim_t k,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta_real,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );

/* ---- hemm, symm, trmm, trmm3 --------------------
 * A is m x m (side = left) or n x n (side = right); only its uploa triangle is read.
 *   hemm:  C := beta*C + alpha * conja(A) * transb(B)   or   alpha * transb(B) * conja(A)     A Hermitian
 *   symm:  the same with A symmetric
 *   trmm3: C := beta*C + alpha * transa(A) * transb(B)  or   alpha * transb(B) * transa(A)    A triangular (diaga)
 *   trmm:  B := alpha * transa(A) * B                   or   alpha * B * transa(A)            in place
 * The imaginary part of a Hermitian diagonal is ignored, a unit diagonal is not read
 * (ref_kernels/1m/bli_packm_cxc_diag_ref.c:36-98).
 *
 * Stand in for bli_hemm_ex / bli_symm_ex / bli_trmm3_ex / bli_trmm_ex
 * (frame/3/bli_l3_oapi_ex.c:349-689); argument order follows the typed API bli_?hemm / bli_?symm
 * (frame/3/bli_l3_tapi.c:114-155), bli_?trmm3 (:298-339) and bli_?trmm (GENTFUNC trmm) with the
 * datatype as the first argument. */
b200_err_t b200_hemm( int dt, int side, int uploa, int conja, int transb,
                      b200_dim_t m, b200_dim_t n,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_symm( int dt, int side, int uploa, int conja, int transb,
                      b200_dim_t m, b200_dim_t n,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_trmm3( int dt, int side, int uploa, int transa, int diaga, int transb,
                      b200_dim_t m, b200_dim_t n,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );
b200_err_t b200_trmm( int dt, int side, int uploa, int transa, int diaga,
                      b200_dim_t m, b200_dim_t n,
                      const void* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      void*       b, b200_inc_t rs_b, b200_inc_t cs_b );

/* ---- mixed-datatype gemm -----------------------
 * C := beta*C + alpha*transa(A)*transb(B) with A, B, C of ANY of the four datatypes and a
 * computation precision comp_prec (BLIS prec_t: 0 = single, 2 = double): what bli_gemm_ex does
 * for operands of different domain/precision (docs/MixedDatatypes.md;
 * frame/3/gemm/bli_gemm_cntl.c:87-392).  A and B are typecast to the computation precision, the
 * product runs in the smallest domain that holds it, the result is typecast and accumulated into C
 * with beta in C's datatype.  alpha and beta point to dcomplex values {real, imag} in host memory
 * (the imaginary part of alpha is ignored when A, B and C are all real, that of beta when C is real).
 * Reached from the BLIS side through bli_gemm_ex_b200 when bli_obj_dt()/bli_obj_comp_prec() differ. */
b200_err_t b200_gemm_md( int dt_a, int dt_b, int dt_c, int comp_prec, int transa, int transb,
                      b200_dim_t m, b200_dim_t n, b200_dim_t k,
                      const double* alpha,
                      const void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      const void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const double* beta,
                      void*       c, b200_inc_t rs_c, b200_inc_t cs_c );

/* ---- batched gemm ------------------------------
 * group_count groups; group i holds group_size[i] independent problems
 *   C_j := beta[i]*C_j + alpha[i]*transa[i](A_j)*transb[i](B_j),   C_j is m[i] x n[i], inner dimension k[i],
 * all with the strides rs/cs_{a,b,c}[i]; the matrix pointers of all groups are concatenated in a[], b[], c[]
 * (sum of group_size entries); alpha/beta are arrays of group_count elements of datatype dt in HOST memory.
 * Parameter order follows ?gemm_batch_ (frame/compat/extra/bla_gemm_batch.c:44-60) with BLIS strides instead of
 * leading dimensions.  Device-resident problems run concurrently on a pool of streams and are ordered after the
 * work already queued on the calling thread's stream; the call returns without synchronising in that case. */
b200_err_t b200_gemm_batch( int dt, int group_count, const int* group_size,
                      const int* transa, const int* transb,
                      const b200_dim_t* m, const b200_dim_t* n, const b200_dim_t* k,
                      const void* alpha,
                      const void* const* a, const b200_inc_t* rs_a, const b200_inc_t* cs_a,
                      const void* const* b, const b200_inc_t* rs_b, const b200_inc_t* cs_b,
                      const void* beta,
                      void* const*       c, const b200_inc_t* rs_c, const b200_inc_t* cs_c );

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------
 * The reference has no distributed layer; these entry points apply ITS partitioning arithmetic across GPUs:
 * bli_thread_partition_2x2 (frame/thread/bli_thread.c:194-320) picks the Pr x Pc grid the way bli_rntm_factorize
 * (frame/base/bli_rntm.c:424-489) picks ic x jc, bli_thread_range_sub (frame/thread/bli_thread_range.c:38-184) cuts
 * the contiguous ranges (ragged edge on the last rank), and k is never split across ranks
 * (frame/3/gemm/bli_gemm_blk_var3.c:110-112), so there is no reduction.
 *
 * Bootstrap: rank 0 calls b200_dist_unique_id() and the caller's launcher (MPI_Bcast, torch.distributed, a file)
 * hands the B200_DIST_ID_BYTES bytes to every rank; every rank then calls b200_dist_init( world, rank, id ) once.
 * NCCL is bound at run time (dlopen of libnccl.so.2). */
#define B200_DIST_ID_BYTES 128
#define B200_DIST_COLS      0     /* 1-D split of C's columns */
#define B200_DIST_ROWS      1     /* 1-D split of C's rows */
#define B200_DIST_AB_STATIC 1     /* flag: the A/B shards are not written by work queued on the calling stream, so the
                                     first gather of this product may start under the previous product's kernels */
#define B200_DIST_TRACE     2     /* flag: record per-step wait events (b200_dist_last_wait_ms) */
typedef struct
{
	int        world, rank;
	int        pr, pc, i, j;        /* process grid (bli_thread_partition_2x2( world, m, n )) and my coordinates, rank = i*pc + j */
	b200_dim_t m0, m1, n0, n1;      /* my block of C: rows [m0, m1), columns [n0, n1) (bli_thread_range_sub, bf = 1) */
	b200_dim_t kb;                  /* k panel width */
	int        L, T, steps;         /* panels per step lcm(pr, pc), panels k/kb, steps T/L */
	int        na, nb;              /* k panels of A / of B this rank holds: panel t of A lives on grid column t % pc,
	                                   panel t of B on grid row t % pr */
} b200_dist_plan_t;

/* The reference's partitioning arithmetic itself (host only, no GPU needed; bit-exact, tests/test_partition.py). */
void       b200_partition_2x2( b200_dim_t n_thread, b200_dim_t work1, b200_dim_t work2, b200_dim_t* nt1, b200_dim_t* nt2 );
void       b200_range_sub( b200_dim_t work_id, b200_dim_t n_way, b200_dim_t n, b200_dim_t bf, int handle_edge_low,
                      b200_dim_t* start, b200_dim_t* end );
b200_err_t b200_dist_plan( int world, int rank, b200_dim_t m, b200_dim_t n, b200_dim_t k, b200_dim_t kb, b200_dist_plan_t* plan );

b200_err_t b200_dist_unique_id( void* id );
b200_err_t b200_dist_init( int world, int rank, const void* id );
b200_err_t b200_dist_finalize( void );

/* C := beta*C + alpha*A*B, global m x n x k, on the Pr x Pc grid of b200_dist_plan (dt = d or z; column-major device
 * shards): a_loc = this rank's plan.na k panels of A's row block, each (m1-m0) x kb with leading dimension m1-m0, one
 * after the other; b_loc = its plan.nb k panels of B's column block, each kb x (n1-n0) with leading dimension kb;
 * c_loc = its block of C with strides (rs_c, cs_c).  Each step all-gathers L k panels in the row/column groups on the
 * engine's communication stream, double buffered under ONE k-panel launch per step (b200_gemm_kpanels' kernel).
 * Asynchronous: returns with the work queued on the calling thread's stream. */
b200_err_t b200_dist_gemm( int dt, b200_dim_t m, b200_dim_t n, b200_dim_t k, b200_dim_t kb,
                      const void* alpha, const void* a_loc, const void* b_loc,
                      const void* beta, void* c_loc, b200_inc_t rs_c, b200_inc_t cs_c, int flags );
double     b200_dist_last_wait_ms( void );
/* One-sided transport for STATIC shards (collective over all ranks): the allocations behind a_loc and b_loc are shared
 * between the ranks (CUDA IPC handles carried by the communicator) and later b200_dist_gemm calls on exactly these
 * pointers with B200_DIST_AB_STATIC move the k panels with peer copies issued by the receiving rank -- the copy engines
 * pull over NVLink, no SM is taken from the DMMA kernel, nothing is asked of the owner.  Returns B200_FAILURE on every
 * rank alike when the memory cannot be shared (the NCCL all-gather transport then stays in use).
 * b200_dist_transport(): what the last b200_dist_gemm of this rank used: 0 nothing to move, 1 NCCL all-gather,
 * 2 copy-engine gets. */
b200_err_t b200_dist_register( const void* a_loc, const void* b_loc );
b200_err_t b200_dist_unregister( const void* a_loc, const void* b_loc );
int        b200_dist_transport( void );

/* Skinny products (m, n >> k: the shapes bli_gemmsup serves, frame/3/bli_l3_sup.c:37-135): 1-D split of C over ALL
 * ranks in units of 128 (b200_range_sub( rank, world, n or m, 128, 0 )).  split = B200_DIST_COLS: c_loc and b are this
 * rank's columns of C and of B, a is the whole m x k operand; split = B200_DIST_ROWS: c_loc and a are its rows, b is the
 * whole k x n operand.  The whole ("small") operand is broadcast from rank `root` into every rank's buffer first
 * (column-major, device resident); root < 0: it is already replicated and no collective runs. */
b200_err_t b200_dist_gemm_1d( int dt, int split, int root, b200_dim_t m, b200_dim_t n, b200_dim_t k,
                      const void* alpha, void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      void* b, b200_inc_t rs_b, b200_inc_t cs_b,
                      const void* beta, void* c_loc, b200_inc_t rs_c, b200_inc_t cs_c );

/* trsm with B split into column blocks (side = left; row blocks for side = right) in units of 128 over all ranks --
 * the reference's only parallel trsm loops run over the columns of B (frame/3/trsm/bli_trsm_cntl.c:446-451,
 * bli_trsm_ll_ker_var2.c:209-212) -- and the triangular A replicated.  b_loc is this rank's block; m, n are GLOBAL.
 * root >= 0: A (column-major, device resident on every rank) is first broadcast from that rank; root < 0: already
 * replicated, no collective at all. */
b200_err_t b200_dist_trsm( int dt, int side, int uploa, int transa, int diaga, int root,
                      b200_dim_t m, b200_dim_t n, const void* alpha,
                      void* a, b200_inc_t rs_a, b200_inc_t cs_a,
                      void* b_loc, b200_inc_t rs_b, b200_inc_t cs_b );

/* ---- blocksizes ------------------------------------------------------------
 * What bli_cntx_init_b200 registers through bli_cntx_set_blkszs()
 * (config/zen3/bli_cntx_init_zen3.c:37-258 is the pattern): the CTA tile
 * shape the engine uses for datatype dt.  bs is one of the bszid_t names
 * below (frame/include/bli_type_defs.h, bszid_t). Returns -1 if unknown. */
#define B200_BS_MR 0
#define B200_BS_NR 1
#define B200_BS_MC 2
#define B200_BS_KC 3
#define B200_BS_NC 4
b200_dim_t b200_blksz( int dt, int bs );

/* ---- measurement helpers ---------------------------------------------------
 * Device-side peak microbenchmarks (no reference analogue; they provide the
 * FP64/FP32 roofline denominators SURVEY.md section 8d asks for).
 * kind: 0 = DFMA (fp64 FMA pipe), 1 = DMMA (mma.sync m8n8k4 f64),
 *       2 = FFMA (fp32 FMA pipe), 3 = FFMA2 (packed fma.rn.f32x2).
 * Runs for about `millis` ms and returns achieved TFLOP/s (< 0 on error). */
double b200_measure_peak( int kind, int millis );

/* Number of CUDA kernels this engine has launched so far in this process
 * (gemm tile kernels, trsm block solves, strided copy/scale helpers); bench.py
 * reports the difference over its timed region as "gpu_launches". */
unsigned long long b200_launch_count( void );

/* ---- schedule arithmetic of the engine itself (host only, no GPU needed; tests/test_host_plans.py) ------------
 * b200_splitk_plan: the split-k tail of mid-size dgemm (csrc/gemm_d.cu).  `tiles` 128x128 output tiles on `grid`
 * persistent CTAs, `kt` 16-wide k steps per tile.  Returns S, the number of k chunks every tile of the partial last
 * wave is cut into (0: whole tiles only), and in *full the number of tiles that stay whole.  The reference has no
 * counterpart: it never splits k inside one gemm (frame/3/gemm/bli_gemm_blk_var3.c:110-112; bli_rntm_factorize,
 * frame/base/bli_rntm.c:424-489, factors threads over ic x jc only); this is the rule that decides when the engine does.
 * b200_trsm_upload_plan: the order in which a host-resident triangular A travels while trsm runs (csrc/host_trsm.cuh):
 * pieces (r0, r1, c0, c1, launches) of the effective m x m view, in the order the recursive solve
 * (bli_trsm_blk_var1's role, frame/3/trsm/bli_trsm_blk_var1.c:40-188) reads them; `launches` = launches of the solve
 * that wait for the piece.  Writes at most `cap` pieces of five numbers each to `out`, returns the number of pieces. */
int        b200_splitk_plan( b200_dim_t tiles, int grid, b200_dim_t kt, int* full );
int        b200_trsm_upload_plan( b200_dim_t m, int leaf_rows, int upper, b200_dim_t* out, int cap );
/* b200_trsm_rowblock_plan: the same list for the row-block (left-looking) pipeline that serves tall systems with pinned
 * host operands (csrc/host_trsm.cuh: trsm_host_rowpipe): per block row of `rb` rows, in processing order (top down for
 * lower, bottom up for upper), the block of A its single update gemm reads, A[j, 0:j] (resp. A[j, j+1:]), followed by the
 * pieces of the diagonal block in the recursion's order.  rb = 0 asks for the engine's own choice (option trsm_host_rb);
 * returns 0 when the shape is served by the column-block pipeline instead. */
int        b200_trsm_rowblock_plan( b200_dim_t m, b200_dim_t n, int leaf_rows, int upper, b200_dim_t rb, b200_dim_t* out, int cap );

/* Tuning knobs, e.g. ("dgemm_cfg", 9); not part of the reference surface.  Every key can be preset in the environment
 * as BLIS_B200_<KEY> (the reference's bli_env convention, frame/base/bli_env.c:68).  Keys that change WHICH schedule
 * serves a call (results stay within the same error bound; exact inputs give identical bits):
 *   dgemm_splitk      1 (default): mid-size dgemm cuts the tiles of a partial last wave into k chunks; 0: never split k
 *   dmma_cst          k up to which dgemm moves D through the TMA unit both ways (default 256; 0: never)
 *   trsm_fused        1 (default): fused 256-row diagonal-panel kernel for dtrsm; 0: 64-row block solves
 *   trsm_host_pipe    1 (default): trsm with pinned host operands runs its transfers under the solve; 0: in sequence
 *   trsm_host_rb      rows per block of that pipeline (0, default: m/32 in whole 256 rows; -1: same as trsm_host_pipe 0)
 *   host_kpipe        1 (default): gemm with host operands and long k is pipelined over k panels
 *   batch_grouped     1 (default): small device-resident problems of b200_gemm_batch share ONE launch;
 *   batch_grouped_max   "small" means m*n*k at most this (default 128^3)
 * Kernel selection / sweeps: dgemm_cfg, zgemm_cfg, sgemm_cfg, cgemm_cfg, grid_mult, dynamic_tiles, raster_group,
 * tma_l2_promotion, transpose_y, ktri_skip, dmma_pp, dist_ab_static, reserve_sms. */
b200_err_t b200_set_option( const char* key, long long value );

#ifdef __cplusplus
}
#endif
#endif /* BLIS_B200_H */
