/*
 * gemmul8_b200_mp.h -- multi-GPU entry points of the B200-native Ozaki-II GEMM emulation (one process per GPU, one node).
 *
 * NEW relative to the reference, which is single-GPU ("Distributed communication backend: none"); it is what
 * BASELINE.json's north_star asks for: "Large problems are partitioned across the 8 x B200 box as a 2D block decomposition
 * of C, with A row-panels and B column-panels distributed ... over NVLink", host code in C++ behind a C ABI.
 *
 * The path shards with no data-path reduction: the shift of a row of op(A) / a column of op(B) is taken over all of k,
 * residues and the CRT are element-wise, so rank (p, q) of a P x Q grid needs row panel p of A and column panel q of B and
 * K is never split.  Every rank starts with 1/Q of its A row panel and 1/P of its B column panel (HPL-like ownership):
 *
 *     a_slice : (m / P) x (k / Q), columns [q k/Q, (q+1) k/Q) of row panel p      (column-major, leading dimension lda)
 *     b_slice : k x (n / Q / P),   columns [p w, (p+1) w) of column panel q, w = n / Q / P            (leading dimension ldb)
 *     c_block : (m / P) x (n / Q)  this rank's block of C                                              (leading dimension ldc)
 *
 * gemmul8_b200_pgemm assembles the panels in grid-owned buffers and runs the single-GPU path (include/gemmul8_b200.h)
 * block-wise while pieces are still arriving: A travels as up to four row pieces, B as P column pieces.  Two transports:
 *     GEMMUL8_MP_EXCHANGE_NCCL  ncclAllGather / ncclBroadcast inside grid rows / columns on two side streams
 *     GEMMUL8_MP_EXCHANGE_COPY  copy engines over peer memory (CUDA IPC): every rank pushes its pieces straight into its
 *                               peers' panel buffers with cudaMemcpy2DAsync and raises a flag there; the consumer's stream
 *                               waits on the flag (stream memory operation).  No SM is used by the exchange, so the
 *                               persistent GEMM keeps its full pipeline depth.
 * Accurate mode adds one tiny exchange: the int32 row / column maxima of the bound product are max-all-reduced inside the grid
 * row / column (ncclAllReduce), so shifts -- and with them every bit of C -- are those of the unpartitioned accurate product.
 *
 * All calls are collective over the ranks of the grid and asynchronous on `stream` (except create / destroy).
 */
#ifndef GEMMUL8_B200_MP_H
#define GEMMUL8_B200_MP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gemmul8_b200_grid gemmul8_b200_grid;

enum { GEMMUL8_MP_EXCHANGE_NCCL = 0, GEMMUL8_MP_EXCHANGE_COPY = 1 };
enum { GEMMUL8_MP_ID_BYTES = 128 };   /* sizeof(ncclUniqueId) */

/* Rank 0 makes the bootstrap id; the caller ships the 128 bytes to every rank (MPI_Bcast, a file, torch.distributed ...). */
int gemmul8_b200_mp_unique_id(void *id_bytes);

/* P x Q grid, rank = p * Q + q, on the CURRENT device of each process.  a_panel_bytes / b_panel_bytes: capacity of the
 * grid-owned panel buffers, i.e. the largest (m/P) x k and k x (n/Q) panels (in bytes) later calls will assemble.
 * exchange: GEMMUL8_MP_EXCHANGE_*; COPY needs peer access between the GPUs and CUDA IPC between the processes and fails
 * with GEMMUL8_ERR_CUDA where that is unavailable (the caller may then create an NCCL grid instead). */
int gemmul8_b200_grid_create(const void *id_bytes, int rank, int nranks, int P, int Q, size_t a_panel_bytes, size_t b_panel_bytes,
                             int exchange, gemmul8_b200_grid **out);
/* The same for a caller that already owns a communicator over exactly the grid's ranks (ncclComm_t passed as void *). */
int gemmul8_b200_grid_create_from_comm(void *nccl_comm, int P, int Q, size_t a_panel_bytes, size_t b_panel_bytes, int exchange,
                                       gemmul8_b200_grid **out);
int gemmul8_b200_grid_destroy(gemmul8_b200_grid *grid);
/* rank coordinates: out[0..3] = P, Q, p, q */
int gemmul8_b200_grid_coords(const gemmul8_b200_grid *grid, int out[4]);

typedef struct {
    size_t m, n, k;              /* the GLOBAL problem; m % P == 0, n % (P Q) == 0, k % Q == 0 */
    const void *alpha;           /* host pointers, type of C (as gemmul8::gemm) */
    const void *a_slice; size_t lda;
    const void *b_slice; size_t ldb;
    const void *beta;
    void *c_block; size_t ldc;
    unsigned num_moduli;
    int fastmode;
    void *work;                  /* device, >= gemmul8_b200_pgemm_worksize bytes */
    int dtype_A, dtype_B, dtype_C; /* gemmul8_dtype_t; complex types: fast mode only (panels are assembled, then one call) */
    int compute_type;            /* gemmul8_compute_t: REAL_DEFAULT for real types, COMPLEX_* for complex types */
    void *stream;                /* cudaStream_t of the compute work */
    unsigned flags;              /* GEMMUL8_FLAG_TIMERS | GEMMUL8_FLAG_PHASE_LOG as for gemmul8_b200_gemm_part */
    double timers_ns[4];
} gemmul8_b200_pargs;

/* bytes of `work` for the rank's (m/P) x (n/Q) block: the single-GPU workSize of that block (real types; for complex types ask
 * the single-GPU library for the block's workSize with the compute type) */
size_t gemmul8_b200_pgemm_worksize(const gemmul8_b200_grid *grid, size_t m, size_t n, size_t k, unsigned num_moduli);
int gemmul8_b200_pgemm(gemmul8_b200_grid *grid, gemmul8_b200_pargs *args);

/* Host logic, exported for tests: up to `want` row ranges of [0, rows) whose starts are multiples of 256 (whole GEMM tiles);
 * bounds[0..count] are the cut points, returns count. */
int gemmul8_b200_mp_row_pieces(size_t rows, int want, size_t bounds[9]);

const char *gemmul8_b200_mp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GEMMUL8_B200_MP_H */
