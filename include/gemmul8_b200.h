/*
 * gemmul8_b200.h -- C ABI of the B200-native Ozaki-scheme-II GEMM emulation.
 *
 * This is the drop-in boundary for the hot path of ptrkgtsch/mixed-GEMMul8:
 *   gemmul8::workSize  (reference GEMMul8/include/gemmul8.hpp:18-22, GEMMul8/src/gemmul8.cu:129-147)
 *   gemmul8::gemm<..>  (reference GEMMul8/include/gemmul8.hpp:29-287, GEMMul8/src/gemmul8.cu:149-1316)
 * Plain C types only: pointers, sizes and enums.  The C++ drop-in header include/gemmul8.hpp
 * (namespace gemmul8, same template specialisations as the reference) is a thin inline shim over
 * these entry points, and INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * All matrices are column-major device pointers exactly as in the reference; alpha/beta are HOST
 * pointers to one element of C's type (reference dereferences them on the host,
 * GEMMul8/src/gemmul8.cu:288).  `work` is a device buffer of at least gemmul8_b200_worksize()
 * bytes, 16-byte aligned.  Every call is asynchronous on `stream` unless GEMMUL8_FLAG_TIMERS is set.
 *
 * There is no CPU fallback: every entry point that computes returns GEMMUL8_ERR_CUDA when no
 * sm_100 device/context is available.
 */
#ifndef GEMMUL8_B200_H
#define GEMMUL8_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of A, B, C (reference: the 12 template specialisations, gemmul8.hpp:49-287) */
typedef enum {
    GEMMUL8_F32 = 0, /* float            */
    GEMMUL8_F64 = 1, /* double           */
    GEMMUL8_C32 = 2, /* cuFloatComplex   */
    GEMMUL8_C64 = 3  /* cuDoubleComplex  */
} gemmul8_dtype_t;

/* same numbering as gemmul8::computeType_t (gemmul8.hpp:7-12) */
typedef enum {
    GEMMUL8_REAL_DEFAULT              = 0,
    GEMMUL8_COMPLEX_BIG_MATRIX_ENCODE = 1,
    GEMMUL8_COMPLEX_CLASSIC_MULT      = 2,
    GEMMUL8_COMPLEX_KARATSUBA_MULT    = 3
} gemmul8_compute_t;

/* same numbering as cublasOperation_t */
typedef enum { GEMMUL8_OP_N = 0, GEMMUL8_OP_T = 1, GEMMUL8_OP_C = 2 } gemmul8_op_t;

typedef enum {
    GEMMUL8_OK              = 0,
    GEMMUL8_ERR_COMPUTETYPE = 1, /* reference prints "Unsupported compute type..." and returns zeros */
    GEMMUL8_ERR_ARGUMENT    = 2, /* num_moduli outside 2..20, k > 2^17, null pointers, bad dtype combo */
    GEMMUL8_ERR_CUDA        = 3  /* CUDA runtime/driver error, or no sm_100 device */
} gemmul8_status_t;

enum {
    GEMMUL8_FLAG_TIMERS = 1u, /* bracket the 4 phases with CUDA events and fill timers_ns (synchronises) */
    GEMMUL8_FLAG_STAGE_SCALING  = 1u << 4, /* run only phase 0: shifts + residue slices (parity tests) */
    GEMMUL8_FLAG_STAGE_RESIDUES = 1u << 5, /* run phases 0-2: ... + per-modulus products mod m_j      */
    GEMMUL8_FLAG_FUSED_CRT      = 1u << 6, /* real types: product, residue reduction, CRT, inverse scaling and alpha / beta in ONE kernel
                                               (CRT accumulators in registers across the modulus walk; no residue matrix in HBM).
                                               Same bits of C; measured slower than the two kernels on B200, see DESIGN.md 3.4 */
    GEMMUL8_FLAG_GEMM_SIMT      = 1u << 8, /* debug: use the CUDA-core int8 GEMM instead of tcgen05   */
    GEMMUL8_FLAG_HOST_SERIAL    = 1u << 9, /* gemm_host: copy in, compute, copy out in series (no wavefront) */
    GEMMUL8_FLAG_STRIPS         = 1u << 10, /* gemm: force the column-strip pipeline on three streams (default: by size, option "strips") */
    GEMMUL8_FLAG_ONLY_SCALE_A   = 1u << 11, /* real types: shifts + residues of A only, then return (B may still be in flight) */
    GEMMUL8_FLAG_SKIP_SCALE_A   = 1u << 12, /* real types: A's shifts + residues are already in `work` (previous flag)        */
    GEMMUL8_FLAG_ONLY_BOUND     = 1u << 14, /* real types, accurate mode: bound product only; its int32 row maxima (m) are left at
                                               work + off_A8i + sizeA, its column maxima (n) at work + off_B8i + sizeB           */
    GEMMUL8_FLAG_SKIP_BOUND     = 1u << 15, /* real types, accurate mode: the maxima are already there (previous flag, possibly
                                               combined with those of other blocks by the caller)                               */
    GEMMUL8_FLAG_DEVICE_SCALARS = 1u << 16, /* alpha and beta point to DEVICE memory (CUBLAS_POINTER_MODE_DEVICE callers): they are read by
                                               the CRT kernel, never on the host (the reference dereferences on the host,
                                               GEMMul8/src/gemmul8.cu:288).  Not for gemm_host.                                  */
    GEMMUL8_FLAG_EXCLUSIVE_SMS  = 1u << 17, /* gemm_part: nothing else needs room on the SMs beside the product (e.g. the panel exchange runs
                                               on copy engines): full pipeline depth instead of the 4 stages that leave room for NCCL's kernels */
    GEMMUL8_FLAG_PHASE_LOG      = 1u << 13  /* record the phase boundaries as events WITHOUT synchronising; the times of all such
                                               calls of this host thread are summed by gemmul8_b200_phase_log_collect()      */
};

/* Arguments of one gemm call, in the reference's argument order (gemmul8.hpp:30-47). */
typedef struct {
    int op_A, op_B;            /* gemmul8_op_t */
    size_t m, n, k;
    const void *alpha;         /* host pointer, type of C (device pointer with GEMMUL8_FLAG_DEVICE_SCALARS) */
    const void *A; size_t lda; /* device */
    const void *B; size_t ldb; /* device */
    const void *beta;          /* host pointer, type of C (device pointer with GEMMUL8_FLAG_DEVICE_SCALARS) */
    void *C; size_t ldc;       /* device */
    unsigned num_moduli;       /* 2..20 */
    int fastmode;              /* 1 = fast (vector-norm bound), 0 = accurate (int8 bound product) */
    void *work;                /* device, >= gemmul8_b200_worksize bytes */
    int compute_type;          /* gemmul8_compute_t */
    int dtype_A, dtype_B, dtype_C; /* gemmul8_dtype_t */
    void *stream;              /* cudaStream_t (NULL = legacy default stream, as the reference) */
    unsigned flags;
    double timers_ns[4];       /* out: {scaling, int8 GEMM, int32->residue, inverse scaling} in ns
                                  (reference returns the same 4 numbers, gemmul8.cu:17,291).  The
                                  residue reduction is fused into the GEMM epilogue here, so [2] = 0
                                  (with GEMMUL8_FLAG_FUSED_CRT also [3] = 0: [1] covers all three). */
} gemmul8_b200_args;

/* Byte offsets of the sub-buffers inside `work`; identical to the reference's carve
 * (GEMMul8/src/gemmul8.cu:229-234, :659-664, :806-815) so that tests can compare them bit-for-bit. */
typedef struct {
    size_t lda8i;   /* = ldb8i: padded inner dimension (multiple of 16)              */
    size_t m_pad;   /* padded row count of the product (multiple of 4)               */
    size_t sizeA;   /* elements per modulus slice of A   (lda8i * m_pad)             */
    size_t sizeB;   /* elements per modulus slice of B   (lda8i * n)                 */
    size_t sizeC;   /* elements per modulus residue matrix (multiple of 16)          */
    size_t off_A8i, off_A8i_imag;   /* int8  [N][sizeA]  (imag: CLASSIC/KARATSUBA only) */
    size_t off_B8i, off_B8i_imag;   /* int8  [N][sizeB]                               */
    size_t off_C8u, off_C8u_imag;   /* uint8 [N][sizeC]                               */
    size_t off_C32i, off_C32i_imag; /* int32 [sizeC] scratch (accurate-mode bound product) */
    size_t off_sftA, off_sftB;      /* int16 shifts (stored negated, as the reference) */
    size_t total;                   /* == gemmul8_b200_worksize()                     */
} gemmul8_b200_layout;

/* Optional, once per device (device < 0: the current one): everything the first gemm call would otherwise do lazily --
 * read the tuning defaults from the environment and probe the SM placement table of the CTA-pair GEMM (one small
 * allocation kept for the life of the process, one launch on a private stream).  A first gemm call that is being
 * captured into a CUDA graph never probes; it runs without the table (same results, work assigned by block index).
 * The reference has no such call (it re-uploads its constant tables on every gemm, gemmul8.cu:236-241). */
int gemmul8_b200_init(int device);

/* Tuning / debug options, process-wide and thread-safe.  Their defaults come from the environment, read once:
 *   "gemm_pair"  (OZ_GEMM_PAIR)  -1 auto | 0 single-CTA kernel | 1 CTA-pair kernel      "band" (OZ_BAND), "pair_band" (OZ_PAIR_BAND)
 *   "pair_stages" (OZ_PAIR_STAGES) 0 auto | 4 | 5 | 6     "encode_reference" (GEMMUL8_B200_ENCODE=reference) 0 | 1
 *   "fused_k" (GEMMUL8_B200_FUSED_K) largest k that takes the single-kernel product + CRT path (0 = never)
 *   "scale_fork" (GEMMUL8_B200_SCALE_FORK) 1 | 0: small operands are scaled on two streams side by side
 *   "tma_store" (GEMMUL8_B200_TMA_STORE) 0 | 1: residues of the pair GEMM leave through TMA bulk tensor stores
 *   "strips" (GEMMUL8_B200_STRIPS) 0 | 1 | 2..8: column-strip pipeline of large real fast-mode calls (encoders of the next
 *                                  strips of B and the CRT of the previous strip run beside the products): 0 = by size,
 *                                  1 = never, 2..8 = that many strips whenever the call allows it
 *   "strip_calls" (read-only)      calls that took that pipeline so far */
int gemmul8_b200_set_option(const char *name, int value);
int gemmul8_b200_get_option(const char *name, int *value);

/* reference: gemmul8::workSize, GEMMul8/src/gemmul8.cu:129-147.  Returns 0 (and prints
 * "Unknown compute type") for an invalid compute_type, as the reference does. */
size_t gemmul8_b200_worksize(size_t m, size_t n, size_t k, unsigned num_moduli, int compute_type);

/* The carve of `work`; returns GEMMUL8_ERR_COMPUTETYPE for an invalid compute_type. */
int gemmul8_b200_work_layout(size_t m, size_t n, size_t k, unsigned num_moduli, int compute_type,
                             gemmul8_b200_layout *out);

/* reference: gemmul8::gemm<TA,TB,TC>, GEMMul8/src/gemmul8.cu:149-1316. */
int gemmul8_b200_gemm(gemmul8_b200_args *args);

/* Same call with HOST matrices: copies op inputs host->device (and C when beta != 0), runs the
 * emulation and copies C back.  `dev_scratch` must hold A, B, C and the workspace
 * (gemmul8_b200_host_scratch_size bytes).  This is the end-to-end plugin call timed by bench.py. */
size_t gemmul8_b200_host_scratch_size(const gemmul8_b200_args *args);
int gemmul8_b200_gemm_host(gemmul8_b200_args *args_with_host_matrices, void *dev_scratch);

/* Block-wise entry for callers whose operands arrive piecewise (multi-GPU panel exchange, streaming from the
 * host): one or more steps of the real fast-mode path restricted to rows [row0, row1) of op(A) / C and columns
 * [col0, col1) of op(B) / C, inside the workspace of the FULL problem described by `args` (reference: the phases
 * of gemm<>, GEMMul8/src/gemmul8.cu:253, :265-273, :288).  row0 and col0 must be multiples of 256.  A PRODUCT
 * step needs the SCALE steps of its rows and columns to have run (in stream order) before. */
enum { GEMMUL8_PART_SCALE_A = 1, GEMMUL8_PART_SCALE_B = 2, GEMMUL8_PART_PRODUCT = 4 };
int gemmul8_b200_gemm_part(gemmul8_b200_args *args, int parts, size_t row0, size_t row1, size_t col0, size_t col1);

/* Low-memory call (real types: fast and accurate mode; complex types: fast mode).  gemm() keeps the residue slices of ALL of op(A) and
 * op(B) resident -- N (m + n) k bytes, 184 GiB at 65536^3 (GEMMul8/src/gemmul8.cu:27-60); the reference's
 * README.md:3 points at a `memory-lt` branch for this, which is not in the tree.  Here C is produced in blocks of
 * block_rows x block_cols (multiples of 256, or the whole dimension) and `args->work` only needs
 * gemmul8_b200_worksize_blocked() bytes: N k (block_rows + block_cols) + N block_rows block_cols + 6 (m + n) + 1024
 * (+ padding).  Result bits are those of gemmul8_b200_gemm (shifts are per row / column, residues and CRT per
 * element, the accurate-mode bound is a maximum over blocks).  gemmul8_b200_plan_blocks picks the block sizes with
 * the least re-encoding for a workspace of at most `max_bytes`.  timers_ns (GEMMUL8_FLAG_TIMERS) are the sums of the
 * phases over all blocks, as in gemmul8_b200_gemm. */
size_t gemmul8_b200_worksize_blocked(size_t m, size_t n, size_t k, unsigned num_moduli, size_t block_rows, size_t block_cols);
/* Complex types (fast mode only): gemm_blocked runs one complete complex call per C block in a workspace of
 * workSize(block_rows, block_cols, k, N, compute_type) bytes -- this function returns exactly that (0 for bad blocks). */
size_t gemmul8_b200_worksize_blocked_complex(size_t m, size_t n, size_t k, unsigned num_moduli, int compute_type, size_t block_rows,
                                             size_t block_cols);
int gemmul8_b200_plan_blocks(size_t m, size_t n, size_t k, unsigned num_moduli, size_t max_bytes,
                             size_t *block_rows, size_t *block_cols, size_t *work_bytes);
int gemmul8_b200_gemm_blocked(gemmul8_b200_args *args, size_t block_rows, size_t block_cols);

/* Debug/parity: raw int32 product of modulus slice `j` (what the reference's cublasGemmEx at
 * gemmul8.cu:265 writes into C32i), column-major with leading dimension m_pad.  Requires the
 * slices to be present in `work` (GEMMUL8_FLAG_STAGE_SCALING or a full gemm call). */
int gemmul8_b200_product_i32(const gemmul8_b200_args *args, unsigned j, int32_t *C32i_out, int imag_part);

/* Constant tables (moduli, CRT weights, ...) for host-side tooling; row = num_moduli - 2. */
int gemmul8_b200_modulus(unsigned j);                /* m_j, j in 0..19 */
double gemmul8_b200_crt_weight(unsigned num_moduli, unsigned j, int part /*0=single,1=hi,2=lo*/);

/* Sum of the phase times (ns, same slots as timers_ns) of every GEMMUL8_FLAG_PHASE_LOG call this host thread has made
 * since the last collect, and how many calls that was.  Synchronises on the recorded events. */
int gemmul8_b200_phase_log_collect(double timers_ns[4], unsigned *calls);

/* Number of CUDA kernels this library has launched so far in this process (bench.py: gpu_launches). */
unsigned long long gemmul8_b200_launch_count(void);

const char *gemmul8_b200_last_error(void);
const char *gemmul8_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GEMMUL8_B200_H */
