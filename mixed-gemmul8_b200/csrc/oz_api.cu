// C ABI + host orchestration of the B200-native Ozaki-II GEMM emulation (see include/gemmul8_b200.h).
//
// Differences from the reference's host path (GEMMul8/src/gemmul8.cu:149-577), all deliberate:
//   * no cudaDeviceSynchronize between phases (reference: 2 + 4N per call, gemmul8.cu:10-18) and no
//     per-call cudaMemcpyToSymbol (gemmul8.cu:236-241): tables are static __constant__ data;
//   * everything is enqueued on the caller's stream; the call is thread-safe (no mutable globals);
//   * the N int8 GEMMs + N int32->uint8 passes are ONE persistent tcgen05 kernel (oz_gemm.cu);
//   * the accurate-mode bound product never materialises C32i: its epilogue reduces row / column
//     maxima directly into two small int32 vectors.
// The carve of `work` is kept identical to the reference (gemmul8.cu:229-234) so that parity tests
// can compare the int8 slices, the shift vectors and the uint8 residues in place.
#include "../../include/gemmul8_b200.h"
#include "oz_common.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace oz {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Options: atomics behind a snapshot struct; the environment is consulted exactly once.
namespace {
struct Option { const char *name; const char *env; std::atomic<int> value; int lo, hi; };
Option g_options[] = {
    {"gemm_pair", "OZ_GEMM_PAIR", {-1}, -1, 1},
    {"band", "OZ_BAND", {16}, 1, 1024},
    {"pair_band", "OZ_PAIR_BAND", {0}, 0, 4096},
    {"pair_stages", "OZ_PAIR_STAGES", {0}, 0, 6},
    {"encode_reference", nullptr, {0}, 0, 1},
    {"fused_k", "GEMMUL8_B200_FUSED_K", {0}, 0, 1 << 17},
    {"scale_fork", "GEMMUL8_B200_SCALE_FORK", {1}, 0, 1},
    {"tma_store", "GEMMUL8_B200_TMA_STORE", {0}, 0, 1},
    {"strips", "GEMMUL8_B200_STRIPS", {0}, 0, 8},
};
std::atomic<int> g_strip_calls{0};   // calls that took the column-strip pipeline (read-only option "strip_calls")
std::once_flag g_options_once;
void load_options_from_env() {
    std::call_once(g_options_once, [] {
        for (auto &o : g_options) {
            const char *e = o.env ? getenv(o.env) : nullptr;
            if (e && *e) {
                const int v = atoi(e);
                if (v >= o.lo && v <= o.hi) o.value.store(v, std::memory_order_relaxed);
            }
        }
        const char *e = getenv("GEMMUL8_B200_ENCODE");
        if (e && !strcmp(e, "reference")) g_options[4].value.store(1, std::memory_order_relaxed);
    });
}
Option *find_option(const char *name) {
    if (!name) return nullptr;
    for (auto &o : g_options) if (!strcmp(o.name, name)) return &o;
    return nullptr;
}
}  // namespace
const Tuning &tuning() {
    load_options_from_env();
    thread_local Tuning t;
    t.gemm_pair        = g_options[0].value.load(std::memory_order_relaxed);
    t.band             = g_options[1].value.load(std::memory_order_relaxed);
    t.pair_band        = g_options[2].value.load(std::memory_order_relaxed);
    t.pair_stages      = g_options[3].value.load(std::memory_order_relaxed);
    t.encode_reference = g_options[4].value.load(std::memory_order_relaxed);
    t.fused_k          = g_options[5].value.load(std::memory_order_relaxed);
    t.scale_fork       = g_options[6].value.load(std::memory_order_relaxed);
    t.tma_store        = g_options[7].value.load(std::memory_order_relaxed);
    t.strips           = g_options[8].value.load(std::memory_order_relaxed);
    return t;
}
}  // namespace oz

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *where) {
    return fail(GEMMUL8_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

inline size_t ceil_to(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Per-launch scratch of the pair GEMM (see GemmProblem::claims): the start of the int32 area that the reference carves for
// its per-modulus product and that this library never fills.  Launches that share a `work` buffer are ordered on one
// stream by every caller in this file; calls that may overlap have their own `work`.
uint32_t *claims_of(uint8_t *work, const oz::Layout &L) {
    return 4 * L.sizeC >= oz::kClaimBytes ? reinterpret_cast<uint32_t *>(work + L.off_C32i) : nullptr;
}
bool dev_scalars(const gemmul8_b200_args *a) { return (a->flags & GEMMUL8_FLAG_DEVICE_SCALARS) != 0; }
bool is_complex(int dt) { return dt == GEMMUL8_C32 || dt == GEMMUL8_C64; }
size_t elem_size(int dt) { return dt == GEMMUL8_F32 ? 4 : dt == GEMMUL8_F64 ? 8 : dt == GEMMUL8_C32 ? 8 : 16; }

// reference: workSize_real / workSize_bigmatrix / workSize_kara, GEMMul8/src/gemmul8.cu:27-127
bool compute_layout(size_t m, size_t n, size_t k, unsigned N, int ct, oz::Layout &L) {
    const bool big = ct == GEMMUL8_COMPLEX_BIG_MATRIX_ENCODE;
    const bool kara = ct == GEMMUL8_COMPLEX_CLASSIC_MULT || ct == GEMMUL8_COMPLEX_KARATSUBA_MULT;
    if (!(ct == GEMMUL8_REAL_DEFAULT || big || kara)) return false;
    L.lda8i = ceil_to(big ? 2 * k : k, 16);
    L.m_pad = ceil_to(big ? 2 * m : m, 4);
    L.sizeA = L.lda8i * L.m_pad;
    L.sizeB = L.lda8i * n;
    L.sizeC = ceil_to(L.m_pad * n, 16);
    const size_t vecA = ceil_to(m, 16), vecB = ceil_to(n, 16);
    size_t off = 0;
    L.off_A8i = off;       off += L.sizeA * N;
    L.off_A8i_imag = off;  if (kara) off += L.sizeA * N;
    L.off_B8i = off;       off += L.sizeB * N;
    L.off_B8i_imag = off;  if (kara) off += L.sizeB * N;
    L.off_C8u = off;       off += L.sizeC * N;
    L.off_C8u_imag = off;  if (kara) off += L.sizeC * N;
    L.off_C32i = off;      off += 4 * L.sizeC;
    L.off_C32i_imag = off; if (kara) off += 4 * L.sizeC;
    L.off_sftA = off;      off += 2 * vecA;
    L.off_sftB = off;      off += 2 * vecB;
    L.total = off;
    return true;
}

int check_args(const gemmul8_b200_args *a, bool need_work = true) {
    if (!a) return fail(GEMMUL8_ERR_ARGUMENT, "null argument block");
    if (a->num_moduli < 2 || a->num_moduli > 20) return fail(GEMMUL8_ERR_ARGUMENT, "num_moduli must be in 2..20");
    if (a->k > (size_t(1) << 17)) return fail(GEMMUL8_ERR_ARGUMENT, "k must be <= 2^17");
    if (a->m > (size_t(1) << 24) || a->n > (size_t(1) << 18)) return fail(GEMMUL8_ERR_ARGUMENT, "m or n too large");
    if (a->dtype_A < 0 || a->dtype_A > 3 || a->dtype_B < 0 || a->dtype_B > 3 || a->dtype_C < 0 || a->dtype_C > 3)
        return fail(GEMMUL8_ERR_ARGUMENT, "bad dtype tag");
    const bool cplx = is_complex(a->dtype_C);
    if (is_complex(a->dtype_A) != cplx || is_complex(a->dtype_B) != cplx)
        return fail(GEMMUL8_ERR_ARGUMENT, "A, B and C must be all real or all complex");
    // reference: real types accept only REAL_DEFAULT, complex types only the COMPLEX_* values
    // (GEMMul8/src/gemmul8.cu:174-177, :1166-1177): message on stderr, zero timers, C untouched
    const bool ct_ok = cplx ? (a->compute_type >= 1 && a->compute_type <= 3) : (a->compute_type == GEMMUL8_REAL_DEFAULT);
    if (!ct_ok) {
        fprintf(stderr, "Unsupported compute type for the argument types.\n");
        return fail(GEMMUL8_ERR_COMPUTETYPE, "unsupported compute type for the argument types");
    }
    if (a->m && a->n && (!a->C || (need_work && !a->work) || !a->alpha || !a->beta)) return fail(GEMMUL8_ERR_ARGUMENT, "null pointer");
    if (a->m && a->n && a->k && (!a->A || !a->B)) return fail(GEMMUL8_ERR_ARGUMENT, "null matrix pointer");
    return GEMMUL8_OK;
}

// Phase boundaries as CUDA events on the caller's stream.
//   GEMMUL8_FLAG_TIMERS   : the call synchronises on its last event and fills timers_ns (the reference's behaviour)
//   GEMMUL8_FLAG_PHASE_LOG: nothing synchronises; the events go to a per-thread log that
//                           gemmul8_b200_phase_log_collect() reads later (bench.py: phase times measured INSIDE the
//                           timed region, with no host wait between or after the calls)
struct PhaseLog {
    struct Entry { cudaEvent_t ev[5]; int n; };
    static constexpr size_t kCapacity = 4096;   // never collected: the most recent calls are kept (ring)
    std::vector<Entry> ring;                     // at most kCapacity entries; `head` is the oldest once it is full
    size_t head = 0;
    std::vector<cudaEvent_t> pool;   // recycled events of device `dev` (an event belongs to the device it was created on)
    int dev = -1;
    cudaEvent_t get() {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != dev) {
            for (auto e : pool) cudaEventDestroy(e);
            pool.clear();
            dev = cur;
        }
        cudaEvent_t e = nullptr;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
    void put(cudaEvent_t e) { if (e) pool.push_back(e); }
    void push(const Entry &e) {
        if (ring.size() < kCapacity) { ring.push_back(e); return; }
        for (auto old_ev : ring[head].ev) put(old_ev);
        ring[head] = e;
        head = (head + 1) % kCapacity;
    }
};
thread_local PhaseLog g_phase_log;

struct PhaseTimer {
    bool on, log; cudaStream_t st; cudaEvent_t ev[5] = {}; int n = 0; bool handed_over = false;
    PhaseTimer(unsigned flags, cudaStream_t s)
        : on((flags & (GEMMUL8_FLAG_TIMERS | GEMMUL8_FLAG_PHASE_LOG)) != 0), log((flags & GEMMUL8_FLAG_PHASE_LOG) != 0), st(s) {
        if (on) for (auto &e : ev) e = g_phase_log.get();
    }
    PhaseTimer(const PhaseTimer &) = delete;
    PhaseTimer &operator=(const PhaseTimer &) = delete;
    ~PhaseTimer() {   // every exit path, error returns included: the events go back to the pool unless the log took them
        if (on && !handed_over) for (auto e : ev) g_phase_log.put(e);
    }
    void mark() { if (on && n < 5) cudaEventRecord(ev[n++], st); }
    // marks: 0 start, 1 after scaling, 2 after gemm(+residues), 3 after crt
    static void spans(const cudaEvent_t *ev, int n, double *out_ns) {
        float ms;
        const int slot[3] = {0, 1, 3};
        for (int i = 0; i + 1 < n && i < 3; ++i) {
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) out_ns[slot[i]] += (double)ms * 1e6;
        }
    }
    void finish(double *out_ns) {
        if (!on || n == 0) return;
        if (log) {
            PhaseLog::Entry e;
            for (int i = 0; i < 5; ++i) e.ev[i] = ev[i];
            e.n = n;
            g_phase_log.push(e);
            handed_over = true;
            return;
        }
        cudaEventSynchronize(ev[n - 1]);
        spans(ev, n, out_ns);
    }
};

#define OZ_CUDA(call, where)                                   \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return fail_cuda(e__, where);  \
    } while (0)

// shifts + residue slices of one real operand.  `rows_are_vectors`: the vectors that get a common
// shift are rows of the stored column-major matrix (op_A == N for A, op_B != N for B).
int scale_operand(int dtype, bool strided, const void *X, size_t ld, size_t nvec, size_t len, int ref_width,
                  float log2M_fast, unsigned N, int8_t *slices, size_t ld8i, size_t inc, int16_t *sft, bool fast,
                  cudaStream_t st) {
    if (fast) OZ_CUDA(oz::launch_fast_shifts(dtype, strided, X, ld, nvec, len, ref_width, log2M_fast, sft, st), "fast shifts");
    OZ_CUDA(oz::launch_encode(dtype, strided, X, ld, nvec, len, sft, N, slices, ld8i, inc, st), "encode");
    return GEMMUL8_OK;
}

// Which product path a real-type block takes: the single kernel (product + residues + CRT, oz_gemm_crt.cu) when the caller
// forces it (GEMMUL8_FLAG_FUSED_CRT) or k is at most the option "fused_k"; product + residues to HBM and the stand-alone
// CRT kernel otherwise.  Both give the same bits of C.
bool take_fused(const gemmul8_b200_args *a) {
    if (a->flags & (GEMMUL8_FLAG_GEMM_SIMT | GEMMUL8_FLAG_STAGE_RESIDUES | GEMMUL8_FLAG_STAGE_SCALING)) return false;
    if (a->flags & GEMMUL8_FLAG_FUSED_CRT) return true;
    const int fk = oz::tuning().fused_k;
    return fk > 0 && a->k <= (size_t)fk;
}
// product of the block described by gp (A8i / B8i / rows / C8u set by the caller) into C: marks the timer between the
// product and the CRT when they are separate kernels
int block_to_c(const gemmul8_b200_args *a, oz::GemmProblem &gp, bool fused, bool split, void *C, const int16_t *sftA, const int16_t *sftB,
               cudaStream_t st, PhaseTimer *timer) {
    if (fused) {
        gp.C = C; gp.ldc = a->ldc; gp.dtype_C = a->dtype_C; gp.split_weights = split; gp.sftA = sftA; gp.sftB = sftB;
        gp.alpha_ptr = a->alpha; gp.beta_ptr = a->beta; gp.device_scalars = dev_scalars(a);
        OZ_CUDA(oz::launch_gemm_crt(gp, st), "int8 gemm + crt");
        if (timer) timer->mark();
        return GEMMUL8_OK;
    }
    auto gemm = (a->flags & GEMMUL8_FLAG_GEMM_SIMT) ? oz::launch_gemm_simt : oz::launch_gemm_tcgen05;
    OZ_CUDA(gemm(gp, oz::EPI_RESIDUE, st), "int8 gemm");
    if (timer) timer->mark();
    if (a->flags & GEMMUL8_FLAG_STAGE_RESIDUES) return GEMMUL8_OK;
    OZ_CUDA(oz::launch_crt(a->dtype_C, split, gp.num_slices, gp.rowsA, gp.rowsB, gp.C8u, gp.ldc8u, gp.sizeC, C, a->ldc, sftA, sftB, a->alpha, a->beta,
                           dev_scalars(a), st), "crt");
    if (timer) timer->mark();
    return GEMMUL8_OK;
}

// ---------------------------------------------------------------------------------------------
// Column-strip pipeline of the device-resident real fast-mode call.  The three phases of a strip
// use different units (encode: FP64 + LSU, product: tensor, CRT: FP64 + LSU), so on three streams
//     side stream B : shifts + residues of B columns [strip j+1]
//     caller stream : shifts + residues of all of A, then the all-moduli product of strip j
//     side stream C : CRT of strip j-1
// run concurrently: a pair CTA of the product owns its SM, so the overlap is between SMs -- where one strip's CTAs have
// finished, the next strip's encoders and the previous strip's CRT run until the next product's CTAs arrive -- and in the
// power budget (the HBM-bound phases no longer run with the tensor pipe idle).  Streams and events are created once per
// host thread and device.
// ---------------------------------------------------------------------------------------------
struct SideStreams {
    int device = -1;
    cudaStream_t sB = nullptr, sC = nullptr;
    cudaEvent_t start = nullptr, done = nullptr, evB[8] = {}, evG[8] = {};
    bool ok = false;
    bool init() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (ok && dev == device) return true;
        if (ok) return false;   // one device per host thread (a second one falls back to the serial path)
        device = dev;
        if (cudaStreamCreateWithFlags(&sB, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&sC, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&start, cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) return false;
        for (int i = 0; i < 8; ++i) {
            if (cudaEventCreateWithFlags(&evB[i], cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&evG[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        ok = true;
        return true;
    }
};
thread_local SideStreams g_side;

// Where the pipeline pays (measured on one B200, 14 moduli unless noted; profiles/r02_ab_strips_shapes.jsonl): 16384^3 +3.2 %
// (20 moduli +8.0 %, 8 moduli +1.3 %), 32768 x 16384 x 16384 +3.4 %, 24576^2 x 8192 +2.2 %, 12288^3 +1.3 %, 16384^2 x 8192 +0.9 %,
// 8192^2 x 32768 +0.4 %; 16384^2 x 4096 -2.7 %, 8192^3 -4.4 %.  What is hidden grows with k n + m n, what the co-running blocks
// cost with every strip launch: long k and a large C.
bool strips_by_size(size_t m, size_t n, size_t k, unsigned N) {
    (void)N;
    return k >= 8192 && m >= 8192 && n >= 8192 && m * n >= (size_t)12288 * 12288;
}

int gemm_real_strips(gemmul8_b200_args *a, const oz::Layout &L, int strips) {
    const size_t m = a->m, n = a->n, k = a->k;
    const unsigned N = a->num_moduli, ti = N - 2;
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    SideStreams &S = g_side;
    uint8_t *work = static_cast<uint8_t *>(a->work);
    int8_t *A8i   = reinterpret_cast<int8_t *>(work + L.off_A8i);
    int8_t *B8i   = reinterpret_cast<int8_t *>(work + L.off_B8i);
    uint8_t *C8u  = work + L.off_C8u;
    int16_t *sftA = reinterpret_cast<int16_t *>(work + L.off_sftA);
    int16_t *sftB = reinterpret_cast<int16_t *>(work + L.off_sftB);
    const bool a_strided = a->op_A == GEMMUL8_OP_N, b_strided = a->op_B != GEMMUL8_OP_N;
    const int ref_width  = oz::ref_reduce_width(a->dtype_A, a->dtype_B, a->dtype_C);
    const float l2       = oz::host_tab::OZ_LOG2M_FAST[ti];
    const bool split     = oz::host_tab::OZ_M_LO[ti] != 0.0 && a->dtype_C == GEMMUL8_F64;
    const size_t esB = elem_size(a->dtype_B), esC = elem_size(a->dtype_C);

    size_t cb[9];
    const size_t tiles = (n + 255) / 256;
    for (int i = 0; i <= strips; ++i) { const size_t x = (tiles * i / strips) * 256; cb[i] = x < n ? x : n; }
    cb[strips] = n;

    // phase marks on the caller's stream: start | first product starts | last product ends | end.  Slot 0 is then the EXPOSED
    // part of the scaling (all of A, and whatever of the first B strip is not ready by then), slot 1 the products back to back,
    // slot 3 the exposed CRT of the last strip; the rest of the scaling and of the CRT runs beside the products.
    PhaseTimer timer(a->flags, st);
    timer.mark();
    oz::g_strip_calls.fetch_add(1, std::memory_order_relaxed);
    OZ_CUDA(cudaEventRecord(S.start, st), "event record");
    OZ_CUDA(cudaStreamWaitEvent(S.sB, S.start, 0), "stream wait");
    for (int j = 0; j < strips; ++j) {
        const size_t c0 = cb[j], c1 = cb[j + 1];
        if (c1 > c0) {
            const uint8_t *Bx = static_cast<const uint8_t *>(a->B) + (b_strided ? c0 : c0 * a->ldb) * esB;
            int rc = scale_operand(a->dtype_B, b_strided, Bx, a->ldb, c1 - c0, k, ref_width, l2, N, B8i + c0 * L.lda8i, L.lda8i, L.sizeB,
                                   sftB + c0, true, S.sB);
            if (rc) return rc;
        }
        OZ_CUDA(cudaEventRecord(S.evB[j], S.sB), "event record");
    }
    int rc = scale_operand(a->dtype_A, a_strided, a->A, a->lda, m, k, ref_width, l2, N, A8i, L.lda8i, L.sizeA, sftA, true, st);
    if (rc) return rc;

    oz::GemmProblem gp{};
    gp.A8i = A8i; gp.rowsA = m; gp.ld8i = L.lda8i; gp.sizeA = L.sizeA; gp.sizeB = L.sizeB; gp.num_slices = N; gp.first_modulus = 0;
    gp.ldc8u = L.m_pad; gp.sizeC = L.sizeC; gp.claims = claims_of(work, L);
    // Full pipeline depth (6 stages): at 95 registers x 640 threads a pair CTA owns its SM anyway, the side streams' blocks run
    // on SMs between two strip launches' CTAs, not beside them; the 4-stage form only lost tensor-pipe time
    // (profiles/r02_ab_strips_stages.jsonl: 16384^3 47.3 -> 45.7 ms, 12288^3 20.6 -> 19.6 ms).
    for (int j = 0; j < strips; ++j) {
        const size_t c0 = cb[j], c1 = cb[j + 1];
        OZ_CUDA(cudaStreamWaitEvent(st, S.evB[j], 0), "stream wait");
        if (j == 0) timer.mark();
        if (c1 <= c0) continue;
        gp.B8i = B8i + c0 * L.lda8i; gp.rowsB = c1 - c0; gp.C8u = C8u + c0 * L.m_pad;
        OZ_CUDA(oz::launch_gemm_tcgen05(gp, oz::EPI_RESIDUE, st), "int8 gemm");
        OZ_CUDA(cudaEventRecord(S.evG[j], st), "event record");
        OZ_CUDA(cudaStreamWaitEvent(S.sC, S.evG[j], 0), "stream wait");
        OZ_CUDA(oz::launch_crt(a->dtype_C, split, N, m, c1 - c0, gp.C8u, L.m_pad, L.sizeC, static_cast<uint8_t *>(a->C) + c0 * a->ldc * esC,
                               a->ldc, sftA, sftB + c0, a->alpha, a->beta, dev_scalars(a), S.sC), "crt");
    }
    timer.mark();
    OZ_CUDA(cudaEventRecord(S.done, S.sC), "event record");
    OZ_CUDA(cudaStreamWaitEvent(st, S.done, 0), "stream wait");
    timer.mark();
    timer.finish(a->timers_ns);
    return GEMMUL8_OK;
}

int gemm_real(gemmul8_b200_args *a) {
    const size_t m = a->m, n = a->n, k = a->k;
    const unsigned N = a->num_moduli, ti = N - 2;
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    oz::Layout L;
    compute_layout(m, n, k, N, GEMMUL8_REAL_DEFAULT, L);
    uint8_t *work = static_cast<uint8_t *>(a->work);
    int8_t *A8i   = reinterpret_cast<int8_t *>(work + L.off_A8i);
    int8_t *B8i   = reinterpret_cast<int8_t *>(work + L.off_B8i);
    uint8_t *C8u  = work + L.off_C8u;
    int16_t *sftA = reinterpret_cast<int16_t *>(work + L.off_sftA);
    int16_t *sftB = reinterpret_cast<int16_t *>(work + L.off_sftB);

    const bool a_strided = a->op_A == GEMMUL8_OP_N;  // rows of column-major A
    const bool b_strided = a->op_B != GEMMUL8_OP_N;  // rows of column-major B
    const int ref_width  = oz::ref_reduce_width(a->dtype_A, a->dtype_B, a->dtype_C);
    const bool simt      = (a->flags & GEMMUL8_FLAG_GEMM_SIMT) != 0;
    auto gemm = simt ? oz::launch_gemm_simt : oz::launch_gemm_tcgen05;

    // The column-strip pipeline above, for large plain fast-mode calls.  With the round-1 kernel it was slower than the phases
    // in series (56.0 vs 51.2 ms at 16384^3); with the pair kernel it wins where
    // the hidden scaling + CRT outweigh what the co-running blocks cost the statically scheduled products
    // (profiles/r02_ab_strips.jsonl: 16384^3 48.97 -> 47.45 ms; 8192^3 5.42 -> 5.67 ms, i.e. not there).
    // Option "strips": 0 = by size (strips_by_size), 1 = never, 2 ... 8 = that many strips; GEMMUL8_FLAG_STRIPS forces it.
    // (GEMMUL8_FLAG_TIMERS does not exclude it: the synchronous drop-in call takes the pipeline too, and its four numbers are
    // then the exposed scaling, the products back to back, 0, and the exposed CRT)
    const unsigned serial_flags = GEMMUL8_FLAG_STAGE_SCALING | GEMMUL8_FLAG_STAGE_RESIDUES | GEMMUL8_FLAG_FUSED_CRT |
                                  GEMMUL8_FLAG_GEMM_SIMT | GEMMUL8_FLAG_SKIP_SCALE_A | GEMMUL8_FLAG_ONLY_SCALE_A | GEMMUL8_FLAG_ONLY_BOUND |
                                  GEMMUL8_FLAG_SKIP_BOUND;
    const int strips_opt = oz::tuning().strips;
    const bool strips_wanted = (a->flags & GEMMUL8_FLAG_STRIPS) || strips_opt >= 2 || (strips_opt == 0 && strips_by_size(m, n, k, N));
    if (strips_wanted && a->fastmode && !(a->flags & serial_flags) && !take_fused(a) && n >= 2048 && oz::tuning().gemm_pair != 0 &&
        g_side.init()) {
        const int strips = strips_opt >= 2 ? (strips_opt < (int)((n + 255) / 256) ? strips_opt : (int)((n + 255) / 256)) : (n >= 8192 ? 4 : 2);
        return gemm_real_strips(a, L, strips);
    }

    PhaseTimer timer(a->flags, st);
    timer.mark();

    // the pair GEMM's claim table is zeroed here, in front of the scaling kernels, so that no memset node sits between the
    // last encoder and the product (small problems are launch-latency-bound); fast mode only: the accurate-mode bound
    // product runs in between on the single-CTA kernel, which does not use the table, but keep that path plain
    uint32_t *claims = claims_of(work, L);
    const bool claims_early = claims != nullptr && a->fastmode && !(a->flags & GEMMUL8_FLAG_ONLY_SCALE_A);
    if (claims_early) OZ_CUDA(cudaMemsetAsync(claims, 0, oz::kClaimBytes, st), "memset claim table");

    // ---------------- phase 0: scaling ----------------
    if (a->fastmode) {
        const float l2 = oz::host_tab::OZ_LOG2M_FAST[ti];
        // (a distributed caller overlaps the arrival of the B panel with the scaling of A: two calls,
        //  GEMMUL8_FLAG_ONLY_SCALE_A then GEMMUL8_FLAG_SKIP_SCALE_A; see mixed-gemmul8_b200/distributed.py)
        int rc = GEMMUL8_OK;
        // Small operands: the four scaling launches (shift A, encode A, shift B, encode B) are latency-bound -- at 1024^3 they
        // are 39 of the call's 64 us -- and A's chain does not depend on B's.  B's chain goes to a side stream (forked and
        // joined with events, so the call stays stream-ordered and capturable) and the two run side by side.
        const bool fork = oz::tuning().scale_fork != 0 && !(a->flags & (GEMMUL8_FLAG_SKIP_SCALE_A | GEMMUL8_FLAG_ONLY_SCALE_A)) &&
                          (m * k + k * n) * 8 <= ((size_t)96 << 20) && g_side.init();
        if (fork) {
            SideStreams &S = g_side;
            OZ_CUDA(cudaEventRecord(S.start, st), "event record");
            OZ_CUDA(cudaStreamWaitEvent(S.sB, S.start, 0), "stream wait");
            rc = scale_operand(a->dtype_B, b_strided, a->B, a->ldb, n, k, ref_width, l2, N, B8i, L.lda8i, L.sizeB, sftB, true, S.sB);
            if (rc) return rc;
            OZ_CUDA(cudaEventRecord(S.done, S.sB), "event record");
            rc = scale_operand(a->dtype_A, a_strided, a->A, a->lda, m, k, ref_width, l2, N, A8i, L.lda8i, L.sizeA, sftA, true, st);
            if (rc) return rc;
            OZ_CUDA(cudaStreamWaitEvent(st, S.done, 0), "stream wait");
        } else {
            if (!(a->flags & GEMMUL8_FLAG_SKIP_SCALE_A))
                rc = scale_operand(a->dtype_A, a_strided, a->A, a->lda, m, k, ref_width, l2, N, A8i, L.lda8i, L.sizeA, sftA, true, st);
            if (rc) return rc;
            if (a->flags & GEMMUL8_FLAG_ONLY_SCALE_A) { timer.mark(); timer.finish(a->timers_ns); return GEMMUL8_OK; }
            rc = scale_operand(a->dtype_B, b_strided, a->B, a->ldb, n, k, ref_width, l2, N, B8i, L.lda8i, L.sizeB, sftB, true, st);
            if (rc) return rc;
        }
    } else {
        if (a->flags & GEMMUL8_FLAG_ONLY_SCALE_A) return GEMMUL8_OK;   // accurate mode needs both operands: all work in the second call
        // reference: int8tc::scaling, GEMMul8/src/scaling.hpp:3053-3136
        // The bound product only needs its row / column maxima.  They live in the (still unused)
        // second modulus slice of A8i / B8i: 4*m_pad <= lda8i*m_pad and 4*n <= lda8i*n always hold.
        int32_t *rowmax = reinterpret_cast<int32_t *>(A8i + L.sizeA);
        int32_t *colmax = reinterpret_cast<int32_t *>(B8i + L.sizeB);
        // A caller that holds one block of a partitioned C (distributed.py) splits the call here: ..._ONLY_BOUND leaves the
        // maxima of ITS block's bound product in `work`, the caller takes the maximum over the blocks that share the rows /
        // the columns (one tiny all-reduce each), and ..._SKIP_BOUND continues with those: the shifts, and with them every
        // bit of C, are then the ones of the unpartitioned call.
        if (!(a->flags & GEMMUL8_FLAG_SKIP_BOUND)) {
            OZ_CUDA(oz::launch_bound_extract(a->dtype_A, a_strided, a->A, a->lda, m, k, A8i, L.lda8i, sftA, st), "bound extract A");
            OZ_CUDA(oz::launch_bound_extract(a->dtype_B, b_strided, a->B, a->ldb, n, k, B8i, L.lda8i, sftB, st), "bound extract B");
            OZ_CUDA(cudaMemsetAsync(rowmax, 0, sizeof(int32_t) * m, st), "memset row maxima");
            OZ_CUDA(cudaMemsetAsync(colmax, 0, sizeof(int32_t) * n, st), "memset col maxima");
            oz::GemmProblem bp{};
            bp.A8i = A8i; bp.B8i = B8i; bp.rowsA = m; bp.rowsB = n; bp.ld8i = L.lda8i; bp.sizeA = L.sizeA; bp.sizeB = L.sizeB;
            bp.num_slices = 1; bp.first_modulus = 0; bp.rowmax = rowmax; bp.colmax = colmax;
            OZ_CUDA(gemm(bp, oz::EPI_ABSMAX, st), "bound product");
        }
        if (a->flags & GEMMUL8_FLAG_ONLY_BOUND) { timer.mark(); timer.finish(a->timers_ns); return GEMMUL8_OK; }
        const float l2 = oz::host_tab::OZ_LOG2M_ACC[ti];
        OZ_CUDA(oz::launch_accurate_shifts(m, rowmax, l2, sftA, st), "accurate shifts A");
        OZ_CUDA(oz::launch_accurate_shifts(n, colmax, l2, sftB, st), "accurate shifts B");
        int rc = scale_operand(a->dtype_A, a_strided, a->A, a->lda, m, k, ref_width, 0.f, N, A8i, L.lda8i, L.sizeA, sftA, false, st);
        if (rc) return rc;
        rc = scale_operand(a->dtype_B, b_strided, a->B, a->ldb, n, k, ref_width, 0.f, N, B8i, L.lda8i, L.sizeB, sftB, false, st);
        if (rc) return rc;
    }
    timer.mark();
    if (a->flags & GEMMUL8_FLAG_STAGE_SCALING) { timer.finish(a->timers_ns); return GEMMUL8_OK; }

    // ---------------- phases 1+2 (+3): all-moduli int8 GEMM with the residue reduction in its epilogue ----------------
    const bool split = oz::host_tab::OZ_M_LO[ti] != 0.0 && a->dtype_C == GEMMUL8_F64;  // numM == 2 (N >= 8) and fp64 out
    oz::GemmProblem gp{};
    gp.A8i = A8i; gp.B8i = B8i; gp.rowsA = m; gp.rowsB = n; gp.ld8i = L.lda8i; gp.sizeA = L.sizeA; gp.sizeB = L.sizeB;
    gp.num_slices = N; gp.first_modulus = 0; gp.C8u = C8u; gp.ldc8u = L.m_pad; gp.sizeC = L.sizeC; gp.claims = claims;
    gp.claims_zeroed = claims_early;
    int rc = block_to_c(a, gp, take_fused(a), split, a->C, sftA, sftB, st, &timer);
    if (rc) return rc;
    timer.finish(a->timers_ns);
    return GEMMUL8_OK;
}

// ---------------------------------------------------------------------------------------------
// complex types: reference gemm_mixed_bigmatrix / gemm_mixed_kara / gemm_mixed_classic,
// GEMMul8/src/gemmul8.cu:579-1052 (dispatch :1146-1178)
// ---------------------------------------------------------------------------------------------
int gemm_complex(gemmul8_b200_args *a) {
    const size_t m = a->m, n = a->n, k = a->k;
    const unsigned N = a->num_moduli, ti = N - 2;
    const int ct = a->compute_type;
    const bool big = ct == GEMMUL8_COMPLEX_BIG_MATRIX_ENCODE;
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    oz::Layout L;
    compute_layout(m, n, k, N, ct, L);
    uint8_t *work    = static_cast<uint8_t *>(a->work);
    int8_t *A_re     = reinterpret_cast<int8_t *>(work + L.off_A8i);
    int8_t *A_im     = reinterpret_cast<int8_t *>(work + L.off_A8i_imag);
    int8_t *B_re     = reinterpret_cast<int8_t *>(work + L.off_B8i);
    int8_t *B_im     = reinterpret_cast<int8_t *>(work + L.off_B8i_imag);
    uint8_t *C_re    = work + L.off_C8u;
    uint8_t *C_im    = big ? C_re + m : work + L.off_C8u_imag;   // big matrix: Im rows follow the m Re rows
    int16_t *sftA    = reinterpret_cast<int16_t *>(work + L.off_sftA);
    int16_t *sftB    = reinterpret_cast<int16_t *>(work + L.off_sftB);

    const bool a_strided = a->op_A == GEMMUL8_OP_N, b_strided = a->op_B != GEMMUL8_OP_N;
    const bool a_conj = a->op_A == GEMMUL8_OP_C, b_conj = a->op_B == GEMMUL8_OP_C;
    const bool simt = (a->flags & GEMMUL8_FLAG_GEMM_SIMT) != 0;
    auto gemm = simt ? oz::launch_gemm_simt : oz::launch_gemm_tcgen05;

    oz::ComplexTarget tA{}, tB{};
    tA.layout = big ? 1 : 0; tA.out_re = A_re; tA.out_im = big ? nullptr : A_im; tA.ld8i = L.lda8i; tA.inc = L.sizeA; tA.k = k; tA.nvec = m;
    tB.layout = big ? 2 : 0; tB.out_re = B_re; tB.out_im = big ? nullptr : B_im; tB.ld8i = L.lda8i; tB.inc = L.sizeB; tB.k = k; tB.nvec = n;
    const size_t rowsA = big ? 2 * m : m;   // rows of the int8 A operand

    PhaseTimer timer(a->flags, st);
    timer.mark();

    // ---------------- phase 0: scaling ----------------
    if (a->fastmode) {
        const float l2 = oz::host_tab::OZ_LOG2M_FAST[ti];
        OZ_CUDA(oz::launch_fast_shifts(a->dtype_A, a_strided, a->A, a->lda, m, k, 128, l2, sftA, st), "fast shifts A");
        OZ_CUDA(oz::launch_fast_shifts(a->dtype_B, b_strided, a->B, a->ldb, n, k, 128, l2, sftB, st), "fast shifts B");
    } else {
        // reference: int8tc::scaling_bigmatrix / scaling_kara, GEMMul8/src/scaling.hpp:3138-3367.
        // Whatever the compute type, the bound product is Re = |Pr||Qr| -+ |Pi||Qi|, Im = |Pi||Qr| +- |Pr||Qi|;
        // in the big-matrix layout that is ONE int8 product, and only its row / column maxima are needed.
        // CLASSIC / KARATSUBA borrow the (still empty) A / B slice stacks for the temporary big bound
        // matrices: ceil16(2k) * ceil4(2m) <= 4 sizeA <= 2 N sizeA.
        const size_t ld_big = ceil_to(2 * k, 16);
        oz::ComplexTarget bA = tA, bB = tB;
        bA.layout = 1; bA.out_im = nullptr; bA.ld8i = ld_big;
        bB.layout = 2; bB.out_im = nullptr; bB.ld8i = ld_big;
        OZ_CUDA(oz::launch_bound_extract_complex(a->dtype_A, a_strided, a->A, a->lda, m, k, sftA, bA, a_conj, st), "bound extract A");
        OZ_CUDA(oz::launch_bound_extract_complex(a->dtype_B, b_strided, a->B, a->ldb, n, k, sftB, bB, b_conj, st), "bound extract B");
        // maxima: 2m + n int32 at the start of the residue area ((N+4) sizeC >= 12 m n bytes: always enough)
        int32_t *rowmax = reinterpret_cast<int32_t *>(C_re);
        int32_t *colmax = rowmax + 2 * m;
        OZ_CUDA(cudaMemsetAsync(rowmax, 0, sizeof(int32_t) * (2 * m + n), st), "memset maxima");
        oz::GemmProblem bp{};
        bp.A8i = A_re; bp.B8i = B_re; bp.rowsA = 2 * m; bp.rowsB = n; bp.ld8i = ld_big;
        bp.sizeA = ld_big * ceil_to(2 * m, 4); bp.sizeB = ld_big * n; bp.num_slices = 1; bp.first_modulus = 0;
        bp.rowmax = rowmax; bp.colmax = colmax;
        OZ_CUDA(gemm(bp, oz::EPI_ABSMAX, st), "bound product");
        const float l2 = oz::host_tab::OZ_LOG2M_ACC[ti];
        OZ_CUDA(oz::launch_accurate_shifts(m, rowmax, l2, sftA, st, m), "accurate shifts A");   // Re row r and Im row r + m
        OZ_CUDA(oz::launch_accurate_shifts(n, colmax, l2, sftB, st), "accurate shifts B");
    }
    OZ_CUDA(oz::launch_encode_complex(a->dtype_A, a_strided, a->A, a->lda, m, k, sftA, N, tA, a_conj, st), "encode A");
    OZ_CUDA(oz::launch_encode_complex(a->dtype_B, b_strided, a->B, a->ldb, n, k, sftB, N, tB, b_conj, st), "encode B");
    timer.mark();
    if (a->flags & GEMMUL8_FLAG_STAGE_SCALING) { timer.finish(a->timers_ns); return GEMMUL8_OK; }

    // ---------------- phases 1+2: products mod m_j ----------------
    oz::GemmProblem gp{};
    gp.rowsA = rowsA; gp.rowsB = n; gp.ld8i = L.lda8i; gp.sizeA = L.sizeA; gp.sizeB = L.sizeB;
    gp.num_slices = N; gp.first_modulus = 0; gp.ldc8u = L.m_pad; gp.sizeC = L.sizeC; gp.claims = claims_of(work, L);
    auto product = [&](const int8_t *A8, const int8_t *B8, uint8_t *C8, int combine, uint8_t *aux) -> cudaError_t {
        gp.A8i = A8; gp.B8i = B8; gp.C8u = C8; gp.combine = combine; gp.C8u_aux = aux;
        return gemm(gp, oz::EPI_RESIDUE, st);
    };
    if (big) {
        OZ_CUDA(product(A_re, B_re, C_re, oz::RC_STORE, nullptr), "big-matrix product");
    } else if (ct == GEMMUL8_COMPLEX_CLASSIC_MULT) {
        OZ_CUDA(product(A_re, B_re, C_re, oz::RC_STORE, nullptr), "Ar Br");
        OZ_CUDA(product(A_im, B_im, C_re, oz::RC_SUB, nullptr), "Ai Bi");
        OZ_CUDA(product(A_im, B_re, C_im, oz::RC_STORE, nullptr), "Ai Br");
        OZ_CUDA(product(A_re, B_im, C_im, oz::RC_ADD, nullptr), "Ar Bi");
    } else {  // Karatsuba: E = ArBr, F = AiBi, G = (Ar+Ai)(Br+Bi); Re = E - F, Im = G - (E + F)
        OZ_CUDA(product(A_re, B_re, C_re, oz::RC_STORE, nullptr), "E");
        // The in-place slice sums Ar <- Ar + Ai, Br <- Br + Bi (as the reference, gemmul8.cu:853-855) only need E to be done: Ar
        // and Br are not read by F, Ai and Bi are only READ by both.  They are HBM-bound (5.6 GB at 8192^3), F is tensor-bound:
        // on a side stream they run beside F (which then leaves room on the SMs: 4 pipeline stages) instead of after it.
        const bool overlap = oz::tuning().scale_fork != 0 && g_side.init();
        if (overlap) {
            SideStreams &S = g_side;
            OZ_CUDA(cudaEventRecord(S.start, st), "event record");
            OZ_CUDA(cudaStreamWaitEvent(S.sB, S.start, 0), "stream wait");
            OZ_CUDA(oz::launch_add_int8_slices(N, L.sizeA, A_re, A_im, S.sB), "Ar + Ai");
            OZ_CUDA(oz::launch_add_int8_slices(N, L.sizeB, B_re, B_im, S.sB), "Br + Bi");
            OZ_CUDA(cudaEventRecord(S.done, S.sB), "event record");
            gp.share_sm = true;
            OZ_CUDA(product(A_im, B_im, C_re, oz::RC_KARATSUBA_F, C_im), "F");
            gp.share_sm = false;
            OZ_CUDA(cudaStreamWaitEvent(st, S.done, 0), "stream wait");
        } else {
            OZ_CUDA(product(A_im, B_im, C_re, oz::RC_KARATSUBA_F, C_im), "F");
            OZ_CUDA(oz::launch_add_int8_slices(N, L.sizeA, A_re, A_im, st), "Ar + Ai");
            OZ_CUDA(oz::launch_add_int8_slices(N, L.sizeB, B_re, B_im, st), "Br + Bi");
        }
        OZ_CUDA(product(A_re, B_re, C_im, oz::RC_RSUB, nullptr), "G");
    }
    timer.mark();
    if (a->flags & GEMMUL8_FLAG_STAGE_RESIDUES) { timer.finish(a->timers_ns); return GEMMUL8_OK; }

    // ---------------- phase 3: CRT + inverse scaling ----------------
    const bool split = oz::host_tab::OZ_M_LO[ti] != 0.0 && a->dtype_C == GEMMUL8_C64;
    OZ_CUDA(oz::launch_crt_complex(a->dtype_C, split, N, m, n, C_re, C_im, L.m_pad, L.sizeC, a->C, a->ldc, sftA, sftB, a->alpha, a->beta, dev_scalars(a), st), "crt");
    timer.mark();
    timer.finish(a->timers_ns);
    return GEMMUL8_OK;
}

}  // namespace

extern "C" {

size_t gemmul8_b200_worksize(size_t m, size_t n, size_t k, unsigned num_moduli, int compute_type) {
    oz::Layout L;
    if (!compute_layout(m, n, k, num_moduli, compute_type, L)) {
        fprintf(stderr, "Unknown compute type\n");  // reference: gemmul8.cu:142-145
        return 0;
    }
    return L.total;
}

int gemmul8_b200_work_layout(size_t m, size_t n, size_t k, unsigned num_moduli, int compute_type, gemmul8_b200_layout *out) {
    oz::Layout L;
    if (!out) return fail(GEMMUL8_ERR_ARGUMENT, "null layout");
    if (!compute_layout(m, n, k, num_moduli, compute_type, L)) return fail(GEMMUL8_ERR_COMPUTETYPE, "unknown compute type");
    static_assert(sizeof(gemmul8_b200_layout) == sizeof(oz::Layout), "layout structs must match");
    memcpy(out, &L, sizeof(L));
    return GEMMUL8_OK;
}

int gemmul8_b200_gemm(gemmul8_b200_args *a) {
    if (a) for (double &t : a->timers_ns) t = 0.0;
    int rc = check_args(a);
    if (rc) return rc;
    if (a->m == 0 || a->n == 0) return GEMMUL8_OK;
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
        return fail(GEMMUL8_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (a->k == 0) {   // empty product: C = beta * C (the reference launches its kernels on empty operands: undefined)
        OZ_CUDA(oz::launch_scale_c(a->dtype_C, a->m, a->n, a->C, a->ldc, a->beta, dev_scalars(a), static_cast<cudaStream_t>(a->stream)), "scale C");
        return GEMMUL8_OK;
    }
    if (is_complex(a->dtype_C)) {
        if (a->k > (size_t(1) << 16)) return fail(GEMMUL8_ERR_ARGUMENT, "complex types: k must be <= 2^16 (int32 accumulation)");
        return gemm_complex(a);
    }
    return gemm_real(a);
}

size_t gemmul8_b200_host_scratch_size(const gemmul8_b200_args *a) {
    if (!a) return 0;
    const size_t colsA = a->op_A == GEMMUL8_OP_N ? a->k : a->m, colsB = a->op_B == GEMMUL8_OP_N ? a->n : a->k;
    size_t s = ceil_to(a->lda * colsA * elem_size(a->dtype_A), 256) + ceil_to(a->ldb * colsB * elem_size(a->dtype_B), 256) +
               ceil_to(a->ldc * a->n * elem_size(a->dtype_C), 256);
    return s + gemmul8_b200_worksize(a->m, a->n, a->k, a->num_moduli, a->compute_type);
}

// Block-wise entry for callers that receive their operands piecewise (multi-GPU panel exchange, host
// streaming): the three steps of the real fast-mode path on a sub-range of the FULL problem's workspace.
//   GEMMUL8_PART_SCALE_A : shifts + residue slices of rows [row0, row1) of op(A)
//   GEMMUL8_PART_SCALE_B : shifts + residue slices of columns [col0, col1) of op(B)
//   GEMMUL8_PART_PRODUCT : all-moduli product, residues, CRT and alpha/beta for C[row0:row1, col0:col1]
// `args` describes the full problem (m, n, k, leading dimensions, pointers to the FULL matrices, work);
// row0 and col0 must be multiples of 256 (whole GEMM tiles; the last block may be ragged).
int gemmul8_b200_gemm_part(gemmul8_b200_args *a, int parts, size_t row0, size_t row1, size_t col0, size_t col1) {
    int rc = check_args(a);
    if (rc) return rc;
    if (is_complex(a->dtype_C) || !a->fastmode) return fail(GEMMUL8_ERR_ARGUMENT, "gemm_part: real types in fast mode only");
    // whole GEMM tiles for a product; the scaling steps work per row / column and take any range
    const bool tile_aligned = !(parts & GEMMUL8_PART_PRODUCT) || ((row0 % 256) == 0 && (col0 % 256) == 0);
    if (row1 > a->m || col1 > a->n || row0 > row1 || col0 > col1 || !tile_aligned)
        return fail(GEMMUL8_ERR_ARGUMENT, "gemm_part: bad block (row0 / col0 of a product must be multiples of 256 inside the matrix)");
    const size_t m = a->m, n = a->n, k = a->k;
    const unsigned N = a->num_moduli, ti = N - 2;
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    oz::Layout L;
    compute_layout(m, n, k, N, GEMMUL8_REAL_DEFAULT, L);
    uint8_t *work = static_cast<uint8_t *>(a->work);
    int8_t *A8i   = reinterpret_cast<int8_t *>(work + L.off_A8i);
    int8_t *B8i   = reinterpret_cast<int8_t *>(work + L.off_B8i);
    uint8_t *C8u  = work + L.off_C8u;
    int16_t *sftA = reinterpret_cast<int16_t *>(work + L.off_sftA);
    int16_t *sftB = reinterpret_cast<int16_t *>(work + L.off_sftB);
    const bool a_strided = a->op_A == GEMMUL8_OP_N, b_strided = a->op_B != GEMMUL8_OP_N;
    const int ref_width  = oz::ref_reduce_width(a->dtype_A, a->dtype_B, a->dtype_C);
    const float l2       = oz::host_tab::OZ_LOG2M_FAST[ti];
    const size_t esA = elem_size(a->dtype_A), esB = elem_size(a->dtype_B), esC = elem_size(a->dtype_C);
    for (double &t : a->timers_ns) t = 0.0;
    PhaseTimer timer(a->flags, st);   // GEMMUL8_FLAG_TIMERS / _PHASE_LOG: the same four slots as gemm()
    timer.mark();
    if ((parts & GEMMUL8_PART_SCALE_A) && row1 > row0 && k > 0) {
        const uint8_t *Ax = static_cast<const uint8_t *>(a->A) + (a_strided ? row0 : row0 * a->lda) * esA;
        rc = scale_operand(a->dtype_A, a_strided, Ax, a->lda, row1 - row0, k, ref_width, l2, N, A8i + row0 * L.lda8i, L.lda8i, L.sizeA,
                           sftA + row0, true, st);
        if (rc) return rc;
    }
    if ((parts & GEMMUL8_PART_SCALE_B) && col1 > col0 && k > 0) {
        const uint8_t *Bx = static_cast<const uint8_t *>(a->B) + (b_strided ? col0 : col0 * a->ldb) * esB;
        rc = scale_operand(a->dtype_B, b_strided, Bx, a->ldb, col1 - col0, k, ref_width, l2, N, B8i + col0 * L.lda8i, L.lda8i, L.sizeB,
                           sftB + col0, true, st);
        if (rc) return rc;
    }
    if ((parts & GEMMUL8_PART_PRODUCT) && row1 > row0 && col1 > col0 && k == 0) {
        OZ_CUDA(oz::launch_scale_c(a->dtype_C, row1 - row0, col1 - col0, static_cast<uint8_t *>(a->C) + (col0 * a->ldc + row0) * esC, a->ldc,
                                   a->beta, dev_scalars(a), st), "scale C");
    } else if ((parts & GEMMUL8_PART_PRODUCT) && row1 > row0 && col1 > col0) {
        const bool split = oz::host_tab::OZ_M_LO[ti] != 0.0 && a->dtype_C == GEMMUL8_F64;
        oz::GemmProblem gp{};
        gp.ld8i = L.lda8i; gp.sizeA = L.sizeA; gp.sizeB = L.sizeB; gp.num_slices = N; gp.first_modulus = 0;
        gp.ldc8u = L.m_pad; gp.sizeC = L.sizeC; gp.claims = claims_of(work, L);
        gp.A8i = A8i + row0 * L.lda8i; gp.rowsA = row1 - row0;
        gp.B8i = B8i + col0 * L.lda8i; gp.rowsB = col1 - col0;
        gp.C8u = C8u + col0 * L.m_pad + row0;
        gp.share_sm = (a->flags & GEMMUL8_FLAG_EXCLUSIVE_SMS) == 0;   // a block-wise caller usually overlaps this product with NCCL transfers
        timer.mark();
        rc = block_to_c(a, gp, take_fused(a), split, static_cast<uint8_t *>(a->C) + (col0 * a->ldc + row0) * esC, sftA + row0, sftB + col0, st, &timer);
        if (rc) return rc;
        timer.finish(a->timers_ns);
        return GEMMUL8_OK;
    }
    timer.mark();
    timer.finish(a->timers_ns);
    return GEMMUL8_OK;
}

}  // extern "C"

namespace {

// ---------------------------------------------------------------------------------------------
// Low-memory call (the reference's README.md:3 points at a `memory-lt` branch that is not
// in the tree).  gemm() keeps the residue slices of ALL of op(A) and op(B): N (m + n) k bytes, 184 GiB at
// 65536^3.  Here C is produced in blocks of block_rows x block_cols and only the slices of one row block
// and one column block (plus the residues of one C block) are resident:
//     work = N k (block_rows + block_cols) + N block_rows block_cols + 6 (m + n) bytes (+ padding).
// Shifts are per row of op(A) / column of op(B), residues and CRT per element, and the accurate-mode bound
// only needs row / column maxima (atomicMax over blocks), so the result bits are those of gemm().  The
// operand of the inner loop is re-encoded once per outer block unless all of it fits in one block.
// ---------------------------------------------------------------------------------------------
struct BlockPlan {
    size_t mb, nb;          // block_rows, block_cols (multiples of 256, or the whole dimension)
    bool outer_rows;        // true: for row blocks { for column blocks }, else the other way round
    size_t lda8i, mb_pad, sizeA, sizeB, sizeC;
    size_t off_A8i, off_B8i, off_C8u, off_sftA, off_sftB, off_rowmax, off_colmax, off_claims, total;
};

bool carve_blocks(size_t m, size_t n, size_t k, unsigned N, size_t mb, size_t nb, BlockPlan &P) {
    if (mb == 0 || nb == 0) return false;
    P.mb = mb < m ? mb : m; P.nb = nb < n ? nb : n;
    if ((P.mb < m && P.mb % 256) || (P.nb < n && P.nb % 256)) return false;
    P.lda8i = ceil_to(k, 16);
    P.mb_pad = ceil_to(P.mb, 4);
    P.sizeA = P.lda8i * P.mb_pad;
    P.sizeB = P.lda8i * P.nb;
    P.sizeC = ceil_to(P.mb_pad * P.nb, 16);
    size_t off = 0;
    P.off_A8i = off;    off += P.sizeA * N;
    P.off_B8i = off;    off += P.sizeB * N;
    P.off_C8u = off;    off += P.sizeC * N;
    P.off_sftA = off;   off += 2 * ceil_to(m, 16);
    P.off_sftB = off;   off += 2 * ceil_to(n, 16);
    P.off_rowmax = off; off += 4 * ceil_to(m, 16);
    P.off_colmax = off; off += 4 * ceil_to(n, 16);
    P.off_claims = off; off += oz::kClaimBytes;
    P.total = off;
    const size_t Rb = (m + P.mb - 1) / P.mb, Cb = (n + P.nb - 1) / P.nb;
    // elements encoded more than once: inner operand, once per outer block (unless it is a single block)
    const size_t cost_rows_outer = Cb > 1 ? (Rb - 1) * n : 0, cost_cols_outer = Rb > 1 ? (Cb - 1) * m : 0;
    P.outer_rows = cost_rows_outer <= cost_cols_outer;
    return true;
}

// largest blocks that fit `max_bytes`: least re-encoding first, then the squarest block
bool plan_blocks(size_t m, size_t n, size_t k, unsigned N, size_t max_bytes, BlockPlan &best) {
    bool found = false;
    double best_cost = 0;
    const size_t tm = (m + 255) / 256, tn = (n + 255) / 256;
    for (size_t Rb = 1; Rb <= tm; ++Rb) {
        const size_t mb = Rb == 1 ? m : ((tm + Rb - 1) / Rb) * 256;
        if (Rb > 1 && (m + mb - 1) / mb != Rb) continue;   // same block size as a smaller Rb
        // widest nb for this mb: total is monotone in nb
        size_t lo = 1, hi = tn, fit = 0;
        while (lo <= hi) {
            const size_t mid = (lo + hi) / 2;
            BlockPlan P;
            const size_t nb = mid == tn ? n : mid * 256;
            if (carve_blocks(m, n, k, N, mb, nb, P) && P.total <= max_bytes) { fit = mid; lo = mid + 1; }
            else hi = mid - 1;
        }
        if (!fit) continue;
        // equalise the column blocks: same count, smallest width
        const size_t Cb = (tn + fit - 1) / fit;
        const size_t nbt = (tn + Cb - 1) / Cb;
        BlockPlan P;
        carve_blocks(m, n, k, N, mb, nbt == tn ? n : nbt * 256, P);
        const double re = P.outer_rows ? (Cb > 1 ? (double)(Rb - 1) * n : 0.0) : (Rb > 1 ? (double)(Cb - 1) * m : 0.0);
        // re-encoded elements (x k), plus a small charge per block for the tails of the persistent GEMM
        const double cost = re * (double)k + 1e-3 * (double)(Rb * Cb) * (double)k * (double)(m + n);
        if (!found || cost < best_cost) { best = P; best_cost = cost; found = true; }
    }
    return found;
}

int gemm_blocked_real(gemmul8_b200_args *a, const BlockPlan &P) {
    const size_t m = a->m, n = a->n, k = a->k;
    const unsigned N = a->num_moduli, ti = N - 2;
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    uint8_t *work = static_cast<uint8_t *>(a->work);
    int8_t *A8i     = reinterpret_cast<int8_t *>(work + P.off_A8i);
    int8_t *B8i     = reinterpret_cast<int8_t *>(work + P.off_B8i);
    uint8_t *C8u    = work + P.off_C8u;
    int16_t *sftA   = reinterpret_cast<int16_t *>(work + P.off_sftA);
    int16_t *sftB   = reinterpret_cast<int16_t *>(work + P.off_sftB);
    int32_t *rowmax = reinterpret_cast<int32_t *>(work + P.off_rowmax);
    int32_t *colmax = reinterpret_cast<int32_t *>(work + P.off_colmax);
    const bool a_strided = a->op_A == GEMMUL8_OP_N, b_strided = a->op_B != GEMMUL8_OP_N;
    const int ref_width  = oz::ref_reduce_width(a->dtype_A, a->dtype_B, a->dtype_C);
    const size_t esA = elem_size(a->dtype_A), esB = elem_size(a->dtype_B), esC = elem_size(a->dtype_C);
    const bool fast = a->fastmode != 0;
    const float l2  = fast ? oz::host_tab::OZ_LOG2M_FAST[ti] : 0.f;
    const bool split = oz::host_tab::OZ_M_LO[ti] != 0.0 && a->dtype_C == GEMMUL8_F64;
    const size_t Rb = (m + P.mb - 1) / P.mb, Cb = (n + P.nb - 1) / P.nb;
    auto rows_of = [&](size_t i, size_t &r0, size_t &r1) { r0 = i * P.mb; r1 = r0 + P.mb < m ? r0 + P.mb : m; };
    auto cols_of = [&](size_t j, size_t &c0, size_t &c1) { c0 = j * P.nb; c1 = c0 + P.nb < n ? c0 + P.nb : n; };
    auto Aptr = [&](size_t r0) { return static_cast<const uint8_t *>(a->A) + (a_strided ? r0 : r0 * a->lda) * esA; };
    auto Bptr = [&](size_t c0) { return static_cast<const uint8_t *>(a->B) + (b_strided ? c0 : c0 * a->ldb) * esB; };

    // phase times over all blocks (GEMMUL8_FLAG_TIMERS): an event at every phase boundary, summed per phase at the end
    struct Marks {
        bool on; cudaStream_t st; std::vector<std::pair<cudaEvent_t, int>> ev;   // (event, phase that ENDS here; -1 = start)
        void mark(int phase) {
            if (!on) return;
            cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.emplace_back(e, phase);
        }
        void finish(double *out_ns) {
            if (!on || ev.empty()) return;
            cudaEventSynchronize(ev.back().first);
            for (size_t i = 1; i < ev.size(); ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i - 1].first, ev[i].first);
                if (ev[i].second >= 0) out_ns[ev[i].second] += (double)ms * 1e6;
            }
            for (auto &e : ev) cudaEventDestroy(e.first);
        }
    } timer{(a->flags & GEMMUL8_FLAG_TIMERS) != 0, st, {}};
    timer.mark(-1);
    if (!fast && k > 0) {
        // accurate mode: row / column maxima of the bound product over all blocks, then all the shifts
        // (reference: int8tc::scaling, GEMMul8/src/scaling.hpp:3053-3136)
        OZ_CUDA(cudaMemsetAsync(rowmax, 0, sizeof(int32_t) * m, st), "memset row maxima");
        OZ_CUDA(cudaMemsetAsync(colmax, 0, sizeof(int32_t) * n, st), "memset col maxima");
        oz::GemmProblem bp{};
        bp.ld8i = P.lda8i; bp.sizeA = P.sizeA; bp.sizeB = P.sizeB; bp.num_slices = 1; bp.first_modulus = 0;
        for (size_t o = 0; o < (P.outer_rows ? Rb : Cb); ++o) {
            for (size_t q = 0; q < (P.outer_rows ? Cb : Rb); ++q) {
                const size_t i = P.outer_rows ? o : q, j = P.outer_rows ? q : o;
                size_t r0, r1, c0, c1;
                rows_of(i, r0, r1); cols_of(j, c0, c1);
                const bool newA = P.outer_rows ? q == 0 : (Rb > 1 || o == 0);
                const bool newB = P.outer_rows ? (Cb > 1 || o == 0) : q == 0;
                if (newA) OZ_CUDA(oz::launch_bound_extract(a->dtype_A, a_strided, Aptr(r0), a->lda, r1 - r0, k, A8i, P.lda8i, sftA + r0, st), "bound extract A");
                if (newB) OZ_CUDA(oz::launch_bound_extract(a->dtype_B, b_strided, Bptr(c0), a->ldb, c1 - c0, k, B8i, P.lda8i, sftB + c0, st), "bound extract B");
                bp.A8i = A8i; bp.B8i = B8i; bp.rowsA = r1 - r0; bp.rowsB = c1 - c0; bp.rowmax = rowmax + r0; bp.colmax = colmax + c0;
                OZ_CUDA(oz::launch_gemm_tcgen05(bp, oz::EPI_ABSMAX, st), "bound product");
            }
        }
        const float l2a = oz::host_tab::OZ_LOG2M_ACC[ti];
        OZ_CUDA(oz::launch_accurate_shifts(m, rowmax, l2a, sftA, st), "accurate shifts A");
        OZ_CUDA(oz::launch_accurate_shifts(n, colmax, l2a, sftB, st), "accurate shifts B");
        timer.mark(0);
    }

    oz::GemmProblem gp{};
    gp.ld8i = P.lda8i; gp.sizeA = P.sizeA; gp.sizeB = P.sizeB; gp.num_slices = N; gp.first_modulus = 0;
    gp.ldc8u = P.mb_pad; gp.sizeC = P.sizeC; gp.A8i = A8i; gp.B8i = B8i; gp.C8u = C8u;
    gp.claims = reinterpret_cast<uint32_t *>(work + P.off_claims);
    for (size_t o = 0; o < (P.outer_rows ? Rb : Cb); ++o) {
        for (size_t q = 0; q < (P.outer_rows ? Cb : Rb); ++q) {
            const size_t i = P.outer_rows ? o : q, j = P.outer_rows ? q : o;
            size_t r0, r1, c0, c1;
            rows_of(i, r0, r1); cols_of(j, c0, c1);
            const bool newA = P.outer_rows ? q == 0 : (Rb > 1 || o == 0);
            const bool newB = P.outer_rows ? (Cb > 1 || o == 0) : q == 0;
            // fast-mode shifts of a re-encoded block are already known after its first visit
            const bool firstA = P.outer_rows ? true : o == 0, firstB = P.outer_rows ? o == 0 : true;
            if (newA && k > 0) {
                int rc = scale_operand(a->dtype_A, a_strided, Aptr(r0), a->lda, r1 - r0, k, ref_width, l2, N, A8i, P.lda8i, P.sizeA, sftA + r0,
                                       fast && firstA, st);
                if (rc) return rc;
            }
            if (newB && k > 0) {
                int rc = scale_operand(a->dtype_B, b_strided, Bptr(c0), a->ldb, c1 - c0, k, ref_width, l2, N, B8i, P.lda8i, P.sizeB, sftB + c0,
                                       fast && firstB, st);
                if (rc) return rc;
            }
            timer.mark(0);
            gp.rowsA = r1 - r0; gp.rowsB = c1 - c0;
            void *Cblk = static_cast<uint8_t *>(a->C) + (c0 * a->ldc + r0) * esC;
            if (take_fused(a)) {
                int rc = block_to_c(a, gp, true, split, Cblk, sftA + r0, sftB + c0, st, nullptr);
                if (rc) return rc;
                timer.mark(1);
            } else {
                OZ_CUDA(oz::launch_gemm_tcgen05(gp, oz::EPI_RESIDUE, st), "int8 gemm");
                timer.mark(1);
                OZ_CUDA(oz::launch_crt(a->dtype_C, split, N, r1 - r0, c1 - c0, C8u, P.mb_pad, P.sizeC, Cblk, a->ldc, sftA + r0, sftB + c0, a->alpha,
                                       a->beta, dev_scalars(a), st), "crt");
                timer.mark(3);
            }
        }
    }
    timer.finish(a->timers_ns);
    return GEMMUL8_OK;
}

}  // namespace

extern "C" {

size_t gemmul8_b200_worksize_blocked_complex(size_t m, size_t n, size_t k, unsigned num_moduli, int compute_type, size_t block_rows,
                                             size_t block_cols) {
    const size_t mb = block_rows < m ? block_rows : m, nb = block_cols < n ? block_cols : n;
    if (mb == 0 || nb == 0 || (mb < m && mb % 256) || (nb < n && nb % 256) || compute_type < 1 || compute_type > 3) return 0;
    return gemmul8_b200_worksize(mb, nb, k, num_moduli, compute_type);
}

size_t gemmul8_b200_worksize_blocked(size_t m, size_t n, size_t k, unsigned num_moduli, size_t block_rows, size_t block_cols) {
    BlockPlan P;
    if (num_moduli < 2 || num_moduli > 20 || !carve_blocks(m, n, k, num_moduli, block_rows, block_cols, P)) return 0;
    return P.total;
}

int gemmul8_b200_plan_blocks(size_t m, size_t n, size_t k, unsigned num_moduli, size_t max_bytes, size_t *block_rows, size_t *block_cols,
                             size_t *work_bytes) {
    BlockPlan P;
    if (num_moduli < 2 || num_moduli > 20) return fail(GEMMUL8_ERR_ARGUMENT, "num_moduli must be in 2..20");
    if (!m || !n) { if (block_rows) *block_rows = m; if (block_cols) *block_cols = n; if (work_bytes) *work_bytes = 0; return GEMMUL8_OK; }
    if (!plan_blocks(m, n, k, num_moduli, max_bytes, P)) return fail(GEMMUL8_ERR_ARGUMENT, "plan_blocks: max_bytes is below the 256 x 256 block minimum");
    if (block_rows) *block_rows = P.mb;
    if (block_cols) *block_cols = P.nb;
    if (work_bytes) *work_bytes = P.total;
    return GEMMUL8_OK;
}

int gemmul8_b200_gemm_blocked(gemmul8_b200_args *a, size_t block_rows, size_t block_cols) {
    if (a) for (double &t : a->timers_ns) t = 0.0;
    int rc = check_args(a);
    if (rc) return rc;
    if (a->m == 0 || a->n == 0) return GEMMUL8_OK;
    if (is_complex(a->dtype_C)) {
        // Complex types, fast mode: every C block is one complete complex call on its row block of op(A) and column block of
        // op(B), in a workspace of workSize(block_rows, block_cols, k, N, compute_type) bytes.  Fast-mode shifts depend on the
        // row / column alone, so the bits are those of gemm(); an operand block is re-encoded for every block it meets
        // (encoding is ~10 % of a complex call).  Accurate mode needs the bound product of the whole matrix: not blocked.
        if (!a->fastmode) return fail(GEMMUL8_ERR_ARGUMENT, "gemm_blocked: complex types in fast mode only");
        if (a->k > (size_t(1) << 16)) return fail(GEMMUL8_ERR_ARGUMENT, "complex types: k must be <= 2^16 (int32 accumulation)");
        const size_t mb = block_rows < a->m ? block_rows : a->m, nb = block_cols < a->n ? block_cols : a->n;
        if (mb == 0 || nb == 0 || (mb < a->m && mb % 256) || (nb < a->n && nb % 256))
            return fail(GEMMUL8_ERR_ARGUMENT, "gemm_blocked: block sizes must be multiples of 256 (or cover the whole dimension)");
        if (a->k == 0) {
            OZ_CUDA(oz::launch_scale_c(a->dtype_C, a->m, a->n, a->C, a->ldc, a->beta, dev_scalars(a), static_cast<cudaStream_t>(a->stream)), "scale C");
            return GEMMUL8_OK;
        }
        const size_t esA = elem_size(a->dtype_A), esB = elem_size(a->dtype_B), esC = elem_size(a->dtype_C);
        const bool a_rows = a->op_A == GEMMUL8_OP_N, b_rows = a->op_B != GEMMUL8_OP_N;   // the block index runs along stored rows
        for (size_t r0 = 0; r0 < a->m; r0 += mb) {
            for (size_t c0 = 0; c0 < a->n; c0 += nb) {
                gemmul8_b200_args s = *a;
                s.m = r0 + mb < a->m ? mb : a->m - r0;
                s.n = c0 + nb < a->n ? nb : a->n - c0;
                s.A = static_cast<const uint8_t *>(a->A) + (a_rows ? r0 : r0 * a->lda) * esA;
                s.B = static_cast<const uint8_t *>(a->B) + (b_rows ? c0 : c0 * a->ldb) * esB;
                s.C = static_cast<uint8_t *>(a->C) + (c0 * a->ldc + r0) * esC;
                s.flags = a->flags & ~(unsigned)GEMMUL8_FLAG_PHASE_LOG;
                rc = gemm_complex(&s);
                if (rc) return rc;
                for (int t = 0; t < 4; ++t) a->timers_ns[t] += s.timers_ns[t];
            }
        }
        return GEMMUL8_OK;
    }
    BlockPlan P;
    if (!carve_blocks(a->m, a->n, a->k, a->num_moduli, block_rows, block_cols, P))
        return fail(GEMMUL8_ERR_ARGUMENT, "gemm_blocked: block sizes must be multiples of 256 (or cover the whole dimension)");
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
        return fail(GEMMUL8_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (a->k == 0) {
        OZ_CUDA(oz::launch_scale_c(a->dtype_C, a->m, a->n, a->C, a->ldc, a->beta, dev_scalars(a), static_cast<cudaStream_t>(a->stream)), "scale C");
        return GEMMUL8_OK;
    }
    return gemm_blocked_real(a, P);
}

}  // extern "C"

// Host-buffer call, real types, fast mode, beta == 0: a wavefront over blocks of C (up to 14 row blocks x 26 column blocks).  The H2D
// stream brings A row blocks and B column blocks alternately (A0, B0, A1, B1, ...); as soon as block
// pair s is on the device the compute stream scales / encodes it and multiplies everything that has
// become computable -- the column strip (rows 0..s, column block s) and the row strip (row block s,
// column blocks 0..s-1) -- and the D2H stream returns each finished strip of C while the next blocks
// are still arriving.  PCIe is full duplex, so the call ends about one strip after the last input
// byte instead of after (inputs + compute + output) in series.
// Measured at 16384^3 (PCIe: 55.6 GB/s H2D, 57.2 GB/s D2H alone, ~53 / ~45 GB/s when both run; strided block copies as fast as
// contiguous ones down to 4 KB runs in isolation, tools/pcie_d2h_width.py): 85.2 ms with the block schedule below.  With
// symmetric blocks (A and B ending together) the last input byte lands at ~81 ms, the last product ends ~5 ms later and the
// D2H queue -- row strips of C -- drains ~4.5 ms after that (profiles/r02_e2e_timeline.txt); "all of B, then A in 16 row
// strips" is slower still (105 ms: products can only start once B is complete).
static int gemm_host_pipelined(gemmul8_b200_args *h, void *dev_scratch) {
    const size_t m = h->m, n = h->n, k = h->k;
    const unsigned N = h->num_moduli, ti = N - 2;
    cudaStream_t st = static_cast<cudaStream_t>(h->stream);
    const size_t esA = elem_size(h->dtype_A), esB = elem_size(h->dtype_B), esC = elem_size(h->dtype_C);
    const size_t colsA = h->op_A == GEMMUL8_OP_N ? k : m, colsB = h->op_B == GEMMUL8_OP_N ? n : k;
    uint8_t *p = static_cast<uint8_t *>(dev_scratch);
    uint8_t *dA = p; p += ceil_to(h->lda * colsA * esA, 256);
    uint8_t *dB = p; p += ceil_to(h->ldb * colsB * esB, 256);
    uint8_t *dC = p; p += ceil_to(h->ldc * n * esC, 256);
    uint8_t *work = p;
    oz::Layout L;
    compute_layout(m, n, k, N, GEMMUL8_REAL_DEFAULT, L);
    int8_t *A8i   = reinterpret_cast<int8_t *>(work + L.off_A8i);
    int8_t *B8i   = reinterpret_cast<int8_t *>(work + L.off_B8i);
    uint8_t *C8u  = work + L.off_C8u;
    int16_t *sftA = reinterpret_cast<int16_t *>(work + L.off_sftA);
    int16_t *sftB = reinterpret_cast<int16_t *>(work + L.off_sftB);
    const bool a_strided = h->op_A == GEMMUL8_OP_N, b_strided = h->op_B != GEMMUL8_OP_N;
    const int ref_width  = oz::ref_reduce_width(h->dtype_A, h->dtype_B, h->dtype_C);
    const float l2       = oz::host_tab::OZ_LOG2M_FAST[ti];
    const bool split     = oz::host_tab::OZ_M_LO[ti] != 0.0 && h->dtype_C == GEMMUL8_F64;

    // Block boundaries: multiples of 256 rows / columns (whole GEMM tiles), in shares of 1/64 of the dimension.
    //   steps 0..13  the wavefront proper: A row blocks SHRINKING towards the end, B column blocks in proportion but only up
    //                to 5/8 of B -- the work a block pair enables grows with everything that arrived before it, so this part
    //                is transfer-bound and the compute stream keeps up;
    //   steps 14..25 A is complete: the remaining 3/8 of B arrives in small column blocks, each enabling one full-height
    //                column strip of C.  Work is now enabled linearly with the bytes that arrive, every strip of C is a
    //                contiguous D2H (row strips are strided copies with 2-4 KB runs, which reach ~32 GB/s inside this call
    //                against ~55 GB/s for column strips), and what is left after the last input byte is one 256-column strip.
    // Measured at 16384^3, one box (profiles/r02_e2e_timeline.txt, r02_ab_e2e_schedule.jsonl): symmetric 12 blocks ending in
    // 1/24 shares 93.0 ms; symmetric 14 blocks ending in 1/32 shares 90.5 - 90.9; A in 16 blocks with the last 1/4 of B behind
    // it 86.9; this schedule 85.2.
    constexpr int kMaxBlocks = 26, kShareSum = 64, kRowBlocks = 14, kColBlocks = 26;
    static const int row_share[kMaxBlocks] = {8, 8, 8, 8, 6, 6, 4, 4, 2, 2, 2, 2, 2, 2};
    static const int col_share[kMaxBlocks] = {5, 5, 5, 5, 4, 4, 2, 2, 2, 1, 1, 2, 1, 1, 3, 3, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1};
    auto bounds = [](size_t len, const int *share, int count, size_t *b) -> int {
        const size_t tiles = (len + 255) / 256;
        if (tiles < 32) {                       // small problems: equal blocks, at most 8
            const int S = (int)(tiles < 8 ? (tiles ? tiles : 1) : 8);
            for (int i = 0; i <= S; ++i) { size_t x = (tiles * i / S) * 256; b[i] = x < len ? x : len; }
            b[S] = len;
            return S;
        }
        size_t acc = 0;
        b[0] = 0;
        for (int i = 0; i < count; ++i) {
            acc += (size_t)share[i];
            size_t x = (tiles * acc / kShareSum) * 256;
            b[i + 1] = x < len ? x : len;
        }
        b[count] = len;
        return count;
    };
    size_t rb[kMaxBlocks + 1], cb[kMaxBlocks + 1];
    const int SR = bounds(m, row_share, kRowBlocks, rb), SC = bounds(n, col_share, kColBlocks, cb);
    const int S = SR > SC ? SR : SC;

    struct Res {
        cudaStream_t in = nullptr, out = nullptr;
        cudaEvent_t evA[kMaxBlocks], evB[kMaxBlocks], evC[2 * kMaxBlocks], start;
        int nev = 0;
        ~Res() {
            for (int i = 0; i < nev; ++i) { cudaEventDestroy(evA[i]); cudaEventDestroy(evB[i]); cudaEventDestroy(evC[2 * i]); cudaEventDestroy(evC[2 * i + 1]); }
            if (nev) cudaEventDestroy(start);
            if (in) cudaStreamDestroy(in);
            if (out) cudaStreamDestroy(out);
        }
    } r;
    OZ_CUDA(cudaStreamCreateWithFlags(&r.in, cudaStreamNonBlocking), "stream");
    OZ_CUDA(cudaStreamCreateWithFlags(&r.out, cudaStreamNonBlocking), "stream");
    OZ_CUDA(cudaEventCreateWithFlags(&r.start, cudaEventDisableTiming), "event");
    for (int i = 0; i < S; ++i) {
        OZ_CUDA(cudaEventCreateWithFlags(&r.evA[i], cudaEventDisableTiming), "event");
        OZ_CUDA(cudaEventCreateWithFlags(&r.evB[i], cudaEventDisableTiming), "event");
        OZ_CUDA(cudaEventCreateWithFlags(&r.evC[2 * i], cudaEventDisableTiming), "event");
        OZ_CUDA(cudaEventCreateWithFlags(&r.evC[2 * i + 1], cudaEventDisableTiming), "event");
        r.nev = i + 1;
    }
    // the copies may not start before earlier work of the caller's stream on this scratch has finished
    OZ_CUDA(cudaEventRecord(r.start, st), "event record");
    OZ_CUDA(cudaStreamWaitEvent(r.in, r.start, 0), "stream wait");

    auto copy_block = [&](uint8_t *dst, const void *src_, size_t ld, size_t es, bool rows, size_t lo, size_t hi, size_t other,
                          cudaMemcpyKind kind, cudaStream_t s) -> cudaError_t {
        // rows == true: rows [lo, hi) of a column-major matrix with `other` columns (strided); else columns [lo, hi)
        if (hi <= lo) return cudaSuccess;
        const uint8_t *src = static_cast<const uint8_t *>(src_);
        if (rows) return cudaMemcpy2DAsync(dst + lo * es, ld * es, src + lo * es, ld * es, (hi - lo) * es, other, kind, s);
        return cudaMemcpyAsync(dst + lo * ld * es, src + lo * ld * es, (hi - lo) * ld * es, kind, s);
    };

    oz::GemmProblem gp{};
    gp.ld8i = L.lda8i; gp.sizeA = L.sizeA; gp.sizeB = L.sizeB; gp.num_slices = N; gp.first_modulus = 0;
    gp.ldc8u = L.m_pad; gp.sizeC = L.sizeC; gp.claims = claims_of(work, L);
    auto strip = [&](size_t r0, size_t r1, size_t c0, size_t c1, cudaEvent_t done) -> int {
        if (r1 <= r0 || c1 <= c0) return GEMMUL8_OK;
        gp.A8i = A8i + r0 * L.lda8i; gp.rowsA = r1 - r0;
        gp.B8i = B8i + c0 * L.lda8i; gp.rowsB = c1 - c0;
        gp.C8u = C8u + c0 * L.m_pad + r0;
        OZ_CUDA(oz::launch_gemm_tcgen05(gp, oz::EPI_RESIDUE, st), "int8 gemm");
        OZ_CUDA(oz::launch_crt(h->dtype_C, split, N, r1 - r0, c1 - c0, gp.C8u, L.m_pad, L.sizeC, dC + (c0 * h->ldc + r0) * esC, h->ldc,
                               sftA + r0, sftB + c0, h->alpha, h->beta, false, st), "crt");
        OZ_CUDA(cudaEventRecord(done, st), "event record");
        OZ_CUDA(cudaStreamWaitEvent(r.out, done, 0), "stream wait");
        // rows [r0, r1) of columns [c0, c1) of C
        OZ_CUDA(cudaMemcpy2DAsync(static_cast<uint8_t *>(h->C) + (c0 * h->ldc + r0) * esC, h->ldc * esC, dC + (c0 * h->ldc + r0) * esC,
                                  h->ldc * esC, (r1 - r0) * esC, c1 - c0, cudaMemcpyDeviceToHost, r.out), "D2H C");
        return GEMMUL8_OK;
    };

    for (int s = 0; s < S; ++s) {
        const size_t r0 = s < SR ? rb[s] : m, r1 = s < SR ? rb[s + 1] : m;
        const size_t c0 = s < SC ? cb[s] : n, c1 = s < SC ? cb[s + 1] : n;
        OZ_CUDA(copy_block(dA, h->A, h->lda, esA, a_strided, r0, r1, colsA, cudaMemcpyHostToDevice, r.in), "H2D A");
        OZ_CUDA(cudaEventRecord(r.evA[s], r.in), "event record");
        OZ_CUDA(copy_block(dB, h->B, h->ldb, esB, b_strided, c0, c1, colsB, cudaMemcpyHostToDevice, r.in), "H2D B");
        OZ_CUDA(cudaEventRecord(r.evB[s], r.in), "event record");
    }
    for (int s = 0; s < S; ++s) {
        const size_t r0 = s < SR ? rb[s] : m, r1 = s < SR ? rb[s + 1] : m;
        const size_t c0 = s < SC ? cb[s] : n, c1 = s < SC ? cb[s + 1] : n;
        OZ_CUDA(cudaStreamWaitEvent(st, r.evA[s], 0), "stream wait");
        if (r1 > r0) {
            const uint8_t *Ax = dA + (a_strided ? r0 : r0 * h->lda) * esA;
            int rc = scale_operand(h->dtype_A, a_strided, Ax, h->lda, r1 - r0, k, ref_width, l2, N, A8i + r0 * L.lda8i, L.lda8i, L.sizeA,
                                   sftA + r0, true, st);
            if (rc) return rc;
        }
        OZ_CUDA(cudaStreamWaitEvent(st, r.evB[s], 0), "stream wait");
        if (c1 > c0) {
            const uint8_t *Bx = dB + (b_strided ? c0 : c0 * h->ldb) * esB;
            int rc = scale_operand(h->dtype_B, b_strided, Bx, h->ldb, c1 - c0, k, ref_width, l2, N, B8i + c0 * L.lda8i, L.lda8i, L.sizeB,
                                   sftB + c0, true, st);
            if (rc) return rc;
        }
        int rc = strip(0, r1, c0, c1, r.evC[2 * s]);          // column strip: all rows so far x the new column block
        if (rc) return rc;
        rc = strip(r0, r1, 0, c0, r.evC[2 * s + 1]);          // row strip: the new row block x the earlier column blocks
        if (rc) return rc;
    }
    OZ_CUDA(cudaStreamSynchronize(r.out), "sync");
    OZ_CUDA(cudaStreamSynchronize(st), "sync");
    return GEMMUL8_OK;
}

extern "C" {

int gemmul8_b200_gemm_host(gemmul8_b200_args *h, void *dev_scratch) {
    int rc = check_args(h, /*need_work=*/false);   // the workspace is carved out of dev_scratch; args->work is not read
    if (rc) return rc;
    if (dev_scalars(h)) return fail(GEMMUL8_ERR_ARGUMENT, "gemm_host: alpha / beta are host scalars here");
    if (h) for (double &t : h->timers_ns) t = 0.0;
    if (!dev_scratch) return fail(GEMMUL8_ERR_ARGUMENT, "null scratch");
    if (h->m == 0 || h->n == 0) return GEMMUL8_OK;
    cudaStream_t st = static_cast<cudaStream_t>(h->stream);
    bool beta_zero = true;
    const size_t es = elem_size(h->dtype_C);
    for (size_t i = 0; i < es; ++i) beta_zero &= static_cast<const unsigned char *>(h->beta)[i] == 0;
    if (!is_complex(h->dtype_C) && h->fastmode && beta_zero && h->k > 0 &&
        !(h->flags & (GEMMUL8_FLAG_TIMERS | GEMMUL8_FLAG_STAGE_SCALING | GEMMUL8_FLAG_STAGE_RESIDUES | GEMMUL8_FLAG_GEMM_SIMT |
                      GEMMUL8_FLAG_FUSED_CRT | GEMMUL8_FLAG_HOST_SERIAL)))
        return gemm_host_pipelined(h, dev_scratch);
    // everything else (accurate mode needs all of A and B before the first shift; complex; beta != 0): in series
    const size_t colsA = h->op_A == GEMMUL8_OP_N ? h->k : h->m, colsB = h->op_B == GEMMUL8_OP_N ? h->n : h->k;
    const size_t bytesA = h->lda * colsA * elem_size(h->dtype_A), bytesB = h->ldb * colsB * elem_size(h->dtype_B);
    const size_t bytesC = h->ldc * h->n * elem_size(h->dtype_C);
    uint8_t *p = static_cast<uint8_t *>(dev_scratch);
    void *dA = p; p += ceil_to(bytesA, 256);
    void *dB = p; p += ceil_to(bytesB, 256);
    void *dC = p; p += ceil_to(bytesC, 256);
    gemmul8_b200_args d = *h;
    d.A = dA; d.B = dB; d.C = dC; d.work = p;
    OZ_CUDA(cudaMemcpyAsync(dA, h->A, bytesA, cudaMemcpyHostToDevice, st), "H2D A");
    OZ_CUDA(cudaMemcpyAsync(dB, h->B, bytesB, cudaMemcpyHostToDevice, st), "H2D B");
    if (!beta_zero) OZ_CUDA(cudaMemcpyAsync(dC, h->C, bytesC, cudaMemcpyHostToDevice, st), "H2D C");
    rc = gemmul8_b200_gemm(&d);
    memcpy(h->timers_ns, d.timers_ns, sizeof(d.timers_ns));
    if (rc) return rc;
    OZ_CUDA(cudaMemcpyAsync(h->C, dC, bytesC, cudaMemcpyDeviceToHost, st), "D2H C");
    OZ_CUDA(cudaStreamSynchronize(st), "sync");
    return GEMMUL8_OK;
}

int gemmul8_b200_product_i32(const gemmul8_b200_args *a, unsigned j, int32_t *C32i_out, int imag_part) {
    int rc = check_args(a);
    if (rc) return rc;
    if (j >= a->num_moduli || !C32i_out || imag_part) return fail(GEMMUL8_ERR_ARGUMENT, "bad slice index / output");
    oz::Layout L;
    compute_layout(a->m, a->n, a->k, a->num_moduli, a->compute_type, L);
    uint8_t *work = static_cast<uint8_t *>(a->work);
    oz::GemmProblem gp{};
    gp.A8i = reinterpret_cast<int8_t *>(work + L.off_A8i) + (size_t)j * L.sizeA;
    gp.B8i = reinterpret_cast<int8_t *>(work + L.off_B8i) + (size_t)j * L.sizeB;
    gp.rowsA = a->compute_type == GEMMUL8_COMPLEX_BIG_MATRIX_ENCODE ? 2 * a->m : a->m;
    gp.rowsB = a->n; gp.ld8i = L.lda8i; gp.sizeA = L.sizeA; gp.sizeB = L.sizeB; gp.num_slices = 1; gp.first_modulus = j;
    gp.C32i = C32i_out; gp.ldc32i = L.m_pad;
    auto gemm = (a->flags & GEMMUL8_FLAG_GEMM_SIMT) ? oz::launch_gemm_simt : oz::launch_gemm_tcgen05;
    OZ_CUDA(gemm(gp, oz::EPI_INT32, static_cast<cudaStream_t>(a->stream)), "int32 product");
    return GEMMUL8_OK;
}

int gemmul8_b200_phase_log_collect(double timers_ns[4], unsigned *calls) {
    PhaseLog &L = g_phase_log;
    if (timers_ns) for (int i = 0; i < 4; ++i) timers_ns[i] = 0.0;
    if (calls) *calls = (unsigned)L.ring.size();
    int rc = GEMMUL8_OK;
    for (auto &e : L.ring) {
        if (e.n > 0 && cudaEventSynchronize(e.ev[e.n - 1]) != cudaSuccess) rc = fail(GEMMUL8_ERR_CUDA, "phase log: event synchronise failed");
        double t[4] = {0, 0, 0, 0};
        if (rc == GEMMUL8_OK) PhaseTimer::spans(e.ev, e.n, t);
        if (timers_ns) for (int i = 0; i < 4; ++i) timers_ns[i] += t[i];
        for (auto ev : e.ev) L.put(ev);
    }
    L.ring.clear();
    L.head = 0;
    return rc;
}

int gemmul8_b200_init(int device) {
    int count = 0, prev = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(GEMMUL8_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device >= count) return fail(GEMMUL8_ERR_ARGUMENT, "init: no such device");
    oz::tuning();                                   // environment defaults, once
    OZ_CUDA(cudaGetDevice(&prev), "get device");
    if (device >= 0 && device != prev) OZ_CUDA(cudaSetDevice(device), "set device");
    oz::gemm_prepare_device(/*allow_probe=*/true);  // SM placement table of the pair GEMM (private stream, once per device)
    if (device >= 0 && device != prev) OZ_CUDA(cudaSetDevice(prev), "set device");
    return GEMMUL8_OK;
}

int gemmul8_b200_set_option(const char *name, int value) {
    oz::load_options_from_env();
    oz::Option *o = oz::find_option(name);
    if (!o) return fail(GEMMUL8_ERR_ARGUMENT, "set_option: unknown option");
    if (value < o->lo || value > o->hi) return fail(GEMMUL8_ERR_ARGUMENT, "set_option: value out of range");
    o->value.store(value, std::memory_order_relaxed);
    return GEMMUL8_OK;
}

int gemmul8_b200_get_option(const char *name, int *value) {
    oz::load_options_from_env();
    if (name && value && !strcmp(name, "strip_calls")) {   // read-only counter: calls that took the column-strip pipeline
        *value = oz::g_strip_calls.load(std::memory_order_relaxed);
        return GEMMUL8_OK;
    }
    oz::Option *o = oz::find_option(name);
    if (!o || !value) return fail(GEMMUL8_ERR_ARGUMENT, "get_option: unknown option");
    *value = o->value.load(std::memory_order_relaxed);
    return GEMMUL8_OK;
}

int gemmul8_b200_modulus(unsigned j) { return j < 20 ? oz::host_tab::OZ_MOD[j] : 0; }

double gemmul8_b200_crt_weight(unsigned N, unsigned j, int part) {
    if (N < 2 || N > 20 || j >= N) return 0.0;
    if (part == 0) return oz::host_tab::OZ_W1[N - 2][j];
    if (N < 8) return 0.0;
    return part == 1 ? oz::host_tab::OZ_W2_HI[N - 8][j] : oz::host_tab::OZ_W2_LO[N - 8][j];
}

unsigned long long gemmul8_b200_launch_count(void) { return oz::g_launches.load(std::memory_order_relaxed); }

const char *gemmul8_b200_last_error(void) { return g_last_error.c_str(); }
const char *gemmul8_b200_version(void) { return "gemmul8_b200 0.1 (sm_100a, tcgen05 kind::i8)"; }

}  // extern "C"
