// libgemmul8_b200_mp.so -- 2-D block decomposition of C over the GPUs of one node, C ABI in include/gemmul8_b200_mp.h.
//
// Host orchestration only: the kernels are those of libgemmul8_b200.so, driven block-wise through gemmul8_b200_gemm_part
// (real types, fast mode) or gemmul8_b200_gemm with the ONLY_BOUND / SKIP_BOUND split (accurate mode).  What this file adds
// is the panel exchange and its overlap with the compute stream:
//
//   exchange stream A (grid row)     piece 0 | piece 1 | piece 2 | piece 3        row pieces of the A panel, own 1/Q from a_slice
//   exchange stream B (grid column)  block 0 | ... | block P-1                    column blocks of the B panel, own 1/P from b_slice
//   compute stream                   scale A_0, scale B_0.., C[0, :] | scale A_1, C[1, :] | ...
//
// NCCL transport: in-place ncclAllGather of every A piece, in-place ncclBroadcast of every B block; events order the compute
// stream behind them.  NCCL's collectives are kernels, so the block-wise entry launches the persistent GEMM with 4 pipeline
// stages to leave them room on the SMs (GemmProblem::share_sm).
// COPY transport: every rank PUSHES its pieces into its peers' panel buffers (mapped through CUDA IPC) with one
// cudaMemcpy2DAsync per piece and peer -- copy engines, no SM -- and then raises a 32-bit flag in the peer's flag array
// (cuMemsetD32Async on the mapped address, stream-ordered behind the copy).  The consumer's compute stream waits for the flag
// with cuStreamWaitValue32.  Flags carry the call's epoch, so nothing is ever reset; before a rank overwrites a peer's buffer
// for epoch e it waits for that peer's acknowledgement of epoch e - 1 (raised behind the peer's last read of the buffer).
#include "../../include/gemmul8_b200.h"
#include "../../include/gemmul8_b200_mp.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define MP_CUDA(call, where)                                                                                   \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) return fail(GEMMUL8_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define MP_NCCL(call, where)                                                                                   \
    do {                                                                                                       \
        ncclResult_t r__ = (call);                                                                             \
        if (r__ != ncclSuccess) return fail(GEMMUL8_ERR_CUDA, std::string(where) + ": " + ncclGetErrorString(r__)); \
    } while (0)
#define MP_CU(call, where)                                                                                     \
    do {                                                                                                       \
        CUresult r__ = (call);                                                                                 \
        if (r__ != CUDA_SUCCESS) return fail(GEMMUL8_ERR_CUDA, std::string(where) + ": driver error " + std::to_string((int)r__)); \
    } while (0)

constexpr int kMaxPieces = 4;    // row pieces of the A panel
constexpr int kMaxSide   = 16;   // P, Q <= 16

size_t elem_size(int dt) { return dt == GEMMUL8_F32 ? 4 : dt == GEMMUL8_C64 ? 16 : 8; }
bool is_complex(int dt) { return dt == GEMMUL8_C32 || dt == GEMMUL8_C64; }

// driver entry points (stream memory operations), resolved once
struct Driver {
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned) = nullptr;
    CUresult (*memset32)(CUdeviceptr, unsigned, size_t, CUstream) = nullptr;
    bool ok = false;
    Driver() {
        cudaDriverEntryPointQueryResult q;
        void *p = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            wait32 = reinterpret_cast<decltype(wait32)>(p);
        if (cudaGetDriverEntryPoint("cuMemsetD32Async", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            memset32 = reinterpret_cast<decltype(memset32)>(p);
        ok = wait32 && memset32;
    }
};
const Driver &driver() { static Driver d; return d; }

}  // namespace

struct gemmul8_b200_grid {
    int rank = 0, nranks = 1, P = 1, Q = 1, p = 0, q = 0, device = 0, exchange = 0;
    ncclComm_t world = nullptr, row = nullptr, col = nullptr;
    bool own_world = false;
    cudaStream_t xa = nullptr, xb = nullptr;             // NCCL transport / local copies: one stream per direction
    cudaStream_t xpa[kMaxSide] = {}, xpb[kMaxSide] = {}; // COPY transport: one stream per peer, so that the pushes to different
                                                         // peers run on different copy engines at the same time
    cudaEvent_t ev_start = nullptr, ev_a[kMaxPieces] = {}, ev_b[kMaxSide] = {}, ev_xa_done = nullptr, ev_xb_done = nullptr;
    cudaEvent_t ev_pa[kMaxSide] = {}, ev_pb[kMaxSide] = {};
    uint8_t *a_buf = nullptr, *b_buf = nullptr;
    size_t a_cap = 0, b_cap = 0;
    // COPY transport
    uint32_t *flags = nullptr;                         // own flag array (layout below)
    uint8_t *peer_a[kMaxSide] = {};                    // a_buf of the ranks of my grid row (index q')
    uint8_t *peer_b[kMaxSide] = {};                    // b_buf of the ranks of my grid column (index p')
    uint32_t *peer_flags_row[kMaxSide] = {}, *peer_flags_col[kMaxSide] = {};
    uint32_t epoch = 0;
    // flag array layout (uint32 words)
    static constexpr int F_A = 0;                                  // [piece][source q']   data of A piece landed
    static constexpr int F_B = F_A + kMaxPieces * kMaxSide;        // [source p']          data of B block landed
    static constexpr int F_ACK_A = F_B + kMaxSide;                 // [reader q']          reader q' has finished with MY pieces in its a_buf
    static constexpr int F_ACK_B = F_ACK_A + kMaxSide;             // [reader p']
    static constexpr int F_TOTAL = F_ACK_B + kMaxSide + 1;         // (+ one scratch word for the barrier in destroy)
};

namespace {

int row_pieces(size_t rows, int want, size_t *b) {
    const size_t tiles = (rows + 255) / 256;
    size_t n = tiles < (size_t)want ? tiles : (size_t)want;
    if (n < 1) n = 1;
    int cnt = 0;
    b[0] = 0;
    for (size_t i = 1; i <= n; ++i) {
        size_t x = i == n ? rows : (tiles * i / n) * 256;
        if (x > rows) x = rows;
        if (x > b[cnt]) b[++cnt] = x;
    }
    if (cnt == 0) { b[1] = rows; cnt = 1; }
    return cnt;
}

int finish_create(gemmul8_b200_grid *g, size_t a_bytes, size_t b_bytes) {
    MP_CUDA(cudaGetDevice(&g->device), "get device");
    MP_NCCL(ncclCommSplit(g->world, g->p, g->q, &g->row, nullptr), "split row communicator");
    MP_NCCL(ncclCommSplit(g->world, g->q, g->p, &g->col, nullptr), "split column communicator");
    MP_CUDA(cudaStreamCreateWithFlags(&g->xa, cudaStreamNonBlocking), "stream");
    MP_CUDA(cudaStreamCreateWithFlags(&g->xb, cudaStreamNonBlocking), "stream");
    MP_CUDA(cudaEventCreateWithFlags(&g->ev_start, cudaEventDisableTiming), "event");
    MP_CUDA(cudaEventCreateWithFlags(&g->ev_xa_done, cudaEventDisableTiming), "event");
    MP_CUDA(cudaEventCreateWithFlags(&g->ev_xb_done, cudaEventDisableTiming), "event");
    for (auto &e : g->ev_a) MP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
    for (auto &e : g->ev_b) MP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
    g->a_cap = (a_bytes + 255) / 256 * 256;
    g->b_cap = (b_bytes + 255) / 256 * 256;
    if (g->Q > 1 && g->a_cap) MP_CUDA(cudaMalloc(&g->a_buf, g->a_cap), "panel buffer A");
    if (g->P > 1 && g->b_cap) MP_CUDA(cudaMalloc(&g->b_buf, g->b_cap), "panel buffer B");
    if (g->exchange != GEMMUL8_MP_EXCHANGE_COPY || g->nranks == 1) return GEMMUL8_OK;

    // ---- COPY transport: map the peers' panel buffers and flag arrays ----
    if (!driver().ok) return fail(GEMMUL8_ERR_CUDA, "stream memory operations are not available in this driver");
    MP_CUDA(cudaMalloc(&g->flags, sizeof(uint32_t) * gemmul8_b200_grid::F_TOTAL), "flag array");
    MP_CUDA(cudaMemset(g->flags, 0, sizeof(uint32_t) * gemmul8_b200_grid::F_TOTAL), "flag array");
    struct Handles { cudaIpcMemHandle_t a, b, f; int has_a, has_b, device, pad; };
    Handles mine{};
    mine.device = g->device;
    if (g->a_buf) { MP_CUDA(cudaIpcGetMemHandle(&mine.a, g->a_buf), "ipc handle A"); mine.has_a = 1; }
    if (g->b_buf) { MP_CUDA(cudaIpcGetMemHandle(&mine.b, g->b_buf), "ipc handle B"); mine.has_b = 1; }
    MP_CUDA(cudaIpcGetMemHandle(&mine.f, g->flags), "ipc handle flags");
    Handles *dev_all = nullptr;
    std::vector<Handles> all((size_t)g->nranks);
    MP_CUDA(cudaMalloc(&dev_all, sizeof(Handles) * (size_t)g->nranks), "handle exchange buffer");
    MP_CUDA(cudaMemcpy(dev_all + g->rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice), "handle upload");
    MP_NCCL(ncclAllGather(dev_all + g->rank, dev_all, sizeof(Handles), ncclInt8, g->world, g->xa), "handle all-gather");
    MP_CUDA(cudaStreamSynchronize(g->xa), "handle all-gather");
    MP_CUDA(cudaMemcpy(all.data(), dev_all, sizeof(Handles) * (size_t)g->nranks, cudaMemcpyDeviceToHost), "handle download");
    cudaFree(dev_all);
    auto open = [&](const cudaIpcMemHandle_t &h, void **out) -> cudaError_t {
        return cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
    };
    for (int qq = 0; qq < g->Q; ++qq) {
        const int r = g->p * g->Q + qq;
        if (r == g->rank) { g->peer_a[qq] = g->a_buf; g->peer_flags_row[qq] = g->flags; continue; }
        if (all[r].has_a) MP_CUDA(open(all[r].a, reinterpret_cast<void **>(&g->peer_a[qq])), "ipc open A (peer access between the GPUs?)");
        MP_CUDA(open(all[r].f, reinterpret_cast<void **>(&g->peer_flags_row[qq])), "ipc open flags");
    }
    for (int pp = 0; pp < g->P; ++pp) {
        const int r = pp * g->Q + g->q;
        if (r == g->rank) { g->peer_b[pp] = g->b_buf; g->peer_flags_col[pp] = g->flags; continue; }
        if (all[r].has_b) MP_CUDA(open(all[r].b, reinterpret_cast<void **>(&g->peer_b[pp])), "ipc open B (peer access between the GPUs?)");
        // a rank that shares both my row and my column is me; otherwise the flag array of a column peer is a second mapping
        MP_CUDA(open(all[r].f, reinterpret_cast<void **>(&g->peer_flags_col[pp])), "ipc open flags");
    }
    for (int qq = 0; qq < g->Q; ++qq)
        if (qq != g->q) {
            MP_CUDA(cudaStreamCreateWithFlags(&g->xpa[qq], cudaStreamNonBlocking), "stream");
            MP_CUDA(cudaEventCreateWithFlags(&g->ev_pa[qq], cudaEventDisableTiming), "event");
        }
    for (int pp = 0; pp < g->P; ++pp)
        if (pp != g->p) {
            MP_CUDA(cudaStreamCreateWithFlags(&g->xpb[pp], cudaStreamNonBlocking), "stream");
            MP_CUDA(cudaEventCreateWithFlags(&g->ev_pb[pp], cudaEventDisableTiming), "event");
        }
    // (flags were zeroed before their handle was published, so no peer can push before that)
    return GEMMUL8_OK;
}

// stream-ordered "flag[word] = value" in a (possibly peer) flag array / wait until own flag[word] >= value
int raise_flag(uint32_t *base, int word, uint32_t value, cudaStream_t st) {
    MP_CU(driver().memset32(reinterpret_cast<CUdeviceptr>(base + word), value, 1, st), "raise flag");
    return GEMMUL8_OK;
}
int wait_flag(uint32_t *base, int word, uint32_t value, cudaStream_t st) {
    MP_CU(driver().wait32(st, reinterpret_cast<CUdeviceptr>(base + word), value, CU_STREAM_WAIT_VALUE_GEQ), "wait flag");
    return GEMMUL8_OK;
}

}  // namespace

extern "C" {

const char *gemmul8_b200_mp_last_error(void) { return g_err.c_str(); }

int gemmul8_b200_mp_row_pieces(size_t rows, int want, size_t bounds[9]) {
    if (want < 1) want = 1;
    if (want > 8) want = 8;
    return row_pieces(rows, want, bounds);
}

int gemmul8_b200_mp_unique_id(void *id_bytes) {
    static_assert(sizeof(ncclUniqueId) == GEMMUL8_MP_ID_BYTES, "ncclUniqueId size");
    if (!id_bytes) return fail(GEMMUL8_ERR_ARGUMENT, "null id");
    ncclUniqueId id;
    MP_NCCL(ncclGetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id_bytes, &id, sizeof(id));
    return GEMMUL8_OK;
}

static int check_shape(int nranks, int P, int Q) {
    if (P < 1 || Q < 1 || P > kMaxSide || Q > kMaxSide || P * Q != nranks) return fail(GEMMUL8_ERR_ARGUMENT, "grid: P * Q must equal the number of ranks (P, Q <= 16)");
    return GEMMUL8_OK;
}

int gemmul8_b200_grid_create(const void *id_bytes, int rank, int nranks, int P, int Q, size_t a_panel_bytes, size_t b_panel_bytes,
                             int exchange, gemmul8_b200_grid **out) {
    if (!id_bytes || !out) return fail(GEMMUL8_ERR_ARGUMENT, "null argument");
    int rc = check_shape(nranks, P, Q);
    if (rc) return rc;
    if (rank < 0 || rank >= nranks) return fail(GEMMUL8_ERR_ARGUMENT, "bad rank");
    auto *g = new gemmul8_b200_grid();
    g->rank = rank; g->nranks = nranks; g->P = P; g->Q = Q; g->p = rank / Q; g->q = rank % Q; g->exchange = exchange;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&g->world, nranks, id, rank);
    if (r != ncclSuccess) { delete g; return fail(GEMMUL8_ERR_CUDA, std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); }
    g->own_world = true;
    rc = finish_create(g, a_panel_bytes, b_panel_bytes);
    if (rc) { gemmul8_b200_grid_destroy(g); return rc; }
    *out = g;
    return GEMMUL8_OK;
}

int gemmul8_b200_grid_create_from_comm(void *nccl_comm, int P, int Q, size_t a_panel_bytes, size_t b_panel_bytes, int exchange,
                                       gemmul8_b200_grid **out) {
    if (!nccl_comm || !out) return fail(GEMMUL8_ERR_ARGUMENT, "null argument");
    auto *g = new gemmul8_b200_grid();
    g->world = static_cast<ncclComm_t>(nccl_comm);
    if (ncclCommCount(g->world, &g->nranks) != ncclSuccess || ncclCommUserRank(g->world, &g->rank) != ncclSuccess) {
        delete g;
        return fail(GEMMUL8_ERR_CUDA, "not a usable NCCL communicator");
    }
    int rc = check_shape(g->nranks, P, Q);
    if (rc) { delete g; return rc; }
    g->P = P; g->Q = Q; g->p = g->rank / Q; g->q = g->rank % Q; g->exchange = exchange;
    rc = finish_create(g, a_panel_bytes, b_panel_bytes);
    if (rc) { gemmul8_b200_grid_destroy(g); return rc; }
    *out = g;
    return GEMMUL8_OK;
}

int gemmul8_b200_grid_destroy(gemmul8_b200_grid *g) {
    if (!g) return GEMMUL8_OK;
    cudaDeviceSynchronize();
    if (g->world && g->flags && g->nranks > 1) {
        // collective: nobody unmaps or frees while a peer may still be pushing into its buffers (a one-word all-reduce as barrier)
        if (ncclAllReduce(g->flags + gemmul8_b200_grid::F_TOTAL - 1, g->flags + gemmul8_b200_grid::F_TOTAL - 1, 1, ncclInt32, ncclMax, g->world, g->xa) == ncclSuccess)
            cudaStreamSynchronize(g->xa);
    }
    for (int i = 0; i < kMaxSide; ++i) {
        if (g->peer_a[i] && g->peer_a[i] != g->a_buf) cudaIpcCloseMemHandle(g->peer_a[i]);
        if (g->peer_b[i] && g->peer_b[i] != g->b_buf) cudaIpcCloseMemHandle(g->peer_b[i]);
        if (g->peer_flags_row[i] && g->peer_flags_row[i] != g->flags) cudaIpcCloseMemHandle(g->peer_flags_row[i]);
        if (g->peer_flags_col[i] && g->peer_flags_col[i] != g->flags) cudaIpcCloseMemHandle(g->peer_flags_col[i]);
    }
    if (g->row) ncclCommDestroy(g->row);
    if (g->col) ncclCommDestroy(g->col);
    if (g->world && g->own_world) ncclCommDestroy(g->world);
    if (g->a_buf) cudaFree(g->a_buf);
    if (g->b_buf) cudaFree(g->b_buf);
    if (g->flags) cudaFree(g->flags);
    for (auto e : g->ev_a) if (e) cudaEventDestroy(e);
    for (auto e : g->ev_b) if (e) cudaEventDestroy(e);
    if (g->ev_start) cudaEventDestroy(g->ev_start);
    if (g->ev_xa_done) cudaEventDestroy(g->ev_xa_done);
    if (g->ev_xb_done) cudaEventDestroy(g->ev_xb_done);
    for (auto s : g->xpa) if (s) cudaStreamDestroy(s);
    for (auto s : g->xpb) if (s) cudaStreamDestroy(s);
    for (auto e : g->ev_pa) if (e) cudaEventDestroy(e);
    for (auto e : g->ev_pb) if (e) cudaEventDestroy(e);
    if (g->xa) cudaStreamDestroy(g->xa);
    if (g->xb) cudaStreamDestroy(g->xb);
    cudaGetLastError();
    delete g;
    return GEMMUL8_OK;
}

int gemmul8_b200_grid_coords(const gemmul8_b200_grid *g, int out[4]) {
    if (!g || !out) return fail(GEMMUL8_ERR_ARGUMENT, "null argument");
    out[0] = g->P; out[1] = g->Q; out[2] = g->p; out[3] = g->q;
    return GEMMUL8_OK;
}

size_t gemmul8_b200_pgemm_worksize(const gemmul8_b200_grid *g, size_t m, size_t n, size_t k, unsigned num_moduli) {
    if (!g || m % (size_t)g->P || n % (size_t)g->Q) return 0;
    return gemmul8_b200_worksize(m / (size_t)g->P, n / (size_t)g->Q, k, num_moduli, GEMMUL8_REAL_DEFAULT);
}

int gemmul8_b200_pgemm(gemmul8_b200_grid *g, gemmul8_b200_pargs *a) {
    if (!g || !a) return fail(GEMMUL8_ERR_ARGUMENT, "null argument");
    for (double &t : a->timers_ns) t = 0.0;
    const size_t P = (size_t)g->P, Q = (size_t)g->Q;
    if (a->m % P || a->n % (P * Q) || a->k % Q) return fail(GEMMUL8_ERR_ARGUMENT, "pgemm: m % P, n % (P Q) and k % Q must be 0");
    if (a->dtype_A > GEMMUL8_C64 || a->dtype_B > GEMMUL8_C64 || a->dtype_C > GEMMUL8_C64 || a->dtype_A < 0 || a->dtype_B < 0 || a->dtype_C < 0)
        return fail(GEMMUL8_ERR_ARGUMENT, "pgemm: bad dtype tag");
    const bool cplx = is_complex(a->dtype_C);
    if (cplx && !a->fastmode && g->nranks > 1)
        return fail(GEMMUL8_ERR_ARGUMENT, "pgemm: complex types are partitioned in fast mode only (the accurate-mode bound exchange is real-only)");
    const size_t m_loc = a->m / P, n_loc = a->n / Q, k = a->k, kq = k / Q, w = n_loc / P;
    const size_t esA = elem_size(a->dtype_A), esB = elem_size(a->dtype_B);
    if (m_loc == 0 || n_loc == 0) return GEMMUL8_OK;
    if (Q > 1 && m_loc * k * esA > g->a_cap) return fail(GEMMUL8_ERR_ARGUMENT, "pgemm: A panel exceeds the grid's panel buffer");
    if (P > 1 && k * n_loc * esB > g->b_cap) return fail(GEMMUL8_ERR_ARGUMENT, "pgemm: B panel exceeds the grid's panel buffer");
    if ((Q > 1 && a->lda < m_loc) || (P > 1 && a->ldb < k)) return fail(GEMMUL8_ERR_ARGUMENT, "pgemm: leading dimension too small");
    cudaStream_t st = static_cast<cudaStream_t>(a->stream);
    const bool copy = g->exchange == GEMMUL8_MP_EXCHANGE_COPY && g->nranks > 1;
    const bool pipelined = a->fastmode != 0 && k > 0 && !cplx;   // the block-wise entry is real-only; complex: whole panels, one call
    const uint32_t epoch = ++g->epoch;
    const bool own_b_in_place = copy && pipelined && P > 1;

    // full-problem argument block of this rank's C block (panels: grid buffers, or the caller's slices where nothing is exchanged)
    gemmul8_b200_args ga{};
    ga.op_A = GEMMUL8_OP_N; ga.op_B = GEMMUL8_OP_N;
    ga.m = m_loc; ga.n = n_loc; ga.k = k;
    ga.alpha = a->alpha; ga.beta = a->beta;
    ga.A = Q > 1 ? (const void *)g->a_buf : a->a_slice; ga.lda = Q > 1 ? m_loc : a->lda;
    ga.B = P > 1 ? (const void *)g->b_buf : a->b_slice; ga.ldb = P > 1 ? k : a->ldb;
    ga.C = a->c_block; ga.ldc = a->ldc;
    ga.num_moduli = a->num_moduli; ga.fastmode = a->fastmode; ga.work = a->work; ga.compute_type = cplx ? a->compute_type : GEMMUL8_REAL_DEFAULT;
    ga.dtype_A = a->dtype_A; ga.dtype_B = a->dtype_B; ga.dtype_C = a->dtype_C;
    ga.stream = a->stream; ga.flags = a->flags | (copy ? (unsigned)GEMMUL8_FLAG_EXCLUSIVE_SMS : 0u);

    // Row pieces of the A panel.  NCCL: four equal pieces (each gather is a kernel that competes with the product before it).
    // COPY: a small first piece (1/8 of the rows, whole tiles) whose products cover the arrival of everything else, then the
    // rest in one piece: fewer, larger scaling and product launches (measured: 4 equal pieces cost ~2 ms of extra scaling time).
    size_t rb[9];
    int npieces = 1;
    rb[0] = 0; rb[1] = m_loc;
    if (Q > 1 && pipelined) {
        if (copy) {
            const size_t first = ((m_loc / 8 + 255) / 256) * 256;
            if (first > 0 && first < m_loc) { rb[1] = first; rb[2] = m_loc; npieces = 2; }
        } else {
            npieces = row_pieces(m_loc, kMaxPieces, rb);
        }
    }

    // ---------------- exchange ----------------
    // The side streams start behind everything already queued on the compute stream: the caller's producers of a_slice /
    // b_slice, and the previous call's readers of the panel buffers.
    MP_CUDA(cudaEventRecord(g->ev_start, st), "event record");
    if (Q > 1) {
        MP_CUDA(cudaStreamWaitEvent(g->xa, g->ev_start, 0), "stream wait");
        const uint8_t *src = static_cast<const uint8_t *>(a->a_slice);
        // piece i is its own column-major rows x k matrix at byte offset r0 * k * es; my k/Q columns sit at column q * k/Q
        auto piece_off = [&](int i) { return (rb[i] * k + (size_t)g->q * kq * (rb[i + 1] - rb[i])) * esA; };
        if (copy) {
            // remote pushes: one stream per peer (right neighbour first), all pieces in order, a flag behind each
            for (size_t d = 1; d < Q; ++d) {
                const size_t qq = ((size_t)g->q + d) % Q;
                cudaStream_t xs = g->xpa[qq];
                MP_CUDA(cudaStreamWaitEvent(xs, g->ev_start, 0), "stream wait");
                if (epoch > 1) { int rc = wait_flag(g->flags, gemmul8_b200_grid::F_ACK_A + (int)qq, epoch - 1, xs); if (rc) return rc; }   // the peer has read what I pushed last time
                for (int i = 0; i < npieces; ++i) {
                    const size_t rows = rb[i + 1] - rb[i];
                    MP_CUDA(cudaMemcpy2DAsync(g->peer_a[qq] + piece_off(i), rows * esA, src + rb[i] * esA, a->lda * esA, rows * esA, kq, cudaMemcpyDeviceToDevice, xs), "push A piece");
                    int rc = raise_flag(g->peer_flags_row[qq], gemmul8_b200_grid::F_A + i * kMaxSide + g->q, epoch, xs);
                    if (rc) return rc;
                }
                MP_CUDA(cudaEventRecord(g->ev_pa[qq], xs), "event record");
            }
        }
        for (int i = 0; i < npieces; ++i) {
            const size_t rows = rb[i + 1] - rb[i];
            // my own part goes into my own buffer (both transports); NCCL then gathers the piece in place
            MP_CUDA(cudaMemcpy2DAsync(g->a_buf + piece_off(i), rows * esA, src + rb[i] * esA, a->lda * esA, rows * esA, kq, cudaMemcpyDeviceToDevice, g->xa), "own part of A piece");
            if (!copy) MP_NCCL(ncclAllGather(g->a_buf + piece_off(i), g->a_buf + rb[i] * k * esA, rows * kq * esA, ncclInt8, g->row, g->xa), "all-gather A piece");
            MP_CUDA(cudaEventRecord(g->ev_a[i], g->xa), "event record");
        }
    }
    if (P > 1) {
        MP_CUDA(cudaStreamWaitEvent(g->xb, g->ev_start, 0), "stream wait");
        const uint8_t *src = static_cast<const uint8_t *>(a->b_slice);
        const size_t off = (size_t)g->p * w * k * esB;       // my w columns of the k x n_loc panel
        if (copy) {
            for (size_t d = 1; d < P; ++d) {
                const size_t pp = ((size_t)g->p + d) % P;
                cudaStream_t xs = g->xpb[pp];
                MP_CUDA(cudaStreamWaitEvent(xs, g->ev_start, 0), "stream wait");
                if (epoch > 1) { int rc = wait_flag(g->flags, gemmul8_b200_grid::F_ACK_B + (int)pp, epoch - 1, xs); if (rc) return rc; }
                MP_CUDA(cudaMemcpy2DAsync(g->peer_b[pp] + off, k * esB, src, a->ldb * esB, k * esB, w, cudaMemcpyDeviceToDevice, xs), "push B block");
                int rc = raise_flag(g->peer_flags_col[pp], gemmul8_b200_grid::F_B + g->p, epoch, xs);
                if (rc) return rc;
                MP_CUDA(cudaEventRecord(g->ev_pb[pp], xs), "event record");
            }
        }
        // my own block: read in place from b_slice by the pipelined COPY path (no local copy: 2 x the block of HBM traffic
        // saved while the encoders compete with the incoming pushes); otherwise it joins the panel buffer
        if (!own_b_in_place)
            MP_CUDA(cudaMemcpy2DAsync(g->b_buf + off, k * esB, src, a->ldb * esB, k * esB, w, cudaMemcpyDeviceToDevice, g->xb), "own B block");
        if (copy) {
            if (!own_b_in_place) MP_CUDA(cudaEventRecord(g->ev_b[g->p], g->xb), "event record");
        } else {
            for (size_t pp = 0; pp < P; ++pp) {
                uint8_t *blk = g->b_buf + pp * w * k * esB;
                MP_NCCL(ncclBroadcast(blk, blk, w * k * esB, ncclInt8, (int)pp, g->col, g->xb), "broadcast B block");
                MP_CUDA(cudaEventRecord(g->ev_b[pp], g->xb), "event record");
            }
        }
    }

    // consumer side of one A piece / one B block: order the compute stream behind its arrival
    auto a_ready = [&](int i) -> int {
        if (Q == 1) return GEMMUL8_OK;
        MP_CUDA(cudaStreamWaitEvent(st, g->ev_a[i], 0), "stream wait");          // my own part (COPY) / the whole gather (NCCL)
        if (copy)
            for (size_t qq = 0; qq < Q; ++qq)
                if ((int)qq != g->q) { int rc = wait_flag(g->flags, gemmul8_b200_grid::F_A + i * kMaxSide + (int)qq, epoch, st); if (rc) return rc; }
        return GEMMUL8_OK;
    };
    auto b_ready = [&](size_t pp) -> int {
        if (P == 1 || (own_b_in_place && (int)pp == g->p)) return GEMMUL8_OK;
        if (copy && (int)pp != g->p) return wait_flag(g->flags, gemmul8_b200_grid::F_B + (int)pp, epoch, st);
        MP_CUDA(cudaStreamWaitEvent(st, g->ev_b[pp], 0), "stream wait");
        return GEMMUL8_OK;
    };
    // ... and tell the producers that their data has been consumed (COPY transport; raised behind the last reader)
    auto ack = [&]() -> int {
        if (!copy) return GEMMUL8_OK;
        for (size_t qq = 0; qq < Q; ++qq)
            if ((int)qq != g->q) { int rc = raise_flag(g->peer_flags_row[qq], gemmul8_b200_grid::F_ACK_A + g->q, epoch, st); if (rc) return rc; }
        for (size_t pp = 0; pp < P; ++pp)
            if ((int)pp != g->p) { int rc = raise_flag(g->peer_flags_col[pp], gemmul8_b200_grid::F_ACK_B + g->p, epoch, st); if (rc) return rc; }
        return GEMMUL8_OK;
    };
    auto check = [&](int rc, const char *what) -> int {
        if (rc != GEMMUL8_OK) return fail(rc, std::string(what) + ": " + gemmul8_b200_last_error());
        return GEMMUL8_OK;
    };

    int rc;
    if (pipelined) {
        // ---------------- fast mode: block-wise, products start while later pieces are in flight ----------------
        bool acked = false;
        int a_scaled = 0;
        size_t b_scaled = 0;
        auto maybe_ack = [&]() -> int {     // once every operand has been encoded the panel buffers are free: tell the producers
            if (acked || a_scaled < npieces || b_scaled < P) return GEMMUL8_OK;
            acked = true;
            return ack();
        };
        auto part = [&](gemmul8_b200_args &pa, int what, size_t r0, size_t r1, size_t c0, size_t c1, const char *name) -> int {
            int rc2 = check(gemmul8_b200_gemm_part(&pa, what, r0, r1, c0, c1), name);
            for (int t = 0; t < 4; ++t) a->timers_ns[t] += pa.timers_ns[t];
            return rc2;
        };
        // one column block of B: wait for it, encode it
        auto scale_b_block = [&](gemmul8_b200_args pa, size_t pp) -> int {
            int rc2;
            if ((rc2 = b_ready(pp))) return rc2;
            const size_t c0 = P > 1 ? pp * w : 0, c1 = P > 1 ? (pp + 1) * w : n_loc;
            if (own_b_in_place && (int)pp == g->p) {   // "column c of the panel" (c0 <= c < c1) is column c - c0 of b_slice
                pa.ldb = a->ldb;
                pa.B   = static_cast<const uint8_t *>(a->b_slice) - c0 * a->ldb * esB;
            }
            if ((rc2 = part(pa, GEMMUL8_PART_SCALE_B, 0, 0, c0, c1, "scale B block"))) return rc2;
            ++b_scaled;
            return maybe_ack();
        };
        for (int i = 0; i < npieces; ++i) {
            const size_t r0 = rb[i], r1 = rb[i + 1];
            gemmul8_b200_args pa = ga;
            if (Q > 1) {   // piece i read in place: "row r of the panel" (r0 <= r < r1) is row r - r0 of the piece
                pa.lda = r1 - r0;
                pa.A   = g->a_buf + r0 * k * esA - r0 * esA;
            }
            // my own B block needs no transfer: it is encoded while the first A piece is still on its way
            const bool own_b_first = i == 0 && own_b_in_place;
            if (own_b_first && (rc = scale_b_block(pa, (size_t)g->p))) return rc;
            if ((rc = a_ready(i))) return rc;
            if ((rc = part(pa, GEMMUL8_PART_SCALE_A, r0, r1, 0, 0, "scale A piece"))) return rc;
            ++a_scaled;
            if (i > 0) {
                if ((rc = maybe_ack())) return rc;
                if ((rc = part(pa, GEMMUL8_PART_PRODUCT, r0, r1, 0, n_loc, "product"))) return rc;
                continue;
            }
            // First piece: B arrives block by block, the own block first (it needs no transfer).  The product of this piece is
            // issued per B block as soon as that block is encoded, so the tensor cores start after one piece of A and 1/P of B
            // instead of after the whole B panel.  (Column blocks need not start on a tile boundary for SCALE_B, whose kernels
            // work per column; a product does.)
            const bool per_block = P > 1 && (w % 256) == 0;
            for (size_t d = 0; d < P; ++d) {
                const size_t pp = ((size_t)g->p + d) % P;
                if (!(own_b_first && d == 0) && (rc = scale_b_block(pa, pp))) return rc;
                const size_t c0 = P > 1 ? pp * w : 0, c1 = P > 1 ? (pp + 1) * w : n_loc;
                if (per_block && (rc = part(pa, GEMMUL8_PART_PRODUCT, r0, r1, c0, c1, "product"))) return rc;
            }
            if (!per_block && (rc = part(pa, GEMMUL8_PART_PRODUCT, r0, r1, 0, n_loc, "product"))) return rc;
        }
    } else {
        // ---------------- accurate mode (and k == 0): whole panels, then the call split at the bound product ----------------
        for (int i = 0; i < npieces; ++i) if ((rc = a_ready(i))) return rc;
        for (size_t pp = 0; pp < P; ++pp) if ((rc = b_ready(pp))) return rc;
        if (a->fastmode || k == 0 || g->nranks == 1) {
            if ((rc = check(gemmul8_b200_gemm(&ga), "gemm"))) return rc;
        } else {
            gemmul8_b200_args b1 = ga;
            b1.flags = (ga.flags & ~(unsigned)(GEMMUL8_FLAG_TIMERS | GEMMUL8_FLAG_PHASE_LOG)) | GEMMUL8_FLAG_ONLY_BOUND;
            if ((rc = check(gemmul8_b200_gemm(&b1), "bound product"))) return rc;
            gemmul8_b200_layout L;
            if ((rc = check(gemmul8_b200_work_layout(m_loc, n_loc, k, a->num_moduli, GEMMUL8_REAL_DEFAULT, &L), "layout"))) return rc;
            uint8_t *work = static_cast<uint8_t *>(a->work);
            int32_t *rowmax = reinterpret_cast<int32_t *>(work + L.off_A8i + L.sizeA);
            int32_t *colmax = reinterpret_cast<int32_t *>(work + L.off_B8i + L.sizeB);
            // the shift of a row of A needs the maximum of its bound-product row over ALL n columns: over the Q blocks of the
            // grid row; the shift of a column of B over the P blocks of its grid column
            if (Q > 1) MP_NCCL(ncclAllReduce(rowmax, rowmax, m_loc, ncclInt32, ncclMax, g->row, st), "all-reduce row maxima");
            if (P > 1) MP_NCCL(ncclAllReduce(colmax, colmax, n_loc, ncclInt32, ncclMax, g->col, st), "all-reduce column maxima");
            gemmul8_b200_args b2 = ga;
            b2.flags = ga.flags | GEMMUL8_FLAG_SKIP_BOUND;
            if ((rc = check(gemmul8_b200_gemm(&b2), "gemm after the bound exchange"))) return rc;
            memcpy(ga.timers_ns, b2.timers_ns, sizeof(ga.timers_ns));
        }
        memcpy(a->timers_ns, ga.timers_ns, sizeof(a->timers_ns));
        if ((rc = ack())) return rc;
    }
    // the exchange streams' work of this call is part of the call: a later operation on `st` is ordered behind it
    if (Q > 1) { MP_CUDA(cudaEventRecord(g->ev_xa_done, g->xa), "event record"); MP_CUDA(cudaStreamWaitEvent(st, g->ev_xa_done, 0), "stream wait"); }
    if (P > 1) { MP_CUDA(cudaEventRecord(g->ev_xb_done, g->xb), "event record"); MP_CUDA(cudaStreamWaitEvent(st, g->ev_xb_done, 0), "stream wait"); }
    if (copy) {   // ... including my pushes to the peers (the caller may overwrite a_slice / b_slice behind this call)
        for (size_t qq = 0; qq < Q; ++qq) if ((int)qq != g->q && Q > 1) MP_CUDA(cudaStreamWaitEvent(st, g->ev_pa[qq], 0), "stream wait");
        for (size_t pp = 0; pp < P; ++pp) if ((int)pp != g->p && P > 1) MP_CUDA(cudaStreamWaitEvent(st, g->ev_pb[pp], 0), "stream wait");
    }
    return GEMMUL8_OK;
}

}  // extern "C"
