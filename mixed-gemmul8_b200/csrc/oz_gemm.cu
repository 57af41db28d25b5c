// All-moduli int8 x int8 -> int32 GEMM for sm_100a with the mod-m_j reduction fused into the epilogue.
//
// Replaces, for every modulus j, the reference's pair
//     cublasGemmEx(OP_T, OP_N, m_pad, n, lda8i, A8i + j*sizeA, B8i + j*sizeB -> C32i)   GEMMul8/src/gemmul8.cu:265
//     conv_32i_2_8u(j, sizeC, C32i, C8u + j*sizeC)                                      GEMMul8/src/conv_32i_2_8u.hpp:7-71
// by ONE persistent kernel: the int32 product lives only in tensor memory; HBM sees the int8 slices
// coming in (TMA, 128-byte swizzle) and one uint8 residue per element and modulus going out.
//
// Two kernels share this file:
//   oz_gemm_pair_kernel    (default)  two CTAs of a TPC share a 256 x 256 tile (tcgen05.mma.cta_group::2); see the banner
//                                     in front of it.  EPI_RESIDUE with every combine mode.
//   oz_gemm_tcgen05_kernel            one CTA per SM, 384 threads, M 128 x N 256 tiles: the raw-int32 tap (EPI_INT32), the
//                                     accurate-mode bound product (EPI_ABSMAX), and EPI_RESIDUE where the pair kernel is
//                                     switched off (option gemm_pair = 0) or no placement table exists:
//     warp 0      TMA producer: 4-stage ring of {A tile 128 x 128 B, B tile 256 x 128 B}
//     warp 1      MMA issuer  : tcgen05.mma.cta_group::1.kind::i8, M=128, N=256, K=32 per instruction,
//                               accumulators double-buffered in TMEM (2 x 256 columns)
//     warp 2      TMEM allocator
//     warps 4-11  epilogue    : tcgen05.ld (lane == row of C), Barrett reduction mod m_j, 32 x 16 byte transpose through a
//                               padded smem scratch, 4-byte stores (4 rows of one column); warps 4-7 columns 0-127 of the
//                               tile, warps 8-11 columns 128-255; complex passes first combine with the stored residue
//                               (ResidueCombine), whose words were prefetched before the accumulator became ready
// A work item is (C tile, modulus), ordered (band of tiles, modulus, ...): the items in flight share one modulus and a
// compact block of panels.  (The single-kernel product + CRT lives in oz_gemm_crt.cu.)
//
// Both operands are K-major exactly as the reference lays them out (A8i[j][row][k], B8i[j][col][k],
// row stride lda8i), so a 3-D tensor map (k, row, modulus) serves all moduli and K / row tails are
// zero-filled by TMA.
#include "oz_tcgen05.cuh"

#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

namespace oz {
using namespace tc;
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int STAGES  = 4;
constexpr int SMEM_A  = BLOCK_M * BLOCK_K;
constexpr int SMEM_B  = BLOCK_N * BLOCK_K;
constexpr int SMEM_STAGE = SMEM_A + SMEM_B;
constexpr int SMEM_BARRIERS = 256;
constexpr int SMEM_SCRATCH  = 8 * 32 * 5 * 4;  // per epilogue warp: 32 rows x 5 words (16 residue bytes + 1 pad word)
constexpr int SMEM_TOTAL = STAGES * SMEM_STAGE + SMEM_BARRIERS + SMEM_SCRATCH + 1024;  // + slack for 1024-B alignment
constexpr int BAND_M = 16;  // row tiles per scheduling band
constexpr int NUM_THREADS = 384;
constexpr uint32_t TMEM_COLS = 512;

struct Sched {
    uint32_t tiles_m, tiles_n, slices, full_bands, per_band_full, total, band_m;
    __host__ __device__ void init(uint32_t tm, uint32_t tn, uint32_t s, uint32_t band = BAND_M) {
        tiles_m = tm; tiles_n = tn; slices = s; band_m = band;
        full_bands = tm / band_m;
        per_band_full = band_m * tn * s;
        total = tm * tn * s;
    }
    __device__ __forceinline__ void decode(uint32_t item, uint32_t &tm, uint32_t &tn, uint32_t &j) const {
        uint32_t band = item / per_band_full, rem, bm;
        if (band < full_bands) { rem = item - band * per_band_full; bm = band_m; }
        else { band = full_bands; rem = item - full_bands * per_band_full; bm = tiles_m - full_bands * band_m; }
        const uint32_t per_j = bm * tiles_n;
        j = rem / per_j;
        const uint32_t r2 = rem - j * per_j;
        tn = r2 / bm;
        tm = band * band_m + (r2 - tn * bm);
    }
};

struct KernelArgs {
    uint32_t rowsA, rowsB, num_kb, first_modulus;
    uint32_t num_slices;
    Sched sched;
    uint8_t *C8u; size_t ldc8u, sizeC; uint32_t rows_store;
    int combine; uint8_t *C8u_aux;
    int32_t *C32i; size_t ldc32i;
    int32_t *rowmax; int32_t *colmax;
};

// the it-th work item of this CTA; false when the CTA has run out of work
__device__ __forceinline__ bool next_work(const KernelArgs &a, uint32_t it, uint32_t &tm, uint32_t &tn, uint32_t &j) {
    const uint32_t unit = blockIdx.x + it * gridDim.x;
    if (unit >= a.sched.total) return false;
    a.sched.decode(unit, tm, tn, j);
    return true;
}

// canonical residue in [0, m) of a (possibly wrapped) int32; modulus index 0 is 256
__device__ __forceinline__ uint32_t reduce_mod(int32_t x, int32_t m, int32_t inv) {
    int32_t r = x - __mulhi(x, inv) * m;  // r in [-m, 2m)
    r -= (r >= m) ? m : 0;
    r += (r < 0) ? m : 0;
    return (uint32_t)r;
}

// r, stored residues in [0, m); see ResidueCombine.  Returns what goes to *out.
__device__ __forceinline__ uint32_t combine_residue(int rc, uint32_t r, const uint8_t *out, uint8_t *aux, int32_t m) {
    const int32_t old = (int32_t)*out, rn = (int32_t)r;
    int32_t t;
    if (rc == RC_ADD) t = old + rn;
    else if (rc == RC_SUB) t = old - rn;
    else if (rc == RC_RSUB) t = rn - old;
    else {  // RC_KARATSUBA_F
        int32_t u = old + rn;
        u -= (u >= m) ? m : 0;
        *aux = (uint8_t)u;
        t = old - rn;
    }
    t -= (t >= m) ? m : 0;
    t += (t < 0) ? m : 0;
    return (uint32_t)t;
}

// the same on four residues packed in a word (four rows of one column), two 16-bit lanes at a time:
// values stay below 2^10, so plain 32-bit adds never carry between lanes
__device__ __forceinline__ uint32_t fold_lanes(uint32_t t, uint32_t m, uint32_t k15) {   // lanes in [0, 2m) -> [0, m)
    const uint32_t ge = ((t + k15) >> 15) & 0x00010001u;   // k15 = 0x8000 - m per lane: bit 15 <=> lane >= m
    return t - ge * m;
}
__device__ __forceinline__ uint32_t combine_word(int rc, uint32_t rnew, uint32_t old, uint32_t &aux, int32_t mi) {
    const uint32_t m = (uint32_t)mi, ml = m * 0x00010001u, k15 = (0x8000u - m) * 0x00010001u;
    const uint32_t r0 = rnew & 0x00ff00ffu, r1 = (rnew >> 8) & 0x00ff00ffu;
    const uint32_t o0 = old & 0x00ff00ffu, o1 = (old >> 8) & 0x00ff00ffu;
    uint32_t t0, t1;
    aux = 0;
    if (rc == RC_ADD)       { t0 = o0 + r0;      t1 = o1 + r1; }
    else if (rc == RC_SUB)  { t0 = o0 + ml - r0; t1 = o1 + ml - r1; }
    else if (rc == RC_RSUB) { t0 = r0 + ml - o0; t1 = r1 + ml - o1; }
    else {                  // RC_KARATSUBA_F: aux = old + new, out = old - new
        aux = fold_lanes(o0 + r0, m, k15) | (fold_lanes(o1 + r1, m, k15) << 8);
        t0 = o0 + ml - r0;  t1 = o1 + ml - r1;
    }
    // (a - b + m lies in [1, 2m - 1] and m - 0 = m folds to 0: one fold is enough everywhere)
    return fold_lanes(t0, m, k15) | (fold_lanes(t1, m, k15) << 8);
}

template <int EPI, bool RMW = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
oz_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const KernelArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base  = smem_base + STAGES * SMEM_STAGE;
    // barrier slots (8 bytes each)
    auto full_bar  = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const uint32_t num_kb = args.num_kb;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, tm, tn, j;
            for (uint32_t it = 0; next_work(args, it, tm, tn, j); ++it) {
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * SMEM_STAGE;
                    mbar_expect_tx(full_bar(stage), SMEM_STAGE);
                    tma_load_3d(sa, &map_a, full_bar(stage), (int)(kb * BLOCK_K), (int)(tm * BLOCK_M), (int)j);
                    tma_load_3d(sa + SMEM_A, &map_b, full_bar(stage), (int)(kb * BLOCK_K), (int)(tn * BLOCK_N), (int)j);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
            uint32_t stage = 0, phase = 0, tm, tn, j;
            for (uint32_t it = 0; next_work(args, it, tm, tn, j); ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * SMEM_STAGE;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + SMEM_A);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
                        umma_i8(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | (uint32_t)k) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(tfull_bar(acc));  // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        uint32_t tm, tn, j;
        // transposed store mapping: this lane stores rows 4*rg .. 4*rg+3 of columns 4*cg .. 4*cg+3 of a 16-column chunk
        const int rg = lane & 7, cg = lane >> 3;
        uint32_t *scr = reinterpret_cast<uint32_t *>(smem_raw + (bar_base + SMEM_BARRIERS - smem_u32(smem_raw))) + (warp - 4) * 160;
        // column chunks (16 wide) of the tile this warp drains: all 16, or one half each for the two epilogue groups
        constexpr int NCH = 8;
        const int ch0     = warp >= 8 ? 8 : 0;
        for (uint32_t it = 0; next_work(args, it, tm, tn, j); ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const uint32_t row  = tm * BLOCK_M + q * 32 + lane;
            const uint32_t col0 = tn * BLOCK_N;
            const bool row_ok   = row < args.rowsA;
            const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16);

            // EPI_RESIDUE state that does not depend on the accumulator
            const uint32_t row4 = tm * BLOCK_M + q * 32 + 4 * rg;
            const bool rows4_ok = row4 < args.rows_store;   // rowsA rounded up to 4 (the residue stacks are padded to 4 rows)
            uint8_t *out4 = args.C8u + (size_t)j * args.sizeC + row4;
            uint8_t *aux4 = args.C8u_aux + (size_t)j * args.sizeC + row4;
            uint32_t old[RMW ? 4 * NCH : 1];
            if constexpr (RMW) {   // complex passes: fetch the stored residues while the MMAs are still running
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const uint32_t col = col0 + 16 * (ch0 + c) + 4 * cg + jj;
                        old[4 * c + jj] = (rows4_ok && col < args.rowsB) ? __ldcg(reinterpret_cast<const uint32_t *>(out4 + (size_t)col * args.ldc8u)) : 0u;
                    }
            }

            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();

            if constexpr (EPI == EPI_RESIDUE) {
                const uint32_t mj  = args.first_modulus + j;
                const int32_t m    = dev_tab::OZ_MOD[mj];
                const int32_t inv  = (int32_t)(4294967296ull / (uint32_t)m);
                const int rc = args.combine;
#pragma unroll(RMW ? NCH : 1)
                for (int c = 0; c < NCH; ++c) {
                    uint32_t v[16];
                    tmem_ld16(taddr + 16 * (ch0 + c), v);
                    tmem_ld_wait();
                    uint32_t r[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) r[e] = (mj == 0) ? (v[e] & 0xffu) : reduce_mod((int32_t)v[e], m, inv);
                    __syncwarp();   // the previous chunk's scratch reads are done
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        scr[5 * lane + w] = r[4 * w] | (r[4 * w + 1] << 8) | (r[4 * w + 2] << 16) | (r[4 * w + 3] << 24);
                    __syncwarp();
                    const uint32_t w0 = scr[5 * (4 * rg) + cg], w1 = scr[5 * (4 * rg + 1) + cg];
                    const uint32_t w2 = scr[5 * (4 * rg + 2) + cg], w3 = scr[5 * (4 * rg + 3) + cg];
                    // 4 x 4 byte transpose: o[jj] = residues of rows 4rg .. 4rg+3 in column 4cg + jj
                    const uint32_t t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                    const uint32_t t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                    uint32_t o[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632),
                                     __byte_perm(t2, t3, 0x5410), __byte_perm(t2, t3, 0x7632)};
                    if (rows4_ok) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const uint32_t col = col0 + 16 * (ch0 + c) + 4 * cg + jj;
                            if (col < args.rowsB) {
                                if constexpr (RMW) {
                                    uint32_t ax;
                                    o[jj] = combine_word(rc, o[jj], old[4 * c + jj], ax, m);
                                    if (rc == RC_KARATSUBA_F) *reinterpret_cast<uint32_t *>(aux4 + (size_t)col * args.ldc8u) = ax;
                                }
                                *reinterpret_cast<uint32_t *>(out4 + (size_t)col * args.ldc8u) = o[jj];
                            }
                        }
                    }
                }
            } else if constexpr (EPI == EPI_INT32) {
                int32_t *__restrict__ out = args.C32i + row;
#pragma unroll 1
                for (int c = 16 * ch0; c < 16 * (ch0 + NCH); c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const uint32_t col = col0 + c + e;
                            if (col < args.rowsB) out[(size_t)col * args.ldc32i] = (int32_t)v[e];
                        }
                    }
                }
            } else {  // EPI_ABSMAX
                int32_t rmax = 0;
#pragma unroll 1
                for (int c = 16 * ch0; c < 16 * (ch0 + NCH); c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const uint32_t col = col0 + c + e;
                        int32_t a = (row_ok && col < args.rowsB) ? abs((int32_t)v[e]) : 0;
                        rmax = max(rmax, a);
                        a = max(a, __shfl_xor_sync(0xffffffffu, a, 16));
                        a = max(a, __shfl_xor_sync(0xffffffffu, a, 8));
                        a = max(a, __shfl_xor_sync(0xffffffffu, a, 4));
                        a = max(a, __shfl_xor_sync(0xffffffffu, a, 2));
                        a = max(a, __shfl_xor_sync(0xffffffffu, a, 1));
                        if (lane == 0 && col < args.rowsB && a > 0) atomicMax(args.colmax + col, a);
                    }
                }
                if (row_ok && rmax > 0) atomicMax(args.rowmax + row, rmax);
            }
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(acc));  // the accumulator is free again: the MMA warp runs ahead

        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// =============================================================================================
// CTA-pair variant (tcgen05 cta_group::2) of the residue kernel: two CTAs on the SMs of one TPC share
// a 256 x 256 tile.  Each CTA keeps 128 rows of the "M-side" operand and HALF of the "N-side" tile (128 rows) in
// shared memory; the tensor cores of the pair exchange the halves, so per 256 x 256 x 128 block of
// MACs the pair pulls 64 KB through L2 instead of the 96 KB two independent CTAs need.  The kernel
// is power-bound on B200, and the saved traffic turns into clock -- PROVIDED the pairs are placed like
// the CTAs of a plain launch (see `placement_slots` and the claim table): DESIGN.md 3.2.
// The operand ROLES ARE SWAPPED with respect to C = A B: the M side walks rows of B8i (columns of C), the N side rows of
// A8i (rows of C), so that a lane of the epilogue holds consecutive ROWS of one column of the column-major residue matrix.
//   every CTA    warp 0 : TMA producer of its 128 M-side rows and its N-side half; the bytes are accounted on the
//                         LEADER's full barrier (cp.async.bulk.tensor .cta_group::2)
//   leader only  warp 1 : claims the pair's work slot, issues tcgen05.mma.cta_group::2 (M 256, N 256, K 32); its commits
//                         are multicast to the empty / accumulator-full barriers of both CTAs
//   every CTA    warp 2 : TMEM allocation (cta_group::2)
//   every CTA    warps 4-19 : epilogue of its own 128 columns of C: tcgen05.ld.x32, Barrett, pack, 256-bit stores (or TMA
//                         bulk stores, option tma_store); one lane per warp tells the leader that the buffer is drained
// =============================================================================================
#ifndef OZ_EPI_DRAIN_FIRST
#define OZ_EPI_DRAIN_FIRST 1   // A/B knob of tools/ab_build.sh; 0 = read and reduce the slab in two halves
#endif
#ifndef OZ_EPI_STORE256
#define OZ_EPI_STORE256 1      // A/B knob: 0 = two 128-bit stores per 32-byte run
#endif
#ifndef OZ_PAIR_STAGES_DEFAULT
#define OZ_PAIR_STAGES_DEFAULT 6
#endif
constexpr int PAIR_STAGES = OZ_PAIR_STAGES_DEFAULT;
constexpr int PAIR_SMEM_A = BLOCK_M * BLOCK_K;          // 128 rows of A
constexpr int PAIR_SMEM_B = 128 * BLOCK_K;              // 128 of the tile's 256 columns of B
constexpr int PAIR_STAGE  = PAIR_SMEM_A + PAIR_SMEM_B;  // 32 KiB per CTA and stage
constexpr int PAIR_EPI_WARPS = 16;                      // 4 per TMEM lane quarter, 64 of the tile's 256 columns each
constexpr int PAIR_THREADS   = (4 + PAIR_EPI_WARPS) * 32;
constexpr int PAIR_SCRATCH   = 0;                       // (the epilogue needs no shared memory any more)
constexpr int PAIR_STAGE_OUT = 4 * 2 * 4096;            // TMA-store epilogue: per lane quarter two boxes of 32 columns x 128 rows
constexpr int pair_smem_total(int stages, bool tma_store) { return stages * PAIR_STAGE + (tma_store ? PAIR_STAGE_OUT : 0) + SMEM_BARRIERS + PAIR_SCRATCH + 1024; }
constexpr int PAIR_BAND = 8;                            // 256-row tiles per scheduling band

struct PairSched {   // items = (256-row tile, column tile, modulus), ordered (band of 8 row tiles, modulus, column tile, row tile)
    uint32_t tiles_m, tiles_n, slices, full_bands, per_band_full, total, band_m;
    __host__ __device__ void init(uint32_t tm, uint32_t tn, uint32_t s, uint32_t band = PAIR_BAND) {
        tiles_m = tm; tiles_n = tn; slices = s; band_m = band;
        full_bands = tm / band_m;
        per_band_full = band_m * tn * s;
        total = tm * tn * s;
    }
    __device__ __forceinline__ void decode(uint32_t item, uint32_t &tm, uint32_t &tn, uint32_t &j) const {
        uint32_t band = item / per_band_full, rem, bm;
        if (band < full_bands) { rem = item - band * per_band_full; bm = band_m; }
        else { band = full_bands; rem = item - full_bands * per_band_full; bm = tiles_m - full_bands * band_m; }
        const uint32_t per_j = bm * tiles_n;
        j = rem / per_j;
        const uint32_t r2 = rem - j * per_j;
        tn = r2 / bm;
        tm = band * band_m + (r2 - tn * bm);
    }
};
struct PairArgs {
    uint32_t rowsA, rowsB, num_kb, first_modulus, rows_store;
    PairSched sched;
    uint8_t *C8u; size_t ldc8u, sizeC;
    int combine; uint8_t *C8u_aux;
    const uint32_t *slot;   // smid -> block index of a plain launch (nullptr: work goes by blockIdx)
    uint32_t *claims;       // with `slot`: one word per pair, zeroed before the launch
};

template <bool RMW, int NSTAGES, bool TMAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
oz_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_c, const PairArgs args) {
    static_assert(!(RMW && TMAST), "the TMA-store epilogue is for plain passes");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_out = smem_base + NSTAGES * PAIR_STAGE;                       // TMAST: 32 KiB of residue staging (1024-B aligned)
    const uint32_t bar_base  = stage_out + (TMAST ? PAIR_STAGE_OUT : 0);
    auto full_bar   = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar  = [&](int s) { return bar_base + 8u * (NSTAGES + s); };
    auto tfull_bar  = [&](int s) { return bar_base + 8u * (2 * NSTAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * NSTAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGES + 4);
    uint32_t *tmem_slot_ptr  = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank    = cluster_ctarank();         // 0 = leader
    const uint32_t npairs = gridDim.x >> 1;
    const uint32_t num_kb  = args.num_kb;
    const uint32_t total   = args.sched.total;
    // Work goes by PLACEMENT, not by block index: a cluster launch fills the SMs GPC by GPC, which would put all pairs
    // that share a B panel behind one GPC port (measured: +5 ms from that alone).  slot[smid] is the block index a plain
    // launch gives this SM (TPC by TPC, round-robin over the GPCs), so the sharers end up spread exactly as in the
    // single-CTA kernel.  The SM id is only a PREFERENCE: nothing guarantees one cluster per TPC (another kernel may hold an
    // SM, so that two clusters of this launch run on the same TPC one after the other), therefore every cluster CLAIMS its
    // slot in a per-launch table and takes the next free one if the preferred slot is gone.  npairs clusters, npairs
    // slots: each slot is worked exactly once whatever the placement.
    const uint32_t pair_slot = bar_base + 8u * (2 * NSTAGES + 5);
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t pair = claim_pair_slot(args.slot, args.claims, npairs, blockIdx.x >> 1);
        asm volatile("st.shared::cta.u32 [%0], %1;" ::"r"(pair_slot), "r"(pair) : "memory");   // read by both CTAs after the cluster barrier
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        if (TMAST) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * PAIR_EPI_WARPS); }   // epilogue warps of both CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();          // barriers of both CTAs are initialised before anybody signals across the pair
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    uint32_t pair;               // the leader's claim (its CTA stays resident until the cluster barrier at the end)
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(pair) : "r"(map_to_cta(pair_slot, 0)) : "memory");

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, tm, tn, j;
            for (uint32_t item = pair; item < total; item += npairs) {
                args.sched.decode(item, tm, tn, j);      // tm: 256-column tile of C (MMA M side), tn: 256-row tile of C (MMA N side)
                const int rowA = (int)(tm * 256 + rank * 128), rowB = (int)(tn * BLOCK_N + rank * 128);   // map_a = B8i, map_b = A8i
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa   = smem_base + stage * PAIR_STAGE;
                    const uint32_t lbar = map_to_cta(full_bar(stage), 0);
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * PAIR_STAGE);   // both CTAs' bytes
                    tma_load_3d_pair(sa, &map_a, lbar, (int)(kb * BLOCK_K), rowA, (int)j);
                    tma_load_3d_pair(sa + PAIR_SMEM_A, &map_b, lbar, (int)(kb * BLOCK_K), rowB, (int)j);
                    if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(256, BLOCK_N);
            uint32_t stage = 0, phase = 0, it = 0;
            for (uint32_t item = pair; item < total; item += npairs, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * PAIR_STAGE;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + PAIR_SMEM_A);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_i8_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | (uint32_t)k) != 0 ? 1u : 0u);
                    tcgen05_commit_pair(empty_bar(stage));   // frees this stage in both CTAs
                    if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit_pair(tfull_bar(acc));         // accumulator complete, both CTAs
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 COLUMNS of C) =====================
        // The operand roles are swapped with respect to C = A B: the "M side" of the MMA (tensor-memory lanes, split over
        // the two CTAs) holds rows of B8i, i.e. COLUMNS of C, and the "N side" (tensor-memory columns) rows of A8i, i.e.
        // ROWS of C.  A thread (lane == one column of C) therefore reads 32 consecutive rows of its column with one
        // tcgen05.ld -- exactly 32 consecutive bytes of the column-major residue matrix: Barrett, pack, two 16-byte stores.
        // No transpose through shared memory, no __syncwarp (the 9.4 instructions per residue of the row-major version
        // were what bounded the kernel below k ~ 2048; this is ~5).
        // 16 warps: warp & 3 = TMEM lane quarter (32 columns of C), (warp - 4) >> 2 = 64-row group, 2 chunks of 32 rows.
        const int q = warp & 3, rowgroup = (warp - 4) >> 2;
        constexpr int NCH = 2;
        const size_t ld   = args.ldc8u;
        const bool vec_ok = (ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(args.C8u) | (uintptr_t)args.sizeC) % 16 == 0) &&
                            ((reinterpret_cast<uintptr_t>(args.C8u_aux)) % 16 == 0);
        // 32 contiguous bytes per lane as ONE 256-bit store where the addresses allow it: a warp's store touches 32 different
        // lines (lane == column), so every store instruction costs 32 LSU wavefronts whatever its width
        const bool vec32_ok = vec_ok && OZ_EPI_STORE256 && (ld % 32 == 0) && ((reinterpret_cast<uintptr_t>(args.C8u) | (uintptr_t)args.sizeC) % 32 == 0);
        uint32_t tu, tv, j, it = 0;
        for (uint32_t item = pair; item < total; item += npairs, ++it) {
            args.sched.decode(item, tu, tv, j);
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const uint32_t col  = tu * 256 + rank * 128 + q * 32 + lane;           // this lane's column of C
            const uint32_t row0 = tv * BLOCK_N + 64 * rowgroup;                    // first of this warp's 64 rows of C
            const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16) + 64 * rowgroup;
            const bool col_ok   = col < args.rowsB;
            const bool interior = vec_ok && (row0 + 64 <= args.rows_store);
            uint8_t *out = args.C8u + (size_t)j * args.sizeC + (size_t)col * ld + row0;
            uint8_t *aux = args.C8u_aux + (size_t)j * args.sizeC + (size_t)col * ld + row0;
            uint32_t old[RMW ? 8 * NCH : 1];
            if constexpr (RMW) {   // complex passes: fetch the stored residues while the MMAs are still running
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (interior && col_ok) {
                        const uint4 a0 = __ldcg(reinterpret_cast<const uint4 *>(out + 32 * c)), a1 = __ldcg(reinterpret_cast<const uint4 *>(out + 32 * c + 16));
                        old[8 * c + 0] = a0.x; old[8 * c + 1] = a0.y; old[8 * c + 2] = a0.z; old[8 * c + 3] = a0.w;
                        old[8 * c + 4] = a1.x; old[8 * c + 5] = a1.y; old[8 * c + 6] = a1.z; old[8 * c + 7] = a1.w;
                    } else {
#pragma unroll
                        for (int w = 0; w < 8; ++w)
                            old[8 * c + w] = (col_ok && row0 + 32 * c + 4 * w < args.rows_store) ? __ldcg(reinterpret_cast<const uint32_t *>(out + 32 * c + 4 * w)) : 0u;
                    }
                }
            }
            const uint32_t mj = args.first_modulus + j;
            const uint32_t m  = (uint32_t)dev_tab::OZ_MOD[mj];
            const uint32_t inv = dev_tab::OZ_BARRETT_INV[mj], negm = dev_tab::OZ_BARRETT_NEGM[mj];
            const uint32_t off = dev_tab::OZ_BARRETT_OFF[mj];
            const int rc       = args.combine;
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            // Plain passes read the whole 64-row slab into registers at once and hand the accumulator buffer back BEFORE any
            // arithmetic: with short k the MMA of item i + 2 waits for exactly this signal (two buffers), so every cycle
            // between "accumulator complete" and "drained" is a cycle the tensor pipe may idle.  The combine passes carry
            // the stored residues in registers as well and read the slab in two halves.
            constexpr bool DRAIN_FIRST = !RMW && OZ_EPI_DRAIN_FIRST;
            uint32_t v2[DRAIN_FIRST ? NCH : 1][32];
            if constexpr (DRAIN_FIRST) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld32(taddr + 32 * c, v2[c]);
                tmem_ld_wait(v2[0], v2[1]);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(tempty_bar(acc), 0));
            }
            if constexpr (TMAST) {
                // Residues leave through the TMA: each lane quarter (4 warps = 32 columns x 256 rows of C) stages its 8 KiB in
                // shared memory in the 128-byte-swizzled layout of two 32 x 128 boxes and ONE thread issues two bulk tensor
                // stores.  A warp's direct store spans 32 lines (lane == column) and costs the LSU 32 wavefronts; here the
                // LSU sees conflict-free 16-byte shared stores, the TMA writes whole 128-byte lines and clips at the edges
                // of the matrix, so there is no bounds logic at all.
                uint32_t pk[NCH][8];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (mj != 0) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v2[c][e] = reduce_mod_u((int32_t)v2[c][e], negm, inv, off);
                    }
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        pk[c][w] = __byte_perm(__byte_perm(v2[c][4 * w], v2[c][4 * w + 1], 0x0040), __byte_perm(v2[c][4 * w + 2], v2[c][4 * w + 3], 0x0040), 0x5410);
                }
                const bool issuer = rowgroup == 0 && lane == 0;
                if (issuer) tma_store_wait_read();            // the previous item's boxes have been read out of shared memory
                named_barrier(1 + q, 128);
                // box h = rowgroup >> 1 (rows 0-127 / 128-255 of the tile); inside a box row (= one column of C, 128 bytes) this
                // warp owns bytes 64 (rowgroup & 1) .. + 63; 16-byte chunk i of row c sits at chunk i ^ (c & 7)
                const uint32_t box = stage_out + q * 8192 + (rowgroup >> 1) * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t i0 = (rowgroup & 1) * 4 + 2 * c;
                    st_shared_v4(box + (((i0 + 0) ^ (lane & 7)) << 4), pk[c][0], pk[c][1], pk[c][2], pk[c][3]);
                    st_shared_v4(box + (((i0 + 1) ^ (lane & 7)) << 4), pk[c][4], pk[c][5], pk[c][6], pk[c][7]);
                }
                fence_proxy_async_smem();
                named_barrier(1 + q, 128);
                if (issuer) {
                    const int ccol = (int)(tu * 256 + rank * 128 + q * 32), crow = (int)(tv * BLOCK_N);
                    tma_store_3d(&map_c, stage_out + q * 8192, crow, ccol, (int)j);
                    tma_store_3d(&map_c, stage_out + q * 8192 + 4096, crow + 128, ccol, (int)j);
                    tma_store_commit();
                }
            } else {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t vr[DRAIN_FIRST ? 1 : 32];
                if constexpr (!DRAIN_FIRST) {
                    tmem_ld32(taddr + 32 * c, vr);
                    tmem_ld_wait(vr);
                    if (c == NCH - 1) {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(tempty_bar(acc), 0));
                    }
                }
                uint32_t (&v)[32] = *reinterpret_cast<uint32_t (*)[32]>(DRAIN_FIRST ? v2[DRAIN_FIRST ? c : 0] : vr);
                if (mj != 0) {          // (modulus 256: the low byte, which the packing below takes anyway)
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = reduce_mod_u((int32_t)v[e], negm, inv, off);
                }
                uint32_t pk[8];
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    pk[w] = __byte_perm(__byte_perm(v[4 * w], v[4 * w + 1], 0x0040), __byte_perm(v[4 * w + 2], v[4 * w + 3], 0x0040), 0x5410);
                if (!col_ok) continue;
                uint8_t *po = out + 32 * c, *pa = aux + 32 * c;
                if constexpr (RMW) {
                    uint32_t ax[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) pk[w] = combine_word(rc, pk[w], old[8 * c + w], ax[w], (int32_t)m);
                    if (rc == RC_KARATSUBA_F) {
                        if (interior) {
                            *reinterpret_cast<uint4 *>(pa)      = make_uint4(ax[0], ax[1], ax[2], ax[3]);
                            *reinterpret_cast<uint4 *>(pa + 16) = make_uint4(ax[4], ax[5], ax[6], ax[7]);
                        } else {
#pragma unroll
                            for (int w = 0; w < 8; ++w)
                                if (row0 + 32 * c + 4 * w < args.rows_store) *reinterpret_cast<uint32_t *>(pa + 4 * w) = ax[w];
                        }
                    }
                }
                if (interior && vec32_ok) {
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(po), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                                 "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                } else if (interior) {
                    *reinterpret_cast<uint4 *>(po)      = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4 *>(po + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                } else {
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (row0 + 32 * c + 4 * w < args.rows_store) *reinterpret_cast<uint32_t *>(po + 4 * w) = pk[w];
                }
            }
            }   // !TMAST
        }
    }

    if (TMAST && warp >= 4 && ((warp - 4) >> 2) == 0 && lane == 0) tma_store_wait_all();   // this thread's bulk stores have landed
    tcgen05_fence_before();
    cluster_sync_all();          // nobody frees tensor memory while the peer may still use the pair's resources
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// CUDA-core cross-check (debug flag GEMMUL8_FLAG_GEMM_SIMT): one thread per C element, dp4a.
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void oz_gemm_simt_kernel(const int8_t *__restrict__ A8i, const int8_t *__restrict__ B8i, size_t sizeA,
                                    size_t sizeB, size_t ld8i, KernelArgs args) {
    const uint32_t row = blockIdx.x * 32 + threadIdx.x;
    const uint32_t col = blockIdx.y * 8 + threadIdx.y;
    const uint32_t j   = blockIdx.z;
    if (row >= args.rowsA || col >= args.rowsB) return;
    const int *a = reinterpret_cast<const int *>(A8i + (size_t)j * sizeA + (size_t)row * ld8i);
    const int *b = reinterpret_cast<const int *>(B8i + (size_t)j * sizeB + (size_t)col * ld8i);
    int acc = 0;
    for (size_t k = 0; k < ld8i / 4; ++k) acc = __dp4a(a[k], b[k], acc);
    if constexpr (EPI == EPI_RESIDUE) {
        const uint32_t mj = args.first_modulus + j;
        const int32_t m   = dev_tab::OZ_MOD[mj];
        const int32_t inv = (int32_t)(4294967296ull / (uint32_t)m);
        uint8_t *out = args.C8u + (size_t)j * args.sizeC + (size_t)col * args.ldc8u + row;
        uint8_t *aux = args.C8u_aux + (size_t)j * args.sizeC + (size_t)col * args.ldc8u + row;
        uint32_t r   = (mj == 0) ? ((uint32_t)acc & 0xffu) : reduce_mod(acc, m, inv);
        if (args.combine != RC_STORE) r = combine_residue(args.combine, r, out, aux, m);
        *out = (uint8_t)r;
    } else if constexpr (EPI == EPI_INT32) {
        args.C32i[(size_t)col * args.ldc32i + row] = acc;
    } else {
        const int a_ = abs(acc);
        if (a_ > 0) { atomicMax(args.rowmax + row, a_); atomicMax(args.colmax + col, a_); }
    }
}

}  // namespace

namespace detail {
namespace {
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static const PFN_cuTensorMapEncodeTiled_v12000 fn = [] {   // initialised once, thread-safe
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
        return (PFN_cuTensorMapEncodeTiled_v12000) nullptr;
    }();
    return fn;
}
}  // namespace

// (k, row, slice) view of a stack of K-major int8 slices; box = 128 bytes of k x box_rows rows
bool make_operand_map(CUtensorMap *map, const int8_t *base, size_t ld8i, size_t rows, size_t slices, size_t slice_stride,
                      uint32_t box_rows) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[3]    = {(cuuint64_t)ld8i, (cuuint64_t)rows, (cuuint64_t)slices};
    cuuint64_t strides[2] = {(cuuint64_t)ld8i, (cuuint64_t)slice_stride};
    cuuint32_t box[3]     = {(cuuint32_t)BLOCK_K, box_rows, 1};
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
bool make_residue_map(CUtensorMap *map, const uint8_t *base, size_t ld, size_t rows, size_t cols, size_t slices, size_t slice_stride) {
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[3]    = {(cuuint64_t)rows, (cuuint64_t)cols, (cuuint64_t)slices};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)slice_stride};
    cuuint32_t box[3]     = {128, 32, 1};
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
}  // namespace detail

namespace {
using detail::make_operand_map;

KernelArgs make_args(const GemmProblem &p) {
    KernelArgs a{};
    a.rowsA = (uint32_t)p.rowsA; a.rowsB = (uint32_t)p.rowsB;
    a.num_kb = (uint32_t)((p.ld8i + BLOCK_K - 1) / BLOCK_K);
    a.first_modulus = p.first_modulus;
    a.num_slices = p.num_slices;
    const int band = tuning().band;
    a.sched.init((uint32_t)((p.rowsA + BLOCK_M - 1) / BLOCK_M), (uint32_t)((p.rowsB + BLOCK_N - 1) / BLOCK_N),
                 p.num_slices, band > 0 ? (uint32_t)band : (uint32_t)BAND_M);
    a.C8u = p.C8u; a.ldc8u = p.ldc8u; a.sizeC = p.sizeC;
    a.rows_store = (uint32_t)((p.rowsA + 3) / 4 * 4);   // a sub-block launch must not touch the rows below it
    a.combine = p.combine; a.C8u_aux = p.C8u_aux ? p.C8u_aux : p.C8u;
    a.C32i = p.C32i; a.ldc32i = p.ldc32i;
    a.rowmax = p.rowmax; a.colmax = p.colmax;
    return a;
}

}  // namespace

namespace detail {
constexpr int kMaxDevices = 64;
int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}
int sm_count() {   // of the current device (cached: an attribute query per launch is measurable at 60 us per call)
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return 0;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
}
}  // namespace detail

namespace {
using detail::sm_count;
using detail::kMaxDevices;
using detail::current_device;

template <int EPI, bool RMW = false>
cudaError_t launch_tc(const GemmProblem &p, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (!make_operand_map(&ma, p.A8i, p.ld8i, p.rowsA, p.num_slices, p.sizeA, BLOCK_M)) return cudaErrorInvalidValue;
    if (!make_operand_map(&mb, p.B8i, p.ld8i, p.rowsB, p.num_slices, p.sizeB, BLOCK_N)) return cudaErrorInvalidValue;
    KernelArgs a = make_args(p);
    auto kern = oz_gemm_tcgen05_kernel<EPI, RMW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    const uint32_t grid = a.sched.total < (uint32_t)sm_count() ? a.sched.total : (uint32_t)sm_count();
    kern<<<grid, NUM_THREADS, SMEM_TOTAL, st>>>(ma, mb, a);
    count_launch();
    return cudaGetLastError();
}

template <int EPI>
cudaError_t launch_simt_t(const GemmProblem &p, cudaStream_t st) {
    KernelArgs a = make_args(p);
    dim3 block(32, 8), grid((unsigned)((p.rowsA + 31) / 32), (unsigned)((p.rowsB + 7) / 8), p.num_slices);
    oz_gemm_simt_kernel<EPI><<<grid, block, 0, st>>>(p.A8i, p.B8i, p.sizeA, p.sizeB, p.ld8i, a);
    count_launch();
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------
// placement probe: which block index does a plain 1-CTA-per-SM launch give each SM?
// Runs once per device, on a private stream, from gemmul8_b200_init() or from the first gemm call that is not being
// captured into a graph; never on the caller's stream, never with a device-wide synchronisation.
// ---------------------------------------------------------------------------------------------
__global__ void placement_probe_kernel(uint32_t *slot) {
    extern __shared__ uint8_t probe_smem[];
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) slot[smid] = blockIdx.x;
    probe_smem[threadIdx.x] = 1;       // the big dynamic allocation keeps it at one CTA per SM
    __nanosleep(200000);               // ... and every CTA resident at the same time
}
const uint32_t *probe_placement(int n) {
    uint32_t *d = nullptr;
    cudaStream_t ps = nullptr;
    if (cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMalloc(&d, sizeof(uint32_t) * (size_t)n) != cudaSuccess) { cudaGetLastError(); cudaStreamDestroy(ps); return nullptr; }
    std::vector<uint32_t> h((size_t)n), seen((size_t)n, 0);
    bool ok = cudaMemsetAsync(d, 0xff, sizeof(uint32_t) * (size_t)n, ps) == cudaSuccess &&
              cudaFuncSetAttribute(placement_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
    if (ok) {
        placement_probe_kernel<<<n, 32, 200 * 1024, ps>>>(d);
        ok = cudaMemcpyAsync(h.data(), d, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ps) == cudaSuccess &&
             cudaStreamSynchronize(ps) == cudaSuccess;
    }
    cudaStreamDestroy(ps);
    // a permutation, with TPC siblings holding consecutive blocks -- anything else (busy GPU, MIG slice, ...): no table
    for (int i = 0; ok && i < n; ++i) { if (h[i] >= (uint32_t)n || seen[h[i]]++) ok = false; }
    for (int i = 0; ok && i + 1 < n; i += 2) ok = (h[i] >> 1) == (h[i + 1] >> 1);
    if (!ok) { cudaGetLastError(); cudaFree(d); return nullptr; }
    return d;                           // n words, kept for the life of the process
}
struct Placement { std::mutex mu; std::atomic<int> state{0}; const uint32_t *table = nullptr; };   // state: 0 unknown, 1 probed
Placement g_placement[kMaxDevices];
}  // namespace

const uint32_t *detail::placement_slots(bool allow_probe) {
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return nullptr;
    Placement &P = g_placement[dev];
    if (P.state.load(std::memory_order_acquire) == 1) return P.table;
    if (!allow_probe) return nullptr;
    std::lock_guard<std::mutex> lock(P.mu);
    if (P.state.load(std::memory_order_acquire) == 0) {
        const int n = sm_count();
        P.table = (n >= 2 && !(n & 1)) ? probe_placement(n) : nullptr;
        P.state.store(1, std::memory_order_release);
    }
    return P.table;
}

namespace {
using detail::placement_slots;

// CTA-pair kernel for EPI_RESIDUE (all combine modes).  Default whenever the placement table exists (whole GPU, even SM
// count); option gemm_pair = 0 selects the single-CTA kernel, 1 forces the pair kernel even without the table.
// Measured at 16384^3, 14 moduli, same box: single-CTA 46.7 / 48.7 ms, pairs placed by block index 53.9 - 57.2 ms,
// pairs placed like a plain launch 44.7 / 45.2 ms (profiles/r01_pair_kernel_notes.md).
bool pair_kernel_enabled(cudaStream_t st) {
    const int mode = tuning().gemm_pair;
    if (mode == 0) return false;
    if (sm_count() < 2 || (sm_count() & 1)) return false;
    if (mode == 1) return true;
    return placement_slots(!detail::stream_is_capturing(st)) != nullptr;
}
template <bool RMW>
cudaError_t launch_pair(const GemmProblem &p, cudaStream_t st) {
    CUtensorMap ma, mb;
    // swapped roles (see the epilogue): the MMA's M side walks rows of B8i = columns of C, its N side rows of A8i = rows of C
    if (!make_operand_map(&ma, p.B8i, p.ld8i, p.rowsB, p.num_slices, p.sizeB, 128)) return cudaErrorInvalidValue;
    if (!make_operand_map(&mb, p.A8i, p.ld8i, p.rowsA, p.num_slices, p.sizeA, 128)) return cudaErrorInvalidValue;
    const Tuning tn = tuning();
    PairArgs a{};
    a.rowsA = (uint32_t)p.rowsA; a.rowsB = (uint32_t)p.rowsB;
    a.num_kb = (uint32_t)((p.ld8i + BLOCK_K - 1) / BLOCK_K);
    a.first_modulus = p.first_modulus;
    a.rows_store = (uint32_t)((p.rowsA + 3) / 4 * 4);
    // Band: the column tiles of C whose B8i rows, for one modulus, the pairs walk together.  Its slice rows (band x 256 x ld8i
    // bytes) should stay in L2 while all row tiles pass: ~32 MB measured best from k = 512 to 16384 (band 64 ... 8 at
    // 16384^2 x k; profiles/r02_ab_band.jsonl).  Option pair_band > 0 overrides.
    uint32_t band = (uint32_t)tn.pair_band;
    if (tn.pair_band <= 0) {
        const size_t fit = ((size_t)32 << 20) / (256 * p.ld8i);
        band = (uint32_t)(fit < 4 ? 4 : fit > 4096 ? 4096 : fit);
    }
    a.sched.init((uint32_t)((p.rowsB + 255) / 256), (uint32_t)((p.rowsA + BLOCK_N - 1) / BLOCK_N), p.num_slices, band);
    a.C8u = p.C8u; a.ldc8u = p.ldc8u; a.sizeC = p.sizeC;
    a.combine = p.combine; a.C8u_aux = p.C8u_aux ? p.C8u_aux : p.C8u;
    int stages = tn.pair_stages ? tn.pair_stages : p.share_sm ? 4 : PAIR_STAGES;   // (7 fit too: no measurable difference)
    if (stages != 4 && stages != 5) stages = 6;
    // TMA-store epilogue (option tma_store, default off): plain passes whose residue matrix has 16-byte-multiple strides (TMA's
    // requirement).  Measured equal to the 256-bit direct stores from k = 512 to 16384 (profiles/r02_ab_tma_store.jsonl: the
    // stores are not what bounds the short-k kernel), so the path with fewer moving parts stays the default.
    CUtensorMap mc{};
    bool tma_store = !RMW && tn.tma_store != 0 && (p.ldc8u % 16) == 0 && (p.sizeC % 16) == 0 && (reinterpret_cast<uintptr_t>(p.C8u) % 16) == 0;
    if (tma_store) tma_store = detail::make_residue_map(&mc, p.C8u, p.ldc8u, a.rows_store, p.rowsB, p.num_slices, p.sizeC);
    void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, PairArgs);
    if constexpr (RMW) {
        kern = stages == 4 ? oz_gemm_pair_kernel<true, 4, false> : stages == 5 ? oz_gemm_pair_kernel<true, 5, false> : oz_gemm_pair_kernel<true, 6, false>;
    } else if (tma_store) {
        kern = stages == 4 ? oz_gemm_pair_kernel<false, 4, true> : stages == 5 ? oz_gemm_pair_kernel<false, 5, true> : oz_gemm_pair_kernel<false, 6, true>;
    } else {
        kern = stages == 4 ? oz_gemm_pair_kernel<false, 4, false> : stages == 5 ? oz_gemm_pair_kernel<false, 5, false> : oz_gemm_pair_kernel<false, 6, false>;
    }
    const int smem_bytes = pair_smem_total(stages, tma_store);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    const uint32_t max_pairs = (uint32_t)sm_count() / 2;
    const uint32_t pairs = a.sched.total < max_pairs ? a.sched.total : max_pairs;
    // work by placement needs one cluster per TPC slot AND the per-launch claim table; without either, by block index
    static_assert(kClaimBytes >= sizeof(uint32_t) * 128, "claim table: one word per pair");
    a.slot = (pairs == max_pairs && p.claims != nullptr && max_pairs <= kClaimBytes / sizeof(uint32_t)) ? placement_slots(false) : nullptr;
    a.claims = p.claims;
    if (a.slot != nullptr && !p.claims_zeroed) {
        e = cudaMemsetAsync(p.claims, 0, sizeof(uint32_t) * max_pairs, st);
        if (e != cudaSuccess) return e;
    }
    kern<<<2 * pairs, PAIR_THREADS, smem_bytes, st>>>(ma, mb, mc, a);   // cluster dims (2,1,1) are a kernel attribute
    count_launch();
    return cudaGetLastError();
}

}  // namespace

void gemm_prepare_device(bool allow_probe) {
    if (tuning().gemm_pair != 0) detail::placement_slots(allow_probe);
}

cudaError_t launch_gemm_tcgen05(const GemmProblem &p, GemmEpilogue epi, cudaStream_t st) {
    if (p.rowsA == 0 || p.rowsB == 0 || p.num_slices == 0) return cudaSuccess;
    switch (epi) {
        case EPI_RESIDUE:
            if (pair_kernel_enabled(st)) return p.combine == RC_STORE ? launch_pair<false>(p, st) : launch_pair<true>(p, st);
            return p.combine == RC_STORE ? launch_tc<EPI_RESIDUE>(p, st) : launch_tc<EPI_RESIDUE, true>(p, st);
        case EPI_INT32:   return launch_tc<EPI_INT32>(p, st);
        case EPI_ABSMAX:  return launch_tc<EPI_ABSMAX>(p, st);
        case EPI_CRT:     return launch_gemm_crt(p, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_simt(const GemmProblem &p, GemmEpilogue epi, cudaStream_t st) {
    if (p.rowsA == 0 || p.rowsB == 0 || p.num_slices == 0) return cudaSuccess;
    switch (epi) {
        case EPI_RESIDUE: return launch_simt_t<EPI_RESIDUE>(p, st);
        case EPI_INT32:   return launch_simt_t<EPI_INT32>(p, st);
        case EPI_ABSMAX:  return launch_simt_t<EPI_ABSMAX>(p, st);
        case EPI_CRT:     return cudaErrorInvalidValue;  // the cross-check path runs EPI_RESIDUE + the stand-alone CRT kernel
    }
    return cudaErrorInvalidValue;
}

}  // namespace oz