// Product, residues, CRT, inverse scaling and alpha / beta in ONE kernel (north_star item 4): the per-modulus int32
// products live in tensor memory, the per-element CRT accumulators live in REGISTERS across the walk over the moduli,
// and HBM sees the int8 slices coming in and C going out -- no uint8 residue matrix, no second pass.
//
// Replaces, for every modulus j, the reference's
//     cublasGemmEx(... A8i_j, B8i_j -> C32i)                  GEMMul8/src/gemmul8.cu:265
//     conv_32i_2_8u(j, C32i -> C8u_j)                         GEMMul8/src/conv_32i_2_8u.hpp:25-56
// and then inverse_scaling(C8u_0..N-1 -> C)                   GEMMul8/src/inverse_scaling.hpp:35-62, :140-172, :823-1063
// with the same arithmetic in the same order (residue in [0, m_j), s1 = fma(w_hi[j], r, s1), s2 = fma(w_lo[j], r, s2) for
// j = 0 .. N-1, then the reduction mod M and the power-of-two scaling), so C is bit-identical to the two-kernel path.
//
// Why this shape.  The CRT state is 16 bytes per element of C (two FP64 accumulators; 8 for single weights) and must stay
// on chip while the CTA walks all N moduli of its tile.  A 128-row x 96-column slab per CTA is 192 registers per thread
// for 256 epilogue threads -- the largest slab the 64 K-register file holds next to the producer warps (setmaxnreg: 40
// registers for the TMA / MMA warps, 232 for the epilogue warps).  Two CTAs of a TPC share a 256 x 96 tile with
// tcgen05.mma.cta_group::2, so each stages its own 128 rows of A and HALF of the B tile: 5.5 KB of shared-memory reads per
// 48-cycle MMA, below the 128 B/clk an SM can feed its tensor core (a single CTA with N = 96 would need 149 B/clk).
// The walk is tile-major: (tile, modulus 0), (tile, modulus 1), ... so one accumulator buffer of 96 TMEM columns per
// modulus, four in flight; the MMA of the next moduli runs while the epilogue warps fold the previous one into the state.
//
// When it wins.  Per element and modulus the epilogue costs ~10 issue slots (Barrett, int -> fp64, two DFMAs) against
// k / 8192 cycles of tensor time, so the kernel is epilogue-bound below k ~ 700 and tensor-bound above; what it removes is
// the 2 N + 8 bytes per element of residue traffic and the epilogue's transposes and stores.  With long k the narrower tile
// costs more L2 traffic per MAC than the 256 x 256 tiles of oz_gemm_pair_kernel, so oz_api.cu selects this kernel by k
// (option "fused_k").
//
// Anatomy (cluster of 2 CTAs, 384 threads each):
//   warp 0      TMA producer: NSTAGES-stage ring of {A 128 x 128 B, B 48 x 128 B}, bytes accounted on the leader's barrier
//   warp 1      leader only: tcgen05.mma.cta_group::2.kind::i8, M 256, N 96, K 32; commits multicast to both CTAs
//   warp 2      TMEM allocation (512 columns: 4 accumulator buffers at 0 / 128 / 256 / 384)
//   warps 4-11  epilogue + CRT: warp & 3 = TMEM lane quarter (32 rows, lane == row), (warp - 4) >> 2 = 48-column half
#include "oz_tcgen05.cuh"
#include "oz_crt.cuh"

namespace oz {
using namespace tc;
namespace {

constexpr int FT_N       = 96;                 // columns of a tile
constexpr int FT_COLS    = 48;                 // columns per epilogue warp
constexpr int FT_SMEM_A  = 128 * BLOCK_K;      // this CTA's 128 rows of A
constexpr int FT_SMEM_B  = (FT_N / 2) * BLOCK_K;   // this CTA's half of the B tile
constexpr int FT_STAGE   = FT_SMEM_A + FT_SMEM_B;  // 22 KiB
constexpr int FT_STAGES  = 8;
constexpr int FT_ACC     = 4;                  // accumulator buffers in TMEM
constexpr int FT_ACC_STRIDE = 128;             // TMEM columns between buffers
constexpr int FT_THREADS = 384;
constexpr int FT_BARRIERS = 256;
constexpr int FT_SMEM_TOTAL = FT_STAGES * FT_STAGE + FT_BARRIERS + 1024;
constexpr int FT_BAND = 8;                     // 256-row tiles per scheduling band

struct FusedArgs {
    uint32_t rowsA, rowsB, num_kb, num_moduli;
    uint32_t tiles_m, tiles_n, band_m, total;   // tiles ordered (band of row tiles, column tile, row tile in band)
    void *C; size_t ldc;
    const int16_t *sftA; const int16_t *sftB;
    double alpha, beta; int ab_mode;            // host scalars, widened exactly
    const void *alpha_dev; const void *beta_dev;
    const uint32_t *slot; uint32_t *claims;
};

__device__ __forceinline__ void decode_tile(const FusedArgs &a, uint32_t t, uint32_t &tm, uint32_t &tn) {
    const uint32_t per_band = a.band_m * a.tiles_n, full_bands = a.tiles_m / a.band_m;
    uint32_t band = t / per_band, bm = a.band_m, rem;
    if (band < full_bands) rem = t - band * per_band;
    else { band = full_bands; rem = t - full_bands * per_band; bm = a.tiles_m - full_bands * a.band_m; }
    tn = rem / bm;
    tm = band * a.band_m + (rem - tn * bm);
}

template <int REGS> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

template <typename T, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FT_THREADS, 1)
oz_gemm_crt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const FusedArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base  = smem_base + FT_STAGES * FT_STAGE;
    auto full_bar   = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar  = [&](int s) { return bar_base + 8u * (FT_STAGES + s); };
    auto tfull_bar  = [&](int s) { return bar_base + 8u * (2 * FT_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * FT_STAGES + FT_ACC + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * FT_STAGES + 2 * FT_ACC);
    const uint32_t pair_slot = tmem_slot + 8u;
    uint32_t *tmem_slot_ptr  = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank   = cluster_ctarank();
    const uint32_t npairs = gridDim.x >> 1;
    const uint32_t num_kb = args.num_kb, N = args.num_moduli;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < FT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < FT_ACC; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 16); }   // 8 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (rank == 0) {
            const uint32_t pair = claim_pair_slot(args.slot, args.claims, npairs, blockIdx.x >> 1);
            asm volatile("st.shared::cta.u32 [%0], %1;" ::"r"(pair_slot), "r"(pair) : "memory");
        }
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    uint32_t pair;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(pair) : "r"(map_to_cta(pair_slot, 0)) : "memory");

    if (warp < 4) {
        reg_dealloc<40>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer (both CTAs) =====================
            uint32_t stage = 0, phase = 0, tm, tn;
            for (uint32_t tile = pair; tile < args.total; tile += npairs) {
                decode_tile(args, tile, tm, tn);
                const int rowA = (int)(tm * 256 + rank * 128), rowB = (int)(tn * FT_N + rank * (FT_N / 2));
                for (uint32_t j = 0; j < N; ++j) {
                    for (uint32_t kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa   = smem_base + stage * FT_STAGE;
                        const uint32_t lbar = map_to_cta(full_bar(stage), 0);
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * FT_STAGE);
                        tma_load_3d_pair(sa, &map_a, lbar, (int)(kb * BLOCK_K), rowA, (int)j);
                        tma_load_3d_pair(sa + FT_SMEM_A, &map_b, lbar, (int)(kb * BLOCK_K), rowB, (int)j);
                        if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (warp == 1 && lane == 0 && rank == 0) {
            // ===================== MMA issuer (leader CTA) =====================
            constexpr uint32_t idesc = make_idesc(256, FT_N);
            uint32_t stage = 0, phase = 0, it = 0;
            for (uint32_t tile = pair; tile < args.total; tile += npairs) {
                for (uint32_t j = 0; j < N; ++j, ++it) {
                    const uint32_t acc = it % FT_ACC, acc_phase = (it / FT_ACC) & 1;
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * FT_ACC_STRIDE;
                    for (uint32_t kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tcgen05_fence_after();
                        const uint32_t sa = smem_base + stage * FT_STAGE;
                        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + FT_SMEM_A);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_i8_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | (uint32_t)k) != 0 ? 1u : 0u);
                        tcgen05_commit_pair(empty_bar(stage));
                        if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tcgen05_commit_pair(tfull_bar(acc));
                }
            }
        }
    } else {
        reg_alloc<232>();
        // ===================== epilogue + CRT (both CTAs, own 128 rows) =====================
        const int q = warp & 3, half = (warp - 4) >> 2;
        const uint32_t ti_w = N - 2;
        T alpha = (T)args.alpha, beta = (T)args.beta;
        int ab_mode = args.ab_mode;
        if (args.alpha_dev != nullptr) {
            alpha = *static_cast<const T *>(args.alpha_dev); beta = *static_cast<const T *>(args.beta_dev);
            ab_mode = alpha_beta_mode(alpha, beta);
        }
        uint32_t tm, tn, it = 0;
        for (uint32_t tile = pair; tile < args.total; tile += npairs) {
            decode_tile(args, tile, tm, tn);
            double s1[FT_COLS], s2[SPLIT ? FT_COLS : 1];
#pragma unroll
            for (int e = 0; e < FT_COLS; ++e) s1[e] = 0.0;
            if constexpr (SPLIT) {
#pragma unroll
                for (int e = 0; e < FT_COLS; ++e) s2[e] = 0.0;
            }
#pragma unroll 1
            for (uint32_t j = 0; j < N; ++j, ++it) {
                const uint32_t acc = it % FT_ACC, acc_phase = (it / FT_ACC) & 1;
                const uint32_t taddr = tmem_base + acc * FT_ACC_STRIDE + ((uint32_t)(q * 32) << 16) + half * FT_COLS;
                const uint32_t inv = dev_tab::OZ_BARRETT_INV[j], negm = dev_tab::OZ_BARRETT_NEGM[j];
                const uint32_t off = dev_tab::OZ_BARRETT_OFF[j];
                double w1, w2 = 0.0;
                if constexpr (SPLIT) { w1 = dev_tab::OZ_W2_HI[N - 8][j]; w2 = dev_tab::OZ_W2_LO[N - 8][j]; }
                else                 { w1 = dev_tab::OZ_W1[ti_w][j]; }
                mbar_wait(tfull_bar(acc), acc_phase);
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < FT_COLS / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(taddr + 16 * c, v);
                    tmem_ld_wait(v);
                    if (c == FT_COLS / 16 - 1) {
                        // every column of this buffer is in registers: hand it back before the arithmetic of the last chunk
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(tempty_bar(acc), 0));
                    }
                    if (j == 0) {          // modulus 256: the low byte (the only product that may wrap)
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] &= 0xffu;
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = reduce_mod_u((int32_t)v[e], negm, inv, off);
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const double rd = byte_to_double(v[e]);
                        s1[16 * c + e] = fma(w1, rd, s1[16 * c + e]);
                        if constexpr (SPLIT) s2[16 * c + e] = fma(w2, rd, s2[16 * c + e]);
                    }
                }
            }
            // ---- all moduli folded: reduce mod M, undo the scaling, alpha / beta, store (lane == row: 256 B per column) ----
            const uint32_t row  = tm * 256 + rank * 128 + q * 32 + lane;
            const uint32_t col0 = tn * FT_N + half * FT_COLS;
            if (row < args.rowsA) {
                const int sa = (int)args.sftA[row];
                T *crow      = static_cast<T *>(args.C) + row;
#pragma unroll
                for (int e = 0; e < FT_COLS; ++e) {
                    const uint32_t col = col0 + e;
                    if (col < args.rowsB) {
                        const double v = scale_pow2(crt_finish<SPLIT>(N, s1[e], SPLIT ? s2[e] : 0.0), sa + (int)args.sftB[col]);
                        T *cptr = crow + (size_t)col * args.ldc;
                        *cptr   = combine<T>(ab_mode, alpha, beta, cast_out<T>(v), cptr);
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <typename T, bool SPLIT>
cudaError_t launch_fused(const GemmProblem &p, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (!detail::make_operand_map(&ma, p.A8i, p.ld8i, p.rowsA, p.num_slices, p.sizeA, 128)) return cudaErrorInvalidValue;
    if (!detail::make_operand_map(&mb, p.B8i, p.ld8i, p.rowsB, p.num_slices, p.sizeB, FT_N / 2)) return cudaErrorInvalidValue;
    FusedArgs a{};
    a.rowsA = (uint32_t)p.rowsA; a.rowsB = (uint32_t)p.rowsB;
    a.num_kb = (uint32_t)((p.ld8i + BLOCK_K - 1) / BLOCK_K);
    a.num_moduli = p.num_slices;
    a.tiles_m = (uint32_t)((p.rowsA + 255) / 256); a.tiles_n = (uint32_t)((p.rowsB + FT_N - 1) / FT_N);
    a.band_m = FT_BAND; a.total = a.tiles_m * a.tiles_n;
    a.C = p.C; a.ldc = p.ldc; a.sftA = p.sftA; a.sftB = p.sftB;
    if (p.device_scalars) {
        a.alpha = 1.0; a.beta = 0.0; a.alpha_dev = p.alpha_ptr; a.beta_dev = p.beta_ptr;
    } else {
        a.alpha = (double)*static_cast<const T *>(p.alpha_ptr); a.beta = (double)*static_cast<const T *>(p.beta_ptr);
    }
    a.ab_mode = alpha_beta_mode((T)a.alpha, (T)a.beta);
    auto kern = oz_gemm_crt_kernel<T, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    const uint32_t max_pairs = (uint32_t)detail::sm_count() / 2;
    if (max_pairs == 0) return cudaErrorInvalidDevice;
    const uint32_t pairs = a.total < max_pairs ? a.total : max_pairs;
    a.slot = (pairs == max_pairs && p.claims != nullptr && max_pairs <= kClaimBytes / sizeof(uint32_t))
                 ? detail::placement_slots(!detail::stream_is_capturing(st)) : nullptr;
    a.claims = p.claims;
    if (a.slot != nullptr && !p.claims_zeroed) {
        e = cudaMemsetAsync(p.claims, 0, sizeof(uint32_t) * max_pairs, st);
        if (e != cudaSuccess) return e;
    }
    kern<<<2 * pairs, FT_THREADS, FT_SMEM_TOTAL, st>>>(ma, mb, a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm_crt(const GemmProblem &p, cudaStream_t st) {
    if (p.rowsA == 0 || p.rowsB == 0 || p.num_slices == 0) return cudaSuccess;
    if (p.combine != RC_STORE || p.first_modulus != 0 || p.num_slices < 2) return cudaErrorInvalidValue;
    if (p.dtype_C == DT_F32) return launch_fused<float, false>(p, st);
    if (p.dtype_C != DT_F64) return cudaErrorInvalidValue;
    return p.split_weights ? launch_fused<double, true>(p, st) : launch_fused<double, false>(p, st);
}

}  // namespace oz
