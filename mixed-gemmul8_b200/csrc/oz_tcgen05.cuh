// tcgen05 / TMA / mbarrier / cluster PTX wrappers and small launch helpers shared by the all-moduli GEMM kernels
// (oz_gemm.cu: product + residues; oz_gemm_crt.cu: product + residues + CRT in one kernel).  sm_100a only.
#pragma once
#include "oz_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace oz {
namespace tc {

constexpr int BLOCK_K = 128;  // int8 elements == bytes == one 128B swizzle atom
constexpr int UMMA_K  = 32;
constexpr uint32_t TMEM_COLS = 512;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.  The bound is wall time
// (%globaltimer), 20 s: far beyond anything a time-sliced or co-scheduled kernel can be held up for, so it only ever
// fires on a genuine deadlock.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xfffu) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// bring a box into L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// TMA store of a shared-memory box (this CTA's) into a 3-D tensor, bulk-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes done
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, s8 x s8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// two 32-register loads in flight, one wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                   "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]),
                   "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]),
                   "+r"(b[16]), "+r"(b[17]), "+r"(b[18]), "+r"(b[19]), "+r"(b[20]), "+r"(b[21]), "+r"(b[22]), "+r"(b[23]),
                   "+r"(b[24]), "+r"(b[25]), "+r"(b[26]), "+r"(b[27]), "+r"(b[28]), "+r"(b[29]), "+r"(b[30]), "+r"(b[31])
                 :
                 : "memory");
}
// wait, with the 32 loaded registers as in/out operands (see the 16-register overload below)
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
// the same, with the loaded registers as in/out operands: every use of v is ordered after the wait by data flow, so
// another tcgen05.ld may be in flight (software-pipelined epilogue) without relying on instruction order alone
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// K-major operand tile, 128-byte swizzle: rows are 128 B apart, 8-row groups (1024 B) are the
// stride dimension; descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)0 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
    d |= (uint64_t)1 << 46;                         // version
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = s32, A = B = s8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// The "accumulator drained" signal hands over tensor memory, not data in global or shared memory: the tcgen05.ld of
// this warp have completed (tcgen05.wait::ld) and are fenced (tcgen05.fence::before_thread_sync) when it is sent.  A
// release at cluster scope would also wait for the warp's residue STOREs to be acknowledged (MEMBAR.GPU + ERRBAR,
// 8 % of the epilogue's time at k = 2048) before the MMA warp may reuse the buffer.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// x mod m in [0, m) for |x| <= 2^17 * 127^2 (every modulus but 256 has |residue| <= 127, k <= 2^17): shift x by a
// multiple of m into the unsigned range, q = umulhi(x', floor(2^32 / m)) is floor(x' / m) or one less, ONE correction.
// Four instructions per residue, two on the integer-multiply pipe (IMAD.HI, IMAD) and two on the ALU (add, add-min): all of
// IMAD* issues on the 16-lane "FMA heavy" pipe, which is what bounded the first version of this epilogue (ncu: 73 % busy,
// four IMAD-class instructions per residue because the compiler negates the quotient and turns adds into IMAD.IADD).
// negm = 2^32 - m comes from a table so that the compiler sees an opaque multiplier and emits q * negm + x' as one IMAD.
constexpr uint32_t kMaxAbsProduct = 2114060288u;   // 2^17 * 127^2
__device__ __forceinline__ uint32_t reduce_mod_u(int32_t x, uint32_t negm, uint32_t inv, uint32_t off) {
    const uint32_t xu = (uint32_t)x + off;
    const uint32_t q  = __umulhi(xu, inv);
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(negm), "r"(xu));   // x' - q m, in [0, 2m)
    return min(r, r + negm);                                                    // r < m: r - m wraps to a huge value
}
// TMA load into this CTA's shared memory; the transaction bytes land on the barrier at `bar_cluster_addr`
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Claim a unique work slot for this cluster (leader's elected thread; see oz_gemm_pair_kernel): the preferred slot is the
// one the SM placement table gives this SM, the next free one otherwise.  claims[] holds one zeroed word per slot.
__device__ __forceinline__ uint32_t claim_pair_slot(const uint32_t *slot, uint32_t *claims, uint32_t npairs, uint32_t fallback) {
    if (slot == nullptr) return fallback;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    uint32_t want = slot[smid] >> 1;
    if (want >= npairs) want = fallback;
    for (uint32_t i = 0; i < npairs; ++i) {
        const uint32_t c = want + i < npairs ? want + i : want + i - npairs;
        if (atomicCAS(claims + c, 0u, 1u) == 0u) return c;
    }
    return fallback;   // unreachable: npairs clusters, npairs slots
}

}  // namespace tc

// host helpers defined in oz_gemm.cu
namespace detail {
bool make_operand_map(CUtensorMap *map, const int8_t *base, size_t ld8i, size_t rows, size_t slices, size_t slice_stride, uint32_t box_rows);
// (row, column, modulus) view of a stack of column-major uint8 residue matrices; box = 128 rows x 32 columns, 128-byte swizzle
bool make_residue_map(CUtensorMap *map, const uint8_t *base, size_t ld, size_t rows, size_t cols, size_t slices, size_t slice_stride);
int sm_count();                                        // of the current device (cached)
const uint32_t *placement_slots(bool allow_probe);     // smid -> block index of a plain launch, or nullptr
bool stream_is_capturing(cudaStream_t st);
}  // namespace detail
}  // namespace oz
