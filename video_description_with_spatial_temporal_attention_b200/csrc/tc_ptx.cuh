// PTX wrappers shared by the fused step kernels (sm_100a): mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma.kind::tf32 with the A operand in tensor memory / st / ld / commit), global-memory flags.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stat {
namespace tcx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// a lost TMA / commit / flag becomes a trap (an error the host sees), never a hang
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar,
                                            unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// L2 cache-hint policy words (the encodings createpolicy.fractional.L2::evict_* 1.0 produces)
constexpr unsigned long long L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr unsigned long long L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr unsigned long long L2_EVICT_LAST = 0x14F0000000000000ull;

// one lane of the (converged) warp; ptxas knows that code predicated on an elect.sync result runs in a single
// thread and issues tcgen05.mma / commit there without wrapping each one in an elect / branch loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, tf32 operands, fp32 accumulation
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The 12 products of one k-atom (32 fp32 of K = 4 k-steps of 8; per k-step lo.hi, hi.lo, hi.hi) into ONE accumulator,
// as a single asm block: the operands cross from vector to uniform registers once per k-atom instead of once per
// MMA.  a_hi: tensor-memory address of the hi half of the A stage (lo half 32 columns further); dqh / dql: shared
// memory descriptors of the hi / lo B tiles (k-step advance = 32 bytes = +2 in the descriptor's address field).
__device__ __forceinline__ void tc_mma_tf32_katom(uint32_t tmem_d, uint32_t a_hi, uint64_t dqh, uint64_t dql,
                                                  uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, pt;\n\t"
      ".reg .b32 ah1, ah2, ah3, al0, al1, al2, al3;\n\t"
      ".reg .b64 qh1, qh2, qh3, ql1, ql2, ql3;\n\t"
      "setp.ne.b32 p0, %5, 0;\n\t"
      "setp.eq.b32 pt, %5, %5;\n\t"
      "add.u32 al0, %1, 32;\n\t"
      "add.u32 ah1, %1, 8;\n\t"
      "add.u32 al1, %1, 40;\n\t"
      "add.u32 ah2, %1, 16;\n\t"
      "add.u32 al2, %1, 48;\n\t"
      "add.u32 ah3, %1, 24;\n\t"
      "add.u32 al3, %1, 56;\n\t"
      "add.s64 qh1, %2, 2;\n\t"
      "add.s64 ql1, %3, 2;\n\t"
      "add.s64 qh2, %2, 4;\n\t"
      "add.s64 ql2, %3, 4;\n\t"
      "add.s64 qh3, %2, 6;\n\t"
      "add.s64 ql3, %3, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al0], %2, %4, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al1], qh1, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], ql1, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], qh1, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al2], qh2, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], ql2, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], qh2, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al3], qh3, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], ql3, %4, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], qh3, %4, pt;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_hi), "l"(dqh), "l"(dql), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t tmem, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout=2 [61,64).  Rows are 128 B; 8-row groups are 1024 B apart (SBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// instruction descriptor of tcgen05.mma.kind::tf32: c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2, both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r & 0xFFFFE000u);
}

// ---- thread-block clusters: rank, barrier, distributed shared memory reads ----------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared-memory location of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void *local, uint32_t rank) {
  uint32_t ra;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
  return ra;
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- global-memory flags between CTAs of one (co-resident) grid ------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until *p >= want (monotonic counters); traps instead of hanging
__device__ __forceinline__ void wait_flag_ge(const unsigned int *p, unsigned int want) {
  unsigned int spins = 0;
  while (ld_acquire_gpu(p) < want) {
    if (++spins > (1u << 24)) __trap();
    __nanosleep(32);
  }
}

}  // namespace tcx
}  // namespace stat
