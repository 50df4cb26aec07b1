// Dense primitive of the STAT decoder: out = post*act(alpha * P.Q^T + bias + addend)
// with fp32-faithful accuracy on the 5th-gen tensor cores.
//
// Precision scheme: every fp32 operand x is split in shared memory
// into hi = tf32(x) (low 13 mantissa bits cleared) and lo = tf32(x - hi); the
// tile product is accumulated in fp32 TMEM as  lo_P.hi_Q + hi_P.lo_Q + hi_P.hi_Q
// ("3xTF32").  The dropped lo.lo term is ~2^-22 relative.
//
// TS variant (default): the operand on the 128-lane axis ("P") never goes back to shared memory --
// the splitter threads read the raw TMA tile, split it in registers and store hi / lo straight
// into tensor memory (tcgen05.st); the MMAs take A from TMEM and only B ("Q") from shared memory.
// That removes the hi/lo write-back and the three A reads per k-step from the shared-memory port,
// which is what bounded the SS variant (224 KB of shared-memory traffic per 32-wide k-slice of a
// 128x128 tile).
//
// Kernel shape (one 128 x BQ output tile per CTA, 320 threads):
//   warp 0      TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) of the raw fp32
//                              P and Q k-slices into a STAGES-deep ring
//   warps 2..5  splitter     : rewrite hi in place, write lo beside it, then
//                              fence.proxy.async + mbarrier arrive
//   warp 1      MMA issuer   : one lane issues tcgen05.mma.kind::tf32 (M=128,N=BQ,K=8),
//                              3 products x 4 k-steps per stage; tcgen05.commit frees
//                              the stage / publishes the accumulator
//   warps 2..5  epilogue     : tcgen05.ld 32x32b -> bias/act -> global
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "stat_common.cuh"
#include "tc_ptx.cuh"

namespace stat {

// ---------------------------------------------------------------------------
// error string (thread local) -- lives here so every TU links one copy
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }

static unsigned long long g_launches = 0;
// STAT_SYNC_DEBUG=1 (fault hunting): every launch is followed by a device synchronisation; the first launch that
// leaves an error behind is reported with its index and the label of its launch site.
static const char *g_launch_label = "";
void set_launch_label(const char *label) { g_launch_label = label; }
void note_launch() {
  ++g_launches;
  static int dbg = -1;
  if (dbg < 0) {
    const char *e = getenv("STAT_SYNC_DEBUG");
    dbg = (e && e[0] == '1') ? 1 : 0;
  }
  if (dbg) {
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      fprintf(stderr, "[stat] launch #%llu (%s) failed: %s\n", g_launches, g_launch_label, cudaGetErrorString(err));
      fflush(stderr);
      dbg = 0;
    }
  }
  g_launch_label = "";
}

int pdl_attr(cudaLaunchAttribute *attr) {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STAT_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on) return 0;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  return 1;
}
unsigned long long launch_count() { return g_launches; }

static long long *g_trace = nullptr;
void gemm_set_trace(long long *p) { g_trace = p; }
long long *gemm_get_trace() { return g_trace; }

static int g_gemm_impl = 0;
void gemm_set_impl(int impl) { g_gemm_impl = impl; }
int gemm_get_impl() { return g_gemm_impl; }

namespace {

constexpr int BP = 128;      // tile rows (TMEM lanes)
constexpr int BK = 32;       // fp32 per k-slice row = 128 B = one swizzle row
constexpr int UMMA_K = 8;    // tf32 MMA K
constexpr int NSPLIT = 256;               // splitter / epilogue threads (8 warps: 2 per scheduler)
constexpr int NTHREADS = 64 + NSPLIT;     // + TMA producer warp + MMA issuer warp

template <int BQ, bool TS>
struct Cfg {
  static constexpr int P_BYTES = BP * BK * 4;
  static constexpr int Q_BYTES = BQ * BK * 4;
  // stage = raw P (-> hi in place) | [lo of P: SS variant only; TS keeps hi / lo of P in tensor memory] | Q hi | Q lo
  static constexpr int Q_OFF = TS ? P_BYTES : 2 * P_BYTES;
  static constexpr int STAGE_BYTES = Q_OFF + 2 * Q_BYTES;
  // TS variant: accumulator in columns [0, BQ), A-operand stage s (hi 32 cols | lo 32 cols) at TS_A_BASE + 64 s
  static constexpr int TS_A_BASE = BQ <= 64 ? 128 : 256;
  static constexpr int TS_TMEM_COLS = 512;
  // Ring depth: what fits in shared memory (TS: and in the tensor-memory columns behind the accumulator).  The skinny
  // per-step products are bound by the latency of a stage's refill (TMA round trip + split + MMAs) divided by the
  // number of stages in flight: 6 stages of 32 KB instead of 4 of 48 KB for BQ = 64.
  static constexpr int SMEM_RING = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 1024 /*static*/;
  static constexpr int BY_SMEM = SMEM_RING / STAGE_BYTES;
  static constexpr int BY_TMEM = (TS_TMEM_COLS - TS_A_BASE) / 64;
  static constexpr int MAX_STAGES = TS ? (BY_SMEM < BY_TMEM ? BY_SMEM : BY_TMEM) : ((BQ >= 128) ? 3 : 4);
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }
  static constexpr int TMEM_COLS = BQ < 32 ? 32 : BQ;
};

struct DevSeg {
  float *C;
  const float *bias;
  const float *addend;
  int ldc, ld_add;
  float alpha, post;
  int act, f0, f1;
};

struct DevArgs {
  int NP, NQ, K;
  int feat_on_p;
  int nseg;
  int ksplit;               // gridDim.z: k-slices, slice z accumulates into plane z of the output
  int stages;               // depth of the shared-memory ring
  int mc;                   // CTAs of a cluster along the P tiles that share (multicast) the Q tile; 1 = off
  long long plane;          // floats between output planes
  long long *trace;         // debug: per-phase clock64 stamps of CTA (0,0,0), or null
  unsigned long long pol_p, pol_q;   // L2 eviction-priority policies of the two operand streams
  DevSeg seg[2];
};

// L2 cache-hint policy words (the encodings createpolicy.fractional.L2::evict_* 1.0 produces)
constexpr unsigned long long L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr unsigned long long L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr unsigned long long L2_EVICT_LAST = 0x14F0000000000000ull;

// ---- PTX wrappers ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// fault hunting (stat_debug_trap_log): a host-mapped word that survives the fault receives the site of the trap
__device__ volatile int *g_trap_log = nullptr;
__device__ __forceinline__ void trap_at(int site) {
  if (g_trap_log) {
    g_trap_log[0] = site;
    g_trap_log[1] = static_cast<int>(blockIdx.x);
    g_trap_log[2] = static_cast<int>(threadIdx.x);
    g_trap_log[3] = static_cast<int>(blockIdx.y) * 65536 + static_cast<int>(blockIdx.z);
    __threadfence_system();
  }
  __trap();
}
#define mbar_wait(bar, parity) mbar_wait_((bar), (parity), 1000000 + __LINE__)
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity, int site) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 26)) trap_at(site);   // a lost TMA / commit becomes an error, never a hang
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1,
                                            uint32_t bar, unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// the Q slice of this CTA goes to the same shared-memory offset of every CTA in `mask` and signals each one's barrier
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar,
                                               uint16_t mask, unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "h"(mask), "l"(policy)
      : "memory");
}
// the stage is free in every CTA of `mask` once the MMAs that read it have completed here
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
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
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 [61,64).
// Rows are 128 B; 8-row groups are 1024 B apart (SBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r & 0xFFFFE000u);   // the 13 bits the tensor core ignores, cleared explicitly
}

// hi/lo split of one 16-byte chunk, in place + sibling buffer
__device__ __forceinline__ void split4(float4 *hi_ptr, float4 *lo_ptr) {
  float4 x = *hi_ptr;
  float4 h, l;
  // hi = x rounded to nearest TF32 (|x - hi| <= 2^-12 |x|), so the dropped lo.lo term is <= 2^-24 relative
  h.x = rna_tf32(x.x);
  h.y = rna_tf32(x.y);
  h.z = rna_tf32(x.z);
  h.w = rna_tf32(x.w);
  uint32_t a, b, c, d;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(x.x - h.x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(x.y - h.y));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(x.z - h.z));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(x.w - h.w));
  l.x = __uint_as_float(a);
  l.y = __uint_as_float(b);
  l.z = __uint_as_float(c);
  l.w = __uint_as_float(d);
  *hi_ptr = h;
  *lo_ptr = l;
}

__device__ __forceinline__ void split_store(const float4 &x, float4 *hi_ptr, float4 *lo_ptr) {
  float4 h, l;
  h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
  // lo = x - hi is exact in fp32; the tensor core reads its top 19 bits (|lo| <= 2^-12 |x|, so what is
  // dropped is <= 2^-22 |x|)
  l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
  *hi_ptr = h;
  *lo_ptr = l;
}

__device__ __forceinline__ float apply_epi(float acc, const DevSeg &s, float bias, float add) {
  float v = fmaf(s.alpha, acc, bias + add);
  if (s.act == 1) v = tanh_fast(v);   // abs. error ~1.5e-7 (stat_common.cuh)
  return v * s.post;
}

template <int BQ, bool TS>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
                   const DevArgs args) {
  using C = Cfg<BQ, TS>;
  const int STAGES = args.stages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B needs the ring 1024-byte aligned; keep the pointer in the shared state space
  // (an integer round trip would turn every access into a generic LD/ST)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * C::STAGE_BYTES);
  // bars: full[STAGES] | split[STAGES] | empty[STAGES] | tmem_full | tmem_ptr(u32)
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_split = bar_full + 8 * STAGES;
  const uint32_t bar_empty = bar_split + 8 * STAGES;
  const uint32_t bar_tmem = bar_empty + 8 * STAGES;
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long *trace = (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? args.trace : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = clock64();
  const int q0 = blockIdx.x * BQ;
  const int p0 = blockIdx.y * BP;
  // this CTA's k-slice: k-blocks [kb0, kb0 + nk)
  const int nk_all = (args.K + BK - 1) / BK;
  const int kz = blockIdx.z;
  const int kb0 = (kz * nk_all) / args.ksplit;
  const int nk = ((kz + 1) * nk_all) / args.ksplit - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_split + 8 * s, TS ? 4 : NSPLIT);
      mbar_init(bar_empty + 8 * s, static_cast<uint32_t>(args.mc));
    }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"(static_cast<uint32_t>(TS ? C::TS_TMEM_COLS : C::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (args.mc > 1) tcx::cluster_sync_all();     // every CTA's barriers exist before a partner signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t crank = args.mc > 1 ? tcx::cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << args.mc) - 1u);
  if (trace && threadIdx.x == 0) trace[1] = clock64();
  // barrier set-up and the tensor-memory allocation above overlap the previous kernel's tail; its
  // results (an operand, an addend) are read from here on.  The successor may be scheduled now:
  // this CTA already holds everything it can block on.
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t stage = smem_u32(smem + s * C::STAGE_BYTES);
        mbar_expect_tx(bar_full + 8 * s, C::P_BYTES + C::Q_BYTES);
        tma_load_2d(stage, &tmP, (kb0 + kb) * BK, p0, bar_full + 8 * s, args.pol_p);
        if (args.mc > 1) {
          // K0 is bound by L2 -> SM traffic (every CTA pulls 32 KB per k-slice: 294 us for ff_local = the L2's
          // ~6 TB/s): the CTAs of a cluster work on different row tiles against the SAME weight tile, so each one
          // fetches 1/mc of it and multicasts it to all
          const uint32_t slice = static_cast<uint32_t>(C::Q_BYTES / args.mc);
          tma_load_2d_mc(stage + C::Q_OFF + crank * slice, &tmQ, (kb0 + kb) * BK,
                         q0 + static_cast<int>(crank) * (BQ / args.mc), bar_full + 8 * s, cmask, args.pol_q);
        } else {
          tma_load_2d(stage + C::Q_OFF, &tmQ, (kb0 + kb) * BK, q0, bar_full + 8 * s, args.pol_q);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();   // lanes 1..31 wait for lane 0: the warp must reach the final barrier converged
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor (cute UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2,
    // b=TF32 [10,13)=2, K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BQ >> 3) << 17) |
                           (static_cast<uint32_t>(BP >> 4) << 24);
    // Every lane computes the (warp-uniform) descriptors; only the MMA / commit instructions run in the lane
    // chosen by elect.sync.  With `if (lane == 0)` around per-MMA asm statements ptxas wraps every tcgen05.mma in
    // an ELECT / R2UR / branch loop that costs 110-160 cycles per MMA -- more than the 64 cycles a 128x128x8
    // tf32 MMA takes -- which is what held this kernel at ~0.57 of the 3xTF32 ceiling in round 1.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t ring_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nk; ++kb) {
      mbar_wait(bar_split + 8 * s, ph);
      tc_fence_after();
      __syncwarp();
      const bool leader = tcx::elect_one();
      if (trace && leader && kb < 16) trace[40 + kb] = clock64();
      const uint32_t stage = ring_u + static_cast<uint32_t>(s) * C::STAGE_BYTES;
      const uint64_t dPh = make_desc(stage);
      const uint64_t dPl = make_desc(stage + C::P_BYTES);
      const uint64_t dQh = make_desc(stage + C::Q_OFF);
      const uint64_t dQl = make_desc(stage + C::Q_OFF + C::Q_BYTES);
      const uint32_t a_hi = tmem_u + C::TS_A_BASE + 64 * s;
      if (leader) {
        if constexpr (TS) {
          tcx::tc_mma_tf32_katom(tmem_u, a_hi, dQh, dQl, idesc, kb ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t adv = static_cast<uint64_t>((ks * UMMA_K * 4) >> 4);  // 32 B per k-step
            tc_mma_tf32(tmem_u, dPl + adv, dQh + adv, idesc, (kb | ks) ? 1u : 0u);
            tc_mma_tf32(tmem_u, dPh + adv, dQl + adv, idesc, 1u);
            tc_mma_tf32(tmem_u, dPh + adv, dQh + adv, idesc, 1u);
          }
        }
        if (args.mc > 1) tc_commit_mc(bar_empty + 8 * s, cmask);
        else tc_commit(bar_empty + 8 * s);
        if (kb == nk - 1) tc_commit(bar_tmem);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================== splitter, then epilogue =====================
    const int t = threadIdx.x - 64;  // 0..NSPLIT-1
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nk; ++kb) {
      // EVERY splitter warp observes EVERY phase of a stage's barrier, also the k-slices its group does not split
      // (TS: the two groups alternate).  A parity wait is only valid for a waiter at most one phase behind: with an
      // odd ring depth a group that skipped the other group's phase of a stage could find that phase still pending
      // when it came back, see the parity of the phase before it and read the stage early -- rare (it needs the
      // loads of two consecutive slices to land out of order), fatal (double arrivals on the split barrier:
      // `unspecified launch failure`; found with cuda-gdb under STAT_PDL=0).
      mbar_wait(bar_full + 8 * s, ph);
      if (trace && t == 0 && kb < 16) trace[2 + kb] = clock64();
      uint8_t *stage = smem + s * C::STAGE_BYTES;
      float4 *Ph = reinterpret_cast<float4 *>(stage);
      float4 *Pl = reinterpret_cast<float4 *>(stage + C::P_BYTES);
      float4 *Qh = reinterpret_cast<float4 *>(stage + C::Q_OFF);
      float4 *Ql = reinterpret_cast<float4 *>(stage + C::Q_OFF + C::Q_BYTES);
      constexpr int NQ4 = (C::Q_BYTES / 16 + NSPLIT - 1) / NSPLIT;  // 1 / 2 / 4
      if constexpr (TS) {
        // The two groups of four splitter warps (one warp per TMEM lane quadrant each) take alternate k-slices, so
        // that the latencies of one slice's split (shared-memory loads, tensor-memory stores and their wait, the
        // proxy fence) overlap the other group's; one arrival per warp.  A thread owns tile row prow (= its TMEM
        // lane) and all 32 k-columns, in two halves of 16.  The TMA tile is SWIZZLE_128B: 16-byte chunk c of row r
        // sits at chunk c ^ (r & 7).
        if ((kb & 1) == ((warp - 2) >> 2)) {
          const int prow = (warp & 3) * 32 + lane;
          const int tg = ((warp - 2) & 3) * 32 + lane;
          const uint8_t *rowp = stage + prow * 128;
          const uint32_t ta = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + C::TS_A_BASE + 64 * s;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 xp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              xp[i] = *reinterpret_cast<const float4 *>(rowp + (((4 * half + i) ^ (prow & 7)) << 4));
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float xs[4] = {xp[i].x, xp[i].y, xp[i].z, xp[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float h = rna_tf32(xs[e]);
                hi[4 * i + e] = __float_as_uint(h);
                lo[4 * i + e] = __float_as_uint(xs[e] - h);
              }
            }
            tc_st16(ta + 16 * half, hi);
            tc_st16(ta + 32 + 16 * half, lo);
          }
          constexpr int NQG = (C::Q_BYTES / 16 + 127) / 128;     // 16-byte chunks of Q per thread of the group
          float4 xq[NQG];
#pragma unroll
          for (int i = 0; i < NQG; ++i) {
            const int idx = tg + i * 128;
            xq[i] = (idx < C::Q_BYTES / 16) ? Qh[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < NQG; ++i) {
            const int idx = tg + i * 128;
            if (idx < C::Q_BYTES / 16) split_store(xq[i], Qh + idx, Ql + idx);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_split + 8 * s);
          if (trace && tg == 0 && kb < 16) trace[20 + kb] = clock64();
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
        continue;
      } else {
        // all 16-byte loads of this thread first, then the arithmetic, then the stores
        constexpr int NP4 = C::P_BYTES / 16 / NSPLIT;                 // 4
        float4 xp[NP4], xq[NQ4];
#pragma unroll
        for (int i = 0; i < NP4; ++i) xp[i] = Ph[t + i * NSPLIT];
#pragma unroll
        for (int i = 0; i < NQ4; ++i) {
          const int idx = t + i * NSPLIT;
          xq[i] = (idx < C::Q_BYTES / 16) ? Qh[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < NP4; ++i) split_store(xp[i], Ph + t + i * NSPLIT, Pl + t + i * NSPLIT);
#pragma unroll
        for (int i = 0; i < NQ4; ++i) {
          const int idx = t + i * NSPLIT;
          if (idx < C::Q_BYTES / 16) split_store(xq[i], Qh + idx, Ql + idx);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_split + 8 * s);
      if (trace && t == 0 && kb < 16) trace[20 + kb] = clock64();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }

    // epilogue: this warp may touch TMEM lanes [32*(warp%4), +32)
    mbar_wait(bar_tmem, 0);
    tc_fence_after();
    if (trace && t == 0) trace[60] = clock64();
    const int wq = warp & 3;                 // TMEM lane quadrant this warp may read
    const int chalf = (warp - 2) >> 2;       // the two warps of a quadrant take half of the columns each
    constexpr int CH = BQ / 2;
    const int cbeg = chalf * CH, cend = cbeg + CH;
    const int p = p0 + wq * 32 + lane;
    const bool sel = args.nseg > 1 && (args.feat_on_p ? p0 : q0) >= args.seg[1].f0;
    // segment parameters into registers (a struct copy indexed at run time would live in local memory);
    // k-slice z > 0 writes its raw partial product into plane z, bias / addend ride on plane 0
    float *const C = (sel ? args.seg[1].C : args.seg[0].C) + static_cast<size_t>(kz) * args.plane;
    const float *const bias = kz > 0 ? nullptr : (sel ? args.seg[1].bias : args.seg[0].bias);
    const float *const addend = kz > 0 ? nullptr : (sel ? args.seg[1].addend : args.seg[0].addend);
    const int ldc = sel ? args.seg[1].ldc : args.seg[0].ldc;
    const int ld_add = sel ? args.seg[1].ld_add : args.seg[0].ld_add;
    const int f0 = sel ? args.seg[1].f0 : args.seg[0].f0;
    const int f1 = sel ? args.seg[1].f1 : args.seg[0].f1;
    const float alpha = sel ? args.seg[1].alpha : args.seg[0].alpha;
    const float post = sel ? args.seg[1].post : args.seg[0].post;
    const bool act = (sel ? args.seg[1].act : args.seg[0].act) == 1;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    if (args.feat_on_p == 0) {
      // rows = p (this lane), features = q: 8 accumulator columns at a time
      const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (((q0 - f0) & 3) == 0);
      float *crow = C + static_cast<size_t>(p) * ldc + (q0 - f0);
      const float *arow = addend ? addend + static_cast<size_t>(p) * ld_add + (q0 - f0) : nullptr;
      const float *brow = bias ? bias + (q0 - f0) : nullptr;
      const int qlim = min(args.NQ, f1) - q0;      // valid columns of this tile
#pragma unroll 1
      for (int c = cbeg; c < cend; c += 8) {
        uint32_t v[8];
        tc_ld8(trow + c, v);
        if (p < args.NP && c < qlim) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const bool ok = c + e < qlim;
            float x = fmaf(alpha, __uint_as_float(v[e]), (ok && brow) ? __ldg(brow + c + e) : 0.f);
            if (ok && arow) x += arow[c + e];
            if (act) x = tanh_fast(x);
            o[e] = x * post;
          }
          if (vec_ok && c + 8 <= qlim) {
            *reinterpret_cast<float4 *>(crow + c) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(crow + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (c + e < qlim) crow[c + e] = o[e];
          }
        }
      }
    } else {
      // features = p (this lane), rows = q: every store is one coalesced 128-byte row segment
      const bool pok = p < args.NP && p < f1;
      const int j = p - f0;
      const float pb = (pok && bias) ? __ldg(bias + j) : 0.f;
      float *ccol = C + static_cast<size_t>(q0) * ldc + j;
      const float *acol = addend ? addend + static_cast<size_t>(q0) * ld_add + j : nullptr;
      const int qlim = args.NQ - q0;
      const size_t ldc_s = static_cast<size_t>(ldc);
      if (!acol && !act) {
        // the per-step projections: plain scaled copy, pointer-bumped
        float *cp = ccol + static_cast<size_t>(cbeg) * ldc_s;
#pragma unroll 1
        for (int c = cbeg; c < cend; c += 8) {
          uint32_t v[8];
          tc_ld8(trow + c, v);
          if (pok) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (c + e < qlim) cp[e * ldc_s] = fmaf(alpha, __uint_as_float(v[e]), pb) * post;
          }
          cp += 8 * ldc_s;
        }
      } else {
#pragma unroll 1
        for (int c = cbeg; c < cend; c += 8) {
          uint32_t v[8];
          tc_ld8(trow + c, v);
          if (pok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (c + e < qlim) {
                float x = fmaf(alpha, __uint_as_float(v[e]), pb);
                if (acol) x += acol[static_cast<size_t>(c + e) * ld_add];
                if (act) x = tanh_fast(x);
                ccol[static_cast<size_t>(c + e) * ldc_s] = x * post;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    if (trace && t == 0) trace[61] = clock64();
  }

  __syncthreads();
  if (args.mc > 1) tcx::cluster_sync_all();     // nobody leaves while a partner may still signal its barriers
  if (trace && threadIdx.x == 0) trace[62] = clock64();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(TS ? C::TS_TMEM_COLS : C::TMEM_COLS))
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// plain fp32 SIMT version of the same contract (device-side cross-check)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float *P, int ldp, const float *Q, int ldq,
                                                        const DevArgs args) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sP[32][33];
  __shared__ float sQ[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int p0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int nk_all = (args.K + 31) / 32, kz = blockIdx.z;
  const int kbeg = ((kz * nk_all) / args.ksplit) * 32;
  const int kend = min(args.K, (((kz + 1) * nk_all) / args.ksplit) * 32);
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + tx;
      sP[r][tx] = (p0 + r < args.NP && k < kend) ? P[static_cast<size_t>(p0 + r) * ldp + k] : 0.f;
      sQ[r][tx] = (q0 + r < args.NQ && k < kend) ? Q[static_cast<size_t>(q0 + r) * ldq + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float qv = sQ[tx][k];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sP[ty + 8 * i][k], qv, acc[i]);
    }
    __syncthreads();
  }
  const int q = q0 + tx;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + 8 * i;
    if (p >= args.NP || q >= args.NQ) continue;
    const int feat = args.feat_on_p ? p : q;
    const int row = args.feat_on_p ? q : p;
    const int sel = (args.nseg > 1 && feat >= args.seg[1].f0) ? 1 : 0;
    const DevSeg &sg = args.seg[sel];
    if (feat < sg.f0 || feat >= sg.f1) continue;
    const int j = feat - sg.f0;
    const float b = (sg.bias && kz == 0) ? sg.bias[j] : 0.f;
    const float add = (sg.addend && kz == 0) ? sg.addend[static_cast<size_t>(row) * sg.ld_add + j] : 0.f;
    sg.C[static_cast<size_t>(kz) * args.plane + static_cast<size_t>(row) * sg.ldc + j] = apply_epi(acc[i], sg, b, add);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

int make_tensor_map(CUtensorMap *tm, const float *base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  STAT_REQUIRE(enc != nullptr, STAT_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  STAT_REQUIRE(r == CUDA_SUCCESS, STAT_ECUDA, "cuTensorMapEncodeTiled failed with %d (rows=%d K=%d ld=%d)",
               static_cast<int>(r), rows, K, ld);
  return STAT_OK;
}

namespace {

template <int BQ, bool TS>
int launch_tc(const GemmArgs &a, const DevArgs &da, cudaStream_t stream) {
  CUtensorMap tmP, tmQ;
  // wide problems (K0): clusters of mc row tiles share the weight tile by TMA multicast
  int mc = 1;
  {
    static int mc_max = -1;
    if (mc_max < 0) {
      const char *e = getenv("STAT_GEMM_MULTICAST");       // 1 = off (default), 2 / 4 = cluster size limit
      mc_max = e ? atoi(e) : 1;       // measured on the B200: no gain at 2, slower at 4 -> off
      if (mc_max != 1 && mc_max != 2 && mc_max != 4) mc_max = 1;
    }
    const int ptiles = (a.NP + BP - 1) / BP;
    if (TS && BQ == 128 && !a.feat_on_p && da.ksplit == 1 && ptiles >= 8) {
      for (int c = mc_max; c > 1; c >>= 1)
        if (ptiles % c == 0) { mc = c; break; }
    }
  }
  STAT_TRY(make_tensor_map(&tmP, a.P, a.NP, a.K, a.ldp, BP));
  STAT_TRY(make_tensor_map(&tmQ, a.Q, a.NQ, a.K, a.ldq, BQ / mc));
  static size_t smem_set[STAT_MAX_DEV] = {};
  STAT_TRY(ensure_dyn_smem(gemm_tf32x3_kernel<BQ, TS>, Cfg<BQ, TS>::smem_bytes(Cfg<BQ, TS>::MAX_STAGES), smem_set));
  DevArgs db = da;
  db.mc = mc;
  const int nk_slice = ((a.K + BK - 1) / BK + da.ksplit - 1) / da.ksplit;
  db.stages = nk_slice < Cfg<BQ, TS>::MAX_STAGES ? nk_slice : Cfg<BQ, TS>::MAX_STAGES;
  dim3 grid((a.NQ + BQ - 1) / BQ, (a.NP + BP - 1) / BP, da.ksplit);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = Cfg<BQ, TS>::smem_bytes(db.stages);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  int na = 0;
  if (mc > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = mc;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  na += pdl_attr(attr + na);
  cfg.numAttrs = na;
  set_launch_label("gemm_tf32x3");
  STAT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<BQ, TS>, tmP, tmQ, db));
  note_launch();
  return STAT_OK;
}

}  // namespace

int gemm_set_trap_log(int *dev_ptr) {
  STAT_CUDA_CHECK(cudaMemcpyToSymbol(g_trap_log, &dev_ptr, sizeof(dev_ptr)));
  return STAT_OK;
}

int gemm_launch(const GemmArgs &a, cudaStream_t stream) {
  STAT_REQUIRE(a.NP > 0 && a.NQ > 0 && a.K > 0, STAT_EINVAL, "gemm: empty problem %d x %d x %d", a.NP, a.NQ, a.K);
  STAT_REQUIRE(a.nseg == 1 || a.nseg == 2, STAT_EINVAL, "gemm: nseg must be 1 or 2");
  DevArgs da;
  da.NP = a.NP;
  da.NQ = a.NQ;
  da.K = a.K;
  da.feat_on_p = a.feat_on_p;
  da.nseg = a.nseg;
  const int nk_all = (a.K + BK - 1) / BK;
  da.ksplit = a.ksplit < 1 ? 1 : (a.ksplit > nk_all ? nk_all : a.ksplit);
  da.stages = 1;
  da.mc = 1;
  da.plane = static_cast<long long>(a.plane);
  da.trace = g_trace;
  // Skinny activations ("swap": the weights ride the 128-lane axis): each weight tile is read by one CTA
  // and not again before the next decode step, with the whole context-block stream of the attention
  // kernel in between -- mark it evict_first so that it does not push those blocks out of the L2; the
  // activation rows are re-read by every CTA.  Wide problems (K0) keep the default priority.
  static int hint = -1;
  if (hint < 0) {
    const char *e = getenv("STAT_GEMM_L2");
    hint = (e && e[0] == '0') ? 0 : 1;
  }
  da.pol_p = (hint && a.feat_on_p && a.NQ <= 128) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  da.pol_q = (hint && a.feat_on_p && a.NQ <= 128) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
  if (da.ksplit > 1) {
    STAT_REQUIRE(a.nseg == 1 && a.seg[0].act == 0, STAT_EINVAL,
                 "gemm: k-split needs a single linear segment (partial sums are combined by the consumer)");
  }
  for (int i = 0; i < a.nseg; ++i) {
    const GemmSeg &s = a.seg[i];
    da.seg[i] = DevSeg{s.C, s.bias, s.addend, s.ldc, s.ld_add, s.alpha, s.post, s.act, s.f0, s.f1};
  }
  if (a.nseg == 1) da.seg[1] = da.seg[0];
  // TMA needs 16-byte aligned rows; operands that are not (odd feature widths of toy
  // configurations) take the fp32 SIMT kernel, as does everything when impl == 1.
  const bool tma_ok = (a.ldp & 3) == 0 && (a.ldq & 3) == 0 && (reinterpret_cast<uintptr_t>(a.P) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(a.Q) & 15) == 0;
  if (g_gemm_impl == 1 || !tma_ok) {
    dim3 grid((a.NQ + 31) / 32, (a.NP + 31) / 32, da.ksplit);
    return launch_pdl(gemm_simt_kernel, grid, dim3(256), 0, stream, a.P, a.ldp, a.Q, a.ldq, da);
  }
  if (a.nseg == 2) {
    const int tile = a.feat_on_p ? BP : 128;
    STAT_REQUIRE(a.seg[1].f0 % tile == 0, STAT_EINVAL, "gemm: segment boundary %d not a multiple of %d",
                 a.seg[1].f0, tile);
  }
  if (g_gemm_impl == 2) {   // SS variant (A and B from shared memory), kept for A/B comparison
    if (a.NQ > 64) return launch_tc<128, false>(a, da, stream);
    if (a.NQ > 32) return launch_tc<64, false>(a, da, stream);
    return launch_tc<32, false>(a, da, stream);
  }
  {
    // 128 x 256 tiles where they still fill the SMs more than twice over (the context projections of K0: 13 312 rows
    // x 1 024 features): an MMA instruction costs ~46 + 0.36 N cycles, so N = 256 runs the tensor pipe at ~0.9 of
    // its peak where N = 128 reaches 0.75, the P tile is split once for twice the columns, and half as many CTAs
    // pay a prologue and an epilogue.   STAT_GEMM_BQ256=0 switches it off.
    static int wide = -1;
    if (wide < 0) {
      const char *e = getenv("STAT_GEMM_BQ256");
      wide = (e && e[0] == '0') ? 0 : 1;
    }
    const long long tiles256 = static_cast<long long>((a.NQ + 255) / 256) * ((a.NP + BP - 1) / BP);
    if (wide && !a.feat_on_p && da.ksplit == 1 && a.NQ % 256 == 0 && tiles256 >= 2 * 148 &&
        (a.nseg == 1 || a.seg[1].f0 % 256 == 0))
      return launch_tc<256, true>(a, da, stream);
  }
  if (a.NQ > 64) return launch_tc<128, true>(a, da, stream);
  if (a.NQ > 32) return launch_tc<64, true>(a, da, stream);
  return launch_tc<32, true>(a, da, stream);
}

}  // namespace stat
