// Shared device-side vocabulary of the tcgen05 kernels (r3d_gemm_tc.cu, r3d_bottom_tc.cu): tile constants, raw PTX wrappers
// (mbarrier, TMA loads/stores, tcgen05.mma / .ld / .st / .commit, cluster helpers) and the UMMA descriptors.  sm_100a only.
#pragma once
#include <cuda.h>

#include <cstdio>

#include "r3d_internal.h"

namespace r3d {

constexpr int TBM = 128;            // UMMA M per CTA (one TMEM lane per output row)
constexpr int TBK = 64;             // K block: 64 bf16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 384;     // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int EPI_WARP0 = 4;
constexpr int SMEM_LIMIT = 227 * 1024;

// Timing experiments (skip stores / residual / epilogue, alternative cache hints ...) exist only in builds made with
// -DR3D_EXPERIMENTS (R3D_BUILD_EXPERIMENTS=1 python -m ray3d_b200.build --force): some of them produce wrong results
// by design, so the shipped library neither reads their environment switches nor contains their branches.
#ifdef R3D_EXPERIMENTS
#define R3D_DBG(bits) (dbg & (bits))
#else
#define R3D_DBG(bits) 0
#endif

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("r3d gemm_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Tagged form (the chained tail kernel): the message names the barrier kind and the waiter's progress.
__device__ __forceinline__ void mbar_wait_tag(uint64_t* bar, uint32_t parity, int tag, int info) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("r3d tail_tc: mbarrier timeout tag %d info %d (block %d thread %d)\n", tag, info, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Same with cluster-scope acquire: the waiter reads data another CTA of the cluster wrote before its (release.cluster)
// arrival -- the tile ids the leader CTA's scheduler stores into the peer's queue.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("r3d gemm_tc: cluster mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// 32-bit store into the shared memory of CTA `rank` of the cluster (same offset as `p` in this CTA)
__device__ __forceinline__ void st_shared_cluster_u32(const volatile void* p, uint32_t rank, uint32_t v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.u32 [ra], %2;\n\t}" ::"r"(smem_u32(const_cast<const void*>(p))),
      "r"(rank), "r"(v)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// orders the generic proxy (ld/st/atomics of this thread, observed writes) with the async proxy (TMA loads/stores), all spaces
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// L2 eviction-priority policies (same encodings CUTLASS uses for TMA::CacheHintSm90)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // activations: streamed once
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;    // weights: re-read by every tile of the problem
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;  // activations the epilogue reads again as the residual
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5, %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1)
               : "memory");
}
// TMA store of one swizzled smem box; completion tracked per thread with bulk groups
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const void* tmap, const void* smem_src, int c0, int c1, int c2, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tmap_prefetch(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// ---- 2-SM (cta_group::2) forms: one MMA spans the CTA pair (M = 256), issued by the leader CTA only ----------------
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// A operand from tensor memory (TS mode): lane = row, two bf16 per 32-bit column, 8 columns per K=16 step
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMA load whose completion is signalled on the LEADER CTA's barrier (peer bit of the cluster address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], "
      "%5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
// Arrive on a barrier of CTA `rank` of the cluster.  These arrivals only hand tensor-memory columns back and forth
// (the tcgen05.wait / tcgen05.fence pair around them orders those accesses), no shared/global data is published, so
// the arrive is .relaxed: the .release.cluster form costs a MEMBAR.ALL.CTA + ERRBAR per arrival, which ncu showed as
// 17% of all stall samples of the first-layer launch (22% of the epilogue warps' time).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank, bool release = false) {
  if (release)
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile ([rows][64] bf16, 8-row groups 1024 bytes apart):
// start address >> 4 | LBO (ignored for swizzled K-major; canonical value 1) | SBO = 1024 >> 4 |
// descriptor version 1 (sm_100) | layout type 2 = SWIZZLE_128B        (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), both K-major, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int n, int m = TBM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- epilogue helpers -------------------------------------------------------------------------------
// Per-warp staging tile in shared memory: 32 rows x 64 bytes (one 32-column bf16 chunk), 16-byte units
// XOR-swizzled so that both the row-per-thread access and the 4-lanes-per-row access are conflict free.
__device__ __forceinline__ int stg_index(int row, int unit) { return row * 4 + (unit ^ ((row >> 1) & 3)); }

// registers (thread = row, 32 bf16 packed in w[16]) -> staging tile in the 64B-swizzle layout the TMA store expects
__device__ __forceinline__ void stage_write(uint4* stg, const uint32_t (&w)[16], int lane) {
#pragma unroll
  for (int u = 0; u < 4; ++u) stg[stg_index(lane, u)] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
}

// Residual tile of one 32x32 chunk.  Loads are coalesced (4 lanes x 16 B per row) and *issued one chunk ahead*
// (software pipelining in registers), so their HBM/L2 latency overlaps the previous chunk's math and stores.
struct ResidualRegs {
  uint4 h[4], l[4];
};
// COHERENT: the residual was written earlier in this same launch by another SM (the chained tail kernel): read it
// through L2 (ld.global.cg) -- the non-coherent path may serve a stale L1 line of a recycled activation buffer.
template <bool COHERENT = false>
__device__ __forceinline__ void residual_issue(ResidualRegs& rr, const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld, int col, int lane,
                                               int m_base, int M) {
  const int u = lane & 3;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int row = m_base + pass * 8 + (lane >> 2);
    rr.h[pass] = make_uint4(0, 0, 0, 0);
    rr.l[pass] = make_uint4(0, 0, 0, 0);
    if (row < M) {
      const int64_t o = (int64_t)row * ld + col + u * 8;
      rr.h[pass] = COHERENT ? __ldcg(reinterpret_cast<const uint4*>(hi + o)) : __ldg(reinterpret_cast<const uint4*>(hi + o));
      if (lo != nullptr) rr.l[pass] = COHERENT ? __ldcg(reinterpret_cast<const uint4*>(lo + o)) : __ldg(reinterpret_cast<const uint4*>(lo + o));
    }
  }
}
// Epilogue arithmetic on pairs of adjacent columns with the packed fp32x2 ALU instructions of sm_100 (FADD2 / FMUL2 /
// FFMA2: two IEEE round-to-nearest results per issue slot, bit-identical to the scalar forms).  The epilogue warps are
// issue-bound (2 warps per scheduler, ~315 instructions per 32-column chunk), so instructions saved are cycles saved.
__device__ __forceinline__ float2 bias_lrelu2(uint32_t r0, uint32_t r1, float b0, float b1, float2 slope2) {
  const float2 x = __fadd2_rn(make_float2(__uint_as_float(r0), __uint_as_float(r1)), make_float2(b0, b1));
  const float2 t = __fmul2_rn(x, slope2);
  return make_float2(fmaxf(x.x, t.x), fmaxf(x.y, t.y));      // LeakyReLU for 0 < slope <= 1 (slope == 1: identity)
}
// fp32 pair -> bf16 hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split_bf16x2(float2 v, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 hh = __floats2bfloat162_rn(v.x, v.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  const float2 d = __ffma2_rn(__bfloat1622float2(hh), make_float2(-1.f, -1.f), v);      // v - hi, one rounding
  const __nv_bfloat162 ll = __floats2bfloat162_rn(d.x, d.y);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}

// registers (4 lanes per row) -> swizzled staging tile -> registers (thread = row), accumulated into v[32]
__device__ __forceinline__ void residual_consume(uint4* stg, const ResidualRegs& rr, bool has_lo, int lane, float (&v)[32]) {
  const int u = lane & 3;
#pragma unroll
  for (int plane = 0; plane < 2; ++plane) {
    if (plane == 1 && !has_lo) break;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) stg[stg_index(pass * 8 + (lane >> 2), u)] = plane ? rr.l[pass] : rr.h[pass];
    __syncwarp();
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
      const uint4 val = stg[stg_index(lane, uu)];
      const uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __fadd2_rn(make_float2(v[uu * 8 + 2 * j], v[uu * 8 + 2 * j + 1]),
                                    __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j])));
        v[uu * 8 + 2 * j] = f.x;
        v[uu * 8 + 2 * j + 1] = f.y;
      }
    }
    __syncwarp();
  }
}


}  // namespace r3d
