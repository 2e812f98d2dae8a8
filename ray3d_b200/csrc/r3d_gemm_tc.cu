// Grouped split-precision GEMM on the Blackwell tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   out[p][m, n] = res[p][m, n] + lrelu( sum_k A[p][m, k] * W[p][n, k] + bias[p][n] )
//
// This is the same contraction as r3d_gemm_ffma.cu (F.conv1d with stride == kernel / F.linear, eval
// BatchNorm folded, LeakyReLU, residual: lib/model/rie.py:86-99, 122-135, 159-169), but every fp32
// operand is carried as two bf16 planes (x = hi + lo, 16 mantissa bits) and the product is formed
// as hi*hi + hi*lo + lo*hi with fp32 accumulation in tensor memory (R3D_PREC_BF16X3, ~1e-5 normwise
// end to end), or hi*hi only (R3D_PREC_BF16).
//
// Structure (one persistent CTA per SM -- a CTA pair per tile in 2-SM mode --, 384 threads; static round-robin unit
// walk, or -- DYN, multi-wave 2-SM launches -- units claimed with atomicAdd and published through a shared-memory queue):
//   warp 0       TMA producer: cp.async.bulk.tensor 2D loads of the A / W planes of one 64-wide K block into a ring of
//                128B-swizzled smem stages, completion on "full" mbarriers; K blocks whose 16-column steps all carry
//                zero weights (GemmProb::kmask) are never staged; L2 hints per operand class
//   warp 1       MMA issuer: one thread issues tcgen05.mma (M=128 or 256 with cta_group::2, N=BLOCK_N, K=16, bf16 -> fp32)
//                for the 1 or 3 products of every non-empty K step; tcgen05.commit releases the smem stage ("empty") and,
//                after the last K block, publishes the accumulator ("tmem_full")
//   warp 2       allocates / frees the TMEM columns (two accumulator stages; acc1/Y + acc2 in the fused conv pair)
//   warps 2, 3   lane 0: store threads -- one TMA tensor store per destination for the 128-row staging tile that the four
//                epilogue warps of a column group have filled (sready / sfree mbarriers)
//   warps 4..11  epilogue: tcgen05.ld (32 lanes x 32 columns per warp and chunk), bias + LeakyReLU + residual in registers
//                (packed fp32x2 ALU operations), re-split to bf16 hi/lo, 64B-swizzled staging tiles in shared memory
// so the epilogue of tile i overlaps the main loop of tile i+1.  FUSED: see gemm_tc_kernel below.
#include <cstdlib>
#include <cstring>

#include "r3d_tc_common.cuh"

namespace r3d {

constexpr int kOpSmemBytes = (sizeof(GemmOpDev) + 256 + 1023) / 1024 * 1024 - 256;   // keeps the staging tiles 1024-byte aligned
// barriers, descriptor, per-warp hi/lo store staging tiles (double buffered in 2-SM mode, where the W half-tiles leave room)
// lean (the fused conv pair in 2-SM mode): ONE staging set per column half and the folded bias broadcast from registers
// (warp shuffles) instead of a shared-memory copy -- its tiles are MMA-bound, the epilogue has slack -- which frees
// exactly the 36 KB a third operand stage needs (3 x 64 KB + 35 KB = 227 KB)
__host__ __device__ constexpr int tc_aux_bytes(int cl, bool lean = false) {
  return 256 + kOpSmemBytes + 8 * 4096 * ((cl == 2 && !lean) ? 2 : 1) + ((cl == 2 && !lean) ? EPI_WARPS * 512 : 0);
}

// per-CTA bytes of one K block: A tile (128 rows) + this CTA's share of the W tile (all of it, or half in 2-SM mode)
__host__ __device__ constexpr int tc_stage_bytes(int block_n, int nsplit, int cl = 1) { return nsplit * (TBM + block_n / cl) * TBK * 2; }
__host__ __device__ constexpr int tc_num_stages(int block_n, int nsplit, int cl = 1, bool lean = false) {
  int s = (SMEM_LIMIT - tc_aux_bytes(cl, lean)) / tc_stage_bytes(block_n, nsplit, cl);
  return s > 6 ? 6 : s;
}
__host__ __device__ constexpr int tc_tmem_cols(int block_n) {
  int c = 2 * block_n;
  return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512;
}

struct TileCoord {
  int p, m0, n0;
};
// Work unit -> (problem, m tile, n tile), m-major: all problems' tiles of one row block are adjacent in the walk, so
// launches whose problems share the A operand (the folded first layer: 6 problems read the same rows) re-use it
// from L2 instead of streaming it from HBM once per problem (measured: 575 MB -> DRAM reads for an 85 MB operand when
// the walk was problem-major, because the 510 MB output stream flushes L2 in between).  With clusters a unit covers
// `cl` consecutive m tiles (one per CTA of the pair).
// The decode runs on the epilogue warps' critical path once per tile (next-tile look-ahead), so it is division- and
// branch-free: per_m is inverted once per thread (exact for unit * per_m < 2^32), and the problem of a unit is found by
// comparing against the packed prefix sums of the problems' n-tile counts (one byte each; <= 16 tiles x 6 problems).
struct TileDecoder {
  uint32_t magic;       // ceil(2^32 / per_m) (0: per_m == 1)
  uint32_t per_m;
  uint64_t cum;         // byte p: n tiles of problems [0, p); bytes >= nprob: 255
  int flip;             // reversed walk: total - 1, else -1
  __device__ __forceinline__ void init(const GemmOpDev& op, int per_m_, int block_n, int total) {
    per_m = (uint32_t)per_m_; magic = per_m_ >= 2 ? 0xFFFFFFFFu / (uint32_t)per_m_ + 1u : 0u;
    flip = op.reverse ? total - 1 : -1;
    cum = 0;
    uint32_t c = 0;
    static_assert(kMaxProb <= 8, "prefix bytes");
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      cum |= (uint64_t)(p < op.nprob ? c : 255u) << (8 * p);
      if (p < kMaxProb && p < op.nprob) c += (uint32_t)(op.prob[p].n_pad / block_n);
    }
  }
  template <int BLOCK_N, int CL>
  __device__ __forceinline__ TileCoord get(int unit, int rank) const {
    if (flip >= 0) unit = flip - unit;           // per_m = total / m_groups: sum over problems of their n tiles
    const uint32_t mg = magic ? __umulhi((uint32_t)unit, magic) : (uint32_t)unit;
    const uint32_t rem = (uint32_t)unit - mg * per_m;
    const uint32_t lo = (uint32_t)cum, hi = (uint32_t)(cum >> 32);
    int p = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i) p += rem >= ((lo >> (8 * i)) & 0xFFu) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) p += rem >= ((hi >> (8 * i)) & 0xFFu) ? 1 : 0;
    const uint32_t base = (uint32_t)(cum >> (8 * p)) & 0xFFu;
    return TileCoord{p, (int)(mg * CL + rank) * TBM, (int)(rem - base) * BLOCK_N};
  }
};

// Diagnostics (R3D_TC_DEBUG bit 32 / r3d_debug_tc_trace): CTA 0 records SM clock stamps per tile for its producer (role 0),
// MMA thread (role 1) and first epilogue warp (role 2): [role][tile index < 64][event < 8].
constexpr int kTraceTiles = 64, kTraceEvents = 8;
__device__ long long g_tc_trace[4 * kTraceTiles * kTraceEvents];   // role 3: store thread of column half 0
#ifdef R3D_TC_TRACE      // built by `R3D_BUILD_TRACE=1 python -m ray3d_b200.build --force`; the stamps cost ~5 % of the epilogue
#define R3D_TRACE(role, ti, ev) \
  do { if (trace && (ti) < kTraceTiles) g_tc_trace[((role) * kTraceTiles + (ti)) * kTraceEvents + (ev)] = clock64(); } while (0)
constexpr bool kTraceBuilt = true;
#else
#define R3D_TRACE(role, ti, ev) do { } while (0)
constexpr bool kTraceBuilt = false;
#endif

// FUSED: every tile runs two GEMMs back to back -- acc1 = A*W^T (K = w*C), Y = lrelu(acc1 + bias) re-split to bf16
// hi/lo IN PLACE in tensor memory (each 32-column fp32 chunk becomes 16 hi + 16 lo packed columns), acc2 = Y*W2^T with
// the A operand read from tensor memory (TS-mode tcgen05.mma), then the usual epilogue on acc2.  TMEM: columns
// [0,256) acc1/Y, [256,512) acc2.  Needs BLOCK_N == 256 == channels.
// DYN: work units after each cluster's first are claimed with atomicAdd on the launch's counter and published to the other
// roles through a small shared-memory queue, instead of the static round-robin walk.  Used for the multi-wave 2-SM
// launches (first layer, fused levels): while the GlobalInfo chain holds a few SMs for hundreds of microseconds (chained
// launch on the side stream) the CTAs of those launches that cannot be resident yet simply take no units -- a static
// share would make the launch wait for them.  Short launches keep the static walk (the hand-offs cost small-batch latency).
template <int BLOCK_N, int NSPLIT, int CL, bool FUSED, bool DYN = false>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const GemmOpDev* __restrict__ opp, const CUtensorMap* __restrict__ tmaps,
                                                                   int M, int total_tiles, int dbg) {
  constexpr int NTHREADS = TC_THREADS;
  constexpr int EW = EPI_WARPS;
  // CL == 2: the CTA pair works as one 256-row tile with cta_group::2 MMAs; each CTA stages its own 128 A rows and
  // HALF of the W tile (the tensor cores read the other half from the peer's shared memory), which cuts the bytes
  // every SM has to receive per MMA by a third -- the measured limiter (~74 GB/s per SM from L2) -- and buys a third
  // pipeline stage.
  constexpr bool LEAN = FUSED && CL == 2;
  constexpr int STAGES = tc_num_stages(BLOCK_N, NSPLIT, CL, LEAN);
  constexpr int A_BYTES = TBM * TBK * 2, W_BYTES = (BLOCK_N / CL) * TBK * 2;
  constexpr int STAGE_BYTES = tc_stage_bytes(BLOCK_N, NSPLIT, CL);
  constexpr int TMEM_COLS = tc_tmem_cols(BLOCK_N);
  constexpr int CH = BLOCK_N >= 32 ? 32 : 16;         // epilogue column chunk
  constexpr int NCHUNK = BLOCK_N / CH;
  constexpr int COL_SPLIT = NCHUNK >= EW / 4 ? EW / 4 : (NCHUNK >= 2 ? 2 : 1);   // epilogue warps sharing a TMEM lane quarter split the columns
  constexpr int CHUNKS_PER_WARP = NCHUNK / COL_SPLIT;
  // Column groups are interleaved chunk by chunk (group g converts chunks g, g + COL_SPLIT, ...): the groups then write
  // neighbouring 64-byte pieces of the same output rows at about the same time, so L2 sees whole 128-byte lines complete
  // quickly instead of half lines waiting a thousand cycles for their other half.
  auto chunk_index = [](int grp, int cc) { return cc * COL_SPLIT + grp; };
  static_assert(STAGES >= 2, "need at least a double-buffered smem ring");
  static_assert(!FUSED || BLOCK_N == 256, "the fused conv pair keeps a full 256-channel row per tile");

  // NB: every pointer below is derived from smem_raw by constant offsets so the compiler keeps the shared address
  // space (LDS/STS); round-tripping through uintptr_t to align by hand degrades them to generic LD/ST.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();      // 128B-swizzled operand tiles need a 1024-byte aligned base
  uint8_t* aux = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint64_t* yready_bar = bars + 2 * STAGES + 4; // [4]  FUSED: channels [64 k, 64 k + 64) of the intermediate (K block k of the second GEMM) are in tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
  constexpr int SQ = 4;                                           // DYN: unit queue depth (roles are at most ~2 tiles apart)
  uint64_t* sq_full = bars + 2 * STAGES + 9;                      // [SQ] unit id published
  uint64_t* sq_empty = sq_full + SQ;                              // [SQ] every consumer (of both CTAs) has read it; the leader's is used
  volatile uint32_t* sq_tile = reinterpret_cast<volatile uint32_t*>(sq_empty + SQ);   // [SQ]
  static_assert((2 * 6 + 9 + 2 * SQ) * 8 + SQ * 4 <= 256, "barrier block overflows its 256 bytes");
  GemmOpDev* sop = reinterpret_cast<GemmOpDev*>(aux + 256);                        // op descriptor, smem resident
  // epilogue warp <-> store thread hand-off, per epilogue warp and staging set: "staged tile ready" / "staging set free"
  static_assert(sizeof(GemmOpDev) % 8 == 0 && 256 + sizeof(GemmOpDev) + 16 * 8 <= 256 + kOpSmemBytes, "no room for the store barriers");
  uint64_t* sready_bar = reinterpret_cast<uint64_t*>(aux + 256 + sizeof(GemmOpDev));   // [column half][staging set], 4 arrivals
  uint64_t* sfree_bar = sready_bar + 8;                                                 // [column group][staging set]
  uint4* stage_s = reinterpret_cast<uint4*>(aux + 256 + kOpSmemBytes);            // [EPI_WARPS][sets][2 planes][32 rows x 64 B]
  float* bias_s = reinterpret_cast<float*>(aux + 256 + kOpSmemBytes + 8 * 4096 * (CL == 2 ? 2 : 1));   // CL == 2 && !LEAN: [EPI_WARPS][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = ((M + TBM - 1) / TBM + CL - 1) / CL;      // m-tile groups (CL tiles each)
  const int per_m = total_tiles / m_tiles;                      // tiles per m group (all problems' n tiles)
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;      // static walk; DYN: unit0 = first unit, the rest is claimed
  TileDecoder decode_unit;                                             // initialised once the descriptor is in shared memory
  // DYN queue consumers per CTA: MMA thread (leader) or TMA producer (peer), the active store threads, 8 epilogue warps
  constexpr int SQ_STORE = CH == 32 ? (COL_SPLIT >= 2 ? 2 : 1) : 0;
  constexpr int SQ_CONSUMERS = (1 + SQ_STORE + EW) * CL;
  constexpr int W_PART_ROWS = BLOCK_N / CL;                      // W rows this CTA stages
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  const bool leader = crank == 0;
  const bool trace = kTraceBuilt && R3D_DBG(32) && blockIdx.x == 0 && lane == 0;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel may start its own setup early
  {   // descriptor -> shared memory (read hundreds of times per tile by the epilogue)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(opp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sop);
    for (int i = threadIdx.x; i < (int)(sizeof(GemmOpDev) / 4); i += NTHREADS) dst[i] = __ldg(src + i);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                               // CL == 2: only the leader's is used (both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EW * CL);                // CL == 2: the peer's epilogue warps arrive remotely
    }
    for (int a = 0; a < 4; ++a) mbar_init(&yready_bar[a], EW * CL);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&sready_bar[i], 4);                             // the four epilogue warps (TMEM lane quarters) of a column group
      mbar_init(&sfree_bar[i], 1);
    }
    if (DYN)
      for (int i = 0; i < SQ; ++i) {
        mbar_init(&sq_full[i], 1);
        mbar_init(&sq_empty[i], SQ_CONSUMERS);                  // CL == 2: the peer's consumers arrive remotely on the leader's
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CL == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {   // same warp id and same destination offset in both CTAs of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  // Programmatic dependent launch: everything above (descriptor copy, barrier init, TMEM allocation) only touches
  // launch-invariant data and overlaps the tail of the previous kernel in the stream; activations written by that
  // kernel are first touched below.
  const GemmOpDev& op = *sop;
  decode_unit.init(op, per_m, BLOCK_N, total_tiles);
  if (lane == 0 && warp < 4 && warp != 1 && !R3D_DBG(32768)) {
    // tensor maps of this CTA's first unit (operand loads: warp 0, stores: warps 2/3) into the descriptor cache while the
    // previous kernel drains -- their first use sits on the first tile's critical path otherwise
    const TileCoord t0 = decode_unit.get<BLOCK_N, CL>(unit0 < total_tiles ? unit0 : 0, crank);
    const CUtensorMap* tm = tmaps + t0.p * kTmapsPerProb;
    if (warp == 0) {
      for (int i = 0; i < 6; ++i) tmap_prefetch(tm + i);
    } else {
      for (int t = warp - 2; t < 2 * op.prob[t0.p].ndst; t += 2) tmap_prefetch(tm + 6 + t);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // DYN: consumer side of the unit queue; every consuming role pops every unit, in order (qc = pops so far)
  uint32_t qc = 0;
  auto sq_pop = [&]() -> int {
    const int slot = (int)(qc % SQ);
    const uint32_t ph = (qc / SQ) & 1u;
    if (CL == 2 && !leader) mbar_wait_cluster(&sq_full[slot], ph);     // written by the leader CTA's scheduler thread
    else mbar_wait(&sq_full[slot], ph);
    const int t = (int)sq_tile[slot];
    ++qc;
    return t;
  };
  auto sq_release = [&](uint32_t popped) {                               // one thread per role, after all its lanes have read the slot
    const int slot = (int)((popped - 1) % SQ);
    if (CL == 2 && !leader) mbar_arrive_cluster(&sq_empty[slot], 0);
    else mbar_arrive(&sq_empty[slot]);
  };
  // next unit of a single-thread role (MMA issuer, store threads, the peer CTA's producer): prev < 0 = first
  auto next_unit = [&](int prev) -> int {
    if (!DYN) return prev < 0 ? unit0 : prev + unit_step;
    const int t = sq_pop();
    sq_release(qc);
    return t;
  };
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int ti = 0;
      bool shared_a = op.nprob > 1 && !R3D_DBG(1024);
      for (int p = 1; p < op.nprob; ++p) shared_a = shared_a && op.prob[p].a.p0 == op.prob[0].a.p0;
      // DYN scheduler (leader CTA): publish a unit, claim the next one (atomicAdd in flight while this unit's loads are issued)
      uint32_t qn = 0;
      const int n_static = (int)gridDim.x / CL;
      auto publish = [&](int tile) {
        const int slot = (int)(qn % SQ);
        mbar_wait(&sq_empty[slot], ((qn / SQ) & 1u) ^ 1u);        // every consumer of both CTAs has read the slot's previous unit
        sq_tile[slot] = (uint32_t)tile;
        if (CL == 2) {
          st_shared_cluster_u32(&sq_tile[slot], 1, (uint32_t)tile);
          mbar_arrive_cluster(&sq_full[slot], 1, true);           // release.cluster: the peer's roles see the id
        }
        mbar_arrive(&sq_full[slot]);
        ++qn;
      };
      int claimed = unit0;                                        // (first unit: the cluster's own index, no L2 round trip)
      for (int tile = DYN && !leader ? next_unit(-1) : unit0;; ++ti) {
        if (DYN && leader) publish(tile);
        if (tile >= total_tiles) break;
        if (DYN && leader) claimed = n_static + (int)atomicAdd(op.sched, 1u);
        const TileCoord tc = decode_unit.get<BLOCK_N, CL>(tile, crank);
        const CUtensorMap* tm = tmaps + tc.p * kTmapsPerProb;
        const int nkb = op.prob[tc.p].K / TBK;
        const uint64_t kmask = op.prob[tc.p].kmask ? op.prob[tc.p].kmask : ~0ull;
        // activations are streamed once (evict first) -- unless every problem of the launch reads the SAME operand (the
        // first layer): then it must survive in L2 from one problem's tile to the next while the output stream passes by
        // ... or the problem has several column tiles (the FC layers): each of them reads the operand again
        const uint64_t a_hint = shared_a ? kEvictLast : ((op.prob[tc.p].n_pad > BLOCK_N && !R3D_DBG(8192)) ? kEvictNormal : kEvictFirst);
        // the residual of a strided conv is the middle tap of its own operand (rie.py:94): those K blocks are read again
        // by the epilogue ~one tile later -> keep them out of the evict-first class so the second read hits L2
        const GemmProb& pp = op.prob[tc.p];
        const bool res_in_a = pp.res.p0 != nullptr && pp.res.p0 == pp.a.p0 && !R3D_DBG(4096);
        const int res_k0 = pp.res_col, res_k1 = pp.res_col + pp.n_pad;
        R3D_TRACE(0, ti, 0);
        for (int kb = 0; kb < nkb; ++kb) {
          if (((kmask >> (kb * (TBK / UMMA_K))) & 0xFull) == 0) continue;      // K block without weights: never staged
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          const uint64_t a_hint_kb = (res_in_a && kb * TBK < res_k1 && (kb + 1) * TBK > res_k0) ? kEvictNormal : a_hint;
          if (CL == 1) {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(st, tm + 0, &full_bar[stage], kb * TBK, tc.m0, a_hint_kb);
            if (NSPLIT == 2) tma_load_2d(st + A_BYTES, tm + 1, &full_bar[stage], kb * TBK, tc.m0, a_hint_kb);
            tma_load_2d(st + NSPLIT * A_BYTES, tm + 2, &full_bar[stage], kb * TBK, tc.n0, kEvictLast);
            if (NSPLIT == 2) tma_load_2d(st + 2 * A_BYTES + W_BYTES, tm + 3, &full_bar[stage], kb * TBK, tc.n0, kEvictLast);
          } else {
            // both CTAs' loads complete on the leader's barrier, which the (leader-only) MMA thread waits on
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            const int wrow = tc.n0 + crank * W_PART_ROWS;
            tma_load_2d_2sm(st, tm + 0, &full_bar[stage], kb * TBK, tc.m0, a_hint_kb);
            if (NSPLIT == 2) tma_load_2d_2sm(st + A_BYTES, tm + 1, &full_bar[stage], kb * TBK, tc.m0, a_hint_kb);
            tma_load_2d_2sm(st + NSPLIT * A_BYTES, tm + 4, &full_bar[stage], kb * TBK, wrow, kEvictLast);
            if (NSPLIT == 2) tma_load_2d_2sm(st + 2 * A_BYTES + W_BYTES, tm + 5, &full_bar[stage], kb * TBK, wrow, kEvictLast);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        R3D_TRACE(0, ti, 1);
        if (FUSED) {   // second GEMM: only W2 K blocks flow through the ring (its A operand is in tensor memory)
          const CUtensorMap* tw = tm + kTmapW2;
          const int nkb2 = op.prob[tc.p].K2 / TBK;
          for (int kb = 0; kb < nkb2; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * STAGE_BYTES;
            if (CL == 1) {
              mbar_expect_tx(&full_bar[stage], NSPLIT * W_BYTES);
              tma_load_2d(st + NSPLIT * A_BYTES, tw + 0, &full_bar[stage], kb * TBK, 0, kEvictLast);
              if (NSPLIT == 2) tma_load_2d(st + 2 * A_BYTES + W_BYTES, tw + 1, &full_bar[stage], kb * TBK, 0, kEvictLast);
            } else {
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * NSPLIT * W_BYTES);
              const int wrow = crank * W_PART_ROWS;
              tma_load_2d_2sm(st + NSPLIT * A_BYTES, tw + 2, &full_bar[stage], kb * TBK, wrow, kEvictLast);
              if (NSPLIT == 2) tma_load_2d_2sm(st + 2 * A_BYTES + W_BYTES, tw + 3, &full_bar[stage], kb * TBK, wrow, kEvictLast);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        tile = !DYN ? tile + unit_step : (leader ? claimed : next_unit(tile));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N, TBM * CL);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      int ti = 0;
      for (int tile = next_unit(-1); tile < total_tiles; tile = next_unit(tile), ++ti) {
        const TileCoord tc = decode_unit.get<BLOCK_N, CL>(tile, crank);
        const int nkb = op.prob[tc.p].K / TBK;
        const uint64_t kmask = op.prob[tc.p].kmask ? op.prob[tc.p].kmask : ~0ull;
        R3D_TRACE(1, ti, 0);
        int last_kb = nkb - 1;                                  // last K block that is staged at all (the mask is never empty)
        while (last_kb > 0 && ((kmask >> (last_kb * (TBK / UMMA_K))) & 0xFull) == 0) --last_kb;
        uint32_t accumulate = 0;                                // first MMA of the tile overwrites the accumulator
        if (!FUSED) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);        // epilogue has drained this accumulator
          tc_fence_after();
        }   // FUSED: acc1 is free again once the previous tile's second GEMM was issued (same thread, in order)
        const uint32_t d_tmem = tmem_base + (FUSED ? 0 : acc * BLOCK_N);
        R3D_TRACE(1, ti, 1);
        for (int kb = 0; kb <= last_kb; ++kb) {
          const uint32_t steps = (uint32_t)(kmask >> (kb * (TBK / UMMA_K))) & 0xFu;   // K steps of this block with weights
          if (steps == 0) continue;                            // the producer skipped it too
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0) R3D_TRACE(1, ti, 2);
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = make_smem_desc(st), w_hi = make_smem_desc(st + NSPLIT * A_BYTES);
          const uint64_t a_lo = make_smem_desc(st + A_BYTES), w_lo = make_smem_desc(st + 2 * A_BYTES + W_BYTES);
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            if (!((steps >> k) & 1u)) continue;
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);     // 32 bytes per K step inside the swizzle atom
            if (CL == 1) {
              umma_bf16(d_tmem, a_hi + koff, w_hi + koff, idesc, accumulate);
              if (NSPLIT == 2) {
                umma_bf16(d_tmem, a_hi + koff, w_lo + koff, idesc, 1);
                umma_bf16(d_tmem, a_lo + koff, w_hi + koff, idesc, 1);
              }
            } else {
              umma_bf16_2sm(d_tmem, a_hi + koff, w_hi + koff, idesc, accumulate);
              if (NSPLIT == 2) {
                umma_bf16_2sm(d_tmem, a_hi + koff, w_lo + koff, idesc, 1);
                umma_bf16_2sm(d_tmem, a_lo + koff, w_hi + koff, idesc, 1);
              }
            }
            accumulate = 1;
          }
          const int tf = FUSED ? 0 : acc;
          if (CL == 1) {
            umma_commit(&empty_bar[stage]);                    // smem stage free once these MMAs retire
            if (kb == last_kb) umma_commit(&tfull_bar[tf]);    // accumulator complete
          } else {                                             // ... signalled in both CTAs of the pair
            umma_commit_2sm(&empty_bar[stage], MC_MASK);
            if (kb == last_kb) umma_commit_2sm(&tfull_bar[tf], MC_MASK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        R3D_TRACE(1, ti, 3);
        if (FUSED) {
          // second GEMM: Y (bf16 hi/lo, written in place over acc1 by the epilogue warps) x W2^T -> acc2
          // The epilogue warps convert Y K block by K block (64 channels: the two warps of a TMEM lane quarter take one
          // 32-channel chunk each): K block k of this GEMM starts as soon as its channels are there, while the later ones
          // are still being converted.
          if (op.flags & 1)
            for (int k = 0; k < 4; ++k) mbar_wait(&yready_bar[k], acc_phase);
          mbar_wait(&tempty_bar[1], acc_phase ^ 1);           // previous tile's epilogue has drained acc2
          tc_fence_after();
          const uint32_t d2 = tmem_base + BLOCK_N;
          const int nkb2 = op.prob[tc.p].K2 / TBK;
          for (int kb = 0; kb < nkb2; ++kb) {
            mbar_wait(&yready_bar[kb & 3], acc_phase);        // channels [64 kb, 64 kb + 64) of Y complete in both CTAs
            tc_fence_after();
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t w_hi = make_smem_desc(st + NSPLIT * A_BYTES), w_lo = make_smem_desc(st + 2 * A_BYTES + W_BYTES);
#pragma unroll
            for (int k = 0; k < TBK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
              const int sidx = kb * (TBK / UMMA_K) + k;                       // K step = Y channels [16 s, 16 s + 16)
              const uint32_t y_hi = tmem_base + 32 * (sidx >> 1) + 8 * (sidx & 1), y_lo = y_hi + 16;
              if (CL == 1) {
                umma_bf16_ts(d2, y_hi, w_hi + koff, idesc, (kb | k) != 0);
                if (NSPLIT == 2) {
                  umma_bf16_ts(d2, y_hi, w_lo + koff, idesc, 1);
                  umma_bf16_ts(d2, y_lo, w_hi + koff, idesc, 1);
                }
              } else {
                umma_bf16_ts_2sm(d2, y_hi, w_hi + koff, idesc, (kb | k) != 0);
                if (NSPLIT == 2) {
                  umma_bf16_ts_2sm(d2, y_hi, w_lo + koff, idesc, 1);
                  umma_bf16_ts_2sm(d2, y_lo, w_hi + koff, idesc, 1);
                }
              }
            }
            if (CL == 1) {
              umma_commit(&empty_bar[stage]);
              if (kb == nkb2 - 1) umma_commit(&tfull_bar[1]);
            } else {
              umma_commit_2sm(&empty_bar[stage], MC_MASK);
              if (kb == nkb2 - 1) umma_commit_2sm(&tfull_bar[1], MC_MASK);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          R3D_TRACE(1, ti, 4);
          acc_phase ^= 1;
        } else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < EPI_WARP0) {
    // =============================== store threads ===============================
    // Lane 0 of warp 2 / warp 3 issues the TMA tensor stores for the epilogue warps of column half 0 / 1.  Issuing a
    // tensor store costs its thread ~300 cycles (more under write back-pressure); inside the epilogue warps that was
    // ~600 cycles per 32-column chunk on their critical path.  The four warps of a column half (one TMEM lane quarter
    // each) stage their 32-row slices into ONE 128-row x 32-column tile per plane, so a single store per destination
    // plane moves all of it (4x fewer store instructions), and the epilogue warps convert the next chunk meanwhile.
    // Protocol per (column half, staging set): the four warps fill their slices, fence, arrive on sready (count 4);
    // this thread issues the stores, commits, and arrives on sfree once the store engine has read the set.
    constexpr int EPI_BUFS = (CL == 2 && !LEAN) ? 2 : 1;
    constexpr int GPS = COL_SPLIT > 2 ? COL_SPLIT / 2 : 1;      // column groups per store thread (1 with 8 epilogue warps)
    const int st = warp - 2;
    if (CH == 32 && lane == 0 && st >= 0 && st * GPS < COL_SPLIT) {
      uint32_t round = 0;                                       // staged chunks per group so far (the client warps count the same)
      uint64_t* prev_free = nullptr;                            // GPS == 2: set handed back one store later (while the next is read)
      int ti = 0;
      const bool strace3 = trace && st == 0;
      for (int tile = next_unit(-1); tile < total_tiles; tile = next_unit(tile), ++ti) {
        const TileCoord tc = decode_unit.get<BLOCK_N, CL>(tile, crank);
        const GemmProb& pr = op.prob[tc.p];
        const CUtensorMap* dmaps = tmaps + tc.p * kTmapsPerProb + 6;      // [dst][hi, lo] store maps
        bool any_bf = false;
        for (int t = 0; t < pr.ndst; ++t) any_bf |= pr.dst[t].f32 == 0;
        if (!any_bf || R3D_DBG(4)) continue;
        for (int cc = 0; cc < CHUNKS_PER_WARP; ++cc) {
          const int b = EPI_BUFS == 2 ? (int)(round & 1) : 0;
          const uint32_t uses = EPI_BUFS == 2 ? round >> 1 : round;
          bool any = false;
          for (int gi = 0; gi < GPS; ++gi) {
            const int grp = st * GPS + gi;
            const int n = tc.n0 + chunk_index(grp, cc) * CH;
            if (n >= pr.N) continue;
            any = true;
            mbar_wait(&sready_bar[grp * 2 + b], uses & 1);
            if (strace3 && gi == 0 && cc < 4) R3D_TRACE(3, ti, 2 * cc);
            if (GPS == 1 && EPI_BUFS == 2 && !R3D_DBG(16384) && prev_free != nullptr) {
              // the previous store was issued a whole chunk ago: the engine has read its set long since -- hand it back
              // BEFORE this chunk's fence/issue (~600 cycles), the client warps are about to ask for it
              bulk_wait_read0();
              mbar_arrive(prev_free);
              prev_free = nullptr;
            }
            // generic-proxy writes of the four client warps (ordered before this point by their mbarrier arrivals) ->
            // async proxy: ONE proxy fence here, on the causality path between the writes and the tensor stores, instead
            // of one per writing warp (the fence drains the SM's shared-memory pipe: 32 of them per tile serialised the
            // whole epilogue at ~250 cycles each)
            if (!R3D_DBG(512)) fence_async_smem();
            if (!R3D_DBG(1)) {
              const uint4* tile_hi = stage_s + (grp * EPI_BUFS + b) * 1024;
              const int srow = R3D_DBG(16) ? 0 : tc.m0;                     // experiment: keep every store in the same L2-resident rows
              for (int t = 0; t < pr.ndst; ++t) {
                const Dst& d = pr.dst[t];
                if (d.f32) continue;
                if (NSPLIT == 2 && R3D_DBG(2048)) tma_store_3d_hint(dmaps + 2 * t, tile_hi, d.col + n, srow, 0, 0x12F0000000000000ull);   // experiment: evict-first output stream
                else if (NSPLIT == 2) tma_store_3d(dmaps + 2 * t, tile_hi, d.col + n, srow, 0);    // both planes, one store
                else tma_store_2d(dmaps + 2 * t, tile_hi, d.col + n, srow);
              }
            }
            bulk_commit();
            // The set is handed back as soon as the store engine has read it (a few hundred cycles), not one round later:
            // the client warps then never wait for their slowest peer's NEXT chunk before reusing a set.
            if (GPS == 1 && (EPI_BUFS == 1 || R3D_DBG(16384))) {
              bulk_wait_read0();
              if (strace3 && cc < 4) R3D_TRACE(3, ti, 2 * cc + 1);
              mbar_arrive(&sfree_bar[grp * 2 + b]);
            } else {                                            // two sets (or two groups) alternate: TWO stores in flight, wait for the previous one only
              bulk_wait_read1();
              if (strace3 && cc < 4) R3D_TRACE(3, ti, 2 * cc + 1);
              if (prev_free != nullptr) mbar_arrive(prev_free);
              prev_free = &sfree_bar[grp * 2 + b];
            }
          }
          if (any) ++round;
        }
      }
      bulk_wait0();                          // every TMA store issued by this thread has landed
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue ===============================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    const int half = ew >> 2;                                 // column group (half of the columns with 8 epilogue warps)
    const bool active = half < COL_SPLIT;
    constexpr int EPI_BUFS = (CL == 2 && !LEAN) ? 2 : 1;    // staging tile sets per column group (hi + lo each)
    // staging: [column group][set][plane] tiles of 128 rows x 64 B (64B-swizzled): the smem image of a (32 columns, 128 rows,
    // 2 planes) box, so ONE 3-D TMA store writes both planes; this warp owns rows [32 q, 32 q + 32)
    uint4* const stage_base = stage_s + half * EPI_BUFS * 1024;
    uint32_t sround = 0;                                      // chunks this warp has handed to its store thread
    int acc = 0;
    uint32_t acc_phase = 0;
    const float2 slope2 = make_float2(op.slope, op.slope);
    const bool etrace = trace && ew == 0;
    int ti = 0;
    constexpr bool BIAS_SMEM = CL == 2 && !LEAN;     // with (almost) all of L1 carved out as smem every bias LDG is an L2 round trip
    constexpr bool BIAS_REG = LEAN;                   // ... LEAN: this tile's bias words stay in registers (lane = column) and are broadcast by shuffles
    constexpr int PB = CHUNKS_PER_WARP;                           // bias words per lane: one per chunk of this warp
    auto chunk_of = [&](int cc) { return 2 * cc + half; };     // FUSED: this warp's cc-th chunk of the first GEMM = half of K block cc of the second
    // The epilogue warps are the critical path of the short-K launches: the coordinates and the folded bias of the NEXT
    // tile are fetched while the current tile is processed (registers), so no tile starts with an L2 round trip.
    float pb[PB], pa[4];
    auto prefetch_bias = [&](const TileCoord& t) {
      const GemmProb& g = op.prob[t.p];
      const float* bp = (FUSED ? g.bias2 : g.bias) + t.n0;
#pragma unroll
      for (int i = 0; i < PB; ++i) pb[i] = lane < CH ? __ldg(bp + chunk_index(half, i) * CH + lane) : 0.f;   // chunk i of this warp
      if (FUSED) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pa[i] = __ldg(g.bias + chunk_of(i) * 32 + lane);
      }
    };
    // DYN: every lane pops (reads the slot), lane 0 hands the slot back once the whole warp has read it
    auto warp_pop = [&]() -> int {
      const int t = sq_pop();
      __syncwarp();
      if (lane == 0) sq_release(qc);
      return t;
    };
    int tile = DYN ? warp_pop() : unit0;
    TileCoord tc = decode_unit.get<BLOCK_N, CL>(tile < total_tiles ? tile : 0, crank);
    if ((BIAS_SMEM || BIAS_REG) && active && tile < total_tiles) prefetch_bias(tc);
    uint32_t pflags = 0;          // per problem: bit 0 any fp32 destination, 1 any bf16 destination, 2 any lo plane, 3 residual
    for (int p = 0; p < op.nprob; ++p) {
      const GemmProb& g = op.prob[p];
      uint32_t f = (g.res.p0 != nullptr && !R3D_DBG(2)) ? 8u : 0u;
      for (int t = 0; t < g.ndst; ++t) {
        f |= g.dst[t].f32 != 0 ? 1u : 2u;
        if (g.dst[t].f32 == 0 && g.dst[t].m.p1 != nullptr) f |= 4u;
      }
      pflags |= f << (4 * p);
    }
    for (; tile < total_tiles; ++ti) {
      const GemmProb& pr = op.prob[tc.p];
      if (etrace) R3D_TRACE(2, ti, 0);
      TileCoord tn = tc;
      int next_tile = DYN ? total_tiles : tile + unit_step;
      bool next_ready = false;                                   // next tile decoded + its bias requested (done inside the chunk loop)
      // the next tile's coordinates and bias are fetched behind the first chunk's TMEM load, where the warp would stall anyway
      // (DYN: the scheduler published the next unit long ago -- before issuing that unit's loads -- so the pop does not wait)
      auto look_ahead = [&]() {
        if (next_ready) return;
        next_ready = true;
        if (DYN) next_tile = warp_pop();
        if (next_tile >= total_tiles) return;
        tn = decode_unit.get<BLOCK_N, CL>(next_tile, crank);
        if ((BIAS_SMEM || BIAS_REG) && active) prefetch_bias(tn);
      };
      const bool ttrace = etrace && R3D_DBG(128);                 // stamps of the per-tile preamble
      if (ttrace) R3D_TRACE(2, ti, 1);
      const int m_base = tc.m0 + q * 32;
      const int row = m_base + lane;
      const bool row_ok = row < M;
      const uint32_t pf = pflags >> (4 * tc.p);
      const bool any_f32 = pf & 1u, any_bf = pf & 2u, any_lo = pf & 4u, has_res = pf & 8u;
      if (ttrace) R3D_TRACE(2, ti, 2);
      const __nv_bfloat16* res_hi = reinterpret_cast<const __nv_bfloat16*>(pr.res.p0);
      const __nv_bfloat16* res_lo = reinterpret_cast<const __nv_bfloat16*>(pr.res.p1);
      ResidualRegs rr;
      if (active && has_res && CH == 32)      // first chunk's residual: in flight while the main loop still runs
        residual_issue(rr, res_hi, res_lo, pr.res.ld, pr.res_col + tc.n0 + chunk_index(half, 0) * CH, lane, m_base, M);
      float* my_bias = bias_s + ew * 128;
      if (FUSED) {
        // ---- epilogue of the first GEMM: acc1 -> Y = lrelu(acc1 + bias) as bf16 hi/lo, in place in tensor memory
        // This warp converts chunk `half` of every 64-channel K block in turn and signals after each: the MMA thread
        // starts the second GEMM on K block 0 while blocks 1..3 are still being converted.
        if (BIAS_SMEM) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) my_bias[lane + 32 * i] = pa[i];
          __syncwarp();
        }
        const float pa0 = pa[0], pa1 = pa[1], pa2 = pa[2], pa3 = pa[3];   // BIAS_REG: this tile's words (pa is refilled for the next tile)
        mbar_wait(&tfull_bar[0], acc_phase);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t r1[32];
        tmem_ld32(ta + chunk_of(0) * 32, r1);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int g = chunk_of(cc);
          float bb[32];
          if (BIAS_REG) {
            const float mine = cc == 0 ? pa0 : cc == 1 ? pa1 : cc == 2 ? pa2 : pa3;
#pragma unroll
            for (int j = 0; j < 32; ++j) bb[j] = __shfl_sync(0xffffffffu, mine, j);
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = BIAS_SMEM ? *reinterpret_cast<const float4*>(my_bias + cc * 32 + j4 * 4)
                                          : __ldg(reinterpret_cast<const float4*>(pr.bias + g * 32) + j4);
              bb[j4 * 4 + 0] = b4.x; bb[j4 * 4 + 1] = b4.y; bb[j4 * 4 + 2] = b4.z; bb[j4 * 4 + 3] = b4.w;
            }
          }
          tmem_ld_wait();
          uint32_t yh[16], yl[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            split_bf16x2(bias_lrelu2(r1[2 * j], r1[2 * j + 1], bb[2 * j], bb[2 * j + 1], slope2), yh[j], yl[j]);
          if (cc + 1 < 4) tmem_ld32(ta + chunk_of(cc + 1) * 32, r1);
          tmem_st16(ta + g * 32, yh);                        // channels [32g, 32g+32) -> 16 packed columns
          if (NSPLIT == 2) tmem_st16(ta + g * 32 + 16, yl);
          {                                                  // this warp's part of K block cc of Y is complete for its rows
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CL == 1) mbar_arrive(&yready_bar[cc]);
              else mbar_arrive_cluster(&yready_bar[cc], 0, (op.flags & 2) != 0);
            }
          }
        }
      }
      const float* const bias_ptr = FUSED ? pr.bias2 : pr.bias;
      const int acc_col = FUSED ? BLOCK_N : acc * BLOCK_N;
      const int fb = FUSED ? 1 : acc;                          // accumulator-full / -empty barrier of this tile
      float pbc[PB];                          // BIAS_REG: this tile's words
#pragma unroll
      for (int i = 0; i < PB; ++i) pbc[i] = pb[i];
      if (BIAS_SMEM && active) {              // this warp's slice of the folded bias (prefetched during the previous tile)
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PB; ++i)
          if (lane < CH) my_bias[i * CH + lane] = pb[i];
        __syncwarp();
        if (ttrace) R3D_TRACE(2, ti, 3);
      }
      if (ttrace) R3D_TRACE(2, ti, 4);
      if (etrace && !ttrace) R3D_TRACE(2, ti, 1);
      mbar_wait(&tfull_bar[fb], acc_phase);
      tc_fence_after();
      if (ttrace) R3D_TRACE(2, ti, 5);
      if (etrace && !ttrace) R3D_TRACE(2, ti, 2);
      if (active) {
        uint32_t r[32];
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc_col;
        if (CH == 32) tmem_ld32(taddr0 + chunk_index(half, 0) * CH, r); else tmem_ld16(taddr0 + chunk_index(half, 0) * CH, r);   // later chunks: issued one ahead
#pragma unroll 1
        for (int cc = 0; cc < CHUNKS_PER_WARP; ++cc) {
          const int n = tc.n0 + chunk_index(half, cc) * CH;
          const bool strace = etrace && R3D_DBG(64) && !R3D_DBG(128) && cc == ((dbg >> 16) & 3);     // sub-step stamps of one chunk (bits 16-17)
          if (strace) R3D_TRACE(2, ti, 1);
          float bb[CH];
          if (BIAS_REG) {
            float mine = pbc[0];
#pragma unroll
            for (int i = 1; i < PB; ++i) mine = cc == i ? pbc[i] : mine;
#pragma unroll
            for (int j = 0; j < CH; ++j) bb[j] = __shfl_sync(0xffffffffu, mine, j);
          } else if (BIAS_SMEM) {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {
              const float4 b4 = *reinterpret_cast<const float4*>(my_bias + cc * CH + j4 * 4);
              bb[j4 * 4 + 0] = b4.x; bb[j4 * 4 + 1] = b4.y; bb[j4 * 4 + 2] = b4.z; bb[j4 * 4 + 3] = b4.w;
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {       // folded bias: warp-uniform 16-byte loads
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_ptr + n) + j4);
              bb[j4 * 4 + 0] = b4.x; bb[j4 * 4 + 1] = b4.y; bb[j4 * 4 + 2] = b4.z; bb[j4 * 4 + 3] = b4.w;
            }
          }
          tmem_ld_wait();
          if (strace) R3D_TRACE(2, ti, 3);
          // staging set of this chunk: probe its "free" barrier now, the answer (a ~150-cycle shared-memory round trip on
          // this warp's in-order critical path) arrives behind the activation math
          const int sbuf = EPI_BUFS == 2 ? (int)(sround & 1) : 0;
          const uint32_t sfree_parity = ((EPI_BUFS == 2 ? sround >> 1 : sround) & 1) ^ 1;
          const bool set_free = CH == 32 && mbar_try_wait(&sfree_bar[half * 2 + sbuf], sfree_parity);
          float v[32];
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) {
            const float2 a = bias_lrelu2(r[2 * j], r[2 * j + 1], bb[2 * j], bb[2 * j + 1], slope2);
            v[2 * j] = a.x;
            v[2 * j + 1] = a.y;
          }
          if (cc + 1 < CHUNKS_PER_WARP) {        // next chunk's accumulator columns: in flight during this chunk's stores
            if (CH == 32) tmem_ld32(taddr0 + chunk_index(half, cc + 1) * CH, r); else tmem_ld16(taddr0 + chunk_index(half, cc + 1) * CH, r);
          }
          if (strace) R3D_TRACE(2, ti, 0);
          if (cc > 0) look_ahead();              // not in chunk 0: the store thread is waiting for that one
          if (strace) R3D_TRACE(2, ti, 2);
          if (n < pr.N && !R3D_DBG(4)) {          // warp-uniform
            // the TMA stores that last used this staging set must have finished reading it (with two sets the store
            // of the previous chunk may still be in flight)
            uint4* const stage_hi = stage_base + sbuf * 1024 + q * 128;   // this warp's rows of the hi-plane tile (also its residual scratch)
            uint4* const stage_lo = stage_hi + 512;                        // lo-plane tile: the next 8 KB
            if (CH == 32) {    // the store thread has released this staging set (its previous stores have read it)
              if (!set_free) mbar_wait(&sfree_bar[half * 2 + sbuf], sfree_parity);
              if (strace) R3D_TRACE(2, ti, 4);
            }
            if (has_res) {
              if (CH == 32) {
                residual_consume(stage_hi, rr, res_lo != nullptr, lane, v);
                if (cc + 1 < CHUNKS_PER_WARP)     // next chunk's residual, one chunk ahead
                  residual_issue(rr, res_hi, res_lo, pr.res.ld, pr.res_col + tc.n0 + chunk_index(half, cc + 1) * CH, lane, m_base, M);
              } else if (row_ok) {
                const int64_t ro = (int64_t)row * pr.res.ld + pr.res_col + n;
                for (int j = 0; j < CH; ++j) {
                  v[j] += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pr.res.p0)[ro + j]);
                  if (pr.res.p1 != nullptr) v[j] += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(pr.res.p1)[ro + j]);
                }
              }
            }
            if (any_f32 && row_ok) {      // network outputs (tiny): direct masked stores
              for (int t = 0; t < pr.ndst; ++t) {
                const Dst& d = pr.dst[t];
                if (!d.f32) continue;
                float* out = reinterpret_cast<float*>(d.m.p0) + (int64_t)row * d.m.ld + d.col + n;
#pragma unroll
                for (int j = 0; j < CH; ++j)
                  if (n + j < pr.N) out[j] = v[j];
              }
            }
            if (any_bf) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < CH / 2; ++j) split_bf16x2(make_float2(v[2 * j], v[2 * j + 1]), hi[j], lo[j]);
              if (CH == 32) {
                // thread = row -> swizzled staging tiles -> one TMA store per destination plane (the TMA engine does
                // the address generation / coalescing; rows past the buffer capacity are clipped by the tensor map,
                // rows in [M, capacity) receive don't-care values nobody reads)
                stage_write(stage_hi, hi, lane);
                if (any_lo) stage_write(stage_lo, lo, lane);
                if (strace) R3D_TRACE(2, ti, 5);
                if (R3D_DBG(512)) fence_async_smem();     // experiment: per-writer fences (the previous scheme)
                __syncwarp();
                if (strace) R3D_TRACE(2, ti, 6);
                if (lane == 0) mbar_arrive(&sready_bar[half * 2 + sbuf]);    // the store thread takes it from here
                ++sround;
              } else if (row_ok) {
                for (int t = 0; t < pr.ndst; ++t) {
                  const Dst& d = pr.dst[t];
                  if (d.f32) continue;
                  const int64_t o = (int64_t)row * d.m.ld + d.col + n;
                  uint4* ph = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.m.p0) + o);
                  ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                  if (d.m.p1 != nullptr) {
                    uint4* pl = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.m.p1) + o);
                    pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                  }
                }
              }
            }
          }
          if (etrace && !R3D_DBG(64 | 128) && cc < 4) R3D_TRACE(2, ti, 3 + cc);
          if (ttrace && cc == 0) R3D_TRACE(2, ti, 6);
          if (strace) R3D_TRACE(2, ti, 7);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (etrace && !R3D_DBG(64)) R3D_TRACE(2, ti, 7);       // (bit 128: stamp 7 = tile done as well)
      if (lane == 0) {
        if (CL == 1) mbar_arrive(&tempty_bar[fb]);
        else mbar_arrive_cluster(&tempty_bar[fb], 0, (op.flags & 2) != 0);         // the MMA thread lives in the leader CTA
      }
      if (FUSED) acc_phase ^= 1;
      else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      look_ahead();                            // (warps without chunks in this tile)
      tc = tn;
      tile = next_tile;
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA exits while a peer may still multicast into its smem
  if (warp == 2) {
    tc_fence_after();
    if (CL == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int encode_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_pitch_elems, uint32_t box_rows,
                     uint32_t box_cols = TBK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  if (base == nullptr) { memset(out, 0, sizeof(*out)); return 0; }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// (32 columns, box_rows rows, 2 planes) box over the hi/lo planes of one destination: the planes live in one allocation
// (workspace slab), so the lo plane is a constant byte offset from the hi plane -> a tensor dimension.
static int encode_planes_3d(CUtensorMap* out, const void* hi, const void* lo, uint64_t inner, uint64_t rows, uint64_t row_pitch_elems,
                            uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  const intptr_t plane = reinterpret_cast<intptr_t>(lo) - reinterpret_cast<intptr_t>(hi);
  if (plane <= 0 || plane % 16) return -6;
  cuuint64_t dims[3] = {inner, rows, 2};
  cuuint64_t strides[2] = {row_pitch_elems * 2, (cuuint64_t)plane};
  cuuint32_t box[3] = {32, box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(hi), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int tc_block_n(const GemmOpDev& h) {
#ifdef R3D_EXPERIMENTS
  if (const char* env = getenv("R3D_TC_NTILE")) {     // experiments: force a narrower tile when it divides every problem
    const int bn = atoi(env);
    bool ok = bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256;
    for (int p = 0; ok && p < h.nprob; ++p) ok = h.prob[p].n_pad % bn == 0;
    if (ok && bn <= h.n_tile) return bn;
  }
#endif
  return h.n_tile;
}

int tc_build_tmaps(const GemmOpDev& h, int precision, int64_t cap_rows, void* out_v) {
  static_assert(sizeof(CUtensorMap) == kTmapBytes, "CUtensorMap size");
  CUtensorMap* out = reinterpret_cast<CUtensorMap*>(out_v);
  const int bn = tc_block_n(h);
  for (int p = 0; p < h.nprob; ++p) {
    const GemmProb& g = h.prob[p];
    if (g.K % TBK || g.a.ld % 8 || g.n_pad % bn) return -2;
    for (int t = 0; t < g.ndst; ++t)   // bf16 destinations are written in whole 16/32-column chunks
      if (!g.dst[t].f32 && (g.N % (bn >= 32 ? 32 : 16) || g.dst[t].col % 16 || g.dst[t].m.ld % 16)) return -3;
    if (g.res.p0 && (g.res_col % 16 || g.res.ld % 16)) return -4;
    int rc = encode_2d(out + p * kTmapsPerProb + 0, g.a.p0, (uint64_t)g.a.ld, (uint64_t)cap_rows, (uint64_t)g.a.ld, TBM);
    if (rc) return rc;
    rc = encode_2d(out + p * kTmapsPerProb + 1, precision == R3D_PREC_BF16X3 ? g.a.p1 : nullptr, (uint64_t)g.a.ld, (uint64_t)cap_rows,
                   (uint64_t)g.a.ld, TBM);
    if (rc) return rc;
    rc = encode_2d(out + p * kTmapsPerProb + 2, g.w0, (uint64_t)g.K, (uint64_t)g.n_pad, (uint64_t)g.K, (uint32_t)bn);
    if (rc) return rc;
    rc = encode_2d(out + p * kTmapsPerProb + 3, precision == R3D_PREC_BF16X3 ? g.w1 : nullptr, (uint64_t)g.K, (uint64_t)g.n_pad,
                   (uint64_t)g.K, (uint32_t)bn);
    if (rc) return rc;
    // epilogue store maps: 32-column x 128-row boxes of every bf16 destination; bf16x3: (32, 128, 2 planes) boxes; 64B swizzle
    for (int t = 0; t < g.ndst; ++t) {
      CUtensorMap* dm = out + p * kTmapsPerProb + 6 + 2 * t;
      if (g.dst[t].f32 || bn < 32) { memset(dm, 0, 2 * sizeof(*dm)); continue; }
      memset(dm + 1, 0, sizeof(*dm));
      if (precision == R3D_PREC_BF16X3)    // one store writes both planes
        rc = encode_planes_3d(dm, g.dst[t].m.p0, g.dst[t].m.p1, (uint64_t)g.dst[t].m.ld, (uint64_t)cap_rows, (uint64_t)g.dst[t].m.ld, TBM);
      else
        rc = encode_2d(dm, g.dst[t].m.p0, (uint64_t)g.dst[t].m.ld, (uint64_t)cap_rows, (uint64_t)g.dst[t].m.ld, TBM, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
    if (h.fused2) {   // second weight matrix of a fused conv pair: full-tile and half-tile (2-SM) boxes
      CUtensorMap* wm = out + p * kTmapsPerProb + kTmapW2;
      if (bn != 256 || g.K2 % TBK || g.w2_0 == nullptr) return -5;
      rc = encode_2d(wm + 0, g.w2_0, (uint64_t)g.K2, 256, (uint64_t)g.K2, 256);
      if (rc) return rc;
      rc = encode_2d(wm + 1, precision == R3D_PREC_BF16X3 ? g.w2_1 : nullptr, (uint64_t)g.K2, 256, (uint64_t)g.K2, 256);
      if (rc) return rc;
      rc = encode_2d(wm + 2, g.w2_0, (uint64_t)g.K2, 256, (uint64_t)g.K2, 128);
      if (rc) return rc;
      rc = encode_2d(wm + 3, precision == R3D_PREC_BF16X3 ? g.w2_1 : nullptr, (uint64_t)g.K2, 256, (uint64_t)g.K2, 128);
      if (rc) return rc;
    }
    if (h.n_tile_tail > 0) {   // W half tiles at the chained tail launch's unit width (r3d_tail_tc.cu)
      if (g.n_pad % h.n_tile_tail) return -7;
      rc = encode_2d(out + p * kTmapsPerProb + kTmapTailW, g.w0, (uint64_t)g.K, (uint64_t)g.n_pad, (uint64_t)g.K, (uint32_t)h.n_tile_tail / 2);
      if (rc) return rc;
      rc = encode_2d(out + p * kTmapsPerProb + kTmapTailW + 1, precision == R3D_PREC_BF16X3 ? g.w1 : nullptr, (uint64_t)g.K, (uint64_t)g.n_pad,
                     (uint64_t)g.K, (uint32_t)h.n_tile_tail / 2);
      if (rc) return rc;
    }
    if (bn >= 32) {   // half-tile W maps for the 2-SM variant
      rc = encode_2d(out + p * kTmapsPerProb + 4, g.w0, (uint64_t)g.K, (uint64_t)g.n_pad, (uint64_t)g.K, (uint32_t)bn / 2);
      if (rc) return rc;
      rc = encode_2d(out + p * kTmapsPerProb + 5, precision == R3D_PREC_BF16X3 ? g.w1 : nullptr, (uint64_t)g.K, (uint64_t)g.n_pad,
                     (uint64_t)g.K, (uint32_t)bn / 2);
      if (rc) return rc;
    }
  }
  return 0;
}

template <int BN, int NS, int CL = 1, bool LEAN = false>
static constexpr int tc_smem_bytes() { return tc_num_stages(BN, NS, CL, LEAN) * tc_stage_bytes(BN, NS, CL) + tc_aux_bytes(CL, LEAN); }

template <int BN, int NS>
static cudaError_t configure_one() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NS, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<BN, NS>());
  if (e != cudaSuccess) return e;
  constexpr int CL2 = BN >= 32 ? 2 : 1;
  if (BN >= 32) e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NS, CL2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<BN, NS, CL2>());
  if (e != cudaSuccess) return e;
  if constexpr (BN == 128) e = cudaFuncSetAttribute(gemm_tc_kernel<BN, NS, CL2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<BN, NS, CL2>());
  if (e != cudaSuccess) return e;
  if (BN == 256) {   // fused conv-pair variants
    constexpr int FB = BN == 256 ? 256 : 256;
    e = cudaFuncSetAttribute(gemm_tc_kernel<FB, NS, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<FB, NS, 1>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<FB, NS, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<FB, NS, 2, true>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<FB, NS, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<FB, NS, 2, true>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<FB, NS, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<FB, NS, 2>());
  }
  return e;
}

static int g_num_sms = 0;
static int g_dbg = 0;            // R3D_EXPERIMENTS builds only (see R3D_DBG): bit mask of timing experiments, R3D_TC_DEBUG env
static int g_pdl = 1;            // programmatic dependent launch between consecutive GEMMs
static int g_cluster_mode = 1;   // 2-SM tiles (CTA pairs, cta_group::2) whenever the op has >= 2 m tiles
static int g_trace_arm = -1;     // >= 0: the launch that many GEMM launches from now records the per-tile clock trace

void tc_trace_arm(int launches_from_now) { g_trace_arm = launches_from_now; }
cudaError_t tc_trace_read(long long* out, int cap) {
  long long h[4 * kTraceTiles * kTraceEvents];
  cudaError_t e = cudaMemcpyFromSymbol(h, g_tc_trace, sizeof(h));
  if (e != cudaSuccess) return e;
  for (int i = 0; i < cap && i < 4 * kTraceTiles * kTraceEvents; ++i) out[i] = h[i];
  return cudaSuccess;
}

cudaError_t tc_configure() {
  cudaError_t e;
#define R3D_CFG(BN)                                               \
  if ((e = configure_one<BN, 1>()) != cudaSuccess) return e;      \
  if ((e = configure_one<BN, 2>()) != cudaSuccess) return e;
  R3D_CFG(16) R3D_CFG(32) R3D_CFG(64) R3D_CFG(128) R3D_CFG(256)
#undef R3D_CFG
#ifdef R3D_EXPERIMENTS
  // R3D_TC_DEBUG bit mask (results are wrong with bits 1/2/4/16): 1 skip the TMA stores, 2 skip the residual, 4 skip the
  // epilogue's staging and stores, 16 fold all stores onto the same rows, 32 per-tile clock stamps, 64 / 128 stamp one
  // chunk's sub-steps / the per-tile preamble instead, 256 prefetch the next tile's tensor maps, 512 per-writer proxy
  // fences, 1024 shared operand evict-first, 2048 evict-first hint on the output stores, 4096 residual K blocks
  // evict-first, 8192 FC operands evict-first
  if (const char* env = getenv("R3D_TC_CLUSTER")) g_cluster_mode = atoi(env);
  if (const char* env = getenv("R3D_TC_PDL")) g_pdl = atoi(env);
  if (const char* env = getenv("R3D_TC_DEBUG")) g_dbg = atoi(env);
#endif
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
}

int tc_num_sms() { return g_num_sms; }
int tc_debug_mask() { return g_dbg; }
int tc_use_pdl() { return g_pdl; }

template <int BN, int NS>
static cudaError_t launch_one(const GemmOpDev* d_op, const CUtensorMap* d_tmaps, int M, const GemmOpDev& h, cudaStream_t s) {
  const int m_tiles = (M + TBM - 1) / TBM;
  constexpr int CL2 = BN >= 32 ? 2 : 1;
  const bool use_cl = CL2 == 2 && g_cluster_mode && m_tiles >= 2;
  const int cl = use_cl ? 2 : 1;
  const int m_groups = (m_tiles + cl - 1) / cl;
  int units = 0;
  for (int p = 0; p < h.nprob; ++p) units += m_groups * (h.prob[p].n_pad / BN);
  int dbg = g_dbg;               // this launch's debug mask
  if (g_trace_arm >= 0 && g_trace_arm-- == 0) dbg |= 32;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = use_cl ? tc_smem_bytes<BN, NS, CL2>() : tc_smem_bytes<BN, NS, 1>();
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (!use_cl) {
    cfg.gridDim = dim3(units < g_num_sms ? units : g_num_sms);
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (BN == 256 && h.fused2) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<256, NS, 1, true>, d_op, d_tmaps, M, units, dbg);
    return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, NS, 1, false>, d_op, d_tmaps, M, units, dbg);
  }
  const int max_clusters = g_num_sms / 2;
  const int clusters = units < max_clusters ? units : max_clusters;
  cfg.gridDim = dim3(clusters * 2);
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  // multi-wave 256-column launches claim their units dynamically (see DYN); the caller zeroes h.sched before the forward
  const bool dyn = (BN == 256 || BN == 128) && h.sched != nullptr && units >= 2 * clusters;
  if (BN == 256 && h.fused2) {
    cfg.dynamicSmemBytes = tc_smem_bytes<256, NS, 2, true>();
    if (dyn) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<256, NS, 2, true, true>, d_op, d_tmaps, M, units, dbg);
    return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<256, NS, 2, true>, d_op, d_tmaps, M, units, dbg);
  }
  if (BN == 256 && dyn) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<256, NS, 2, false, true>, d_op, d_tmaps, M, units, dbg);
  if (BN == 128 && dyn) return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<128, NS, 2, false, true>, d_op, d_tmaps, M, units, dbg);
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, NS, CL2, false>, d_op, d_tmaps, M, units, dbg);
}

cudaError_t launch_gemm_tc(const GemmOpDev* d_op, const GemmOpDev& h, const void* d_tmaps, int M, int precision, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  if (g_num_sms <= 0) return cudaErrorNotReady;
  const int bn = tc_block_n(h);
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(d_tmaps);
  const int ns = precision == R3D_PREC_BF16X3 ? 2 : 1;
#define R3D_LAUNCH(BN)                                                             \
  case BN:                                                                         \
    return ns == 2 ? launch_one<BN, 2>(d_op, tm, M, h, s) : launch_one<BN, 1>(d_op, tm, M, h, s);
  switch (bn) {
    R3D_LAUNCH(16) R3D_LAUNCH(32) R3D_LAUNCH(64) R3D_LAUNCH(128) R3D_LAUNCH(256)
    default: return cudaErrorInvalidValue;
  }
#undef R3D_LAUNCH
}

}  // namespace r3d
