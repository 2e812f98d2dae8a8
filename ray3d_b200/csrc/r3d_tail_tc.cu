// Chained tail launch: every GEMM of the path that has ONE row per window in a single persistent tcgen05 kernel.
//
// Reference ops covered (lib/model/rie.py): the last level of the temporal tree when it is down to one frame (:94-97),
// shrink (:99), GlobalInfo (:362), FuseBlocks (:388-394), Integration_* (:410-414) and their trajectory-net twins
// (:543-555) -- 13 grouped launches at stage 1, 17 at stage 3.  Each has only batch x 1 rows per problem: one to three waves
// of latency-bound tiles per launch plus ~8-10 us of fixed cost (CTA relaunch, TMEM allocation, cold first loads, store
// drain), i.e. ~230 us for ~8 % of the flops when launched one by one.
//
// Here the units of all those ops (128-column x 128/256-row tiles, m-major per op, ops in dependency order with the
// independent GlobalInfo chain interleaved) form ONE sequence that the CTAs (pairs) claim dynamically.  A unit of
// (op, problem, row group) starts when the units producing its input rows have landed: the store threads of a finished
// unit bump a per-(op, problem, row group) counter (release), the claiming CTA's producer thread spins on the counters
// of the unit's producers (acquire) before it publishes the unit to the CTA's other roles and issues its TMA loads.
// Claimed units are always held by resident CTAs and the sequence is topologically ordered, so a waiting unit's producers
// have been claimed before it: no deadlock, also when another kernel of the plan's second lane holds part of the GPU.
//
// Arithmetic, tile code and the split-precision scheme are those of gemm_tc_kernel (r3d_gemm_tc.cu): results are
// bit-identical to the one-launch-per-layer form.
#include <cstring>

#include "r3d_tc_common.cuh"

namespace r3d {

// one queue slot: everything the roles need to know about a claimed unit
struct TileDesc {
  GemmProb prob;
  float slope;
  int32_t p, m0, n0;
  int32_t tmap0;       // first tensor map of (op, problem)
  int32_t done_idx;    // completion counter of (op, problem, row group)
  int32_t bn;          // unit width of the op: 128 or 256 columns
  int32_t _pad;
};
static_assert(sizeof(TileDesc) % 8 == 0, "TileDesc is copied word-wise into 8-byte aligned shared memory");

constexpr int TAIL_SQ = 4;                                       // queue depth
constexpr int kTailDescBytes = 2816;                             // TileDesc queue + store barriers (keeps the staging tiles 1024-byte aligned)
static_assert(TAIL_SQ * sizeof(TileDesc) + (16 + TAIL_SQ) * 8 <= kTailDescBytes, "descriptor block too small");
__host__ __device__ constexpr int tail_aux_bytes(int cl) { return 256 + kTailDescBytes + 8 * 4096 * (cl == 2 ? 2 : 1) + EPI_WARPS * 512; }
constexpr int kTailMaxN = 256;                                    // widest unit; narrower ops (the heads) use 128 of the W slot
__host__ __device__ constexpr int tail_stage_bytes(int nsplit, int cl) { return nsplit * (TBM + kTailMaxN / cl) * TBK * 2; }
__host__ __device__ constexpr int tail_num_stages(int nsplit, int cl) {
  int s = (SMEM_LIMIT - tail_aux_bytes(cl)) / tail_stage_bytes(nsplit, cl);
  return s > 6 ? 6 : s;
}

// Experiment builds: where the leader CTAs' producer / MMA threads spend their cycles (summed over CTAs):
// [0] dependency spin, [1] unit-queue slot wait, [2] smem ring slot wait, [3] MMA: operands not yet landed, [4] MMA: accumulator
// not yet drained, [5] MMA: next unit not yet published, [6] units, [7] kernel cycles per CTA
__device__ unsigned long long g_tail_stats[8];
#ifdef R3D_EXPERIMENTS
#define R3D_STAT_T0() const long long _t0 = clock64()
#define R3D_STAT_ADD(i) st_acc[i] += clock64() - _t0
#else
#define R3D_STAT_T0() do { } while (0)
#define R3D_STAT_ADD(i) do { } while (0)
#endif

template <int NSPLIT, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1) tail_tc_kernel(const GemmOpDev* __restrict__ ops, const CUtensorMap* __restrict__ tmaps,
                                                                   const MultiOpDev* __restrict__ mo, int M) {
  static_assert(CL == 2, "the chained tail launch runs on CTA pairs (batches of >= 256 windows)");
  constexpr int MAXN = kTailMaxN;
  constexpr int EW = EPI_WARPS;
  constexpr int STAGES = tail_num_stages(NSPLIT, CL);
  constexpr int A_BYTES = TBM * TBK * 2, W_BYTES = (MAXN / CL) * TBK * 2;     // W slot of a stage (a 128-wide unit fills half of it)
  constexpr int STAGE_BYTES = tail_stage_bytes(NSPLIT, CL);
  constexpr int TMEM_COLS = 2 * MAXN;                              // two accumulator stages
  constexpr int CH = 32, COL_SPLIT = 2, MAX_CPW = MAXN / CH / COL_SPLIT;       // chunks per warp: bn / 64
  constexpr int EPI_BUFS = CL == 2 ? 2 : 1;                        // staging tile sets per column group (hi + lo each)
  constexpr int SQ = TAIL_SQ;
  auto chunk_index = [](int grp, int cc) { return cc * COL_SPLIT + grp; };
  static_assert(STAGES >= 2, "need at least a double-buffered smem ring");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* aux = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* sq_full = bars + 2 * STAGES + 5;    // [SQ] unit published in THIS CTA (descriptor slot filled)
  uint64_t* sq_empty = sq_full + SQ;            // [SQ] every consumer of both CTAs is done with the slot (the leader's is used)
  uint64_t* idq_full = sq_empty + SQ;           // [SQ] peer CTA: the leader has sent the unit id
  volatile uint32_t* sq_tile = reinterpret_cast<volatile uint32_t*>(idq_full + SQ);   // [SQ]
  static_assert((2 * 6 + 5 + 3 * SQ) * 8 + SQ * 4 <= 256, "barrier block overflows its 256 bytes");
  TileDesc* tq = reinterpret_cast<TileDesc*>(aux + 256);                                  // [SQ]
  uint64_t* sready_bar = reinterpret_cast<uint64_t*>(aux + 256 + SQ * sizeof(TileDesc));  // [column half][staging set], 4 arrivals
  uint64_t* sfree_bar = sready_bar + 8;
  uint4* stage_s = reinterpret_cast<uint4*>(aux + 256 + kTailDescBytes);
  float* bias_s = reinterpret_cast<float*>(aux + 256 + kTailDescBytes + 8 * 4096 * (CL == 2 ? 2 : 1));   // [EPI_WARPS][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const bool leader = crank == 0;
  constexpr int SQ_CONSUMERS = (1 + 2 + EW) * CL;                // MMA thread (leader) / producer (peer), 2 store threads, 8 epilogue warps
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EW * CL);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&sready_bar[i], 4);
      mbar_init(&sfree_bar[i], 1);
    }
    for (int i = 0; i < SQ; ++i) {
      mbar_init(&sq_full[i], 1);
      mbar_init(&sq_empty[i], SQ_CONSUMERS);
      mbar_init(&idq_full[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CL == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  long long st_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long st_begin = clock64();
  (void)st_acc; (void)st_begin;
  const int nops = __ldg(&mo->nops);
  const int m_groups = ((M + TBM - 1) / TBM + CL - 1) / CL;        // row groups (CL row tiles each)
  const int total = __ldg(&mo->unit0[nops]) * m_groups;

  // consumer side of the unit queue: every consuming role walks every unit in order; a slot (unit id + descriptor) stays
  // valid until the role hands it back
  uint32_t qc = 0;
  auto sq_pop = [&](int& slot) -> int {
    slot = (int)(qc % SQ);
    mbar_wait_tag(&sq_full[slot], (qc / SQ) & 1u, 1, (int)qc);
    ++qc;
    return (int)sq_tile[slot];
  };
  auto sq_release = [&](int slot) {
    if (CL == 2 && !leader) mbar_arrive_cluster(&sq_empty[slot], 0);
    else mbar_arrive(&sq_empty[slot]);
  };

  if (warp == 0) {
    // =============================== scheduler + TMA producer ===============================
    int stage = 0;
    uint32_t phase = 0, qn = 0;
    const int n_static = (int)gridDim.x / CL;
    int next_tile = (int)blockIdx.x / CL;                          // first unit: this cluster's index (no L2 round trip)
    uint32_t* const done = mo->done;
    uint32_t* const sched = mo->sched;
    const int cap = __ldg(&mo->m_groups_cap);
    for (;;) {
      const int slot = (int)(qn % SQ);
      const uint32_t qph = (qn / SQ) & 1u;
      int tile = 0, oi = 0, p = 0, mg = 0, nt = 0;
      if (lane == 0) {
        if (leader) {
          tile = next_tile;
          { R3D_STAT_T0(); mbar_wait_tag(&sq_empty[slot], qph ^ 1u, 2, (int)qn); R3D_STAT_ADD(1); }   // every consumer of both CTAs is done with the slot's previous unit
        } else {
          mbar_wait_cluster(&idq_full[slot], qph);
          tile = (int)sq_tile[slot];
          fence_proxy_async_all();                                 // the leader observed the unit's inputs; this thread's TMA loads follow
        }
        if (tile < total) {
          while (oi + 1 < nops && tile >= __ldg(&mo->unit0[oi + 1]) * m_groups) ++oi;
          const int local = tile - __ldg(&mo->unit0[oi]) * m_groups;
          const int per_m = __ldg(&mo->unit0[oi + 1]) - __ldg(&mo->unit0[oi]);
          mg = local / per_m;
          int rem = local - mg * per_m;
          for (p = 0;; ++p) {
            const int n_tiles = __ldg(&mo->ntiles[oi][p]);
            if (rem < n_tiles) break;
            rem -= n_tiles;
          }
          nt = rem;
          if (leader) {
            // the unit's input rows: wait until every producing unit of this row group has landed
            const int nd = __ldg(&mo->ndep[oi][p]);
            R3D_STAT_T0();
            for (int d = 0; d < nd; ++d) {
              const int row = __ldg(&mo->dep[oi][p][d]);
              const uint32_t tgt = (uint32_t)__ldg(&mo->ntiles[row / kMaxProb][row % kMaxProb]) * 2u * CL;   // column tiles x store threads x CTAs
              const uint32_t* c = done + (size_t)row * cap + mg;
              if (ld_acquire_gpu_u32(c) < tgt) {
                const long long t0 = clock64();
                while (ld_acquire_gpu_u32(c) < tgt) {
                  if (clock64() - t0 > 4000000000LL) {
                    printf("r3d tail_tc: dependency timeout (block %d unit %d op %d problem %d dep row %d)\n", blockIdx.x, tile, oi, p, row);
                    __trap();
                  }
                }
              }
            }
            R3D_STAT_ADD(0);
            if (nd) fence_proxy_async_all();                       // the loads below go through the async proxy
          }
        }
        if (leader) {
          if (CL == 2) {
            st_shared_cluster_u32(&sq_tile[slot], 1, (uint32_t)tile);
            mbar_arrive_cluster(&idq_full[slot], 1, true);         // release.cluster: the id (and the observed counters) reach the peer
          }
          sq_tile[slot] = (uint32_t)tile;
          if (tile < total) next_tile = n_static + (int)atomicAdd(sched, 1u);
        }
      }
      tile = __shfl_sync(0xffffffffu, tile, 0);
      oi = __shfl_sync(0xffffffffu, oi, 0);
      p = __shfl_sync(0xffffffffu, p, 0);
      mg = __shfl_sync(0xffffffffu, mg, 0);
      nt = __shfl_sync(0xffffffffu, nt, 0);
      TileDesc& td = tq[slot];
      if (tile < total) {                                            // descriptor -> this CTA's slot (warp-wide copy)
        const int opi = __ldg(&mo->op_index[oi]);
        const GemmOpDev* gop = ops + opi;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&gop->prob[p]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&td.prob);
        for (int i = lane; i < (int)(sizeof(GemmProb) / 4); i += 32) dst[i] = __ldg(src + i);
        if (lane == 0) {
          td.slope = __ldg(&gop->slope);
          td.p = p;
          td.m0 = (mg * CL + crank) * TBM;
          const int bn = kTailN * (int)__ldg(&mo->width[oi]);
          td.bn = bn;
          td.n0 = nt * bn;
          td.tmap0 = (opi * kMaxProb + p) * kTmapsPerProb;
          td.done_idx = (oi * kMaxProb + p) * cap + mg;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sq_full[slot]);                    // this CTA's other roles may take the unit
      ++qn;
      if (tile >= total) break;
      if (lane == 0) {
        const CUtensorMap* tm = tmaps + td.tmap0;
        const int nkb = td.prob.K / TBK;
        const int m0 = td.m0, n0 = td.n0, bn = td.bn;
        const uint32_t stage_tx = (uint32_t)(NSPLIT * (A_BYTES + (bn / CL) * TBK * 2));     // bytes this CTA receives per K block
        for (int kb = 0; kb < nkb; ++kb) {
          { R3D_STAT_T0(); mbar_wait_tag(&empty_bar[stage], phase ^ 1, 4, kb); R3D_STAT_ADD(2); }
          uint8_t* st = smem + stage * STAGE_BYTES;
          // the operand rows are read again by the unit's sibling column tiles: keep them in L2; weights are shared by all row groups
          // both CTAs' loads complete on the leader's barrier, which the (leader-only) MMA thread waits on
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * stage_tx);
          const int wrow = n0 + crank * (bn / CL);
          tma_load_2d_2sm(st, tm + 0, &full_bar[stage], kb * TBK, m0, kEvictNormal);
          if (NSPLIT == 2) tma_load_2d_2sm(st + A_BYTES, tm + 1, &full_bar[stage], kb * TBK, m0, kEvictNormal);
          tma_load_2d_2sm(st + NSPLIT * A_BYTES, tm + kTmapTailW, &full_bar[stage], kb * TBK, wrow, kEvictLast);
          if (NSPLIT == 2) tma_load_2d_2sm(st + NSPLIT * A_BYTES + W_BYTES, tm + kTmapTailW + 1, &full_bar[stage], kb * TBK, wrow, kEvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (CL == 2 && !leader) sq_release(slot);                   // the peer's producer counts as a consumer of the slot
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (;;) {
        int slot;
        R3D_STAT_T0();
        const int tile = sq_pop(slot);
        R3D_STAT_ADD(5);
        if (tile >= total) break;
        st_acc[6] += 1;
        const int nkb = tq[slot].prob.K / TBK;
        const uint32_t idesc = make_idesc(tq[slot].bn, TBM * CL);
        sq_release(slot);                                            // (nothing else of the descriptor is needed here)
        { R3D_STAT_T0(); mbar_wait_tag(&tempty_bar[acc], acc_phase ^ 1, 6, tile); R3D_STAT_ADD(4); }   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * MAXN;
        for (int kb = 0; kb < nkb; ++kb) {
          { R3D_STAT_T0(); mbar_wait_tag(&full_bar[stage], phase, 5, kb); R3D_STAT_ADD(3); }
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = make_smem_desc(st), w_hi = make_smem_desc(st + NSPLIT * A_BYTES);
          const uint64_t a_lo = make_smem_desc(st + A_BYTES), w_lo = make_smem_desc(st + NSPLIT * A_BYTES + W_BYTES);
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
            const uint32_t accumulate = (kb | k) != 0;
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
          }
          if (CL == 1) {
            umma_commit(&empty_bar[stage]);
            if (kb == nkb - 1) umma_commit(&tfull_bar[acc]);
          } else {
            umma_commit_2sm(&empty_bar[stage], MC_MASK);
            if (kb == nkb - 1) umma_commit_2sm(&tfull_bar[acc], MC_MASK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < EPI_WARP0) {
    // =============================== store threads ===============================
    // lane 0 of warp 2 / 3: TMA tensor stores of column half 0 / 1 (see gemm_tc_kernel), then the unit's completion signal
    const int st = warp - 2;
    if (lane == 0) {
      uint32_t round = 0;
      uint32_t* const done = mo->done;
      for (;;) {
        int slot;
        const int tile = sq_pop(slot);
        if (tile >= total) break;
        const TileDesc& td = tq[slot];
        const GemmProb& pr = td.prob;
        const CUtensorMap* dmaps = tmaps + td.tmap0 + 6;
        bool any_bf = false;
        for (int t = 0; t < pr.ndst; ++t) any_bf |= pr.dst[t].f32 == 0;
        if (any_bf) {
          const int cpw = td.bn / (CH * COL_SPLIT);
          for (int cc = 0; cc < cpw; ++cc) {
            const int n = td.n0 + chunk_index(st, cc) * CH;
            if (n >= pr.N) continue;
            const int b = EPI_BUFS == 2 ? (int)(round & 1) : 0;
            const uint32_t uses = EPI_BUFS == 2 ? round >> 1 : round;
            mbar_wait_tag(&sready_bar[st * 2 + b], uses & 1, 8, tile);
            fence_async_smem();
            const uint4* tile_hi = stage_s + (st * EPI_BUFS + b) * 1024;
            for (int t = 0; t < pr.ndst; ++t) {
              const Dst& d = pr.dst[t];
              if (d.f32) continue;
              if (NSPLIT == 2) tma_store_3d(dmaps + 2 * t, tile_hi, d.col + n, td.m0, 0);
              else tma_store_2d(dmaps + 2 * t, tile_hi, d.col + n, td.m0);
            }
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(&sfree_bar[st * 2 + b]);
            ++round;
          }
          bulk_wait0();                                              // this half's rows have landed ...
          fence_proxy_async_all();
          __threadfence();
        }
        red_release_gpu_add_u32(done + td.done_idx, 1u);             // ... and every unit waiting for this row group may count it
        sq_release(slot);
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    const int half = ew >> 2;                                 // column group
    uint4* const stage_base = stage_s + half * EPI_BUFS * 1024;
    uint32_t sround = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr bool BIAS_SMEM = CL == 2;
    constexpr int PB = MAX_CPW;
    float pb[PB];
    auto prefetch_bias = [&](const TileDesc& t) {
      const float* bp = t.prob.bias + t.n0;
      const int cpw = t.bn / (CH * COL_SPLIT);
#pragma unroll
      for (int i = 0; i < PB; ++i) pb[i] = i < cpw ? __ldg(bp + chunk_index(half, i) * CH + lane) : 0.f;
    };
    // Every lane observes the slot; it is handed back at the end of the unit.  `block` = false only probes: the next unit
    // is published AFTER its inputs have landed, and those may (transitively) depend on the unit this warp is still
    // working on -- waiting for it in the middle of the current unit would deadlock the chain.
    auto warp_pop = [&](int& slot, bool block, bool& got) -> int {
      const int sl = (int)(qc % SQ);
      const uint32_t ph = (qc / SQ) & 1u;
      if (block) {
        mbar_wait_tag(&sq_full[sl], ph, 1, (int)qc);
      } else if (!__all_sync(0xffffffffu, mbar_try_wait(&sq_full[sl], ph))) {
        got = false;
        return 0;
      }
      got = true;
      slot = sl;
      ++qc;
      __syncwarp();
      return (int)sq_tile[sl];
    };
    int slot = 0;
    bool got_first = false;
    int tile = warp_pop(slot, true, got_first);
    if (BIAS_SMEM && tile < total) prefetch_bias(tq[slot]);
    while (tile < total) {
      const TileDesc& td = tq[slot];
      const GemmProb& pr = td.prob;
      int next_slot = 0, next_tile = total;
      bool next_ready = false;
      auto look_ahead = [&](bool block) {                        // next unit's descriptor and bias, behind this unit's TMEM loads
        if (next_ready) return;
        bool got = false;
        const int t = warp_pop(next_slot, block, got);
        if (!got) return;
        next_ready = true;
        next_tile = t;
        if (next_tile >= total) return;
        if (BIAS_SMEM) prefetch_bias(tq[next_slot]);
      };
      const CUtensorMap* dmaps = tmaps + td.tmap0 + 6;
      (void)dmaps;
      const int m_base = td.m0 + q * 32;
      const int row = m_base + lane;
      const bool row_ok = row < M;
      bool any_f32 = false, any_bf = false, any_lo = false;
      for (int t = 0; t < pr.ndst; ++t) {
        any_f32 |= pr.dst[t].f32 != 0;
        any_bf |= pr.dst[t].f32 == 0;
        any_lo |= pr.dst[t].f32 == 0 && pr.dst[t].m.p1 != nullptr;
      }
      const bool has_res = pr.res.p0 != nullptr;
      const float2 slope2 = make_float2(td.slope, td.slope);
      const __nv_bfloat16* res_hi = reinterpret_cast<const __nv_bfloat16*>(pr.res.p0);
      const __nv_bfloat16* res_lo = reinterpret_cast<const __nv_bfloat16*>(pr.res.p1);
      ResidualRegs rr;
      if (has_res) residual_issue<true>(rr, res_hi, res_lo, pr.res.ld, pr.res_col + td.n0 + chunk_index(half, 0) * CH, lane, m_base, M);
      float* my_bias = bias_s + ew * 128;
      if (BIAS_SMEM) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PB; ++i) my_bias[i * CH + lane] = pb[i];
        __syncwarp();
      }
      mbar_wait_tag(&tfull_bar[acc], acc_phase, 7, tile);
      tc_fence_after();
      {
        uint32_t r[32];
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MAXN);
        const int cpw = td.bn / (CH * COL_SPLIT);
        tmem_ld32(taddr0 + chunk_index(half, 0) * CH, r);
#pragma unroll 1
        for (int cc = 0; cc < cpw; ++cc) {
          const int n = td.n0 + chunk_index(half, cc) * CH;
          float bb[CH];
          if (BIAS_SMEM) {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {
              const float4 b4 = *reinterpret_cast<const float4*>(my_bias + cc * CH + j4 * 4);
              bb[j4 * 4 + 0] = b4.x; bb[j4 * 4 + 1] = b4.y; bb[j4 * 4 + 2] = b4.z; bb[j4 * 4 + 3] = b4.w;
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < CH / 4; ++j4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(pr.bias + n) + j4);
              bb[j4 * 4 + 0] = b4.x; bb[j4 * 4 + 1] = b4.y; bb[j4 * 4 + 2] = b4.z; bb[j4 * 4 + 3] = b4.w;
            }
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) {
            const float2 a = bias_lrelu2(r[2 * j], r[2 * j + 1], bb[2 * j], bb[2 * j + 1], slope2);
            v[2 * j] = a.x;
            v[2 * j + 1] = a.y;
          }
          if (cc + 1 < cpw) tmem_ld32(taddr0 + chunk_index(half, cc + 1) * CH, r);
          if (cc > 0) look_ahead(false);
          if (n < pr.N) {                        // warp-uniform
            const int sbuf = EPI_BUFS == 2 ? (int)(sround & 1) : 0;
            uint4* const stage_hi = stage_base + sbuf * 1024 + q * 128;
            uint4* const stage_lo = stage_hi + 512;
            if (any_bf || has_res)               // the staging set doubles as the residual's transpose scratch
              mbar_wait_tag(&sfree_bar[half * 2 + sbuf], (((EPI_BUFS == 2 ? sround >> 1 : sround) & 1) ^ 1), 9, tile);
            if (has_res) {
              residual_consume(stage_hi, rr, res_lo != nullptr, lane, v);
              if (cc + 1 < cpw)
                residual_issue<true>(rr, res_hi, res_lo, pr.res.ld, pr.res_col + td.n0 + chunk_index(half, cc + 1) * CH, lane, m_base, M);
            }
            if (any_f32 && row_ok) {             // network outputs (tiny): direct masked stores
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
              stage_write(stage_hi, hi, lane);
              if (any_lo) stage_write(stage_lo, lo, lane);
              __syncwarp();
              if (lane == 0) mbar_arrive(&sready_bar[half * 2 + sbuf]);    // the store thread takes it from here
              ++sround;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(&tempty_bar[acc], 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      __syncwarp();
      if (lane == 0) sq_release(slot);
      look_ahead(true);                                          // (this unit is complete: waiting for the next one is safe now)
      slot = next_slot;
      tile = next_tile;
    }
    __syncwarp();
  }

#ifdef R3D_EXPERIMENTS
  if (leader && lane == 0 && (warp == 0 || warp == 1)) {
    if (warp == 1) st_acc[7] = clock64() - st_begin;
    for (int i = 0; i < 8; ++i)
      if (st_acc[i]) atomicAdd(&g_tail_stats[i], (unsigned long long)st_acc[i]);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if (CL == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
cudaError_t tail_stats_read(unsigned long long* out, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_tail_stats, sizeof(g_tail_stats));
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(g_tail_stats, z, sizeof(z));
  }
  return e;
}

template <int NS, int CL>
static constexpr int tail_smem_bytes() { return tail_num_stages(NS, CL) * tail_stage_bytes(NS, CL) + tail_aux_bytes(CL); }

cudaError_t tail_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(tail_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem_bytes<1, 2>())) != cudaSuccess) return e;
  return cudaFuncSetAttribute(tail_tc_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem_bytes<2, 2>());
}

template <int NS>
static cudaError_t launch_tail(const GemmOpDev* d_ops, const CUtensorMap* tm, const MultiOpDev* d_mo, int units, int M, int cap, cudaStream_t s) {
  const int sms = tc_num_sms();
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const int max_clusters = (cap > 0 && cap < sms / 2) ? cap : sms / 2;
  cfg.gridDim = dim3(2 * (units < max_clusters ? units : max_clusters));
  cfg.dynamicSmemBytes = tail_smem_bytes<NS, 2>();
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, tail_tc_kernel<NS, 2>, d_ops, tm, d_mo, M);
}

cudaError_t launch_tail_tc(const GemmOpDev* d_ops, const void* d_tmaps, const MultiOpDev* d_mo, const MultiOpDev& h_mo, int M, int precision,
                           int max_clusters, cudaStream_t s) {
  if (M <= 0 || h_mo.nops <= 0) return cudaSuccess;
  if (tc_num_sms() <= 0) return cudaErrorNotReady;
  if (!tail_uses_pairs(M)) return cudaErrorInvalidValue;         // small batches keep one launch per op (narrower tiles)
  const int units = h_mo.unit0[h_mo.nops] * tail_row_groups(M);
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(d_tmaps);
  return precision == R3D_PREC_BF16X3 ? launch_tail<2>(d_ops, tm, d_mo, units, M, max_clusters, s) : launch_tail<1>(d_ops, tm, d_mo, units, M, max_clusters, s);
}

}  // namespace r3d
