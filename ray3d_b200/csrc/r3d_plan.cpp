// Host side of the C ABI: plan construction, BatchNorm folding + weight packing, workspace layout,
// and the launch sequence of one forward pass.  See include/ray3d_b200.h for the contract and the
// reference interfaces each entry point replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "r3d_internal.h"

using namespace r3d;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(R3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                         \
  } while (0)

// Experiment switches are read from the environment only in builds made with -DR3D_EXPERIMENTS; the shipped library
// takes its (few, result-neutral) tuning options through r3d_plan_set_option.
static const char* exp_env(const char* name) {
#ifdef R3D_EXPERIMENTS
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// Selects the plan's device for the duration of a call and restores the caller's on every exit path.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      switched = err == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

extern "C" R3D_API const char* r3d_last_error(void) { return g_err.c_str(); }
extern "C" R3D_API int r3d_abi_version(void) { return R3D_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------
// static network description (mirrors ray3d_b200/spec.py; reference lines cited there)
// ------------------------------------------------------------------------------------------------
static const char* kGroupNames[5] = {"Torso", "LArm", "RArm", "LLeg", "RLeg"};

struct GroupTable {
  std::vector<int> joints[5];
  std::vector<std::pair<int, int>> slots;   // output slot -> (group, joint in head)
};

static bool group_table(int J, GroupTable& g) {
  auto set = [&](int i, std::initializer_list<int> l) { g.joints[i] = l; };
  struct Seg { int grp, first, n; };
  std::vector<Seg> order;
  if (J == 17) {   // lib/model/rie.py:308-315, 426-427
    set(0, {0, 7, 8, 9, 10}); set(1, {14, 15, 16}); set(2, {11, 12, 13}); set(3, {1, 2, 3}); set(4, {4, 5, 6});
    order = {{0, 0, 1}, {3, 0, 3}, {4, 0, 3}, {0, 1, 4}, {2, 0, 3}, {1, 0, 3}};
  } else if (J == 15) {   // rie.py:316-323, 428-429
    set(0, {0, 1, 14}); set(1, {2, 3, 4}); set(2, {5, 6, 7}); set(3, {8, 9, 10}); set(4, {11, 12, 13});
    order = {{0, 0, 2}, {3, 0, 3}, {4, 0, 3}, {2, 0, 3}, {1, 0, 3}, {0, 2, 1}};
  } else if (J == 14) {   // rie.py:324-331, 430-431
    set(0, {0, 7}); set(1, {8, 9, 10}); set(2, {11, 12, 13}); set(3, {4, 5, 6}); set(4, {1, 2, 3});
    order = {{0, 0, 1}, {3, 0, 3}, {4, 0, 3}, {2, 0, 3}, {1, 0, 3}, {0, 1, 1}};
  } else {
    return false;
  }
  g.slots.clear();
  for (auto& s : order)
    for (int i = 0; i < s.n; ++i) g.slots.push_back({s.grp, s.first + i});
  return (int)g.slots.size() == J;
}

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
static constexpr int kFcWidth = 1024;   // rie.py:226,232,245-253,483,494
static constexpr int kEmbedMid = 32;    // embedding.py:5
static constexpr double kBnEps = 1e-5;

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct TensorEntry {
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool set = false;
};

struct PackedLayer {
  int n = 0, k = 0, n_pad = 0, k_pad = 0;
  std::vector<float> w, b;        // [n_pad][k_pad], [n_pad]  BN folded, fp32
  size_t off_w0 = 0, off_w1 = 0, off_b = 0;   // offsets in the device weight slab
  bool plain = false;             // embedder matrices: no padding, fp32 only
  int alg_k = 0;                  // K of the reference's own op (flop accounting); 0 => k
  uint64_t kmask = 0;             // 16-column K steps holding at least one non-zero weight (0: K > 1024, no elision)
};

struct MatReq {          // one activation matrix in the workspace slab
  int rows_per_seq;
  int ld;
  bool f32;
  size_t off0 = 0, off1 = 0;
};

struct OpHost {
  GemmOpDev dev;                         // pointers filled by bind_workspace()
  std::string name;
  bool side = false;                     // runs on the side stream (independent of the temporal tree)
  bool join_before = false;              // first op that consumes side-stream results
  // symbolic bindings resolved to pointers once the slabs exist
  struct Bind { int a = -1, a_ld = 0, res = -1, res_ld = 0, res_col = 0; std::string layer, layer2;
                std::vector<std::pair<int, int>> dst; std::vector<int> dst_f32; };
  Bind bind[kMaxProb];
};

struct r3d_plan {
  r3d_config cfg{};
  GroupTable groups;
  int J = 0, Cin = 0, JC = 0, C = 0, L = 0, T = 0, tc = 0, E = 0, ext = 0;
  bool embed = false, has_pos = false, has_trj = false;
  std::vector<int> widths, lens;
  int feat_pos = 0, feat_trj = 0;
  // Column layout of the first-layer operand shared by all problems (build_a0_layout).  Slot 0 = the Torso joints
  // (root first), slots 1..4 = [copy of the root joint | the limb's joints]; inside a slot the columns are frame-major
  // [frame w0*t' | ... | frame w0*t'+w0-1 | frame tc] x joints x Cin; every slot starts on a 16-column K step.
  std::vector<int> a0_src;        // [a0_kpad] column -> source index (see PrologueDev::a0_map), -1 = zero
  std::vector<int> a0_home;       // [J] column of the joint's (tap 0, coordinate 0) entry
  std::vector<int> a0_tap;        // [5] columns per frame inside slot g (tap stride)
  int a0_slot_of[32] = {};        // joint -> slot
  int a0_root[5] = {0, 0, 0, 0, 0};   // column of the root joint's (tap 0, coordinate 0) entry in slot g
  int a0_k = 0, a0_kpad = 0;

  std::map<std::string, TensorEntry> tensors[2];     // expected entries per net (0 pos, 1 trj)
  std::map<std::string, PackedLayer> layers;         // key "<net>:<module path>"
  bool finalized = false;

  // temporal-block problems of this plan
  struct TB { int net; std::string prefix; std::vector<int> joints; int head; int group; };
  std::vector<TB> tbs;

  // device state
  int device = -1;
  bool uploaded = false;
  char* d_weights = nullptr;
  size_t weight_bytes = 0;
  char* d_ws = nullptr;
  size_t ws_bytes = 0;
  int cap = 0;                                        // sequences the workspace holds
  std::vector<MatReq> mats;
  std::vector<OpHost> ops;
  PrologueDev pro{};
  AssembleDev asmb{};
  char* d_desc = nullptr;                             // ops + prologue + assemble + tmaps
  size_t off_ops = 0, off_pro = 0, off_asm = 0, off_tmaps = 0, off_sched = 0;
  static constexpr size_t kSchedStride = 128;
  // symbolic ids used while building
  int m_inc = -1, m_heads[kMaxProb] = {-1, -1, -1, -1, -1, -1};
  struct EmbBind { int net; std::vector<std::pair<int, int>> dst; };
  std::vector<EmbBind> emb_binds;
  std::vector<int> m_a0;
  // optional per-launch timing (r3d_plan_set_profiling): ring of event sets, one per forward chunk
  bool profiling = false;
  bool use_side_stream = true;
  // chained launches (r3d_tail_tc.cu): a dependency chain of one-row ops as ONE persistent kernel.
  //   chain 0 (option "tail_fusion", off): top tree level, shrink, FuseBlocks, Integration on the main stream
  //   chain 1 (option "side_chain", on):   the GlobalInfo chain on the side stream, on a FEW CTA pairs
  struct Chain {
    std::vector<int> ops;                             // plan op indices in unit order (empty: every op is its own launch)
    MultiOpDev mo{};                                  // host copy; the device copy lives in the descriptor slab
    size_t off_mo = 0, off_done = 0;
    int max_clusters = 0;                             // 0: the whole GPU
  };
  Chain chain[2];
  bool tail_fusion = false;                           // (off: measured slower than one launch per layer)
  bool side_chain = true, side_chain_force = false;   // option "side_chain": 0 off, 1 when worth it (default), 2 always
  int tail_width = 128;                               // option "tail_width": unit width of chain 0 (128 or 256 columns)
  int tail_clusters = 0;                              // option "tail_clusters": CTA pairs chain 0 may hold (0: the whole GPU)
  int side_clusters = 0;                              // option "side_clusters": CTA pairs the GlobalInfo chain may hold (0: from its flop share)
  int tile_policy = 0;                                // option "tile_policy": 0 auto, 1 latency (wave-count model), 2 throughput (widest tiles)
  size_t zero_bytes = 0;
  struct Launch { std::string name; std::vector<int> ops; int chain; };     // chain: -1 = one op
  std::vector<Launch> launches[4];                    // GEMM launches of one forward; variant = (chain 0 used) | (chain 1 used) << 1
  int prof_variant = 0;                               // variant the last profiled forward used
  std::vector<int> flip_in, flip_out;                 // joint permutations of the flip augmentation (empty: not set)
  std::vector<cudaEvent_t> prof_ev;                   // [kProfRing][nops + 3]
  int prof_runs = 0;
  // host-call staging
  cudaStream_t s_copy = nullptr, s_comp = nullptr, s_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  static constexpr int kSlots = 4;           // device staging slots of the host-buffer path (slot & 1 = lane)
  cudaEvent_t ev_in[kSlots] = {}, ev_done[kSlots] = {};
  static constexpr int kTicketRing = 8;      // asynchronous host submissions in flight (r3d_submit_*_host / r3d_wait)
  cudaEvent_t ev_ticket[kTicketRing][2] = {};   // per lane
  uint8_t ticket_lanes[kTicketRing] = {};        // lanes a submission ran on (bit mask)
  bool use_lanes = true;                         // R3D_LANES=1 keeps every submission on lane 0
  uint64_t submit_seq = 0;                   // tickets handed out so far
  uint64_t slot_seq = 0;                     // staging-slot uses so far (alternates the two slots across calls)
  char* d_stage = nullptr;
  size_t stage_bytes = 0;
  int host_chunk = 0;                        // windows per chunk of a host-buffer call (0: 512 blocking / 1024 streamed)
  char* d_vid = nullptr;                     // per lane: [param row | ray-encoded frames] of the video being evaluated
  size_t vid_bytes = 0;
  cudaEvent_t ev_ws = nullptr;               // per lane: recorded after the last forward that used this lane's buffers
  // small batches are launch-latency bound (~20 launches, 2 streams): their launch sequence is captured once into a
  // CUDA graph per (batch, input kind, output set) and replayed with one cudaGraphLaunch
  struct GraphEntry {
    int batch = 0, mask = 0;
    int64_t src_stride = 0, prm_stride = 0;
    cudaGraphExec_t exec = nullptr;
    char* buf = nullptr;                         // static input/output buffers the captured kernels point at
    size_t off_prm = 0, off_pos = 0, off_sum = 0, off_trj = 0;
  };
  std::vector<GraphEntry> graphs;
  int graph_max_batch = 64;                      // R3D_GRAPH_MAX_BATCH; 0 disables
  uint64_t graph_launches = 0;
  std::shared_ptr<std::mutex> mu = std::make_shared<std::mutex>();
  // Second lane: a twin plan with its own workspace, descriptors and streams that shares this plan's weight slab.
  // Asynchronous submissions (r3d_submit_*) alternate between the two lanes, so the under-filled tail launches of one
  // batch (upper tree levels, FC heads) run while the other batch's large launches keep the remaining SMs busy.
  r3d_plan* twin = nullptr;
  bool is_twin = false;
  cudaEvent_t ev_sub = nullptr;                // orders a device submission after the caller's stream
};

static void clear_graphs(r3d_plan* p) {
  for (auto& g : p->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.buf) cudaFree(g.buf);
  }
  p->graphs.clear();
}


// ---- expected state_dict entries ---------------------------------------------------------------
static void expect(r3d_plan* p, int net, const std::string& name, std::vector<int64_t> shape) {
  TensorEntry e;
  e.shape = std::move(shape);
  p->tensors[net][name] = std::move(e);
}
static void expect_bn(r3d_plan* p, int net, const std::string& pre, int c) {
  for (const char* s : {".weight", ".bias", ".running_mean", ".running_var"}) expect(p, net, pre + s, {c});
}
static void expect_linear(r3d_plan* p, int net, const std::string& pre, int cin, int cout) {
  expect(p, net, pre + ".weight", {cout, cin});
  expect(p, net, pre + ".bias", {cout});
}
static void expect_tblock(r3d_plan* p, int net, const std::string& pre, int in_ch) {
  const int C = p->C;
  expect_bn(p, net, pre + ".expand_bn", C);
  expect(p, net, pre + ".shrink.weight", {p->L, C, 1});
  expect(p, net, pre + ".shrink.bias", {p->L});
  expect(p, net, pre + ".expand_conv.weight", {C, in_ch, p->widths[0]});
  for (size_t i = 1; i < p->widths.size(); ++i) {
    const std::string a = std::to_string(2 * (i - 1)), b = std::to_string(2 * (i - 1) + 1);
    expect(p, net, pre + ".layers_conv." + a + ".weight", {C, C, p->widths[i]});
    expect(p, net, pre + ".layers_conv." + b + ".weight", {C, C, 1});
    expect_bn(p, net, pre + ".layers_bn." + a, C);
    expect_bn(p, net, pre + ".layers_bn." + b, C);
  }
}
static void expect_fcblock(r3d_plan* p, int net, const std::string& pre, int cin, int cout, int nblocks) {
  expect_linear(p, net, pre + ".fc_1", cin, kFcWidth);
  expect_bn(p, net, pre + ".bn_1", kFcWidth);
  expect_linear(p, net, pre + ".fc_2", kFcWidth, cout);
  for (int i = 0; i < nblocks; ++i) {
    const std::string q = pre + ".layers." + std::to_string(i);
    expect_linear(p, net, q + ".w1", kFcWidth, kFcWidth);
    expect_bn(p, net, q + ".batch_norm1", kFcWidth);
    expect_linear(p, net, q + ".w2", kFcWidth, kFcWidth);
    expect_bn(p, net, q + ".batch_norm2", kFcWidth);
  }
}
static void expect_embed(r3d_plan* p, int net) {
  expect_linear(p, net, "embedder.w1", p->ext, kEmbedMid);
  expect_bn(p, net, "embedder.b1", kEmbedMid);
  expect_linear(p, net, "embedder.w2", kEmbedMid, p->E);
  expect_bn(p, net, "embedder.b2", p->E);
}

// The folded expand_conv weights of a joint group (pack_expand_folded) are non-zero only in the columns of the group's
// own joints and of the root joint.  Ordering the shared operand's columns group by group turns that sparsity into
// whole 16-column K steps that the tensor-core GEMM neither loads nor multiplies (GemmProb::kmask): at J=17, Cin=3,
// w0=3 a limb problem issues 3 of 16 K steps, the Torso 4, the trajectory net (all joints) 16.
static void build_a0_layout(r3d_plan* p) {
  const int Cin = p->Cin, JC = p->JC, w0 = p->widths[0];
  p->a0_src.clear();
  p->a0_home.assign(p->J, -1);
  p->a0_tap.assign(5, 0);
  // one slot: frame-major [tap 0: joints x Cin | tap 1 | ... | tap w0-1 | frame tc], so consecutive columns mostly
  // read consecutive input words (the input stage gathers them from shared memory)
  auto put_slot = [&](const std::vector<int>& js, int g) {
    while (p->a0_src.size() % 16) p->a0_src.push_back(-1);            // every slot starts on a 16-column K step
    const int first = (int)p->a0_src.size(), per_tap = (int)js.size() * Cin;
    for (int tap = 0; tap <= w0; ++tap)
      for (int j : js)
        for (int c = 0; c < Cin; ++c) p->a0_src.push_back(tap * JC + j * Cin + c);   // tap == w0: frame tc
    p->a0_tap[g] = per_tap;
    return first;
  };
  for (int g = 0; g < 5; ++g) {
    std::vector<int> js = p->groups.joints[g];
    if (g != 0) js.insert(js.begin(), 0);                              // limb slots carry their own copy of the root joint
    const int first = put_slot(js, g);
    for (size_t i = 0; i < js.size(); ++i) {
      if (g != 0 && i == 0) continue;
      p->a0_home[js[i]] = first + (int)i * Cin;
      p->a0_slot_of[js[i]] = g;
    }
    p->a0_root[g] = first;                                             // root (joint 0) is the first joint of every slot
  }
  p->a0_k = (int)p->a0_src.size();
  p->a0_kpad = round_up(p->a0_k, kKAlign);
  p->a0_src.resize(p->a0_kpad, -1);
}

extern "C" R3D_API int r3d_plan_create(const r3d_config* cfg, r3d_plan** out) {
  if (!cfg || !out) return fail(R3D_ERR_BAD_ARG, "r3d_plan_create: null argument");
  *out = nullptr;
  std::unique_ptr<r3d_plan> p(new r3d_plan);
  p->cfg = *cfg;
  if (!group_table(cfg->num_joints, p->groups))
    return fail(R3D_ERR_UNSUPPORTED, "num_joints=%d: the reference defines joint groups only for 17/15/14 (rie.py:306-357)",
                cfg->num_joints);
  if (cfg->in_features != 2 && cfg->in_features != 3)
    return fail(R3D_ERR_UNSUPPORTED, "in_features=%d: must be 3 (ray) or 2 (rie.py:306,333)", cfg->in_features);
  if (cfg->n_widths < 1 || cfg->n_widths > R3D_MAX_WIDTHS) return fail(R3D_ERR_BAD_ARG, "n_widths=%d out of range", cfg->n_widths);
  if (cfg->channels <= 0 || cfg->channels % kKAlign || cfg->latent <= 0 || cfg->latent % kKAlign)
    return fail(R3D_ERR_UNSUPPORTED, "channels=%d / latent=%d must be positive multiples of %d", cfg->channels, cfg->latent, kKAlign);
  if (!(cfg->nets & (R3D_NET_POS | R3D_NET_TRJ)) || (cfg->nets & ~3)) return fail(R3D_ERR_BAD_ARG, "nets=%d", cfg->nets);
  if (cfg->precision < R3D_PREC_FP32 || cfg->precision > R3D_PREC_BF16) return fail(R3D_ERR_BAD_ARG, "precision=%d", cfg->precision);
  p->J = cfg->num_joints; p->Cin = cfg->in_features; p->JC = p->J * p->Cin; p->C = cfg->channels; p->L = cfg->latent;
  p->T = 1;
  for (int i = 0; i < cfg->n_widths; ++i) {
    const int w = cfg->widths[i];
    if (w < 1 || w % 2 == 0 || w > 63) return fail(R3D_ERR_UNSUPPORTED, "filter width %d must be odd and in 1..63", w);
    p->widths.push_back(w);
    p->T *= w;
  }
  int t = p->T;
  for (int w : p->widths) { t /= w; p->lens.push_back(t); }
  p->tc = p->T / p->Cin;                                       // rie.py:290,304
  p->embed = cfg->extrinsic_dim > 0 && cfg->embed_dim > 0;     // rie.py:235
  p->ext = p->embed ? cfg->extrinsic_dim : 0;
  p->E = p->embed ? cfg->embed_dim : 0;
  if (p->ext > 8) return fail(R3D_ERR_UNSUPPORTED, "extrinsic_dim=%d > 8", p->ext);
  p->has_pos = cfg->nets & R3D_NET_POS;
  p->has_trj = cfg->nets & R3D_NET_TRJ;
  p->feat_pos = p->L * (cfg->stage == 1 ? 2 : 3) + p->E;     // rie.py:241-242
  p->feat_trj = p->L * 2 + p->E;                              // rie.py:491-492
  if ((size_t)p->T * p->JC * 4 > 200 * 1024) return fail(R3D_ERR_UNSUPPORTED, "receptive field %d too large for the input stage", p->T);
  build_a0_layout(p.get());
  if ((p->widths[0] + 1) * p->JC > 32767) return fail(R3D_ERR_UNSUPPORTED, "first filter width %d too large for the input stage", p->widths[0]);

  if (p->has_pos) {
    for (int g = 0; g < 5; ++g) {
      r3d_plan::TB tb{0, std::string("LocalLayer_") + kGroupNames[g], p->groups.joints[g], g, g};
      expect_tblock(p.get(), 0, tb.prefix, 3 * (int)tb.joints.size() * p->Cin);
      p->tbs.push_back(tb);
    }
    expect_fcblock(p.get(), 0, "GlobalInfo", p->JC, p->L, 2);
    if (cfg->stage != 1)
      for (int i = 0; i < 5; ++i) expect_fcblock(p.get(), 0, "FuseBlocks." + std::to_string(i), 4 * p->L, p->L, 1);
    if (p->embed) expect_embed(p.get(), 0);
    for (int g = 0; g < 5; ++g)
      expect_fcblock(p.get(), 0, std::string("Integration_") + kGroupNames[g], p->feat_pos, 3 * (int)p->groups.joints[g].size(), 1);
  }
  if (p->has_trj) {
    std::vector<int> all;
    for (int j = 0; j < p->J; ++j) all.push_back(j);
    r3d_plan::TB tb{1, "LocalLayer", all, kMaxProb - 1, -1};
    expect_tblock(p.get(), 1, tb.prefix, 3 * p->JC);
    p->tbs.push_back(tb);
    expect_fcblock(p.get(), 1, "GlobalInfo", p->JC, p->L, 2);
    if (p->embed) expect_embed(p.get(), 1);
    expect_fcblock(p.get(), 1, "Integration", p->feat_trj, 3, 1);
  }
  *out = p.release();
  return R3D_OK;
}

extern "C" R3D_API int r3d_plan_set_tensor(r3d_plan* p, int net, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!p || !name || !data || (ndim > 0 && !shape)) return fail(R3D_ERR_BAD_ARG, "r3d_plan_set_tensor: null argument");
  if (net != R3D_NET_POS && net != R3D_NET_TRJ) return fail(R3D_ERR_BAD_ARG, "net must be R3D_NET_POS or R3D_NET_TRJ");
  if (!(p->cfg.nets & net)) return fail(R3D_ERR_BAD_ARG, "plan was not created for net %d", net);
  std::string key(name);
  if (key.rfind("module.", 0) == 0) key = key.substr(7);   // nn.DataParallel checkpoints
  if (key.size() > 20 && key.compare(key.size() - 20, 20, ".num_batches_tracked") == 0) return R3D_OK;
  auto& tab = p->tensors[net == R3D_NET_POS ? 0 : 1];
  auto it = tab.find(key);
  if (it == tab.end()) return fail(R3D_ERR_BAD_ARG, "unknown state_dict key '%s' for this configuration", key.c_str());
  TensorEntry& e = it->second;
  bool same = (int)e.shape.size() == ndim;
  int64_t n = 1;
  for (int i = 0; same && i < ndim; ++i) { same = e.shape[i] == shape[i]; n *= shape[i]; }
  if (!same) {
    std::string want, got;
    for (auto d : e.shape) want += std::to_string(d) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    return fail(R3D_ERR_BAD_ARG, "shape mismatch for '%s': expected (%s) got (%s)", key.c_str(), want.c_str(), got.c_str());
  }
  e.data.assign(data, data + n);
  e.set = true;
  p->finalized = false;
  return R3D_OK;
}

// ---- folding + packing -----------------------------------------------------------------------------
namespace {
struct Folder {
  r3d_plan* p;
  int net;
  const float* get(const std::string& k) const { return p->tensors[net].at(k).data.data(); }
  // scale/shift of an eval-mode BatchNorm1d:  y = x*scale + shift   (running stats, eps 1e-5)
  void bn(const std::string& pre, int c, std::vector<double>& scale, std::vector<double>& shift) const {
    const float *g = get(pre + ".weight"), *b = get(pre + ".bias"), *m = get(pre + ".running_mean"), *v = get(pre + ".running_var");
    scale.resize(c); shift.resize(c);
    for (int i = 0; i < c; ++i) {
      scale[i] = (double)g[i] / std::sqrt((double)v[i] + kBnEps);
      shift[i] = (double)b[i] - (double)m[i] * scale[i];
    }
  }
  // conv weight (n, cin, w) [+ BN]  ->  [n_pad][k_pad] with column = tap*cin + c
  PackedLayer conv(const std::string& wkey, int n, int cin, int w, const std::string& bnpre, const std::string& biaskey) const {
    PackedLayer pl;
    pl.n = n; pl.k = cin * w; pl.k_pad = round_up(pl.k, kKAlign);
    pl.n_pad = (p->tail_fusion && n < kTailN) ? kTailN : round_up(n, 16);   // chained tail launch: narrow heads fill one 128-column unit
    pl.w.assign((size_t)pl.n_pad * pl.k_pad, 0.f); pl.b.assign(pl.n_pad, 0.f);
    std::vector<double> sc(n, 1.0), sh(n, 0.0);
    if (!bnpre.empty()) bn(bnpre, n, sc, sh);
    const float* W = get(wkey);
    const float* B = biaskey.empty() ? nullptr : get(biaskey);
    for (int o = 0; o < n; ++o) {
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < w; ++t)
          pl.w[(size_t)o * pl.k_pad + t * cin + c] = (float)((double)W[((size_t)o * cin + c) * w + t] * sc[o]);
      pl.b[o] = (float)((B ? (double)B[o] * sc[o] : 0.0) + sh[o]);
    }
    return pl;
  }
  PackedLayer linear(const std::string& pre, int n, int k, const std::string& bnpre) const {
    return conv(pre + ".weight", n, k, 1, bnpre, pre + ".bias");
  }
  PackedLayer plain(const std::string& pre, int n, int k, const std::string& bnpre) const {
    PackedLayer pl = conv(pre + ".weight", n, k, 1, bnpre, pre + ".bias");
    PackedLayer q;
    q.n = q.n_pad = n; q.k = q.k_pad = k; q.plain = true;
    q.w.resize((size_t)n * k); q.b.assign(pl.b.begin(), pl.b.begin() + n);
    for (int o = 0; o < n; ++o)
      for (int i = 0; i < k; ++i) q.w[(size_t)o * k + i] = pl.w[(size_t)o * pl.k_pad + i];
    return q;
  }
};
}  // namespace

// expand_conv is linear in its input [x_g | x_g - root | x_g - x_g[tc]] (rie.py:301-357), so the positional and temporal
// differences are folded into the weights instead of being materialised per joint group:
//   sum_{tap,ch} W[o,ch,tap] * in[ch, w0*t'+tap]
//     = sum_{tap,s} Wf[o,tap,s] * x[s, w0*t'+tap]  +  sum_s Wc[o,s] * x[s, tc]
//   Wf[o,tap,(j,c)] = [j in group] (Wx+Wd+Wt)[o,(jj,c),tap]  -  [j == root] sum_jj Wd[o,(jj,c),tap]
//   Wc[o,(j,c)]     = -[j in group] sum_tap Wt[o,(jj,c),tap]
// Every problem then reads the SAME compact operand row  {x[:, w0*t' .. w0*t'+w0-1], x[:, tc]}  (K = (w0+1)*(J+4)*Cin
// instead of w0*3*|group|*Cin per group: ~4x fewer bytes at T=243), summed in float64 with the BatchNorm scale.  The
// columns are ordered by build_a0_layout so that a group's non-zero weights sit in a few 16-column K steps.
static PackedLayer pack_expand_folded(r3d_plan* p, int net, const std::string& pre, const std::vector<int>& joints, int group) {
  Folder f{p, net};
  const int C = p->C, Cin = p->Cin, w0 = p->widths[0], nj = (int)joints.size(), cg = 3 * nj * Cin;
  PackedLayer pl;
  pl.n = C; pl.k = p->a0_k; pl.n_pad = round_up(C, 16); pl.k_pad = p->a0_kpad;
  pl.alg_k = cg * w0;
  pl.w.assign((size_t)pl.n_pad * pl.k_pad, 0.f); pl.b.assign(pl.n_pad, 0.f);
  std::vector<double> sc, sh;
  f.bn(pre + ".expand_bn", C, sc, sh);
  const float* W = f.get(pre + ".expand_conv.weight");     // (C, cg, w0), channel = part*nj*Cin + jj*Cin + c
  // operand columns (build_a0_layout): joint j, frame tap (tap == w0: frame tc), coordinate c
  const int root = group >= 0 ? p->a0_root[group] : p->a0_home[0];
  const int root_tap = p->a0_tap[group >= 0 ? group : 0];
  auto col = [&](int j, int tap, int c) { return p->a0_home[j] + tap * p->a0_tap[p->a0_slot_of[j]] + c; };
  std::vector<double> row(pl.k_pad);
  for (int o = 0; o < C; ++o) {
    std::fill(row.begin(), row.end(), 0.0);
    auto w = [&](int part, int jj, int c, int tap) { return (double)W[((size_t)o * cg + part * nj * Cin + jj * Cin + c) * w0 + tap]; };
    for (int jj = 0; jj < nj; ++jj)
      for (int c = 0; c < Cin; ++c)
        for (int tap = 0; tap < w0; ++tap) {
          row[col(joints[jj], tap, c)] += w(0, jj, c, tap) + w(1, jj, c, tap) + w(2, jj, c, tap);
          row[root + tap * root_tap + c] -= w(1, jj, c, tap);            // root joint 0, same coordinate (rie.py:301)
          row[col(joints[jj], w0, c)] -= w(2, jj, c, tap);               // x[:, tc] term (rie.py:304)
        }
    for (int k = 0; k < pl.k; ++k) pl.w[(size_t)o * pl.k_pad + k] = (float)(row[k] * sc[o]);
    pl.b[o] = (float)sh[o];
  }
  return pl;
}

static void pack_tblock(r3d_plan* p, int net, const std::string& pre, const std::vector<int>& joints, int group) {
  Folder f{p, net};
  const std::string key = std::to_string(net) + ":" + pre;
  const int C = p->C;
  p->layers[key + ".expand_conv"] = pack_expand_folded(p, net, pre, joints, group);
  for (size_t i = 1; i < p->widths.size(); ++i) {
    const std::string a = std::to_string(2 * (i - 1)), b = std::to_string(2 * (i - 1) + 1);
    p->layers[key + ".layers_conv." + a] = f.conv(pre + ".layers_conv." + a + ".weight", C, C, p->widths[i], pre + ".layers_bn." + a, "");
    p->layers[key + ".layers_conv." + b] = f.conv(pre + ".layers_conv." + b + ".weight", C, C, 1, pre + ".layers_bn." + b, "");
  }
  p->layers[key + ".shrink"] = f.conv(pre + ".shrink.weight", p->L, C, 1, "", pre + ".shrink.bias");
}
static void pack_fcblock(r3d_plan* p, int net, const std::string& pre, int cin, int cout, int nblocks) {
  Folder f{p, net};
  const std::string key = std::to_string(net) + ":" + pre;
  p->layers[key + ".fc_1"] = f.linear(pre + ".fc_1", kFcWidth, cin, pre + ".bn_1");
  for (int i = 0; i < nblocks; ++i) {
    const std::string q = ".layers." + std::to_string(i);
    p->layers[key + q + ".w1"] = f.linear(pre + q + ".w1", kFcWidth, kFcWidth, pre + q + ".batch_norm1");
    p->layers[key + q + ".w2"] = f.linear(pre + q + ".w2", kFcWidth, kFcWidth, pre + q + ".batch_norm2");
  }
  p->layers[key + ".fc_2"] = f.linear(pre + ".fc_2", cout, kFcWidth, "");
}
static void pack_embed(r3d_plan* p, int net) {
  Folder f{p, net};
  const std::string key = std::to_string(net) + ":embedder";
  p->layers[key + ".w1"] = f.plain("embedder.w1", kEmbedMid, p->ext, "embedder.b1");
  p->layers[key + ".w2"] = f.plain("embedder.w2", p->E, kEmbedMid, "embedder.b2");
}

// ---- graph construction --------------------------------------------------------------------------
static int add_mat(r3d_plan* p, int rows_per_seq, int ld, bool f32 = false) {
  MatReq m;
  m.rows_per_seq = rows_per_seq; m.ld = ld; m.f32 = f32;
  p->mats.push_back(m);
  return (int)p->mats.size() - 1;
}

static OpHost& add_op(r3d_plan* p, const std::string& name, int nprob, int rows_per_seq, float slope) {
  p->ops.emplace_back();
  OpHost& op = p->ops.back();
  memset(&op.dev, 0, sizeof(op.dev));
  op.name = name;
  op.dev.nprob = nprob;
  op.dev.rows_per_seq = rows_per_seq;
  op.dev.slope = slope;
  return op;
}

static void build_graph(r3d_plan* p) {
  p->mats.clear(); p->ops.clear(); p->emb_binds.clear(); p->m_a0.clear();
  const int C = p->C, L = p->L, nl = (int)p->widths.size(), ntb = (int)p->tbs.size();
  const float act = 0.2f;   // nn.LeakyReLU(0.2), rie.py:27,113,156

  // --- first-layer operand shared by every problem: row (b, t') = [x[b, w0*t' .. w0*t'+w0-1, :] | x[b, tc, :] | 0-pad]
  {
    const int a0 = add_mat(p, p->lens[0], p->a0_kpad);
    for (int q = 0; q < ntb; ++q) p->m_a0.push_back(a0);
  }
  p->m_inc = add_mat(p, 1, round_up(p->JC, kKAlign));

  // --- feature / head buffers
  int m_feat[kMaxProb] = {-1, -1, -1, -1, -1, -1}, m_fuse[5] = {-1, -1, -1, -1, -1};
  for (int q = 0; q < ntb; ++q) {
    const auto& tb = p->tbs[q];
    m_feat[q] = add_mat(p, 1, round_up(tb.net == 0 ? p->feat_pos : p->feat_trj, kKAlign));
    p->m_heads[tb.head] = add_mat(p, 1, 16, true);
  }
  const bool fuse = p->has_pos && p->cfg.stage != 1;
  if (fuse) for (int i = 0; i < 5; ++i) m_fuse[i] = add_mat(p, 1, 4 * L);

  // --- temporal tree: ping-pong X buffers, one Y scratch
  // tensor-core precisions with 256 channels: the k=w conv and the 1x1 conv of a level run as ONE launch whose
  // intermediate stays in tensor memory (gemm_tc_kernel FUSED); otherwise two launches through the Y scratch.
  bool fuse_pairs = p->cfg.precision != R3D_PREC_FP32 && C == 256;
  if (const char* env = exp_env("R3D_TC_FUSE")) fuse_pairs = fuse_pairs && atoi(env) != 0;
  // Levels with few rows per window (the top of the tree) are one-wave launches: a fused tile there serialises
  // GEMM -> convert -> GEMM -> store on a fraction of the SMs, while two narrow-tile launches spread over all of them.
  int fuse_min_rows = 3;
  if (const char* env = exp_env("R3D_TC_FUSE_MIN_ROWS")) fuse_min_rows = atoi(env);
  std::vector<char> lvl_fused(nl, 0);
  int y_len = 0;
  for (int i = 1; i < nl; ++i) {
    lvl_fused[i] = fuse_pairs && p->lens[i] >= fuse_min_rows;
    if (!lvl_fused[i]) y_len = std::max(y_len, p->lens[i]);
  }
  std::vector<int> xa(ntb), xb(ntb), yb(ntb);
  for (int q = 0; q < ntb; ++q) {
    xa[q] = add_mat(p, p->lens[0], C);
    xb[q] = nl > 1 ? add_mat(p, p->lens[1], C) : -1;
    yb[q] = y_len > 0 ? add_mat(p, y_len, C) : -1;
  }
  auto lkey = [&](int q, const std::string& s) { return std::to_string(p->tbs[q].net) + ":" + p->tbs[q].prefix + s; };
  {
    OpHost& op = add_op(p, "expand_conv", ntb, p->lens[0], act);                       // rie.py:86
    for (int q = 0; q < ntb; ++q) {
      auto& b = op.bind[q];
      b.a = p->m_a0[q]; b.a_ld = p->mats[p->m_a0[q]].ld; b.layer = lkey(q, ".expand_conv");
      b.dst = {{xa[q], 0}};
    }
  }
  std::vector<int> cur = xa, nxt = xb;
  for (int i = 1; i < nl; ++i) {
    const int w = p->widths[i];
    const std::string a = std::to_string(2 * (i - 1)), bb = std::to_string(2 * (i - 1) + 1);
    if (lvl_fused[i]) {
      OpHost& o = add_op(p, "layers_conv." + a + "+" + bb, ntb, p->lens[i], act);       // rie.py:94-97
      o.dev.fused2 = 1;
      for (int q = 0; q < ntb; ++q) {
        auto& b = o.bind[q];
        b.a = cur[q]; b.a_ld = w * C; b.layer = lkey(q, ".layers_conv." + a); b.layer2 = lkey(q, ".layers_conv." + bb);
        b.res = cur[q]; b.res_ld = w * C; b.res_col = (w / 2) * C;
        b.dst = {{nxt[q], 0}};
      }
      std::swap(cur, nxt);
      continue;
    }
    OpHost& o1 = add_op(p, "layers_conv." + a, ntb, p->lens[i], act);                  // rie.py:96
    for (int q = 0; q < ntb; ++q) {
      auto& b = o1.bind[q];
      b.a = cur[q]; b.a_ld = w * C; b.layer = lkey(q, ".layers_conv." + a); b.dst = {{yb[q], 0}};
    }
    OpHost& o2 = add_op(p, "layers_conv." + bb, ntb, p->lens[i], act);                 // rie.py:94,97
    for (int q = 0; q < ntb; ++q) {
      auto& b = o2.bind[q];
      b.a = yb[q]; b.a_ld = C; b.layer = lkey(q, ".layers_conv." + bb);
      b.res = cur[q]; b.res_ld = w * C; b.res_col = (w / 2) * C;                          // x[:, :, w//2::w]
      b.dst = {{nxt[q], 0}};
    }
    std::swap(cur, nxt);
  }
  {
    OpHost& op = add_op(p, "shrink", ntb, 1, 1.0f);                                     // rie.py:99
    for (int q = 0; q < ntb; ++q) {
      const auto& tb = p->tbs[q];
      auto& b = op.bind[q];
      b.a = cur[q]; b.a_ld = C; b.layer = lkey(q, ".shrink");
      b.dst = {{m_feat[q], 0}};
      if (tb.net == 0 && fuse)                                                            // rie.py:392-394
        for (int i = 0; i < 5; ++i)
          if (i != tb.group) b.dst.push_back({m_fuse[i], (tb.group < i ? tb.group : tb.group - 1) * L});
    }
  }

  // --- FC chains.  Hidden buffers: 3 x [B][1024] per problem slot, shared by successive chains.
  // GlobalInfo runs on the side stream concurrently with the tree / fuse chains, so it owns its own set.
  int hb_main[3][kMaxProb], hb_side[3][kMaxProb];
  for (int j = 0; j < 3; ++j)
    for (int q = 0; q < kMaxProb; ++q) {
      hb_main[j][q] = add_mat(p, 1, kFcWidth);
      hb_side[j][q] = q < 2 ? add_mat(p, 1, kFcWidth) : -1;
    }
  int (*hb)[kMaxProb] = hb_main;
  struct Chain { std::string key; int a; int a_ld; int nblocks; std::vector<std::pair<int, int>> dst; bool f32; };
  auto fc_chain = [&](const std::string& name, std::vector<Chain>& ch) {
    const int n = (int)ch.size();
    if (!n) return;
    {
      OpHost& op = add_op(p, name + ".fc_1", n, 1, act);                                // rie.py:161-163
      for (int q = 0; q < n; ++q) { auto& b = op.bind[q]; b.a = ch[q].a; b.a_ld = ch[q].a_ld; b.layer = ch[q].key + ".fc_1"; b.dst = {{hb[0][q], 0}}; }
    }
    int h = 0;
    const int nb = ch[0].nblocks;
    for (int i = 0; i < nb; ++i) {                                                      // rie.py:122-135
      const std::string ls = ".layers." + std::to_string(i);
      const int y = (h + 1) % 3, hn = (h + 2) % 3;
      OpHost& o1 = add_op(p, name + ls + ".w1", n, 1, act);
      for (int q = 0; q < n; ++q) { auto& b = o1.bind[q]; b.a = hb[h][q]; b.a_ld = kFcWidth; b.layer = ch[q].key + ls + ".w1"; b.dst = {{hb[y][q], 0}}; }
      OpHost& o2 = add_op(p, name + ls + ".w2", n, 1, act);
      for (int q = 0; q < n; ++q) {
        auto& b = o2.bind[q];
        b.a = hb[y][q]; b.a_ld = kFcWidth; b.layer = ch[q].key + ls + ".w2";
        b.res = hb[h][q]; b.res_ld = kFcWidth; b.res_col = 0; b.dst = {{hb[hn][q], 0}};
      }
      h = hn;
    }
    OpHost& op = add_op(p, name + ".fc_2", n, 1, 1.0f);                                 // rie.py:167
    for (int q = 0; q < n; ++q) {
      auto& b = op.bind[q];
      b.a = hb[h][q]; b.a_ld = kFcWidth; b.layer = ch[q].key + ".fc_2"; b.dst = ch[q].dst;
      b.dst_f32.assign(ch[q].dst.size(), ch[q].f32 ? 1 : 0);
    }
  };

  const int inc_ld = p->mats[p->m_inc].ld;
  const int gcol_pos = (p->cfg.stage == 1 ? 1 : 2) * L;
  {
    std::vector<Chain> ch;                                                               // rie.py:362, :543
    if (p->has_pos) {
      Chain c{"0:GlobalInfo", p->m_inc, inc_ld, 2, {}, false};
      for (int q = 0; q < ntb; ++q) if (p->tbs[q].net == 0) c.dst.push_back({m_feat[q], gcol_pos});
      ch.push_back(c);
    }
    if (p->has_trj) {
      Chain c{"1:GlobalInfo", p->m_inc, inc_ld, 2, {}, false};
      for (int q = 0; q < ntb; ++q) if (p->tbs[q].net == 1) c.dst.push_back({m_feat[q], L});
      ch.push_back(c);
    }
    const size_t first = p->ops.size();
    hb = hb_side;
    fc_chain("GlobalInfo", ch);
    hb = hb_main;
    for (size_t i = first; i < p->ops.size(); ++i) p->ops[i].side = true;   // depends only on the input stage
  }
  if (fuse) {                                                                            // rie.py:388-394
    std::vector<Chain> ch;
    for (int i = 0; i < 5; ++i) ch.push_back({"0:FuseBlocks." + std::to_string(i), m_fuse[i], 4 * L, 1, {{m_feat[i], L}}, false});
    fc_chain("FuseBlocks", ch);
  }
  {
    std::vector<Chain> ch;                                                               // rie.py:410-414, :555
    for (int q = 0; q < ntb; ++q) {
      const auto& tb = p->tbs[q];
      const std::string key = tb.net == 0 ? std::string("0:Integration_") + kGroupNames[tb.group] : std::string("1:Integration");
      ch.push_back({key, m_feat[q], p->mats[m_feat[q]].ld, 1, {{p->m_heads[tb.head], 0}}, true});
    }
    const size_t first = p->ops.size();
    fc_chain("Integration", ch);
    p->ops[first].join_before = true;                                          // needs x_global from the side stream
  }

  // --- embedder destinations (rie.py:375-380, 395-401, :549-551)
  if (p->embed) {
    if (p->has_pos) {
      r3d_plan::EmbBind e{0, {}};
      for (int q = 0; q < ntb; ++q) if (p->tbs[q].net == 0) e.dst.push_back({m_feat[q], gcol_pos + L});
      p->emb_binds.push_back(e);
    }
    if (p->has_trj) {
      r3d_plan::EmbBind e{1, {}};
      for (int q = 0; q < ntb; ++q) if (p->tbs[q].net == 1) e.dst.push_back({m_feat[q], 2 * L});
      p->emb_binds.push_back(e);
    }
  }
}

// One chained launch (r3d_tail_tc.cu): the ops `order` (plan indices, dependency order), their unit counts, and which
// earlier (op, problem) pairs each (op, problem) reads from.  Host only; pointers are resolved in bind_workspace.
static bool build_chain(r3d_plan* p, const std::vector<int>& order, int width_pref, MultiOpDev& mo) {
  memset(&mo, 0, sizeof(mo));
  bool ok = order.size() >= 2 && order.size() <= (size_t)kMaxTailOps;
  for (size_t li = 0; ok && li < order.size(); ++li) {
    const OpHost& op = p->ops[order[li]];
    ok = ok && op.dev.rows_per_seq == 1 && !op.dev.fused2;
    for (int q = 0; ok && q < op.dev.nprob; ++q) {
      const PackedLayer& l = p->layers.at(op.bind[q].layer);
      ok = ok && l.n_pad % kTailN == 0 && l.n_pad / kTailN <= 255 && l.k_pad % kKAlign == 0;
    }
  }
  if (!ok) return false;
  mo.nops = (int)order.size();
  for (int li = 0; ok && li < mo.nops; ++li) {
    const OpHost& op = p->ops[order[li]];
    mo.op_index[li] = order[li];
    mo.nprob[li] = (uint8_t)op.dev.nprob;
    int per_m = 0, width = width_pref >= 256 ? 2 : 1;            // 256-column units unless some problem of the op is narrower
    for (int q = 0; q < op.dev.nprob; ++q)
      if (p->layers.at(op.bind[q].layer).n_pad % (2 * kTailN)) width = 1;
    mo.width[li] = (uint8_t)width;
    for (int q = 0; q < op.dev.nprob; ++q) {
      mo.ntiles[li][q] = (uint8_t)(p->layers.at(op.bind[q].layer).n_pad / (kTailN * width));
      per_m += mo.ntiles[li][q];
      // producers: for every matrix this problem reads (operand, residual) and every column offset written into it, the
      // LAST earlier (op, problem) of the launch that writes there (activation buffers are recycled along a chain)
      std::map<std::pair<int, int>, int> writer;                   // (matrix, column) -> counter row
      for (int lj = 0; lj < li; ++lj) {
        const OpHost& pj = p->ops[order[lj]];
        for (int r = 0; r < pj.dev.nprob; ++r)
          for (auto& d : pj.bind[r].dst)
            if (d.first == op.bind[q].a || (op.bind[q].res >= 0 && d.first == op.bind[q].res)) writer[{d.first, d.second}] = lj * kMaxProb + r;
      }
      std::vector<int> rows;
      for (auto& kv : writer)
        if (std::find(rows.begin(), rows.end(), kv.second) == rows.end()) rows.push_back(kv.second);
      int nd = 0;
      for (int row : rows) {
        if (nd == kMaxDeps) { ok = false; break; }
        mo.dep[li][q][nd++] = (int16_t)row;
      }
      mo.ndep[li][q] = (uint8_t)nd;
    }
    mo.unit0[li + 1] = mo.unit0[li] + per_m;
  }
  if (!ok) memset(&mo, 0, sizeof(mo));
  return ok;
}

static void plan_chains(r3d_plan* p) {
  const int nops = (int)p->ops.size();
  const bool tc = p->cfg.precision != R3D_PREC_FP32;
  for (auto& c : p->chain) { c.ops.clear(); memset(&c.mo, 0, sizeof(c.mo)); c.max_clusters = 0; }
  if (tc && p->tail_fusion) {
    // main chain: the maximal suffix of one-row ops (a fused conv pair keeps its own launch).  The GlobalInfo chain is NOT
    // part of it: inside, its six dependent layers would sit in front of Integration.fc_1 (measured: 62 % of the kernel's
    // cycles spent waiting for inputs).
    int first = nops;
    for (int i = nops - 1; i >= 0; --i) {
      if (p->ops[i].side) continue;
      if (p->ops[i].dev.rows_per_seq != 1 || p->ops[i].dev.fused2) break;
      first = i;
    }
    std::vector<int> mains;
    for (int i = first; i < nops; ++i)
      if (!p->ops[i].side) mains.push_back(i);
    if (build_chain(p, mains, p->tail_width, p->chain[0].mo)) {
      p->chain[0].ops = mains;
      p->chain[0].max_clusters = p->tail_clusters;
    }
  }
  if (tc && p->side_chain) {
    // The GlobalInfo chain only depends on the input stage and has ~600 us of slack before Integration.fc_1 needs it (at
    // B=1024, T=243).  As six under-filled launches racing the tree's persistent kernels for SMs it costs the step 90 us
    // (measured by skipping it); as ONE chained launch on a few CTA pairs it takes longer but occupies far fewer SM-microseconds.
    std::vector<int> sides;
    for (int i = 0; i < nops; ++i)
      if (p->ops[i].side) sides.push_back(i);
    // Only worth it when the chain is a small fraction of the work the tree does before Integration.fc_1 needs its output
    // (T = 243: 11 %): on few CTA pairs it takes as long as (share x 74 / pairs) of that time.  Short receptive fields keep
    // one launch per layer.  CTA pairs: 1.5x the flop share unless the option "side_clusters" fixes it.
    auto op_flops = [&](const OpHost& op) {
      double f = 0;
      for (int q = 0; q < op.dev.nprob; ++q) {
        const PackedLayer& l = p->layers.at(op.bind[q].layer);
        f += (double)op.dev.rows_per_seq * l.n * (l.alg_k ? l.alg_k : l.k);
        if (!op.bind[q].layer2.empty()) f += (double)op.dev.rows_per_seq * p->layers.at(op.bind[q].layer2).n * p->layers.at(op.bind[q].layer2).k;
      }
      return f;
    };
    double f_side = 0, f_tree = 0;
    for (const OpHost& op : p->ops) {
      if (op.side) f_side += op_flops(op);
      else if (!op.join_before && f_tree >= 0) f_tree += op_flops(op);
      if (op.join_before) break;                                  // ops from the first consumer on do not hide the chain
    }
    const double share = f_tree > 0 ? f_side / f_tree : 1.0;
    if ((share <= 0.2 || p->side_chain_force) && build_chain(p, sides, 256, p->chain[1].mo)) {
      p->chain[1].ops = sides;
      p->chain[1].max_clusters = p->side_clusters > 0 ? p->side_clusters : std::min(24, std::max(4, (int)std::ceil(74.0 * share * 1.5)));
    }
  }
  // the GEMM launches of one forward, for every combination of chains in use (a chain needs batches of >= 256 windows)
  for (int v = 0; v < 4; ++v) {
    auto& out = p->launches[v];
    out.clear();
    std::vector<int> of(nops, -1);
    for (int c = 0; c < 2; ++c)
      if ((v >> c) & 1)
        for (int i : p->chain[c].ops) of[i] = c;
    bool listed[2] = {false, false};
    for (int i = 0; i < nops; ++i) {
      const int c = of[i];
      if (c < 0) { out.push_back({p->ops[i].name, {i}, -1}); continue; }
      // a chained launch goes where its LAST op stood: every launch it reads from has been enqueued (and joined) before it
      const auto& co = p->chain[c].ops;
      if (listed[c] || i != *std::max_element(co.begin(), co.end())) continue;
      listed[c] = true;
      out.push_back({std::string(c ? "side[" : "tail[") + p->ops[co.front()].name + " .. " + p->ops[co.back()].name + "]", co, c});
    }
  }
}

static int pick_n_tile(int n_pad) {
  for (int t : {256, 128, 64, 32, 16})
    if (n_pad % t == 0) return t;
  return 16;
}

extern "C" R3D_API int r3d_plan_finalize(r3d_plan* p) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  for (int net = 0; net < 2; ++net)
    for (auto& kv : p->tensors[net])
      if (!kv.second.set)
        return fail(R3D_ERR_MISSING_WEIGHT, "state_dict entry '%s' (%s net) was never supplied", kv.first.c_str(), net ? "trj" : "pos");
  p->layers.clear();
  if (p->has_pos) {
    for (int g = 0; g < 5; ++g)
      pack_tblock(p, 0, std::string("LocalLayer_") + kGroupNames[g], p->groups.joints[g], g);
    pack_fcblock(p, 0, "GlobalInfo", p->JC, p->L, 2);
    if (p->cfg.stage != 1)
      for (int i = 0; i < 5; ++i) pack_fcblock(p, 0, "FuseBlocks." + std::to_string(i), 4 * p->L, p->L, 1);
    if (p->embed) pack_embed(p, 0);
    for (int g = 0; g < 5; ++g)
      pack_fcblock(p, 0, std::string("Integration_") + kGroupNames[g], p->feat_pos, 3 * (int)p->groups.joints[g].size(), 1);
  }
  if (p->has_trj) {
    {
      std::vector<int> all;
      for (int j = 0; j < p->J; ++j) all.push_back(j);
      pack_tblock(p, 1, "LocalLayer", all, -1);
    }
    pack_fcblock(p, 1, "GlobalInfo", p->JC, p->L, 2);
    if (p->embed) pack_embed(p, 1);
    pack_fcblock(p, 1, "Integration", p->feat_trj, 3, 1);
  }
  build_graph(p);
  plan_chains(p);
  // weight slab layout
  const int prec = p->cfg.precision;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  for (auto& kv : p->layers) {
    PackedLayer& l = kv.second;
    l.kmask = 0;
    if (!l.plain && l.k_pad <= 1024) {           // which 16-column K steps carry weight at all
      for (int o = 0; o < l.n_pad; ++o)
        for (int k = 0; k < l.k_pad; ++k)
          if (l.w[(size_t)o * l.k_pad + k] != 0.f) l.kmask |= 1ull << (k / 16);
      if (l.kmask == 0) l.kmask = 1ull;          // an all-zero layer still has to write its accumulator
    }
    const size_t ne = l.w.size();
    if (l.plain || prec == R3D_PREC_FP32) { l.off_w0 = take(ne * 4); l.off_w1 = 0; }
    else { l.off_w0 = take(ne * 2); l.off_w1 = prec == R3D_PREC_BF16X3 ? take(ne * 2) : 0; }
    l.off_b = take(l.b.size() * 4);
  }
  p->weight_bytes = off;
  p->finalized = true;
  p->uploaded = false;
  return R3D_OK;
}

extern "C" R3D_API int r3d_plan_packed_layer(const r3d_plan* p, int net, const char* layer, int32_t* n_pad, int32_t* k_pad,
                                     float* w_out, int64_t cap, float* b_out, int64_t cap_b) {
  if (!p || !layer) return fail(R3D_ERR_BAD_ARG, "null argument");
  if (!p->finalized) return fail(R3D_ERR_STATE, "plan not finalized");
  const std::string key = std::to_string(net == R3D_NET_POS ? 0 : 1) + ":" + layer;
  auto it = p->layers.find(key);
  if (it == p->layers.end()) return fail(R3D_ERR_BAD_ARG, "no packed layer '%s'", key.c_str());
  const PackedLayer& l = it->second;
  if (n_pad) *n_pad = l.n_pad;
  if (k_pad) *k_pad = l.k_pad;
  if (w_out) memcpy(w_out, l.w.data(), sizeof(float) * std::min<int64_t>(cap, (int64_t)l.w.size()));
  if (b_out) memcpy(b_out, l.b.data(), sizeof(float) * std::min<int64_t>(cap_b, (int64_t)l.b.size()));
  return R3D_OK;
}

// JSON description of the launch graph (buffers, ops, bindings, input-stage tables).  Lets the CPU test-suite
// replay the exact wiring with numpy against the oracle without a GPU.
extern "C" R3D_API int r3d_plan_describe(const r3d_plan* p, char* out, int64_t cap, int64_t* needed) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  if (!p->finalized) return fail(R3D_ERR_STATE, "plan not finalized");
  std::string j = "{";
  auto kv = [&](const char* k, long v) { j += std::string("\"") + k + "\":" + std::to_string(v) + ","; };
  kv("T", p->T); kv("J", p->J); kv("Cin", p->Cin); kv("tc", p->tc); kv("w0", p->widths[0]); kv("L0", p->lens[0]);
  kv("inc", p->m_inc); kv("ext", p->ext); kv("emb_mid", p->embed ? kEmbedMid : 0); kv("emb_dim", p->E);
  kv("has_pos", p->has_pos); kv("has_trj", p->has_trj);
  j += "\"mats\":[";
  for (size_t i = 0; i < p->mats.size(); ++i)
    j += std::string(i ? "," : "") + "[" + std::to_string(p->mats[i].rows_per_seq) + "," + std::to_string(p->mats[i].ld) + "," +
         std::to_string((int)p->mats[i].f32) + "]";
  j += "],\"a0\":[";
  for (size_t i = 0; i < p->m_a0.size(); ++i) j += std::string(i ? "," : "") + std::to_string(p->m_a0[i]);
  j += "],\"a0_map\":[";
  for (size_t i = 0; i < p->a0_src.size(); ++i) j += std::string(i ? "," : "") + std::to_string(p->a0_src[i]);
  j += "],\"heads\":[";
  for (int q = 0; q < kMaxProb; ++q) j += std::string(q ? "," : "") + std::to_string(p->m_heads[q]);
  j += "],\"slots\":[";
  for (int s2 = 0; s2 < p->J; ++s2)
    j += std::string(s2 ? "," : "") + "[" + std::to_string(p->groups.slots[s2].first) + "," + std::to_string(p->groups.slots[s2].second) + "]";
  j += "],\"embed\":[";
  for (size_t e = 0; e < p->emb_binds.size(); ++e) {
    j += std::string(e ? "," : "") + "{\"net\":" + std::to_string(p->emb_binds[e].net) + ",\"dst\":[";
    for (size_t d = 0; d < p->emb_binds[e].dst.size(); ++d)
      j += std::string(d ? "," : "") + "[" + std::to_string(p->emb_binds[e].dst[d].first) + "," + std::to_string(p->emb_binds[e].dst[d].second) + "]";
    j += "]}";
  }
  j += "],\"launches\":[";                            // the variant with every planned chain in use (batches >= 256)
  {
    const int v = (p->chain[0].ops.empty() ? 0 : 1) | (p->chain[1].ops.empty() ? 0 : 2);
    for (size_t k = 0; k < p->launches[v].size(); ++k) {
      const auto& L = p->launches[v][k];
      j += std::string(k ? "," : "") + "{\"name\":\"" + L.name + "\",\"tail\":" + (L.chain >= 0 ? "1" : "0") + ",\"chain\":" + std::to_string(L.chain) + ",\"ops\":[";
      for (size_t q = 0; q < L.ops.size(); ++q) j += std::string(q ? "," : "") + std::to_string(L.ops[q]);
      j += "]}";
    }
  }
  j += "],\"tail\":{\"unit0\":[";
  const MultiOpDev& tmo = p->chain[0].ops.empty() ? p->chain[1].mo : p->chain[0].mo;
  for (int i = 0; i <= tmo.nops && tmo.nops > 0; ++i) j += std::string(i ? "," : "") + std::to_string(tmo.unit0[i]);
  j += "],\"deps\":[";                               // per chained op, per problem: [local producer op, producer problem] pairs
  for (int i = 0; i < tmo.nops; ++i) {
    j += std::string(i ? "," : "") + "[";
    for (int q = 0; q < tmo.nprob[i]; ++q) {
      j += std::string(q ? "," : "") + "[";
      for (int d = 0; d < tmo.ndep[i][q]; ++d)
        j += std::string(d ? "," : "") + "[" + std::to_string(tmo.dep[i][q][d] / kMaxProb) + "," + std::to_string(tmo.dep[i][q][d] % kMaxProb) + "]";
      j += "]";
    }
    j += "]";
  }
  j += "]},\"ops\":[";
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const OpHost& op = p->ops[i];
    char sl[32];
    snprintf(sl, sizeof(sl), "%.9g", op.dev.slope);
    j += std::string(i ? "," : "") + "{\"name\":\"" + op.name + "\",\"rows_per_seq\":" + std::to_string(op.dev.rows_per_seq) +
         ",\"slope\":" + sl + ",\"n_tile\":" + std::to_string(op.dev.n_tile) + ",\"prob\":[";
    for (int q = 0; q < op.dev.nprob; ++q) {
      const auto& b = op.bind[q];
      const PackedLayer& pl = p->layers.at(b.layer);
      j += std::string(q ? "," : "") + "{\"a\":" + std::to_string(b.a) + ",\"a_ld\":" + std::to_string(b.a_ld) + ",\"n\":" + std::to_string(pl.n) +
           ",\"k\":" + std::to_string(pl.k) + ",\"n_pad\":" + std::to_string(pl.n_pad) + ",\"k_pad\":" + std::to_string(pl.k_pad) + ",\"alg_k\":" + std::to_string(pl.alg_k ? pl.alg_k : pl.k) + ",\"k_steps\":" + std::to_string(pl.kmask ? __builtin_popcountll(pl.kmask) : pl.k_pad / 16) +
           (b.layer2.empty() ? std::string() : ",\"layer2\":\"" + b.layer2 + "\",\"n2\":" + std::to_string(p->layers.at(b.layer2).n) +
                                                    ",\"k2\":" + std::to_string(p->layers.at(b.layer2).k)) +
           ",\"layer\":\"" + b.layer +
           "\",\"res\":" + std::to_string(b.res) + ",\"res_ld\":" + std::to_string(b.res_ld) + ",\"res_col\":" + std::to_string(b.res_col) +
           ",\"dst\":[";
      for (size_t d = 0; d < b.dst.size(); ++d)
        j += std::string(d ? "," : "") + "[" + std::to_string(b.dst[d].first) + "," + std::to_string(b.dst[d].second) + "]";
      j += "]}";
    }
    j += "]}";
  }
  j += "]}";
  if (needed) *needed = (int64_t)j.size() + 1;
  if (out && cap > 0) {
    const size_t n = std::min<size_t>((size_t)cap - 1, j.size());
    memcpy(out, j.data(), n);
    out[n] = 0;
  }
  return R3D_OK;
}

extern "C" R3D_API int64_t r3d_plan_weight_bytes(const r3d_plan* p) { return p ? (int64_t)p->weight_bytes : 0; }
extern "C" R3D_API int64_t r3d_plan_workspace_bytes(const r3d_plan* p) { return p ? (int64_t)p->ws_bytes : 0; }
extern "C" R3D_API int r3d_plan_receptive_field(const r3d_plan* p) { return p ? p->T : 0; }
extern "C" R3D_API int r3d_plan_kernel_launches(const r3d_plan* p) {     // at batches >= 256 (chained launches in use)
  return p ? (int)p->launches[(p->chain[0].ops.empty() ? 0 : 1) | (p->chain[1].ops.empty() ? 0 : 2)].size() + 2 : 0;
}
extern "C" R3D_API int64_t r3d_plan_graph_launches(const r3d_plan* p) { return p ? (int64_t)p->graph_launches : 0; }

// ---- device upload -------------------------------------------------------------------------------
static uint16_t f2bf(float f) {   // round-to-nearest-even, like __float2bfloat16_rn (finite inputs)
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static void free_device(r3d_plan* p) {
  clear_graphs(p);
  if (p->device < 0) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  if (p->twin) {                                 // the second lane goes with the state it was cloned from
    cudaDeviceSynchronize();
    free_device(p->twin);
    delete p->twin;
    p->twin = nullptr;
  }
  if (p->ev_sub) cudaEventDestroy(p->ev_sub);
  p->ev_sub = nullptr;
  if (p->d_weights && !p->is_twin) cudaFree(p->d_weights);
  if (p->d_ws) cudaFree(p->d_ws);
  if (p->d_desc) cudaFree(p->d_desc);
  if (p->d_stage) cudaFree(p->d_stage);
  if (p->d_vid) cudaFree(p->d_vid);
  p->d_vid = nullptr; p->vid_bytes = 0;
  if (p->ev_ws) cudaEventDestroy(p->ev_ws);
  p->ev_ws = nullptr;
  for (int i = 0; i < r3d_plan::kSlots; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
    p->ev_in[i] = p->ev_done[i] = nullptr;
  }
  for (int i = 0; i < r3d_plan::kTicketRing; ++i)
    for (int l = 0; l < 2; ++l) {
      if (p->ev_ticket[i][l]) cudaEventDestroy(p->ev_ticket[i][l]);
      p->ev_ticket[i][l] = nullptr;
    }
  for (auto& e : p->prof_ev) cudaEventDestroy(e);
  p->prof_ev.clear();
  p->prof_runs = 0;
  if (p->s_copy) cudaStreamDestroy(p->s_copy);
  if (p->s_comp) cudaStreamDestroy(p->s_comp);
  if (p->s_side) cudaStreamDestroy(p->s_side);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  p->s_side = nullptr; p->ev_fork = p->ev_join = nullptr;
  p->d_weights = p->d_ws = p->d_desc = p->d_stage = nullptr;
  p->s_copy = p->s_comp = nullptr;
  p->ws_bytes = p->stage_bytes = 0;
  p->cap = 0;
  p->uploaded = false;
  cudaSetDevice(prev);
}

extern "C" R3D_API void r3d_plan_destroy(r3d_plan* p) {
  if (!p) return;
  free_device(p);
  delete p;
}

static int create_side_stream(r3d_plan* p) {
  int prio = 0;
  if (const char* env = exp_env("R3D_SIDE_PRIO")) prio = atoi(env);
  int lo = 0, hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // lo = least urgent (numerically largest)
  CUDA_TRY(cudaStreamCreateWithPriority(&p->s_side, cudaStreamNonBlocking, prio > 0 ? lo : prio < 0 ? hi : 0));
  return R3D_OK;
}

extern "C" R3D_API int r3d_plan_upload(r3d_plan* p, int device) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  if (!p->finalized) return fail(R3D_ERR_STATE, "r3d_plan_upload before r3d_plan_finalize");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(R3D_ERR_NO_DEVICE, "no CUDA device visible: ray3d_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(R3D_ERR_BAD_ARG, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(R3D_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
  std::lock_guard<std::mutex> lk(*p->mu);
  free_device(p);
  p->device = device;
  DeviceGuard dg(device);                  // the caller's current device is restored on every exit path
  CUDA_TRY(dg.err);
  CUDA_TRY(cudaMalloc(&p->d_weights, p->weight_bytes));
  std::vector<char> slab(p->weight_bytes, 0);
  const int prec = p->cfg.precision;
  for (auto& kv : p->layers) {
    PackedLayer& l = kv.second;
    if (l.plain || prec == R3D_PREC_FP32) {
      memcpy(slab.data() + l.off_w0, l.w.data(), l.w.size() * 4);
    } else {
      uint16_t* hi = reinterpret_cast<uint16_t*>(slab.data() + l.off_w0);
      uint16_t* lo = prec == R3D_PREC_BF16X3 ? reinterpret_cast<uint16_t*>(slab.data() + l.off_w1) : nullptr;
      for (size_t i = 0; i < l.w.size(); ++i) {
        hi[i] = f2bf(l.w[i]);
        if (lo) lo[i] = f2bf(l.w[i] - bf2f(hi[i]));
      }
    }
    memcpy(slab.data() + l.off_b, l.b.data(), l.b.size() * 4);
  }
  CUDA_TRY(cudaMemcpy(p->d_weights, slab.data(), p->weight_bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaDeviceSynchronize());   // (same: the weights must have landed before any non-blocking stream reads them)
  CUDA_TRY(prologue_configure(200 * 1024 + 1024));
  if (prec != R3D_PREC_FP32) {
    CUDA_TRY(tc_configure());
    CUDA_TRY(tail_configure());
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
  {   // experiment knobs: R3D_SIDE_STREAM=0 serialises the GlobalInfo chain; R3D_SIDE_PRIO=-1/0/1 sets its stream priority
    if (const char* env = exp_env("R3D_SIDE_STREAM")) p->use_side_stream = atoi(env) != 0;
    if (const char* env = exp_env("R3D_GRAPH_MAX_BATCH")) p->graph_max_batch = std::max(0, atoi(env));
    const int rc = create_side_stream(p);
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < r3d_plan::kSlots; ++i) {
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < r3d_plan::kTicketRing; ++i)
    for (int l = 0; l < 2; ++l) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_ticket[i][l], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&p->ev_sub, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&p->ev_ws, cudaEventDisableTiming));
  if (const char* env = exp_env("R3D_LANES")) p->use_lanes = atoi(env) >= 2;
  p->uploaded = true;
  return R3D_OK;
}

// Lane `lane` of an uploaded plan (0 = the plan itself).  The twin is cloned on first use: same packed layers / launch
// graph / weight slab, its own workspace, descriptors, compute + side streams.  Caller holds p->mu.
static int get_lane(r3d_plan* p, int lane, r3d_plan** out) {
  *out = p;
  if (lane == 0) return R3D_OK;
  if (!p->twin) {
    std::unique_ptr<r3d_plan> t(new r3d_plan(*p));
    t->mu = std::make_shared<std::mutex>();
    t->is_twin = true;
    t->twin = nullptr;
    for (int net = 0; net < 2; ++net) t->tensors[net].clear();
    for (auto& kv : t->layers) { std::vector<float>().swap(kv.second.w); std::vector<float>().swap(kv.second.b); }
    t->d_ws = nullptr; t->ws_bytes = 0; t->cap = 0; t->d_desc = nullptr; t->d_stage = nullptr; t->stage_bytes = 0;
    t->graphs.clear(); t->graph_max_batch = 0; t->graph_launches = 0;
    t->prof_ev.clear(); t->prof_runs = 0; t->profiling = false;
    t->s_copy = t->s_comp = t->s_side = nullptr;
    t->ev_fork = t->ev_join = t->ev_sub = t->ev_ws = nullptr;
    t->d_vid = nullptr; t->vid_bytes = 0;
    for (int i = 0; i < r3d_plan::kSlots; ++i) t->ev_in[i] = t->ev_done[i] = nullptr;
    for (int i = 0; i < r3d_plan::kTicketRing; ++i) t->ev_ticket[i][0] = t->ev_ticket[i][1] = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&t->s_comp, cudaStreamNonBlocking));
    int rc = create_side_stream(t.get());
    if (rc) return rc;
    CUDA_TRY(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&t->ev_ws, cudaEventDisableTiming));
    p->twin = t.release();
  }
  *out = p->twin;
  return R3D_OK;
}

// (re)allocate the activation workspace for `cap` sequences and resolve every symbolic binding
static int bind_workspace(r3d_plan* p, int cap) {
  const int prec = p->cfg.precision;
  if (p->d_ws) { CUDA_TRY(cudaDeviceSynchronize()); CUDA_TRY(cudaFree(p->d_ws)); p->d_ws = nullptr; }
  clear_graphs(p);                               // captured launches point into the old workspace / descriptors
  if (p->d_desc) { CUDA_TRY(cudaFree(p->d_desc)); p->d_desc = nullptr; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  for (auto& m : p->mats) {
    const size_t ne = (size_t)cap * m.rows_per_seq * m.ld;
    if (m.f32 || prec == R3D_PREC_FP32) { m.off0 = take(ne * 4); m.off1 = 0; }
    else { m.off0 = take(ne * 2); m.off1 = prec == R3D_PREC_BF16X3 ? take(ne * 2) : 0; }
  }
  p->ws_bytes = off;
  CUDA_TRY(cudaMalloc(&p->d_ws, p->ws_bytes));
  // zero padding must be in place before the first kernel touches it: cudaMemset runs on the legacy stream and may
  // return early, while the plan's own streams are non-blocking (no implicit ordering with it) -> wait for it here
  CUDA_TRY(cudaMemset(p->d_ws, 0, p->ws_bytes));
  CUDA_TRY(cudaDeviceSynchronize());
  p->cap = cap;
  auto mat = [&](int id, int ld) {
    Mat m{};
    if (id < 0) return m;
    const MatReq& r = p->mats[id];
    m.p0 = p->d_ws + r.off0;
    m.p1 = (r.f32 || prec != R3D_PREC_BF16X3) ? nullptr : p->d_ws + r.off1;
    m.ld = ld > 0 ? ld : r.ld;
    return m;
  };
  auto dsts = [&](const std::vector<std::pair<int, int>>& v, const std::vector<int>& f32, Dst* out) {
    for (size_t i = 0; i < v.size(); ++i) {
      out[i].m = mat(v[i].first, 0);
      out[i].col = v[i].second;
      out[i].f32 = (i < f32.size() && f32[i]) || p->mats[v[i].first].f32;
    }
    return (int)v.size();
  };
  for (auto& op : p->ops) {
    int ntile = 256;
    for (int q = 0; q < op.dev.nprob; ++q) {
      const auto& b = op.bind[q];
      const PackedLayer& l = p->layers.at(b.layer);
      GemmProb& g = op.dev.prob[q];
      g.a = mat(b.a, b.a_ld);
      g.w0 = p->d_weights + l.off_w0;
      g.w1 = (prec == R3D_PREC_BF16X3) ? p->d_weights + l.off_w1 : nullptr;
      g.bias = reinterpret_cast<const float*>(p->d_weights + l.off_b);
      g.res = mat(b.res, b.res_ld);
      g.res_col = b.res_col;
      g.K = l.k_pad; g.N = l.n; g.n_pad = l.n_pad;
      g.kmask = l.kmask;                           // K steps whose weights are all zero are never loaded or multiplied
      if (const char* env = exp_env("R3D_TC_KMASK")) if (atoi(env) == 0) g.kmask = 0;
      g.ndst = dsts(b.dst, b.dst_f32, g.dst);
      ntile = std::min(ntile, pick_n_tile(l.n_pad));
      if (!b.layer2.empty()) {
        const PackedLayer& l2 = p->layers.at(b.layer2);
        g.w2_0 = p->d_weights + l2.off_w0;
        g.w2_1 = (prec == R3D_PREC_BF16X3) ? p->d_weights + l2.off_w1 : nullptr;
        g.bias2 = reinterpret_cast<const float*>(p->d_weights + l2.off_b);
        g.K2 = l2.k_pad;
        g.N = l2.n;
      }
    }
    // Tile width for the tensor path.  The widest tile maximises operand reuse (least L2->SM traffic per flop), but the
    // small-M launches (upper tree levels, FC chains) then occupy a fraction of the 148 SMs.
    //  * latency policy: minimise waves x (tile cost) with a fixed per-tile overhead, wave count at the plan's capacity
    //    batch -- the shortest launch when the forward has the GPU to itself (small batches);
    //  * throughput policy: keep the widest tile -- with two batches in flight (lanes) the SMs a narrow-grid launch leaves
    //    free run the other batch's kernels, so SM-seconds per flop decide, not the launch's own duration (measured at
    //    B=1024: every one of these launches 8-25 % longer in isolation, the step 1-2 % shorter).
    // auto: throughput from 256 windows of capacity (where the chained launches start as well).
    bool tile_heuristic = p->tile_policy == 1 || (p->tile_policy == 0 && !tail_uses_pairs(cap));
    if (const char* env = exp_env("R3D_TC_TILE_HEUR")) tile_heuristic = atoi(env) != 0;
    op.dev.n_tile_tail = 0;
    for (const auto& c : p->chain) {
      const auto it = std::find(c.ops.begin(), c.ops.end(), (int)(&op - p->ops.data()));
      if (it != c.ops.end()) op.dev.n_tile_tail = kTailN * c.mo.width[it - c.ops.begin()];
    }
    if (prec != R3D_PREC_FP32 && ntile > 64 && !op.dev.fused2 && tile_heuristic) {
      const int64_t m_tiles = ((int64_t)cap * op.dev.rows_per_seq + 127) / 128;
      auto cost = [&](int bn) {
        int64_t tiles = 0;
        for (int q = 0; q < op.dev.nprob; ++q) tiles += m_tiles * (op.dev.prob[q].n_pad / bn);
        return ((tiles + 147) / 148) * (int64_t)(bn + 48);
      };
      int best = ntile;
      for (int bn = ntile / 2; bn >= 64; bn /= 2)
        if (cost(bn) < cost(best)) best = bn;
      ntile = best;
    }
    op.dev.n_tile = ntile;
  }
  {   // alternate the tile walk direction along each dependency chain (input stage writes front-to-back)
    int main_i = 0, side_i = 0;
    for (auto& op : p->ops) {
      int& i = op.side ? side_i : main_i;
      op.dev.reverse = (i % 2 == 0) ? 1 : 0;
      ++i;
    }
    if (const char* env = exp_env("R3D_TC_YSPLIT")) if (atoi(env) == 0) for (auto& op : p->ops) op.dev.flags |= 1;
    if (const char* env = exp_env("R3D_TC_RELEASE_ARRIVE")) if (atoi(env) != 0) for (auto& op : p->ops) op.dev.flags |= 2;
    if (const char* env = exp_env("R3D_TC_REVERSE")) if (atoi(env) == 0) for (auto& op : p->ops) op.dev.reverse = 0;
  }
  // prologue
  PrologueDev& pd = p->pro;
  memset(&pd, 0, sizeof(pd));
  pd.T = p->T; pd.J = p->J; pd.Cin = p->Cin; pd.JC = p->JC; pd.tc = p->tc; pd.w0 = p->widths[0]; pd.L0 = p->lens[0];
  pd.a0 = mat(p->m_a0[0], 0);
  pd.k_pad = p->mats[p->m_a0[0]].ld;
  pd.a0_off = nullptr;                           // patched below once the descriptor slab's address is known
  for (int j = 0; j < 32; ++j) pd.flip_perm[j] = (int8_t)(j < (int)p->flip_in.size() ? p->flip_in[j] : j);
  pd.inc = mat(p->m_inc, 0);
  pd.n_embed = (int)p->emb_binds.size(); pd.ext_dim = p->ext; pd.emb_mid = p->embed ? kEmbedMid : 0; pd.emb_dim = p->E;
  for (int e = 0; e < pd.n_embed; ++e) {
    const std::string key = std::to_string(p->emb_binds[e].net) + ":embedder";
    const PackedLayer &l1 = p->layers.at(key + ".w1"), &l2 = p->layers.at(key + ".w2");
    pd.embed[e].w1 = reinterpret_cast<const float*>(p->d_weights + l1.off_w0);
    pd.embed[e].b1 = reinterpret_cast<const float*>(p->d_weights + l1.off_b);
    pd.embed[e].w2 = reinterpret_cast<const float*>(p->d_weights + l2.off_w0);
    pd.embed[e].b2 = reinterpret_cast<const float*>(p->d_weights + l2.off_b);
    pd.embed[e].ndst = dsts(p->emb_binds[e].dst, {}, pd.embed[e].dst);
  }
  // assemble
  AssembleDev& ad = p->asmb;
  memset(&ad, 0, sizeof(ad));
  for (int q = 0; q < kMaxProb; ++q)
    ad.heads[q] = p->m_heads[q] >= 0 ? reinterpret_cast<const float*>(p->d_ws + p->mats[p->m_heads[q]].off0) : nullptr;
  ad.head_ld = 16; ad.J = p->J; ad.has_pos = p->has_pos; ad.has_trj = p->has_trj;
  for (int s = 0; s < p->J; ++s) {
    ad.slot_prob[s] = (int16_t)p->groups.slots[s].first;
    ad.slot_joint[s] = (int16_t)p->groups.slots[s].second;
    ad.flip_slot[s] = (int16_t)(s < (int)p->flip_out.size() ? p->flip_out[s] : s);
  }

  // descriptor slab: [ops][prologue][assemble][tensor maps]
  const size_t nops = p->ops.size();
  off = 0;
  p->off_ops = take(nops * sizeof(GemmOpDev));
  p->off_pro = take(sizeof(PrologueDev));
  p->off_asm = take(sizeof(AssembleDev));
  p->off_tmaps = take(nops * kMaxProb * kTmapsPerProb * kTmapBytes);
  p->off_sched = take(nops * r3d_plan::kSchedStride);                    // one work-unit counter per GEMM launch, 128 bytes apart
  const int m_groups_cap = (cap + 127) / 128;
  for (auto& c : p->chain)                                     // completion counters, contiguous with the work-unit counters: one memset
    c.off_done = take((size_t)std::max(1, c.mo.nops) * kMaxProb * m_groups_cap * sizeof(uint32_t));
  p->zero_bytes = off - p->off_sched;
  for (auto& c : p->chain) c.off_mo = take(sizeof(MultiOpDev));
  const size_t off_map = take(p->a0_src.size() * sizeof(int32_t));
  CUDA_TRY(cudaMalloc(&p->d_desc, off));
  pd.a0_off = reinterpret_cast<const int32_t*>(p->d_desc + off_map);
  std::vector<char> h(off, 0);
  {   // a0_src (index into [w0 frames | frame tc]) -> offset into the staged window, row-relative flag in bit 30
    const int k_frames = p->widths[0] * p->JC;
    for (size_t k = 0; k < p->a0_src.size(); ++k) {
      const int sidx = p->a0_src[k];
      reinterpret_cast<int32_t*>(h.data() + off_map)[k] =
          sidx < 0 ? p->T * p->JC : (sidx < k_frames ? (sidx | (1 << 30)) : sidx - k_frames + p->tc * p->JC);
    }
  }
  for (size_t i = 0; i < nops; ++i) {
    p->ops[i].dev.sched = reinterpret_cast<uint32_t*>(p->d_desc + p->off_sched + i * r3d_plan::kSchedStride);
    memcpy(h.data() + p->off_ops + i * sizeof(GemmOpDev), &p->ops[i].dev, sizeof(GemmOpDev));
  }
  for (auto& c : p->chain) {
    if (c.mo.nops == 0) continue;
    c.mo.m_groups_cap = m_groups_cap;
    c.mo.done = reinterpret_cast<uint32_t*>(p->d_desc + c.off_done);
    c.mo.sched = reinterpret_cast<uint32_t*>(p->d_desc + p->off_sched + (size_t)c.ops.front() * r3d_plan::kSchedStride);
    memcpy(h.data() + c.off_mo, &c.mo, sizeof(MultiOpDev));
  }
  memcpy(h.data() + p->off_pro, &pd, sizeof(pd));
  memcpy(h.data() + p->off_asm, &ad, sizeof(ad));
  if (prec != R3D_PREC_FP32) {
    for (size_t i = 0; i < nops; ++i) {
      const int rc = tc_build_tmaps(p->ops[i].dev, prec, (int64_t)cap * p->ops[i].dev.rows_per_seq,
                                    h.data() + p->off_tmaps + i * kMaxProb * kTmapsPerProb * kTmapBytes);
      if (rc != 0) return fail(R3D_ERR_CUDA, "cuTensorMapEncodeTiled failed for op %s (code %d)", p->ops[i].name.c_str(), rc);
    }
  }
  CUDA_TRY(cudaMemcpy(p->d_desc, h.data(), off, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaDeviceSynchronize());   // a pageable-memory copy may return before its DMA lands; the plan's streams do not wait for the legacy stream
  return R3D_OK;
}

static int ensure_capacity(r3d_plan* p, int batch) {
  if (batch <= p->cap) return R3D_OK;
  int cap = std::max(batch, 16);
  return bind_workspace(p, cap);
}

// ---- forward -------------------------------------------------------------------------------------
static constexpr int kMaxChunk = 8192;
static constexpr int kProfRing = 64;

static int64_t window_floats(const r3d_plan* p, int src_kind) {
  return src_kind == R3D_SRC_UV ? (int64_t)p->T * p->J * 2 : (int64_t)p->T * p->JC;
}
// floats / camera-row elements spanned by `n` consecutive windows (windows of a video overlap: stride < window length)
static int64_t src_span(const r3d_plan* p, const InputSpec& in, int64_t n) { return n <= 0 ? 0 : (n - 1) * in.src_stride + window_floats(p, in.src_kind); }
static int64_t cam_span(const r3d_plan* p, const InputSpec& in, int64_t n) {
  return (n <= 0 || in.cam == nullptr) ? 0 : (n - 1) * in.cam_stride + cam_row_elems(in.cam_kind, p->ext);
}
static InputSpec advance(const InputSpec& in, int64_t b0) {
  InputSpec o = in;
  o.src = in.src + b0 * in.src_stride;
  if (in.cam) o.cam = reinterpret_cast<const char*>(in.cam) + b0 * in.cam_stride * cam_elem_bytes(in.cam_kind);
  return o;
}

// `batch` windows are read; with tta the launch graph runs on 2*batch windows (direct + mirrored copies)
static int run_chunk(r3d_plan* p, const InputSpec& in, float* pos, float* trj, float* sum, int batch_in, cudaStream_t s, bool tta = false) {
  const int batch = tta ? 2 * batch_in : batch_in;
  const int prec = p->cfg.precision;
  // chained launches work on 256-row units (CTA pairs); small batches keep one launch per op (narrow tiles)
  // (the GlobalInfo chain: only while its per-layer launches would under-fill the GPU, i.e. up to 2048 windows)
  const int variant = tail_uses_pairs(batch) ? ((p->chain[0].ops.empty() ? 0 : 1) | ((p->chain[1].ops.empty() || batch > 2048) ? 0 : 2)) : 0;
  const std::vector<r3d_plan::Launch>& launches = p->launches[variant];
  const int nl = (int)launches.size() + 2, nev = 2 * nl;      // start/end event per launch
  cudaEvent_t* ev = nullptr;
  if (p->profiling) {
    p->prof_variant = variant;
    if (p->prof_ev.empty()) {
      p->prof_ev.resize((size_t)kProfRing * 2 * (p->ops.size() + 2));
      for (auto& e : p->prof_ev) CUDA_TRY(cudaEventCreate(&e));
    }
    ev = p->prof_ev.data() + (size_t)(p->prof_runs % kProfRing) * 2 * (p->ops.size() + 2);
    ++p->prof_runs;
    CUDA_TRY(cudaEventRecord(ev[0], s));
  }
  // work-unit counters of the dynamically scheduled launches (multi-wave 2-SM GEMMs: never below 64 windows) and the
  // chained launches' completion counters
  if (prec != R3D_PREC_FP32 && (variant || batch > 64))
    CUDA_TRY(cudaMemsetAsync(p->d_desc + p->off_sched, 0, p->zero_bytes, s));
  CUDA_TRY(launch_prologue(reinterpret_cast<const PrologueDev*>(p->d_desc + p->off_pro), p->pro, prec, in, batch, tta ? batch_in : batch, s));
  if (ev) CUDA_TRY(cudaEventRecord(ev[1], s));
  // fork: the GlobalInfo chain only depends on the input stage and runs on the side stream, filling the SMs the
  // (small-M) upper levels of the temporal tree leave idle; it joins before the first Integration GEMM.  (With the
  // chained tail launch that chain is part of the tail kernel's unit sequence instead: no side stream.)
  // per-launch timing serialises the launches on one stream: a side-stream launch's start/end events would also span the
  // time it spends waiting for SMs held by the main stream's kernels
  const bool use_side = p->use_side_stream && !p->profiling;
  bool forked = false, fork_recorded = false;
  if (use_side) {
    for (const auto& L : launches) fork_recorded |= p->ops[L.ops[0]].side;
    if (fork_recorded) CUDA_TRY(cudaEventRecord(p->ev_fork, s));   // right after the input stage, before any GEMM is enqueued
  }
  const GemmOpDev* d_ops = reinterpret_cast<const GemmOpDev*>(p->d_desc + p->off_ops);
  for (size_t k = 0; k < launches.size(); ++k) {
    const auto& L = launches[k];
    const size_t i = (size_t)L.ops[0];
    const OpHost& oh = p->ops[i];
    cudaStream_t st = s;
    if (use_side && oh.side) {
      if (!forked) {
        CUDA_TRY(cudaStreamWaitEvent(p->s_side, p->ev_fork, 0));
        forked = true;
      }
      st = p->s_side;
    }
    bool joins = false;
    for (int j : L.ops) joins |= p->ops[j].join_before;
    if (use_side && joins && forked) {
      CUDA_TRY(cudaEventRecord(p->ev_join, p->s_side));
      CUDA_TRY(cudaStreamWaitEvent(s, p->ev_join, 0));
      forked = false;
    }
    if (ev) CUDA_TRY(cudaEventRecord(ev[2 * (k + 1)], st));
    if (oh.side && exp_env("R3D_SKIP_SIDE")) {}                        // experiment (wrong results): what the GlobalInfo chain costs the step
    else if (L.chain >= 0)
      CUDA_TRY(launch_tail_tc(d_ops, p->d_desc + p->off_tmaps, reinterpret_cast<const MultiOpDev*>(p->d_desc + p->chain[L.chain].off_mo),
                              p->chain[L.chain].mo, batch, prec, p->chain[L.chain].max_clusters, st));
    else if (prec == R3D_PREC_FP32)
      CUDA_TRY(launch_gemm_ffma(d_ops + i, oh.dev, batch * oh.dev.rows_per_seq, st));
    else
      CUDA_TRY(launch_gemm_tc(d_ops + i, oh.dev, p->d_desc + p->off_tmaps + i * kMaxProb * kTmapsPerProb * kTmapBytes, batch * oh.dev.rows_per_seq, prec, st));
    if (ev) CUDA_TRY(cudaEventRecord(ev[2 * (k + 1) + 1], st));
  }
  if (forked) {
    CUDA_TRY(cudaEventRecord(p->ev_join, p->s_side));
    CUDA_TRY(cudaStreamWaitEvent(s, p->ev_join, 0));
  }
  if (ev) CUDA_TRY(cudaEventRecord(ev[nev - 2], s));
  CUDA_TRY(launch_assemble(reinterpret_cast<const AssembleDev*>(p->d_desc + p->off_asm), p->asmb, pos, trj, sum, batch_in, tta ? 1 : 0, s));
  if (ev) CUDA_TRY(cudaEventRecord(ev[nev - 1], s));
  return R3D_OK;
}

// r3d_input -> InputSpec, with the argument checks every entry point shares (mirrors the asserts at rie.py:285-287)
static int make_input(r3d_plan* p, const r3d_input* in, float* pos, float* trj, float* sum, int batch, InputSpec* out, bool* tta) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "forward: null plan");
  if (!in) return fail(R3D_ERR_BAD_ARG, "forward: null input descriptor");
  if (!p->uploaded) return fail(R3D_ERR_STATE, "forward before r3d_plan_upload");
  if (batch < 0) return fail(R3D_ERR_BAD_ARG, "batch=%d", batch);
  if (!in->src && batch != 0) return fail(R3D_ERR_BAD_ARG, "forward: null input");
  if ((pos || sum) && !p->has_pos) return fail(R3D_ERR_BAD_ARG, "plan has no pose net but pos/sum output requested");
  if (trj && !p->has_trj) return fail(R3D_ERR_BAD_ARG, "plan has no trajectory net but trj output requested");
  if (sum && !p->has_trj) return fail(R3D_ERR_BAD_ARG, "sum output needs both nets");
  if (in->window_stride <= 0) return fail(R3D_ERR_BAD_ARG, "window_stride=%lld must be positive", (long long)in->window_stride);
  if (in->cam_stride < 0) return fail(R3D_ERR_BAD_ARG, "cam_stride=%lld", (long long)in->cam_stride);
  if (in->flags & ~(R3D_IN_UNDISTORT | R3D_IN_FLIP_TTA)) return fail(R3D_ERR_BAD_ARG, "unknown input flags 0x%x", in->flags);
  InputSpec s{};
  s.src = in->src; s.src_stride = in->window_stride; s.src_kind = in->src_kind;
  s.cam = in->cam; s.cam_stride = in->cam_stride; s.cam_kind = in->cam_kind;
  s.undistort = (in->flags & R3D_IN_UNDISTORT) ? 1 : 0;
  if (in->src_kind == R3D_SRC_UV) {
    if (p->Cin != 3) return fail(R3D_ERR_UNSUPPORTED, "pixel-keypoint input needs in_features == 3 (ray encoding, utils.py:91-96)");
    if (p->embed && p->ext != 2) return fail(R3D_ERR_UNSUPPORTED, "pixel-keypoint input derives param=[height,pitch]: extrinsic_dim must be 2");
    if (in->cam_kind != R3D_CAM_F32 && in->cam_kind != R3D_CAM_F64) return fail(R3D_ERR_BAD_ARG, "pixel-keypoint input needs R3D_CAM_F32 or R3D_CAM_F64 camera rows");
    if (!in->cam && batch != 0) return fail(R3D_ERR_BAD_ARG, "cam is null");
    if (s.undistort && in->cam_kind != R3D_CAM_F64) return fail(R3D_ERR_BAD_ARG, "R3D_IN_UNDISTORT needs R3D_CAM_F64 rows (they carry the distortion coefficients)");
  } else if (in->src_kind == R3D_SRC_RAYS) {
    if (in->cam_kind != R3D_CAM_PARAM) return fail(R3D_ERR_BAD_ARG, "encoded input takes R3D_CAM_PARAM rows");
    if (s.undistort) return fail(R3D_ERR_BAD_ARG, "R3D_IN_UNDISTORT applies to pixel-keypoint input only");
    if (p->embed && !in->cam && batch != 0) return fail(R3D_ERR_BAD_ARG, "camera embedding enabled but param/cam pointer is null");
    if (!p->embed) s.cam = nullptr;
  } else {
    return fail(R3D_ERR_BAD_ARG, "src_kind=%d", in->src_kind);
  }
  *tta = (in->flags & R3D_IN_FLIP_TTA) != 0;
  if (*tta && p->flip_in.empty()) return fail(R3D_ERR_STATE, "flip augmentation requested before r3d_plan_set_flip");
  *out = s;
  return R3D_OK;
}

// Overlapping pixel-keypoint windows that share one camera row are the frames of ONE video (trainer.py:47-58, :323-324):
// every frame is ray-encoded once into the lane's scratch and the windows are indexed there by the input stage, instead
// of each window re-encoding its RF frames.  Rewrites `in` to the encoded form.  Caller holds the lane's mutex / device.
static int encode_video_once(r3d_plan* p, InputSpec& in, int batch, cudaStream_t s) {
  const int64_t frame = (int64_t)p->J * 2;
  if (in.src_kind != R3D_SRC_UV || in.cam_stride != 0 || batch < 2 || in.src_stride >= window_floats(p, R3D_SRC_UV) || in.src_stride % frame) return R3D_OK;
  const int64_t step = in.src_stride / frame, frames = (int64_t)(batch - 1) * step + p->T;
  const size_t need = 256 + (size_t)frames * p->JC * 4;
  if (p->vid_bytes < need) {
    if (p->d_vid) { CUDA_TRY(cudaDeviceSynchronize()); CUDA_TRY(cudaFree(p->d_vid)); p->d_vid = nullptr; p->vid_bytes = 0; }
    const size_t cap = need + need / 4;
    CUDA_TRY(cudaMalloc(&p->d_vid, cap));
    p->vid_bytes = cap;
  }
  float* prm = reinterpret_cast<float*>(p->d_vid);
  float* rays = reinterpret_cast<float*>(p->d_vid + 256);
  CUDA_TRY(launch_video_encode(in.src, rays, prm, frames * p->J, in.cam, in.cam_kind, in.undistort, s));
  in.src = rays; in.src_stride = step * p->JC; in.src_kind = R3D_SRC_RAYS;
  in.cam = p->embed ? prm : nullptr; in.cam_kind = R3D_CAM_PARAM; in.cam_stride = 0; in.undistort = 0;
  return R3D_OK;
}

// Small-batch path: replay the captured launch sequence.  Returns 1 when the graph route is not available (the caller
// then launches directly), R3D_OK / an error code otherwise.  Caller holds p->mu and has selected the device.
static int forward_graph(r3d_plan* p, const InputSpec& in, float* pos, float* trj, float* sum, int batch, cudaStream_t s) {
  const int mask = (in.cam ? 1 : 0) | (pos ? 2 : 0) | (trj ? 4 : 0) | (sum ? 8 : 0) | (in.src_kind << 4) | (in.cam_kind << 6) | (in.undistort << 8);
  int rc = ensure_capacity(p, batch);           // may rebind the workspace and drop every captured graph
  if (rc) return rc;
  r3d_plan::GraphEntry* g = nullptr;
  for (auto& e : p->graphs)
    if (e.batch == batch && e.mask == mask && e.src_stride == in.src_stride && e.prm_stride == in.cam_stride) g = &e;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  // windows may overlap (video: batch stride of one frame) and the parameter row may be shared (stride 0)
  const size_t in_b = (size_t)src_span(p, in, batch) * 4, prm_b = (size_t)cam_span(p, in, batch) * cam_elem_bytes(in.cam_kind);
  const size_t out_b = (size_t)batch * p->J * 3 * 4, trj_b = (size_t)batch * 3 * 4;
  if (g == nullptr) {
    if (p->graphs.size() >= 16) clear_graphs(p);
    r3d_plan::GraphEntry e;
    e.batch = batch; e.mask = mask; e.src_stride = in.src_stride; e.prm_stride = in.cam_stride;
    e.off_prm = al(in_b); e.off_pos = e.off_prm + al(prm_b); e.off_sum = e.off_pos + al(out_b); e.off_trj = e.off_sum + al(out_b);
    CUDA_TRY(cudaMalloc(&e.buf, e.off_trj + al(trj_b)));
    InputSpec gi = in;
    gi.src = reinterpret_cast<float*>(e.buf);
    gi.cam = in.cam ? e.buf + e.off_prm : nullptr;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(p->s_comp, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
      rc = run_chunk(p, gi, pos ? reinterpret_cast<float*>(e.buf + e.off_pos) : nullptr, trj ? reinterpret_cast<float*>(e.buf + e.off_trj) : nullptr,
                     sum ? reinterpret_cast<float*>(e.buf + e.off_sum) : nullptr, batch, p->s_comp);
      ce = cudaStreamEndCapture(p->s_comp, &graph);          // always ends the capture, also after a failed launch
      if (rc == R3D_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&e.exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
    }
    if (rc != R3D_OK || ce != cudaSuccess || e.exec == nullptr) {   // capture unsupported here: launch directly from now on
      cudaGetLastError();
      cudaFree(e.buf);
      p->graph_max_batch = 0;
      return 1;
    }
    p->graphs.push_back(e);
    g = &p->graphs.back();
  }
  CUDA_TRY(cudaMemcpyAsync(g->buf, in.src, in_b, cudaMemcpyDeviceToDevice, s));
  if (in.cam) CUDA_TRY(cudaMemcpyAsync(g->buf + g->off_prm, in.cam, prm_b, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaGraphLaunch(g->exec, s));
  ++p->graph_launches;
  if (pos) CUDA_TRY(cudaMemcpyAsync(pos, g->buf + g->off_pos, out_b, cudaMemcpyDeviceToDevice, s));
  if (sum) CUDA_TRY(cudaMemcpyAsync(sum, g->buf + g->off_sum, out_b, cudaMemcpyDeviceToDevice, s));
  if (trj) CUDA_TRY(cudaMemcpyAsync(trj, g->buf + g->off_trj, trj_b, cudaMemcpyDeviceToDevice, s));
  return R3D_OK;
}

// ---- tickets: one ring slot per asynchronous submission, one event per lane it ran on -------------------------------
static int ticket_slot_reuse(r3d_plan* p) {          // the ring slot's previous owner has completed
  const int idx = (int)(p->submit_seq % r3d_plan::kTicketRing);
  if (p->submit_seq >= (uint64_t)r3d_plan::kTicketRing)
    for (int l = 0; l < 2; ++l)
      if (p->ticket_lanes[idx] & (1 << l)) CUDA_TRY(cudaEventSynchronize(p->ev_ticket[idx][l]));
  return R3D_OK;
}
static int ticket_issue(r3d_plan* p, int lanes, uint64_t* ticket) {
  const int idx = (int)(p->submit_seq % r3d_plan::kTicketRing);
  for (int l = 0; l < 2; ++l)
    if (lanes & (1 << l)) CUDA_TRY(cudaEventRecord(p->ev_ticket[idx][l], l == 0 ? p->s_comp : p->twin->s_comp));
  p->ticket_lanes[idx] = (uint8_t)lanes;
  *ticket = p->submit_seq++;
  return R3D_OK;
}

// One forward on the lane `p` (workspace, descriptors, side stream, captured graphs and the video scratch are the
// lane's), enqueued on stream `s`.  A lane's resources are single-buffered, so whatever stream used them last -- another
// caller stream, the lane's own compute stream (r3d_submit_*, r3d_*_host) -- is waited for on the device first
// (ev_ws = "lane last used"); the host never blocks.  Caller holds the lane owner's mutex and has selected the device.
static int forward_on_lane(r3d_plan* p, InputSpec in, float* pos, float* trj, float* sum, int batch, cudaStream_t s, bool tta) {
  int rc = R3D_OK;
  CUDA_TRY(cudaStreamWaitEvent(s, p->ev_ws, 0));
  rc = encode_video_once(p, in, batch, s);
  bool done = false;
  if (rc == R3D_OK && !tta && !p->profiling && batch <= p->graph_max_batch) {
    rc = forward_graph(p, in, pos, trj, sum, batch, s);
    if (rc != 1) done = true; else rc = R3D_OK;
  }
  if (!done && rc == R3D_OK) {
    const int max_in = tta ? kMaxChunk / 2 : kMaxChunk;     // the mirrored copies double the rows in flight
    rc = ensure_capacity(p, std::min(batch, max_in) * (tta ? 2 : 1));
    for (int b0 = 0; rc == R3D_OK && b0 < batch; b0 += max_in) {
      const int nb = std::min(max_in, batch - b0);
      rc = run_chunk(p, advance(in, b0), pos ? pos + (int64_t)b0 * p->J * 3 : nullptr, trj ? trj + (int64_t)b0 * 3 : nullptr,
                     sum ? sum + (int64_t)b0 * p->J * 3 : nullptr, nb, s, tta);
    }
  }
  CUDA_TRY(cudaEventRecord(p->ev_ws, s));        // (also after a failed enqueue: whatever was launched still uses the lane)
  return rc;
}

extern "C" R3D_API int r3d_forward(r3d_plan* p, const r3d_input* in, float* pos, float* trj, float* sum, int32_t batch, void* stream) {
  InputSpec is;
  bool tta = false;
  int rc = make_input(p, in, pos, trj, sum, batch, &is, &tta);
  if (rc || batch == 0) return rc;
  std::lock_guard<std::mutex> lk(*p->mu);
  DeviceGuard dg(p->device);
  CUDA_TRY(dg.err);
  return forward_on_lane(p, is, pos, trj, sum, batch, (cudaStream_t)stream, tta);
}

// Asynchronous device-buffer submission: ordered after the work already enqueued on `stream`, executed on one of the
// plan's two lanes (alternating), completion observed through r3d_join / r3d_wait.  Inputs and outputs must stay valid
// until then.  Two submissions in flight keep the GPU busy across the under-filled tail launches of each batch.
extern "C" R3D_API int r3d_submit(r3d_plan* p, const r3d_input* in, float* pos, float* trj, float* sum, int32_t batch, void* stream,
                                  uint64_t* ticket) {
  if (!ticket) return fail(R3D_ERR_BAD_ARG, "null ticket");
  InputSpec is;
  bool tta = false;
  int rc = make_input(p, in, pos, trj, sum, batch, &is, &tta);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(*p->mu);
  DeviceGuard dg(p->device);
  CUDA_TRY(dg.err);
  rc = ticket_slot_reuse(p);
  r3d_plan* lp = p;
  if (rc == R3D_OK) rc = get_lane(p, p->use_lanes ? (int)(p->slot_seq++ & 1) : 0, &lp);
  if (rc == R3D_OK) {
    CUDA_TRY(cudaEventRecord(p->ev_sub, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamWaitEvent(lp->s_comp, p->ev_sub, 0));
    if (batch > 0) rc = forward_on_lane(lp, is, pos, trj, sum, batch, lp->s_comp, tta);
  }
  if (rc == R3D_OK) rc = ticket_issue(p, lp == p ? 1 : 2, ticket);
  return rc;
}

// Host-buffer forward: chunks of the batch flow H2D (copy stream) -> a lane's compute stream -> D2H, through four device
// staging slots so the PCIe transfer of one chunk overlaps the kernels of the others.
// ticket == nullptr: synchronous (results are in host memory on return).  Otherwise the copies and launches are only
// enqueued and *ticket identifies the submission for r3d_wait.
static int forward_host(r3d_plan* p, const r3d_input* in, float* pos, float* trj, float* sum, int batch, uint64_t* ticket) {
  InputSpec hs;
  bool tta = false;
  int rc = make_input(p, in, pos, trj, sum, batch, &hs, &tta);
  if (rc) return rc;
  if (batch == 0 && ticket == nullptr) return R3D_OK;
  std::lock_guard<std::mutex> lk(*p->mu);
  DeviceGuard dg(p->device);
  CUDA_TRY(dg.err);
  if (batch == 0) {   // empty submission: a ticket that completes with everything enqueued before it (on both lanes)
    rc = ticket_slot_reuse(p);
    if (rc == R3D_OK) rc = ticket_issue(p, p->twin ? 3 : 1, ticket);
    return rc;
  }
  // chunking trades PCIe/compute overlap against per-launch efficiency (small batches under-fill the GPU)
  // measured on B200 (T=243): 1024-sequence chunks keep the kernels efficient; smaller chunks lose more in
  // under-filled launches than they gain in copy/compute overlap (H2D of 1024 windows is 0.6 ms at 55 GB/s)
  // ... for streamed submissions.  A blocking call has nothing else to overlap with, so it splits into 512-window chunks
  // that alternate between the two lanes: the second chunk's copy runs under the first chunk's kernels (+16 %).
  // A video (overlapping windows) ships 136 B per frame: no copy to hide, whole launches win.
  const bool video = hs.src_stride < window_floats(p, hs.src_kind);
  int per = p->host_chunk > 0 ? p->host_chunk : ((ticket != nullptr || video) ? 1024 : 512);
  const int max_in = tta ? kMaxChunk / 2 : kMaxChunk;
  const int parts = std::max(1, batch / per);
  const int chunk = std::min(batch, std::max(64, std::min(max_in, (batch + parts - 1) / parts)));
  const size_t cam_eb = cam_elem_bytes(hs.cam_kind);
  const size_t in_b = (size_t)src_span(p, hs, chunk) * 4, prm_b = std::max<size_t>((size_t)cam_span(p, hs, chunk) * cam_eb, 8);
  const size_t out_b = (size_t)chunk * p->J * 3 * 4, trj_b = (size_t)chunk * 3 * 4;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t slot = al(in_b) + al(prm_b) + 2 * al(out_b) + al(trj_b);
  if (p->stage_bytes < r3d_plan::kSlots * slot) {
    if (p->d_stage) { CUDA_TRY(cudaDeviceSynchronize()); CUDA_TRY(cudaFree(p->d_stage)); p->d_stage = nullptr; }
    CUDA_TRY(cudaMalloc(&p->d_stage, r3d_plan::kSlots * slot));
    p->stage_bytes = r3d_plan::kSlots * slot;
  }
  if (ticket != nullptr) rc = ticket_slot_reuse(p);
  int lanes_used = 0;
  for (int b0 = 0; rc == R3D_OK && b0 < batch; b0 += chunk) {
    // four staging slots, lane = slot & 1: consecutive chunks / submissions alternate between the two lanes, so the
    // copy of one overlaps the kernels of the others AND the under-filled tail launches of one batch overlap the
    // large launches of the next
    const int nb = std::min(chunk, batch - b0), sl = (int)(p->slot_seq++ % r3d_plan::kSlots);
    r3d_plan* lp = p;
    rc = get_lane(p, p->use_lanes ? (sl & 1) : 0, &lp);
    if (rc) break;
    lanes_used |= lp == p ? 1 : 2;
    char* base = p->d_stage + (size_t)sl * (p->stage_bytes / r3d_plan::kSlots);   // fixed stride: submissions of other sizes may be in flight
    float* d_in = reinterpret_cast<float*>(base);
    char* d_prm = base + al(in_b);
    float* d_pos = reinterpret_cast<float*>(base + al(in_b) + al(prm_b));
    float* d_sum = reinterpret_cast<float*>(base + al(in_b) + al(prm_b) + al(out_b));
    float* d_trj = reinterpret_cast<float*>(base + al(in_b) + al(prm_b) + 2 * al(out_b));
    const InputSpec hc = advance(hs, b0);
    CUDA_TRY(cudaStreamWaitEvent(p->s_copy, p->ev_done[sl], 0));   // slot's previous results have left (no-op on first use)
    CUDA_TRY(cudaMemcpyAsync(d_in, hc.src, (size_t)src_span(p, hc, nb) * 4, cudaMemcpyHostToDevice, p->s_copy));
    if (hc.cam) CUDA_TRY(cudaMemcpyAsync(d_prm, hc.cam, (size_t)cam_span(p, hc, nb) * cam_eb, cudaMemcpyHostToDevice, p->s_copy));
    CUDA_TRY(cudaEventRecord(p->ev_in[sl], p->s_copy));
    cudaStream_t sc = lp->s_comp;
    CUDA_TRY(cudaStreamWaitEvent(sc, p->ev_in[sl], 0));
    InputSpec dc = hc;
    dc.src = d_in;
    dc.cam = hc.cam ? d_prm : nullptr;
    const int graph_cap = lp->graph_max_batch;       // staged chunks are launched directly (their buffers are static already)
    lp->graph_max_batch = 0;
    rc = forward_on_lane(lp, dc, pos ? d_pos : nullptr, trj ? d_trj : nullptr, sum ? d_sum : nullptr, nb, sc, tta);
    lp->graph_max_batch = graph_cap;
    if (rc) break;
    if (pos) CUDA_TRY(cudaMemcpyAsync(pos + (int64_t)b0 * p->J * 3, d_pos, (size_t)nb * p->J * 12, cudaMemcpyDeviceToHost, sc));
    if (sum) CUDA_TRY(cudaMemcpyAsync(sum + (int64_t)b0 * p->J * 3, d_sum, (size_t)nb * p->J * 12, cudaMemcpyDeviceToHost, sc));
    if (trj) CUDA_TRY(cudaMemcpyAsync(trj + (int64_t)b0 * 3, d_trj, (size_t)nb * 12, cudaMemcpyDeviceToHost, sc));
    CUDA_TRY(cudaEventRecord(p->ev_done[sl], sc));
  }
  if (rc == R3D_OK && ticket != nullptr) {
    rc = ticket_issue(p, lanes_used, ticket);
  } else if (rc == R3D_OK) {
    CUDA_TRY(cudaStreamSynchronize(p->s_comp));
    if (p->twin) CUDA_TRY(cudaStreamSynchronize(p->twin->s_comp));
    CUDA_TRY(cudaStreamSynchronize(p->s_copy));
  }
  return rc;
}

extern "C" R3D_API int r3d_forward_host(r3d_plan* p, const r3d_input* in, float* pos, float* trj, float* sum, int32_t batch) {
  return forward_host(p, in, pos, trj, sum, batch, nullptr);
}
extern "C" R3D_API int r3d_submit_host(r3d_plan* p, const r3d_input* in, float* pos, float* trj, float* sum, int32_t batch, uint64_t* ticket) {
  if (!ticket) return fail(R3D_ERR_BAD_ARG, "null ticket");
  return forward_host(p, in, pos, trj, sum, batch, ticket);
}

// ---- named forms of the generic calls (the reference interface each stands in for: include/ray3d_b200.h) -------------
static r3d_input in_rays(const r3d_plan* p, const float* x, const float* param, int64_t stride, int32_t flags = 0) {
  r3d_input in{};
  in.src = x; in.window_stride = stride; in.src_kind = R3D_SRC_RAYS;
  in.cam = param; in.cam_kind = R3D_CAM_PARAM; in.cam_stride = stride == (int64_t)p->T * p->JC ? p->ext : 0; in.flags = flags;
  return in;
}
static r3d_input in_uv(const r3d_plan* p, const float* uv, const void* cam, int cam_kind, int64_t stride, int32_t flags = 0) {
  r3d_input in{};
  in.src = uv; in.window_stride = stride; in.src_kind = R3D_SRC_UV;
  in.cam = cam; in.cam_kind = cam_kind;
  in.cam_stride = stride == (int64_t)p->T * p->J * 2 ? (cam_kind == R3D_CAM_F64 ? R3D_CAM64_STRIDE : 6) : 0; in.flags = flags;
  return in;
}
#define R3D_NEED_PLAN(p) do { if (!(p)) return fail(R3D_ERR_BAD_ARG, "null plan"); } while (0)

extern "C" R3D_API int r3d_forward_rays(r3d_plan* p, const float* x, const float* param, float* pos, float* trj, float* sum,
                                int32_t batch, void* stream) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_rays(p, x, param, (int64_t)p->T * p->JC);
  return r3d_forward(p, &in, pos, trj, sum, batch, stream);
}
extern "C" R3D_API int r3d_forward_uv(r3d_plan* p, const float* uv, const float* cam, float* pos, float* trj, float* sum,
                              int32_t batch, void* stream) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_uv(p, uv, cam, R3D_CAM_F32, (int64_t)p->T * p->J * 2);
  return r3d_forward(p, &in, pos, trj, sum, batch, stream);
}
extern "C" R3D_API int r3d_forward_uv_cam64(r3d_plan* p, const float* uv, const double* cam64, int32_t undistort, float* pos, float* trj,
                                            float* sum, int32_t batch, void* stream) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_uv(p, uv, cam64, R3D_CAM_F64, (int64_t)p->T * p->J * 2, undistort ? R3D_IN_UNDISTORT : 0);
  return r3d_forward(p, &in, pos, trj, sum, batch, stream);
}
extern "C" R3D_API int r3d_submit_rays(r3d_plan* p, const float* x, const float* param, float* pos, float* trj, float* sum,
                                       int32_t batch, void* stream, uint64_t* ticket) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_rays(p, x, param, (int64_t)p->T * p->JC);
  return r3d_submit(p, &in, pos, trj, sum, batch, stream, ticket);
}
extern "C" R3D_API int r3d_submit_uv(r3d_plan* p, const float* uv, const float* cam, float* pos, float* trj, float* sum,
                                     int32_t batch, void* stream, uint64_t* ticket) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_uv(p, uv, cam, R3D_CAM_F32, (int64_t)p->T * p->J * 2);
  return r3d_submit(p, &in, pos, trj, sum, batch, stream, ticket);
}
extern "C" R3D_API int r3d_forward_video(r3d_plan* p, const float* seq, const float* param, float* pos, float* trj, float* sum,
                                 int32_t frames_out, void* stream) {
  R3D_NEED_PLAN(p);
  // window f = frames [f, f+RF) of the padded video: batch stride of one frame, shared param row
  const r3d_input in = in_rays(p, seq, param, (int64_t)p->JC);
  return r3d_forward(p, &in, pos, trj, sum, frames_out, stream);
}
extern "C" R3D_API int r3d_forward_video_uv(r3d_plan* p, const float* uv_seq, const double* cam64_row, int32_t flags, float* pos, float* trj,
                                            float* sum, int32_t frames_out, void* stream) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_uv(p, uv_seq, cam64_row, R3D_CAM_F64, (int64_t)p->J * 2, flags);
  return r3d_forward(p, &in, pos, trj, sum, frames_out, stream);
}
extern "C" R3D_API int r3d_forward_video_uv_host(r3d_plan* p, const float* uv_seq, const double* cam64_row, int32_t flags, float* pos, float* trj,
                                                 float* sum, int32_t frames_out) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_uv(p, uv_seq, cam64_row, R3D_CAM_F64, (int64_t)p->J * 2, flags);
  return forward_host(p, &in, pos, trj, sum, frames_out, nullptr);
}
extern "C" R3D_API int r3d_submit_video_uv_host(r3d_plan* p, const float* uv_seq, const double* cam64_row, int32_t flags, float* pos, float* trj,
                                                float* sum, int32_t frames_out, uint64_t* ticket) {
  R3D_NEED_PLAN(p);
  if (!ticket) return fail(R3D_ERR_BAD_ARG, "null ticket");
  const r3d_input in = in_uv(p, uv_seq, cam64_row, R3D_CAM_F64, (int64_t)p->J * 2, flags);
  return forward_host(p, &in, pos, trj, sum, frames_out, ticket);
}

extern "C" R3D_API int r3d_plan_set_flip(r3d_plan* p, const int32_t* in_perm, const int32_t* out_perm) {
  if (!p || !in_perm || !out_perm) return fail(R3D_ERR_BAD_ARG, "r3d_plan_set_flip: null argument");
  std::vector<int> a(in_perm, in_perm + p->J), b(out_perm, out_perm + p->J);
  for (int j = 0; j < p->J; ++j)
    if (a[j] < 0 || a[j] >= p->J || b[j] < 0 || b[j] >= p->J) return fail(R3D_ERR_BAD_ARG, "flip permutation entry out of range at joint %d", j);
  std::lock_guard<std::mutex> lk(*p->mu);
  p->flip_in = a;
  p->flip_out = b;
  if (p->cap > 0 || p->twin) {            // descriptors already on the device: rebuild them with the new tables
    DeviceGuard dg(p->device);
    CUDA_TRY(dg.err);
    if (p->twin) {               // the second lane is re-cloned (with the new tables) on its next use
      cudaDeviceSynchronize();
      free_device(p->twin);
      delete p->twin;
      p->twin = nullptr;
    }
    if (p->cap > 0) return bind_workspace(p, p->cap);
  }
  return R3D_OK;
}

extern "C" R3D_API int r3d_forward_rays_tta(r3d_plan* p, const float* x, const float* param, float* pos, float* trj, float* sum,
                                    int32_t batch, void* stream) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_rays(p, x, param, (int64_t)p->T * p->JC, R3D_IN_FLIP_TTA);
  return r3d_forward(p, &in, pos, trj, sum, batch, stream);
}
extern "C" R3D_API int r3d_forward_video_tta(r3d_plan* p, const float* seq, const float* param, float* pos, float* trj, float* sum,
                                     int32_t frames_out, void* stream) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_rays(p, seq, param, (int64_t)p->JC, R3D_IN_FLIP_TTA);
  return r3d_forward(p, &in, pos, trj, sum, frames_out, stream);
}

extern "C" R3D_API int r3d_submit_rays_host(r3d_plan* p, const float* x, const float* param, float* pos, float* trj, float* sum, int32_t batch,
                                            uint64_t* ticket) {
  R3D_NEED_PLAN(p);
  if (!ticket) return fail(R3D_ERR_BAD_ARG, "null ticket");
  const r3d_input in = in_rays(p, x, param, (int64_t)p->T * p->JC);
  return forward_host(p, &in, pos, trj, sum, batch, ticket);
}
extern "C" R3D_API int r3d_submit_uv_host(r3d_plan* p, const float* uv, const float* cam, float* pos, float* trj, float* sum, int32_t batch,
                                          uint64_t* ticket) {
  R3D_NEED_PLAN(p);
  if (!ticket) return fail(R3D_ERR_BAD_ARG, "null ticket");
  const r3d_input in = in_uv(p, uv, cam, R3D_CAM_F32, (int64_t)p->T * p->J * 2);
  return forward_host(p, &in, pos, trj, sum, batch, ticket);
}
extern "C" R3D_API int r3d_forward_rays_host(r3d_plan* p, const float* x, const float* param, float* pos, float* trj, float* sum, int32_t batch) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_rays(p, x, param, (int64_t)p->T * p->JC);
  return forward_host(p, &in, pos, trj, sum, batch, nullptr);
}
extern "C" R3D_API int r3d_forward_uv_host(r3d_plan* p, const float* uv, const float* cam, float* pos, float* trj, float* sum, int32_t batch) {
  R3D_NEED_PLAN(p);
  const r3d_input in = in_uv(p, uv, cam, R3D_CAM_F32, (int64_t)p->T * p->J * 2);
  return forward_host(p, &in, pos, trj, sum, batch, nullptr);
}

extern "C" R3D_API int r3d_wait(r3d_plan* p, uint64_t ticket) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  cudaEvent_t ev[2] = {nullptr, nullptr};
  {
    std::lock_guard<std::mutex> lk(*p->mu);
    if (ticket >= p->submit_seq) return fail(R3D_ERR_BAD_ARG, "r3d_wait: unknown ticket");
    if (ticket + r3d_plan::kTicketRing < p->submit_seq) return R3D_OK;   // recycled: its successor in the ring was submitted after it completed
    const int idx = (int)(ticket % r3d_plan::kTicketRing);
    for (int l = 0; l < 2; ++l)
      if (p->ticket_lanes[idx] & (1 << l)) ev[l] = p->ev_ticket[idx][l];
  }
  for (int l = 0; l < 2; ++l)
    if (ev[l]) CUDA_TRY(cudaEventSynchronize(ev[l]));
  return R3D_OK;
}

// Stream-ordered completion: `stream` waits (on the device) for the submission; the host does not block.
extern "C" R3D_API int r3d_join(r3d_plan* p, uint64_t ticket, void* stream) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  std::lock_guard<std::mutex> lk(*p->mu);
  if (ticket >= p->submit_seq) return fail(R3D_ERR_BAD_ARG, "r3d_join: unknown ticket");
  if (ticket + r3d_plan::kTicketRing < p->submit_seq) return R3D_OK;     // completed before its ring slot was reused
  const int idx = (int)(ticket % r3d_plan::kTicketRing);
  for (int l = 0; l < 2; ++l)
    if (p->ticket_lanes[idx] & (1 << l)) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, p->ev_ticket[idx][l], 0));
  return R3D_OK;
}

// Result-neutral tuning options (everything else is fixed at build time; experiment switches need -DR3D_EXPERIMENTS).
extern "C" R3D_API int r3d_plan_set_option(r3d_plan* p, const char* name, int32_t value) {
  if (!p || !name) return fail(R3D_ERR_BAD_ARG, "r3d_plan_set_option: null argument");
  std::lock_guard<std::mutex> lk(*p->mu);
  const std::string k(name);
  auto both = [&](auto&& f) { f(p); if (p->twin) f(p->twin); };
  if (k == "graph_max_batch") p->graph_max_batch = std::max(0, (int)value);          // lane 0 only: submissions launch directly
  else if (k == "lanes") p->use_lanes = value >= 2;
  else if (k == "side_stream") both([&](r3d_plan* q) { q->use_side_stream = value != 0; });
  else if (k == "host_chunk") p->host_chunk = std::max(0, (int)value);
  else if (k == "tail_fusion" || k == "tail_width" || k == "tail_clusters" || k == "side_chain" || k == "side_clusters" || k == "tile_policy") {
    if (p->uploaded || p->finalized) return fail(R3D_ERR_STATE, "%s must be set before r3d_plan_finalize", name);
    if (k == "tail_fusion") p->tail_fusion = value != 0;
    else if (k == "tail_width") p->tail_width = value >= 256 ? 256 : 128;
    else if (k == "tail_clusters") p->tail_clusters = std::max(0, (int)value);
    else if (k == "side_chain") { p->side_chain = value != 0; p->side_chain_force = value >= 2; }
    else if (k == "tile_policy") p->tile_policy = std::min(2, std::max(0, (int)value));
    else p->side_clusters = std::max(0, (int)value);
  }
  else return fail(R3D_ERR_BAD_ARG, "unknown option '%s' (graph_max_batch, lanes, side_stream, host_chunk, tail_fusion, tail_width, side_chain, side_clusters, tile_policy)", name);
  return R3D_OK;
}

extern "C" R3D_API int r3d_plan_set_profiling(r3d_plan* p, int enable) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  std::lock_guard<std::mutex> lk(*p->mu);
  p->profiling = enable != 0;
  p->prof_runs = 0;
  return R3D_OK;
}

extern "C" R3D_API int r3d_plan_launch_times(r3d_plan* p, float* ms_out, int32_t cap, int32_t* n_launches, int32_t* n_runs) {
  if (!p) return fail(R3D_ERR_BAD_ARG, "null plan");
  std::lock_guard<std::mutex> lk(*p->mu);
  const int nl = (int)p->launches[p->prof_variant].size() + 2, nev = 2 * nl;
  if (n_launches) *n_launches = nl;
  const int runs = std::min(p->prof_runs, kProfRing);
  if (n_runs) *n_runs = runs;
  if (!ms_out || runs == 0) return R3D_OK;
  std::vector<double> acc(nl, 0.0);
  for (int r = 0; r < runs; ++r) {
    cudaEvent_t* ev = p->prof_ev.data() + (size_t)r * 2 * (p->ops.size() + 2);
    CUDA_TRY(cudaEventSynchronize(ev[nev - 1]));
    for (int i = 0; i < nl; ++i) {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
      acc[i] += ms;
    }
  }
  for (int i = 0; i < nl && i < cap; ++i) ms_out[i] = (float)(acc[i] / runs);
  return R3D_OK;
}

extern "C" R3D_API const char* r3d_plan_launch_name(const r3d_plan* p, int32_t i) {
  if (!p) return "";
  const auto& launches = p->launches[p->prof_variant];
  const int nl = (int)launches.size() + 2;
  if (i == 0) return "input_stage";
  if (i == nl - 1) return "output_stage";
  if (i > 0 && i < nl - 1) return launches[i - 1].name.c_str();
  return "";
}

extern "C" R3D_API int r3d_ray_encode_f64(const double* uv, double* ray, int64_t n, double fx, double fy, double ppx, double ppy,
                                  double c, double s, void* stream) {
  if (!uv || !ray || n < 0) return fail(R3D_ERR_BAD_ARG, "r3d_ray_encode_f64: bad argument");
  CUDA_TRY(launch_ray_encode_f64(uv, ray, n, fx, fy, ppx, ppy, c, s, (cudaStream_t)stream));
  return R3D_OK;
}

extern "C" R3D_API int r3d_undistort_points_f64(const double* uv, double* out, int64_t n, double fx, double fy, double cx, double cy,
                                        const double* dist5_host, void* stream) {
  if (!uv || !out || !dist5_host || n < 0) return fail(R3D_ERR_BAD_ARG, "r3d_undistort_points_f64: bad argument");
  CUDA_TRY(launch_undistort_points_f64(uv, out, n, fx, fy, cx, cy, dist5_host, (cudaStream_t)stream));
  return R3D_OK;
}

extern "C" R3D_API int r3d_normalize_screen_f64(const double* xy, double* out, int64_t n, double w, double h, void* stream) {
  if (!xy || !out || n < 0 || w == 0.0) return fail(R3D_ERR_BAD_ARG, "r3d_normalize_screen_f64: bad argument");
  CUDA_TRY(launch_normalize_screen_f64(xy, out, n, w, h, (cudaStream_t)stream));
  return R3D_OK;
}

extern "C" R3D_API int r3d_eval_metrics(const float* pred, const float* target, int32_t frames, int32_t joints, const double* rn2w_tn2w,
                                double* sums_dev, void* stream) {
  if (!pred || !target || !sums_dev || frames < 0 || joints < 1 || joints > 32) return fail(R3D_ERR_BAD_ARG, "r3d_eval_metrics: bad argument");
  CUDA_TRY(launch_eval_metrics(pred, target, frames, joints, rn2w_tn2w, sums_dev, (cudaStream_t)stream));
  return R3D_OK;
}

// ---- on-device self test: tensor-core GEMM vs FP32 FFMA GEMM -----------------------------------------
#ifdef R3D_EXPERIMENTS
// Experiment builds: cycle accounting of the chained tail launch (see g_tail_stats in r3d_tail_tc.cu); resets the counters.
extern "C" R3D_API int r3d_debug_tail_stats(uint64_t* out8) {
  CUDA_TRY(cudaDeviceSynchronize());
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "stat word");
  CUDA_TRY(tail_stats_read(reinterpret_cast<unsigned long long*>(out8), 1));
  return R3D_OK;
}
#endif

#ifdef R3D_TC_TRACE
// Diagnostics of trace builds only (R3D_BUILD_TRACE=1 python -m ray3d_b200.build --force; scripts/tile_trace.py): out == NULL
// arms a per-tile SM-clock trace for the tensor-core GEMM launch `arm_after_launches` launches from now; out != NULL
// synchronises the device and copies the last trace ([4 roles][64 tiles][8 events] clock64 stamps of CTA 0).
extern "C" R3D_API int r3d_debug_tc_trace(int32_t arm_after_launches, int64_t* out, int32_t cap) {
  if (out == nullptr) { tc_trace_arm(arm_after_launches); return R3D_OK; }
  static_assert(sizeof(long long) == sizeof(int64_t), "trace word");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(tc_trace_read(reinterpret_cast<long long*>(out), cap));
  return R3D_OK;
}
#endif

extern "C" R3D_API int r3d_selftest_gemm(int32_t m, int32_t n, int32_t k, int32_t nprob, int32_t precision, int32_t device,
                                 double* rel_err, double* ms_tc, double* ms_ffma) {
  if (m <= 0 || n <= 0 || k <= 0 || k % kKAlign || n % 16 || nprob < 1 || nprob > kMaxProb)
    return fail(R3D_ERR_BAD_ARG, "selftest: need k%%64==0, n%%16==0, 1<=nprob<=6");
  if (precision != R3D_PREC_BF16X3 && precision != R3D_PREC_BF16) return fail(R3D_ERR_BAD_ARG, "selftest precision");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(tc_configure());
  const size_t na = (size_t)m * k, nw = (size_t)n * k, nc = (size_t)m * n;
  std::vector<float> hA(na * nprob), hW(nw * nprob), hB((size_t)n * nprob), hR(nc * nprob);
  uint32_t st = 12345u + m * 7 + n * 3 + k;
  auto rnd = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xffff) / 32768.0f - 1.0f; };
  for (auto& v : hA) v = rnd();
  for (auto& v : hW) v = rnd() * 0.05f;
  for (auto& v : hB) v = rnd() * 0.1f;
  for (auto& v : hR) v = rnd();
  // device buffers: fp32 set for FFMA, bf16 planes for the tensor path
  float *dA, *dW, *dB, *dR, *dC0, *dC1;
  uint16_t *dAh, *dAl, *dWh, *dWl, *dRh, *dRl;
  CUDA_TRY(cudaMalloc(&dA, na * nprob * 4)); CUDA_TRY(cudaMalloc(&dW, nw * nprob * 4)); CUDA_TRY(cudaMalloc(&dB, n * nprob * 4));
  CUDA_TRY(cudaMalloc(&dR, nc * nprob * 4)); CUDA_TRY(cudaMalloc(&dC0, nc * nprob * 4)); CUDA_TRY(cudaMalloc(&dC1, nc * nprob * 4));
  CUDA_TRY(cudaMalloc(&dAh, na * nprob * 2)); CUDA_TRY(cudaMalloc(&dAl, na * nprob * 2)); CUDA_TRY(cudaMalloc(&dWh, nw * nprob * 2));
  CUDA_TRY(cudaMalloc(&dWl, nw * nprob * 2)); CUDA_TRY(cudaMalloc(&dRh, nc * nprob * 2)); CUDA_TRY(cudaMalloc(&dRl, nc * nprob * 2));
  auto split = [&](const std::vector<float>& src, std::vector<float>& rounded, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo) {
    hi.resize(src.size()); lo.resize(src.size()); rounded.resize(src.size());
    for (size_t i = 0; i < src.size(); ++i) {
      hi[i] = f2bf(src[i]);
      lo[i] = precision == R3D_PREC_BF16X3 ? f2bf(src[i] - bf2f(hi[i])) : 0;
      rounded[i] = bf2f(hi[i]) + bf2f(lo[i]);   // what the tensor path actually multiplies
    }
  };
  std::vector<float> rA, rW, rR;
  std::vector<uint16_t> Ah, Al, Wh, Wl, Rh, Rl;
  split(hA, rA, Ah, Al); split(hW, rW, Wh, Wl); split(hR, rR, Rh, Rl);
  // FFMA reference multiplies the *same rounded operands* for BF16 (single product); for BF16X3 it uses the
  // original fp32 values, so rel_err includes the dropped lo*lo term (expected ~1e-5).
  const std::vector<float>& fa = precision == R3D_PREC_BF16 ? rA : hA;
  const std::vector<float>& fw = precision == R3D_PREC_BF16 ? rW : hW;
  CUDA_TRY(cudaMemcpy(dA, fa.data(), na * nprob * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dW, fw.data(), nw * nprob * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dB, hB.data(), n * nprob * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dR, rR.data(), nc * nprob * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dAh, Ah.data(), na * nprob * 2, cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(dAl, Al.data(), na * nprob * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dWh, Wh.data(), nw * nprob * 2, cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(dWl, Wl.data(), nw * nprob * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dRh, Rh.data(), nc * nprob * 2, cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(dRl, Rl.data(), nc * nprob * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(dC0, 0, nc * nprob * 4)); CUDA_TRY(cudaMemset(dC1, 0, nc * nprob * 4));
  // the tensor path writes its result the way the network does: bf16 hi (+lo) planes when n is a multiple of 32
  const bool bf_dst = n % 32 == 0;
  uint16_t *dOh = nullptr, *dOl = nullptr;
  if (bf_dst) {
    CUDA_TRY(cudaMalloc(&dOh, nc * nprob * 4));       // hi and lo planes in one allocation, like the plan's workspace
    dOl = dOh + nc * nprob;
    CUDA_TRY(cudaMemset(dOh, 0, nc * nprob * 4));
  }
  GemmOpDev f{}, t{};
  f.nprob = t.nprob = nprob; f.rows_per_seq = t.rows_per_seq = 1; f.slope = t.slope = 0.2f; f.n_tile = t.n_tile = pick_n_tile(n);
  for (int q = 0; q < nprob; ++q) {
    GemmProb& a = f.prob[q];
    a.a = Mat{dA + q * na, nullptr, k, 0}; a.w0 = dW + q * nw; a.bias = dB + (size_t)q * n;
    a.res = Mat{dR + q * nc, nullptr, n, 0}; a.res_col = 0; a.K = k; a.N = n; a.n_pad = n; a.ndst = 1;
    a.dst[0] = Dst{Mat{dC0 + q * nc, nullptr, n, 0}, 0, 1};
    GemmProb& b = t.prob[q];
    b = a;
    b.a = Mat{dAh + q * na, precision == R3D_PREC_BF16X3 ? dAl + q * na : nullptr, k, 0};
    b.w0 = dWh + q * nw; b.w1 = precision == R3D_PREC_BF16X3 ? dWl + q * nw : nullptr;
    b.res = Mat{dRh + q * nc, precision == R3D_PREC_BF16X3 ? dRl + q * nc : nullptr, n, 0};
    b.dst[0] = Dst{Mat{dC1 + q * nc, nullptr, n, 0}, 0, 1};
    if (bf_dst) b.dst[0] = Dst{Mat{dOh + q * nc, precision == R3D_PREC_BF16X3 ? dOl + q * nc : nullptr, n, 0}, 0, 0};
  }
  GemmOpDev *dF, *dT;
  uint32_t* dSched;
  CUDA_TRY(cudaMalloc(&dSched, 128));
  t.sched = dSched;
  void* dMaps;
  std::vector<char> maps((size_t)kMaxProb * kTmapsPerProb * kTmapBytes, 0);
  if (tc_build_tmaps(t, precision, m, maps.data()) != 0) return fail(R3D_ERR_CUDA, "selftest: tensor map encode failed");
  CUDA_TRY(cudaMalloc(&dF, sizeof(f))); CUDA_TRY(cudaMalloc(&dT, sizeof(t))); CUDA_TRY(cudaMalloc(&dMaps, maps.size()));
  CUDA_TRY(cudaMemcpy(dF, &f, sizeof(f), cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(dT, &t, sizeof(t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dMaps, maps.data(), maps.size(), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, 0));
    CUDA_TRY(launch_gemm_ffma(dF, f, m, 0));
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  }
  if (ms_ffma) *ms_ffma = ms;
  for (int rep = 0; rep < 2; ++rep) {
    CUDA_TRY(cudaMemsetAsync(dSched, 0, 128, 0));
    CUDA_TRY(cudaEventRecord(e0, 0));
    CUDA_TRY(launch_gemm_tc(dT, t, dMaps, m, precision, 0));
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  }
  if (ms_tc) *ms_tc = ms;
  std::vector<float> c0(nc * nprob), c1(nc * nprob);
  CUDA_TRY(cudaMemcpy(c0.data(), dC0, nc * nprob * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(c1.data(), dC1, nc * nprob * 4, cudaMemcpyDeviceToHost));
  if (bf_dst) {
    std::vector<uint16_t> oh(nc * nprob), ol(nc * nprob);
    CUDA_TRY(cudaMemcpy(oh.data(), dOh, nc * nprob * 2, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(ol.data(), dOl, nc * nprob * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < c1.size(); ++i) c1[i] = bf2f(oh[i]) + (precision == R3D_PREC_BF16X3 ? bf2f(ol[i]) : 0.f);
  }
  double mx = 0, md = 0;
  for (size_t i = 0; i < c0.size(); ++i) {
    mx = std::max(mx, (double)std::fabs(c0[i]));
    const double d = std::fabs((double)c0[i] - (double)c1[i]);
    md = (d != d) ? 1e30 : std::max(md, d);
  }
  if (rel_err) *rel_err = md / std::max(mx, 1e-30);
  for (void* q : {(void*)dA, (void*)dW, (void*)dB, (void*)dR, (void*)dC0, (void*)dC1, (void*)dAh, (void*)dAl, (void*)dWh, (void*)dWl,
                  (void*)dRh, (void*)dRl, (void*)dF, (void*)dT, dMaps, (void*)dOh, (void*)dSched})
    cudaFree(q);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return R3D_OK;
}
