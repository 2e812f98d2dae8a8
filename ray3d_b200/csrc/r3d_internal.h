// Internal types shared by the host plan builder and the sm_100a kernels.
// Vocabulary: a "problem" is one independent GEMM of a grouped launch (one joint group's
// temporal block, or one FC head); a "level" is one stride-w stage of the temporal tree.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "ray3d_b200.h"

namespace r3d {

constexpr int kMaxProb = 6;    // 5 joint groups (pose net) + 1 trajectory net
constexpr int kMaxDst = 6;     // an epilogue may scatter one tile into up to 6 feature buffers
constexpr int kKAlign = 64;    // every GEMM K (and every activation row pitch) is a multiple of 64

// Activation matrix in HBM, row-major [rows][ld].
//   FP32 precision  : p0 = float*
//   BF16X3 precision: p0 = bf16 "hi" plane, p1 = bf16 "lo" plane (value = hi + lo)
//   BF16 precision  : p0 = bf16 plane
struct Mat {
  void* p0;
  void* p1;
  int32_t ld;
  int32_t _pad;
};

struct Dst {
  Mat m;
  int32_t col;    // first column written inside m
  int32_t f32;    // 1: p0 is float* regardless of the plan precision (network outputs)
};

struct GemmProb {
  Mat a;                 // [M][K] (row pitch a.ld >= K)
  const void* w0;        // packed weights [n_pad][K] K-major: float (FP32) or bf16 hi
  const void* w1;        // bf16 lo plane (BF16X3) or null
  const float* bias;     // [n_pad] BN-folded bias
  Mat res;               // optional residual, same row index; null p0 => none
  int32_t res_col;
  int32_t K;             // multiple of 64
  int32_t N;             // valid output columns
  int32_t n_pad;         // rows of the packed weight matrix (multiple of the op's n tile)
  int32_t ndst;
  Dst dst[kMaxDst];
  // fused second GEMM (op.fused2): out = res + lrelu( lrelu(A*W^T + bias) * W2^T + bias2 ), the k=w conv followed by
  // the 1x1 conv of one temporal level (rie.py:96-97); the intermediate never leaves tensor memory
  const void* w2_0;      // [256][K2] K-major bf16 hi
  const void* w2_1;      // bf16 lo plane or null
  const float* bias2;
  int32_t K2;
  int32_t _pad2;
  uint64_t kmask;        // bit s set: the 16-wide K step s of the first GEMM holds non-zero weights and is multiplied;
                         // clear steps (zero padding, joints outside the problem's group in the shared first-layer
                         // operand) are neither loaded nor issued.  0 => every step (K > 1024 or elision disabled)
};

struct GemmOpDev {
  int32_t nprob;
  int32_t rows_per_seq;  // M = batch * rows_per_seq
  float slope;           // LeakyReLU slope; 1.0f = linear
  int32_t n_tile;
  int32_t reverse;       // walk the tiles last-to-first (alternates per layer: the tail of the previous layer's
                         // output is what is still L2 resident when this one starts)
  int32_t fused2;        // 1: every problem carries a second weight matrix (see GemmProb::w2_0)
  int32_t flags;         // experiments: bit 0 fused pair waits for the whole intermediate before the second GEMM;
                         // bit 1 release (instead of relaxed) remote barrier arrivals
  int32_t n_tile_tail;   // unit width of this op inside the chained tail launch (0: not part of it)
  uint32_t* sched;       // work-unit counter of this launch (zeroed before every forward): dynamically scheduled launches claim units with atomicAdd
  GemmProb prob[kMaxProb];
};

// ---- chained tail launch (r3d_tail_tc.cu) -----------------------------------------------------------------
// Every GEMM with ONE row per window -- the top of the temporal tree, shrink, GlobalInfo, FuseBlocks, Integration
// (rie.py:99-105, 159-169, 362-414) -- runs in a single persistent kernel: its work units (128-column tiles of every op,
// in dependency order) are claimed dynamically, and a unit starts as soon as the units that produce its input rows
// have landed (per (op, problem, row group) completion counters), instead of one under-filled launch per layer.
constexpr int kMaxTailOps = 24;
constexpr int kMaxDeps = 4;
constexpr int kTailN = 128;      // tile width of every unit
struct MultiOpDev {
  int32_t nops, m_groups_cap;
  uint32_t* done;                                  // [nops * kMaxProb][m_groups_cap] arrivals, zeroed before every forward
  uint32_t* sched;                                 // the launch's work-unit counter (zeroed with it)
  int32_t unit0[kMaxTailOps + 1];                  // units PER ROW GROUP of the ops before op i; op i's first unit = unit0[i] * row groups
  int32_t op_index[kMaxTailOps];                   // index into the plan's op / tensor-map arrays
  uint8_t nprob[kMaxTailOps];
  uint8_t width[kMaxTailOps];                      // unit width of the op in 128-column steps (1 or 2)
  uint8_t ntiles[kMaxTailOps][kMaxProb];           // column tiles (units) per problem
  uint8_t ndep[kMaxTailOps][kMaxProb];
  int16_t dep[kMaxTailOps][kMaxProb][kMaxDeps];    // counter rows (local op * kMaxProb + problem) a unit of (op, problem) waits for;
                                                   // a row is complete at ntiles x 2 store threads x CTAs-per-tile arrivals
};

// ---- input stage --------------------------------------------------------------------------------
struct EmbedDev {
  const float* w1;   // [mid][ext]  BN folded
  const float* b1;   // [mid]
  const float* w2;   // [emb][mid]  BN folded
  const float* b2;   // [emb]
  int32_t ndst;
  int32_t _pad;
  Dst dst[kMaxDst];
};

struct PrologueDev {
  int32_t T, J, Cin, JC, tc, w0, L0;
  int32_t k_pad;         // row pitch of a0
  Mat a0;                // first-layer operand shared by all problems: [B*L0][k_pad], columns per a0_map
  const int32_t* a0_off; // [k_pad] decoded source of every operand column (r3d_plan.cpp:build_a0_layout): offset into the
                         // window staged in shared memory, | 1 << 30 when relative to the row's first frame; padding
                         // columns point at a zero word
  Mat inc;               // in_current, [B][roundup(J*Cin,64)]
  int32_t n_embed, ext_dim, emb_mid, emb_dim;
  EmbedDev embed[2];
  int8_t flip_perm[32];   // flip test-time augmentation: source joint of every input joint (L/R swap)
};

struct AssembleDev {
  const float* heads[kMaxProb];   // [B][16] fp32 per problem; index 5 = trajectory head
  int32_t head_ld;
  int32_t J;
  int32_t has_pos, has_trj;
  int16_t slot_prob[32];          // output joint slot -> problem index
  int16_t slot_joint[32];         // output joint slot -> joint index inside that head
  int16_t flip_slot[32];          // flip augmentation: output slot whose flipped prediction lands in this slot
};

// ---- kernel launchers (defined in the .cu files) ----------------------------------------------
// What one forward reads: windows of encoded rays or of pixel keypoints, plus one camera/param row per window (stride 0:
// one row shared by all windows).  Pointers are device pointers by the time a kernel sees them.
struct InputSpec {
  const float* src;
  int64_t src_stride;     // floats between consecutive windows: T*J*C for materialised windows, J*C for a video
  int32_t src_kind;       // R3D_SRC_RAYS / R3D_SRC_UV
  int32_t cam_kind;       // R3D_CAM_PARAM / R3D_CAM_F32 / R3D_CAM_F64
  const void* cam;
  int64_t cam_stride;     // elements (float or double) between rows
  int32_t undistort;      // host-side knowledge: some R3D_CAM_F64 row has its undistort flag set
  int32_t _pad;
};
inline int64_t cam_row_elems(int cam_kind, int ext_dim) { return cam_kind == R3D_CAM_F64 ? R3D_CAM64_STRIDE : cam_kind == R3D_CAM_F32 ? 6 : ext_dim; }
inline int64_t cam_elem_bytes(int cam_kind) { return cam_kind == R3D_CAM_F64 ? 8 : 4; }

cudaError_t launch_prologue(const PrologueDev* d_desc, const PrologueDev& h_desc, int precision, const InputSpec& in, int batch,
                            int flip_from, cudaStream_t s);
cudaError_t launch_video_encode(const float* uv, float* rays, float* param, int64_t n_points, const void* cam, int cam_kind,
                                int undistort, cudaStream_t s);
cudaError_t launch_assemble(const AssembleDev* d_desc, const AssembleDev& h_desc, float* pos, float* trj, float* sum,
                            int batch, int flip, cudaStream_t s);
cudaError_t launch_gemm_ffma(const GemmOpDev* d_op, const GemmOpDev& h_op, int M, cudaStream_t s);
cudaError_t launch_ray_encode_f64(const double* uv, double* ray, int64_t n, double fx, double fy, double ppx,
                                  double ppy, double c, double s, cudaStream_t st);
cudaError_t launch_normalize_screen_f64(const double* xy, double* out, int64_t n, double w, double h, cudaStream_t st);
cudaError_t launch_eval_metrics(const float* pred, const float* target, int frames, int J, const double* rt, double* acc, cudaStream_t st);
cudaError_t launch_undistort_points_f64(const double* uv, double* out, int64_t n, double fx, double fy, double cx, double cy,
                                        const double* dist5, cudaStream_t st);
cudaError_t prologue_configure(int max_smem_bytes);

// tensor-core path (r3d_gemm_tc.cu)
struct TcOpHost;   // per-op TMA descriptors etc.
cudaError_t launch_gemm_tc(const GemmOpDev* d_op, const GemmOpDev& h_op, const void* d_tmaps, int M, int precision,
                           cudaStream_t s);
int tc_build_tmaps(const GemmOpDev& h_op, int precision, int64_t cap_rows, void* h_tmaps_out /* kMaxProb*4 maps */);
cudaError_t tc_configure();
cudaError_t tail_configure();
cudaError_t tail_stats_read(unsigned long long* out, int reset);   // experiment builds: cycle accounting of the chained tail launch
int tc_num_sms();
// one launch for every op listed in `mo` (device copy d_mo); ops / tensor maps are the plan's arrays
cudaError_t launch_tail_tc(const GemmOpDev* d_ops, const void* d_tmaps, const MultiOpDev* d_mo, const MultiOpDev& h_mo, int M, int precision,
                           int max_clusters, cudaStream_t s);
inline bool tail_uses_pairs(int M) { return M >= 256; }                       // the chained launch works on 256-row units (CTA pairs)
inline int tail_row_groups(int M) { return (M + 255) / 256; }
void tc_trace_arm(int launches_from_now);              // diagnostics: per-tile clock trace of one GEMM launch
cudaError_t tc_trace_read(long long* out, int cap);     // 3 roles x 64 tiles x 8 events
constexpr int kTmapsPerProb = 6 + 2 * kMaxDst + 4 + 2;   // A hi/lo, W hi/lo, W hi/lo (half tile), {hi, lo} store maps per destination, W2 hi/lo full + half,
                                                         // W hi/lo half tile at the chained tail launch's unit width
constexpr int kTmapW2 = 6 + 2 * kMaxDst;
constexpr int kTmapTailW = kTmapW2 + 4;
constexpr int kTmapBytes = 128;

}  // namespace r3d
