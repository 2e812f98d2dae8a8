// Input stage (camera-ray encode + positional/temporal differences + joint-group gather +
// camera embedding) and output stage (joint permutation, pos + trj) of the lifting path.
//
// Reference behaviour restated here (paths relative to the reference checkout):
//   ray encode ................. lib/camera/camera.py:423-441, 460-471 (undistort=False)
//   in_current / diff / diff_t . lib/model/rie.py:289-304
//   group gather + cat ......... lib/model/rie.py:306-357 (pose), :540 (trajectory)
//   Embedding .................. lib/model/embedding.py:15-19 (LeakyReLU slope 0.01)
//   output joint order ......... lib/model/rie.py:415-432; pos += trj trainer.py:353
#include <cstdlib>

#include "r3d_internal.h"

namespace r3d {

__device__ __forceinline__ void store_act(const Mat& m, int precision, int64_t row, int col, float v) {
  const int64_t idx = row * m.ld + col;
  if (precision == R3D_PREC_FP32) {
    reinterpret_cast<float*>(m.p0)[idx] = v;
  } else {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(m.p0)[idx] = hi;
    if (precision == R3D_PREC_BF16X3)
      reinterpret_cast<__nv_bfloat16*>(m.p1)[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

__device__ __forceinline__ void store_act2(const Mat& m, int precision, int64_t row, int col, float v0, float v1) {
  const int64_t idx = row * m.ld + col;   // col even, ld even => 8-byte (fp32) / 4-byte (bf16) aligned
  if (precision == R3D_PREC_FP32) {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(m.p0) + idx) = make_float2(v0, v1);
  } else {
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
    *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(m.p0) + idx) = hi;
    if (precision == R3D_PREC_BF16X3) {
      const float2 hf = __bfloat1622float2(hi);
      *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(m.p1) + idx) =
          __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
    }
  }
}

// IEEE-754 correctly rounded a / b from rb = RN(1 / b) with five FMA-pipe operations instead of the ~30-instruction
// __ddiv_rn sequence: after one residual correction q is faithful, and with a correctly rounded reciprocal the second
// correction rounds to RN(a / b) (Markstein's theorem; scripts/check_fma_division.c compares 4e8 operand pairs of
// this path's ranges with the hardware quotient: no mismatch).  No overflow/underflow handling: operands are pixel
// offsets and focal lengths.
__device__ __forceinline__ double div_by(double a, double b, double rb) {
  double q = __dmul_rn(a, rb);
  double r = __fma_rn(-q, b, a);
  q = __fma_rn(r, rb, q);
  r = __fma_rn(-q, b, a);
  return __fma_rn(r, rb, q);
}

// ---- camera row of one window ---------------------------------------------------------------------------------------
// R3D_CAM_F32: 6 floats [fx, fy, cx, cy, pitch, height]; sin/cos of the pitch are evaluated on the device.
// R3D_CAM_F64: 16 doubles [fx, fy, ppx, ppy, cos(pitch), sin(pitch), pitch, height, k1, k2, p1, p2, k3, undistort, K02, K12]
//   (pp = the principal point the encode subtracts, camera.py:253-259: the UNDISTORTED one for a distorted lens; K02/K12 =
//   the raw principal point the lens model itself uses):
//   the reference's own float64 intrinsics (camera.py:438-439 divides by float64 K entries), cos/sin as the host's libm
//   produced them for Rc2n (camera.py:333-338), and the lens model of cv2.undistortPoints (camera.py:412-421, :435-436).
struct CamRow {
  double fx, fy, cx, cy, rfx, rfy, cp, sp;     // cx, cy: principal point of the encode (pp_cam)
  double k1, k2, p1, p2, k3, kcx, kcy;        // lens model; kcx, kcy: K's own principal point
  float height, pitch;
  bool undistort;
};

__device__ __forceinline__ CamRow load_cam(const void* cam, int64_t stride, int kind, int64_t row) {
  CamRow c;
  c.k1 = c.k2 = c.p1 = c.p2 = c.k3 = c.kcx = c.kcy = 0.0;
  c.undistort = false;
  if (kind == R3D_CAM_F64) {
    const double* r = reinterpret_cast<const double*>(cam) + row * stride;
    c.fx = r[0]; c.fy = r[1]; c.cx = r[2]; c.cy = r[3]; c.cp = r[4]; c.sp = r[5];
    c.pitch = (float)r[6]; c.height = (float)r[7];
    c.k1 = r[8]; c.k2 = r[9]; c.p1 = r[10]; c.p2 = r[11]; c.k3 = r[12];
    c.undistort = r[13] != 0.0;
    c.kcx = r[14]; c.kcy = r[15];
  } else {
    const float* r = reinterpret_cast<const float*>(cam) + row * stride;
    c.fx = r[0]; c.fy = r[1]; c.cx = r[2]; c.cy = r[3];
    sincos((double)r[4], &c.sp, &c.cp);
    c.pitch = r[4]; c.height = r[5];
  }
  c.rfx = __ddiv_rn(1.0, c.fx);
  c.rfy = __ddiv_rn(1.0, c.fy);
  return c;
}

// cv2.undistortPoints(pts, K, dist, P=K) with the 5-coefficient model (k1, k2, p1, p2, k3), pixels in -> pixels out.
// OpenCV's default termination for this overload is exactly 5 fixed-point iterations; the arithmetic follows
// cvUndistortPointsInternal operation by operation with round-to-nearest multiplies/adds (no FMA contraction), which
// reproduces opencv-python 4.13 bit for bit (tests/golden/camera_undistort.npz).
__device__ __forceinline__ void undistort_px(double& u, double& v, const CamRow& c) {
  double x = __dmul_rn(__dsub_rn(u, c.kcx), c.rfx), y = __dmul_rn(__dsub_rn(v, c.kcy), c.rfy);
  const double x0 = x, y0 = y;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    // numerator 1 + ((k6 r2 + k5) r2 + k4) r2 with k4..k6 = 0 evaluates to exactly 1
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
    const double icdist = __ddiv_rn(1.0, den);
    const double two_xx = __dmul_rn(__dmul_rn(2.0, x), x), two_yy = __dmul_rn(__dmul_rn(2.0, y), y);
    const double dx = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p1), x), y), __dmul_rn(c.p2, __dadd_rn(r2, two_xx)));
    const double dy = __dadd_rn(__dmul_rn(c.p1, __dadd_rn(r2, two_yy)), __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p2), x), y));
    x = __dmul_rn(__dsub_rn(x0, dx), icdist);
    y = __dmul_rn(__dsub_rn(y0, dy), icdist);
  }
  // P = K:  xx = fx*x + 0*y + cx,  ww = 1/(0*x + 0*y + 1) = 1
  u = __dadd_rn(__dmul_rn(c.fx, x), c.kcx);
  v = __dadd_rn(__dmul_rn(c.fy, y), c.kcy);
}

// One keypoint: camera.py:435-439 (optional undistortion, (uv - pp) / f) and :471 ([xn, yn, 1] @ Rx(pitch)^T) in float64,
// then .astype(float32) (trainer.py:298).
template <bool UNDIST>
__device__ __forceinline__ float3 encode_keypoint(float2 p, const CamRow& c) {
  double u = (double)p.x, v = (double)p.y;
  if (UNDIST && c.undistort) undistort_px(u, v, c);
  const double xn = div_by(__dsub_rn(u, c.cx), c.fx, c.rfx);
  const double yn = div_by(__dsub_rn(v, c.cy), c.fy, c.rfy);
  return make_float3((float)xn, (float)__dadd_rn(__dmul_rn(c.cp, yn), c.sp), (float)__dadd_rn(__dmul_rn(-c.sp, yn), c.cp));
}

// Hidden layer of the camera embedding (embedding.py:15-16, BN folded, fp32 FFMA): one thread per hidden unit of every net.
__device__ __forceinline__ void embed_hidden(const PrologueDev& d, const float (&prm)[8], float* scratch, int tid, int nthr) {
  const int mid = d.emb_mid;
  for (int t = tid; t < d.n_embed * mid; t += nthr) {
    const EmbedDev& em = d.embed[t / mid];
    const int j = t % mid;
    float acc = __ldg(em.b1 + j);
    for (int i = 0; i < d.ext_dim; ++i) acc = fmaf(__ldg(em.w1 + j * d.ext_dim + i), prm[i], acc);
    scratch[t] = acc > 0.f ? acc : 0.01f * acc;
  }
}

// Output layer of the camera embedding (embedding.py:17-19) by the thread group [0, nthr): one thread per output of every net.
__device__ __forceinline__ void embed_output(const PrologueDev& d, int precision, int b, const float* scratch, int tid, int nthr) {
  const int mid = d.emb_mid;
  for (int t = tid; t < d.n_embed * d.emb_dim; t += nthr) {
    const int e = t / d.emb_dim, j = t % d.emb_dim;
    const EmbedDev& em = d.embed[e];
    const float* w = em.w2 + j * mid;
    const float* h = scratch + e * mid;
    float acc = __ldg(em.b2 + j);
    int i = 0;
    // the weight row is fetched 16 floats (four 16-byte loads) per L2 round trip
    for (; i + 16 <= mid; i += 16) {
      float4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[u] = __ldg(reinterpret_cast<const float4*>(w + i) + u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {                                       // same summation order as the scalar loop
        acc = fmaf(wv[u].x, h[i + 4 * u], acc);
        acc = fmaf(wv[u].y, h[i + 4 * u + 1], acc);
        acc = fmaf(wv[u].z, h[i + 4 * u + 2], acc);
        acc = fmaf(wv[u].w, h[i + 4 * u + 3], acc);
      }
    }
    for (; i < mid; ++i) acc = fmaf(__ldg(w + i), h[i], acc);
    acc = acc > 0.f ? acc : 0.01f * acc;
    for (int q = 0; q < em.ndst; ++q) store_act(em.dst[q].m, precision, b, em.dst[q].col + j, acc);
  }
}

// One CTA per sequence (window).  Dynamic smem: T*J*Cin floats (+ embed scratch).
// src_kind R3D_SRC_UV: pixel keypoints, camera rows per cam_kind; R3D_SRC_RAYS: encoded input, `cam` = param rows (float).
template <bool UNDIST>
__global__ void __launch_bounds__(320) prologue_kernel(const __grid_constant__ PrologueDev d, int precision,
                                                       const float* __restrict__ src, int64_t src_batch_stride,
                                                       int src_is_uv, const void* __restrict__ cam, int64_t cam_stride,
                                                       int cam_kind, int batch, int flip_from) {
  extern __shared__ float smem[];
  // The descriptor (1.6 KB: shapes, destination pointers, embedder pointers) is a kernel PARAMETER (constant bank): read
  // through a pointer its fields were L2 round trips whenever the streamed keypoints had pushed them out of the small L1,
  // and the camera embedding chained three of those (pointer -> pointer -> value).
  __shared__ int kp_next;                  // next 256-keypoint chunk of the window (claimed warp by warp)
  if (threadIdx.x == 0) kp_next = 0;
  const int b = blockIdx.x;
  // flip test-time augmentation (trainer.py:299-302): windows [flip_from, batch) are the mirrored copies of
  // windows [0, batch - flip_from): x component negated, left/right joints swapped, same camera parameters
  const bool flip = b >= flip_from;
  const int bs = flip ? b - flip_from : b;
  CamRow c{};
  if (src_is_uv) c = load_cam(cam, cam_stride, cam_kind, bs);      // camera.py:438-439,471 in float64
  __syncthreads();                                                 // kp_next
  const int T = d.T, J = d.J, JC = d.JC;
  float* xs = smem;                    // [T][JC]
  float* scratch = smem + T * JC + 8;  // [emb_mid] embed hidden
  if (threadIdx.x == 0) xs[T * JC] = 0.f;   // zero word read for the padding columns of the first-layer operand
  float prm[8];                        // the embedder's input row (trainer.py:297: [height, pitch])

  // ---- 1. stage the (ray-encoded) window in shared memory -----------------------------------
  if (src_is_uv) {
    // ray encode in float64, then .astype(float32) (trainer.py:298)
    prm[0] = c.height;
    prm[1] = c.pitch;
    // The camera embedding first, by warps 0-3 only (own named barrier between its two layers): its dependent L2 round
    // trips (bias / weight rows) are then hidden by the other warps' keypoint work instead of sitting at the end of the
    // CTA with nothing beside them (7 of the stage's 44 us).  The keypoints are claimed in chunks of 256 by whichever
    // warp is free, so the embedding warps simply encode fewer of them.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (d.n_embed > 0 && warp < 4) {
      embed_hidden(d, prm, scratch, threadIdx.x, 128);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      embed_output(d, precision, b, scratch, threadIdx.x, 128);
    }
    const float2* uv = reinterpret_cast<const float2*>(src + (int64_t)bs * src_batch_stride);
    // keypoints are fetched in batches of 8 independent loads per thread before any float64 math touches them
    // (one HBM round trip per batch instead of one per keypoint)
    const int nkp = T * J;
    for (;;) {
      int chunk = 0;
      if (lane == 0) chunk = atomicAdd(&kp_next, 1);
      chunk = __shfl_sync(0xffffffffu, chunk, 0);
      const int i0 = chunk * 256 + lane;
      if (chunk * 256 >= nkp) break;
      float2 pv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 32;
        if (i < nkp) {
          int is = i;
          if (flip) {                                   // (the modulo costs three XU operations: mirrored windows only)
            const int jj = i % J;
            is = i - jj + d.flip_perm[jj];
          }
          pv[u] = __ldg(uv + is);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 32;
        if (i < nkp) {
          const float3 r = encode_keypoint<UNDIST>(pv[u], c);
          xs[i * 3 + 0] = flip ? -r.x : r.x;
          xs[i * 3 + 1] = r.y;
          xs[i * 3 + 2] = r.z;
        }
      }
    }
  } else {
    if (d.n_embed > 0 && threadIdx.x < 128) {
      for (int i = 0; i < d.ext_dim && i < 8; ++i) prm[i] = reinterpret_cast<const float*>(cam)[(int64_t)bs * cam_stride + i];
      embed_hidden(d, prm, scratch, threadIdx.x, 128);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      embed_output(d, precision, b, scratch, threadIdx.x, 128);
    }
    const float* x = src + (int64_t)bs * src_batch_stride;
    if (!flip) {
      for (int i = threadIdx.x; i < T * JC; i += blockDim.x) xs[i] = __ldg(x + i);
    } else {
      const int Cin = d.Cin;
      for (int i = threadIdx.x; i < T * JC; i += blockDim.x) {
        const int c = i % Cin, jj = (i / Cin) % J;
        const float v = __ldg(x + i + (d.flip_perm[jj] - jj) * Cin);
        xs[i] = c == 0 ? -v : v;
      }
    }
  }
  __syncthreads();

  // ---- 2. first-layer operand, shared by every joint group (the x-root / x-x[tc] differences are folded into the
  // weights, see pack_expand_folded): row (b, tq) holds the w0 frames [w0*tq, w0*tq+w0) and frame tc of the window,
  // columns ordered joint group by joint group so that each group's weights are non-zero in a few 16-column K steps
  // only.  a0_off[k] (decoded on the host) = shared-memory offset of column k's source | kRowRel when it is relative to
  // the row's first frame.  A lane produces 2 adjacent columns: warp stores cover 128 contiguous bytes per plane.
  {
    const int kp = d.k_pad, k_frames = d.w0 * JC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    constexpr int kRowRel = 1 << 30;
    auto row_out = [&](int tq, auto&& src) {                   // src(g, j) -> value of column lane*2 + 64*g + j
      const int64_t rb = ((int64_t)b * d.L0 + tq) * d.a0.ld + lane * 2;
      if (precision == R3D_PREC_FP32) {
        float* o = reinterpret_cast<float*>(d.a0.p0) + rb;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (lane * 2 + g * 64 < kp) *reinterpret_cast<float2*>(o + g * 64) = make_float2(src(g, 0), src(g, 1));
      } else {
        __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(d.a0.p0) + rb;
        __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(d.a0.p1) + rb;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (lane * 2 + g * 64 < kp) {
            const float v0 = src(g, 0), v1 = src(g, 1);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
            *reinterpret_cast<__nv_bfloat162*>(oh + g * 64) = hi;
            if (precision == R3D_PREC_BF16X3) {
              const float2 hf = __bfloat1622float2(hi);
              *reinterpret_cast<__nv_bfloat162*>(ol + g * 64) = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
            }
          }
      }
    };
    if (kp == 256 && precision != R3D_PREC_FP32) {
      // the shipped configurations on the tensor path: a lane owns 8 ADJACENT columns -- one 16-byte store per plane and
      // row (512 contiguous bytes per warp), the hi/lo split on packed fp32x2 operations
      int cur[8], step[8];        // source offset of this lane's 8 columns for the warp's current row / its advance per row
      {
        const int4 e0 = __ldg(reinterpret_cast<const int4*>(d.a0_off + lane * 8)), e1 = __ldg(reinterpret_cast<const int4*>(d.a0_off + lane * 8) + 1);
        const int e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rel = (e[j] & kRowRel) ? 1 : 0;                 // relative to the row's first frame, or fixed (frame tc, zero word)
          cur[j] = (e[j] & ~kRowRel) + rel * warp * k_frames;
          step[j] = rel * nwarp * k_frames;
        }
      }
      __nv_bfloat16* const oh = reinterpret_cast<__nv_bfloat16*>(d.a0.p0);
      __nv_bfloat16* const ol = reinterpret_cast<__nv_bfloat16*>(d.a0.p1);
      const bool x3 = precision == R3D_PREC_BF16X3;
      for (int tq = warp; tq < d.L0; tq += nwarp) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v = make_float2(xs[cur[2 * j]], xs[cur[2 * j + 1]]);
          cur[2 * j] += step[2 * j];
          cur[2 * j + 1] += step[2 * j + 1];
          const __nv_bfloat162 hh = __floats2bfloat162_rn(v.x, v.y);
          hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
          const float2 dlt = __ffma2_rn(__bfloat1622float2(hh), make_float2(-1.f, -1.f), v);      // v - hi, one rounding
          const __nv_bfloat162 ll = __floats2bfloat162_rn(dlt.x, dlt.y);
          lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const int64_t rb = ((int64_t)b * d.L0 + tq) * d.a0.ld + lane * 8;
        *reinterpret_cast<uint4*>(oh + rb) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (x3) *reinterpret_cast<uint4*>(ol + rb) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else if (kp <= 512) {                                    // other widths / fp32: map in registers
      int off[16];
      uint32_t relbits = 0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int k = lane * 2 + g * 64;
        const int2 e = k < kp ? __ldg(reinterpret_cast<const int2*>(d.a0_off + k)) : make_int2(T * JC, T * JC);
        off[2 * g] = e.x & ~kRowRel;
        off[2 * g + 1] = e.y & ~kRowRel;
        relbits |= (e.x & kRowRel ? 1u : 0u) << (2 * g) | (e.y & kRowRel ? 1u : 0u) << (2 * g + 1);
      }
      for (int tq = warp; tq < d.L0; tq += nwarp) {          // one warp per row: no index division, 128-byte warp stores
        const int base = tq * k_frames;
        row_out(tq, [&](int g, int j) { return xs[off[2 * g + j] + base * (int)((relbits >> (2 * g + j)) & 1u)]; });
      }
    } else {
      for (int tq = warp; tq < d.L0; tq += nwarp) {
        const int base = tq * k_frames;
        for (int kk = lane * 2; kk < kp; kk += 64) {
          const int e0 = d.a0_off[kk], e1 = d.a0_off[kk + 1];
          store_act2(d.a0, precision, (int64_t)b * d.L0 + tq, kk, xs[(e0 & ~kRowRel) + (e0 & kRowRel ? base : 0)],
                     xs[(e1 & ~kRowRel) + (e1 & kRowRel ? base : 0)]);
        }
      }
    }
  }

  // ---- 3. in_current (rie.py:290-292), zero padded to the row pitch ------------------------------
  for (int i = threadIdx.x; i < d.inc.ld; i += blockDim.x)
    store_act(d.inc, precision, b, i, i < JC ? xs[d.tc * JC + i] : 0.f);

}

static int g_prologue_smem_cap = 48 * 1024;

cudaError_t prologue_configure(int max_smem_bytes) {
  for (int u = 0; u < 2; ++u) {
    const void* fn = u ? (const void*)prologue_kernel<true> : (const void*)prologue_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
    if (e == cudaSuccess)   // several windows per SM: ask for the largest shared-memory carve-out
      e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  g_prologue_smem_cap = max_smem_bytes;
  return cudaSuccess;
}

cudaError_t launch_prologue(const PrologueDev* d_desc, const PrologueDev& h, int precision, const InputSpec& in, int batch,
                            int flip_from, cudaStream_t s) {
  const size_t smem = (size_t)(h.T * h.JC + 2 * h.emb_mid + 16) * sizeof(float);
  if ((int)smem > g_prologue_smem_cap) return cudaErrorInvalidValue;
  const int threads = 256;   // measured: 4 CTAs x 256 threads per SM (register-limited) beat 3 x 320 (57 vs 64 us at B=1024, T=243)
  const int is_uv = in.src_kind == R3D_SRC_UV;
  if (is_uv && in.undistort)   // lens undistortion inside the encode: a separate instantiation keeps the common path's registers
    prologue_kernel<true><<<batch, threads, smem, s>>>(h, precision, in.src, in.src_stride, is_uv, in.cam, in.cam_stride, in.cam_kind,
                                                     batch, flip_from);
  else
    prologue_kernel<false><<<batch, threads, smem, s>>>(h, precision, in.src, in.src_stride, is_uv, in.cam, in.cam_stride, in.cam_kind,
                                                      batch, flip_from);
  return cudaGetLastError();
}

// ---- per-frame ray encode of a whole video (r3d_forward_video_uv): every frame is encoded ONCE -- the windows that
// contain it are indexed in the input stage afterwards (trainer.py:47-58 copies each frame RF times instead).  One camera
// row for the whole video; also writes the embedder's param row [height, pitch] (trainer.py:297, :324).
template <bool UNDIST>
__global__ void video_encode_kernel(const float2* __restrict__ uv, float* __restrict__ rays, float* __restrict__ param, int64_t n,
                                    const void* __restrict__ cam, int cam_kind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CamRow c = load_cam(cam, 0, cam_kind, 0);
  const float3 r = encode_keypoint<UNDIST>(__ldg(uv + i), c);
  rays[i * 3 + 0] = r.x;
  rays[i * 3 + 1] = r.y;
  rays[i * 3 + 2] = r.z;
  if (i == 0 && param != nullptr) { param[0] = c.height; param[1] = c.pitch; }
}

cudaError_t launch_video_encode(const float* uv, float* rays, float* param, int64_t n_points, const void* cam, int cam_kind,
                                int undistort, cudaStream_t s) {
  if (n_points <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n_points + 255) / 256);
  if (undistort) video_encode_kernel<true><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(uv), rays, param, n_points, cam, cam_kind);
  else video_encode_kernel<false><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(uv), rays, param, n_points, cam, cam_kind);
  return cudaGetLastError();
}

// ---- output stage ------------------------------------------------------------------------------
// flip != 0: head rows [batch, 2*batch) hold the predictions for the mirrored windows; un-mirror them (negate x,
// swap left/right slots) and average with the direct prediction (trainer.py:338-353; torch.mean of two values).
// (descriptor as a kernel parameter: the kernel is the last link of every forward's dependency chain, and reading the head
// pointers through a descriptor pointer added a dependent L2 round trip to it)
__global__ void assemble_kernel(const __grid_constant__ AssembleDev d, float* __restrict__ pos, float* __restrict__ trj,
                                float* __restrict__ sum, int batch, int flip) {
  const int per = d.J * 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)batch * per) return;
  const int b = (int)(i / per), r = (int)(i % per), slot = r / 3, c = r % 3;
  const float sgn = c == 0 ? -1.f : 1.f;
  float t = 0.f;
  if (d.has_trj) {
    const float* h = d.heads[kMaxProb - 1];
    t = h[(int64_t)b * d.head_ld + c];
    if (flip) t = (t + sgn * h[(int64_t)(b + batch) * d.head_ld + c]) / 2.f;
    if (trj != nullptr && slot == 0) trj[(int64_t)b * 3 + c] = t;
  }
  if (d.has_pos) {
    float v = d.heads[d.slot_prob[slot]][(int64_t)b * d.head_ld + d.slot_joint[slot] * 3 + c];
    if (flip) {
      const int fs = d.flip_slot[slot];
      v = (v + sgn * d.heads[d.slot_prob[fs]][(int64_t)(b + batch) * d.head_ld + d.slot_joint[fs] * 3 + c]) / 2.f;
    }
    if (pos != nullptr) pos[i] = v;
    if (sum != nullptr) sum[i] = v + t;
  }
}

cudaError_t launch_assemble(const AssembleDev* d_desc, const AssembleDev& h, float* pos, float* trj, float* sum,
                            int batch, int flip, cudaStream_t s) {
  const int64_t n = (int64_t)batch * h.J * 3;
  assemble_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h, pos, trj, sum, batch, flip);
  return cudaGetLastError();
}

// ---- standalone float64 ray encode (CameraInfoPacket.get_cam_ray_given_uv, camera.py:460-471) ----
__global__ void ray_encode_f64_kernel(const double2* __restrict__ uv, double* __restrict__ ray, int64_t n, double fx,
                                      double fy, double ppx, double ppy, double c, double s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 p = uv[i];
  const double xn = __ddiv_rn(__dsub_rn(p.x, ppx), fx);
  const double yn = __ddiv_rn(__dsub_rn(p.y, ppy), fy);
  ray[i * 3 + 0] = xn;
  ray[i * 3 + 1] = __dadd_rn(__dmul_rn(c, yn), s);     // no FMA contraction: matches numpy
  ray[i * 3 + 2] = __dadd_rn(__dmul_rn(-s, yn), c);
}

cudaError_t launch_ray_encode_f64(const double* uv, double* ray, int64_t n, double fx, double fy, double ppx,
                                  double ppy, double c, double s, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ray_encode_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const double2*>(uv), ray, n, fx,
                                                                    fy, ppx, ppy, c, s);
  return cudaGetLastError();
}

// ---- lens undistortion: CameraInfoPacket.undistort_point (camera.py:412-421), float64 pixels in -> float64 pixels out
__global__ void undistort_points_f64_kernel(const double2* __restrict__ uv, double2* __restrict__ out, int64_t n, double fx, double fy,
                                            double cx, double cy, double k1, double k2, double p1, double p2, double k3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CamRow c;
  c.fx = fx; c.fy = fy; c.kcx = cx; c.kcy = cy;
  c.rfx = __ddiv_rn(1.0, fx); c.rfy = __ddiv_rn(1.0, fy);
  c.k1 = k1; c.k2 = k2; c.p1 = p1; c.p2 = p2; c.k3 = k3;
  double2 p = uv[i];
  undistort_px(p.x, p.y, c);
  out[i] = p;
}

cudaError_t launch_undistort_points_f64(const double* uv, double* out, int64_t n, double fx, double fy, double cx, double cy,
                                        const double* dist5, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  undistort_points_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const double2*>(uv), reinterpret_cast<double2*>(out),
                                                                          n, fx, fy, cx, cy, dist5[0], dist5[1], dist5[2], dist5[3], dist5[4]);
  return cudaGetLastError();
}

// ---- normalize_screen_coordinates (camera.py:11-18): X / w * 2 - [1, h / w], float64 ---------------
__global__ void normalize_screen_f64_kernel(const double2* __restrict__ xy, double2* __restrict__ out, int64_t n, double w,
                                            double hw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 p = xy[i];
  out[i] = make_double2(__dsub_rn(__dmul_rn(__ddiv_rn(p.x, w), 2.0), 1.0), __dsub_rn(__dmul_rn(__ddiv_rn(p.y, w), 2.0), hw));
}

cudaError_t launch_normalize_screen_f64(const double* xy, double* out, int64_t n, double w, double h, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  normalize_screen_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const double2*>(xy),
                                                                          reinterpret_cast<double2*>(out), n, w, h / w);
  return cudaGetLastError();
}

// ---- evaluation tail: normalized2world (camera.py:401-410) + MPJPE / MRPE / N-MPJPE / MPJVE (lib/loss/loss.py) ------
// One warp per frame, lane = joint (J <= 32), float64 like the reference (numpy promotes the float32 predictions when
// they are multiplied by the float64 Rn2w).  acc[0..3] += sum of per-joint errors: position, root, scale-normalised
// position, velocity; acc[4] += sum of per-joint errors after the similarity (Procrustes) alignment of p_mpjpe
// (loss.py:30-69).  Means are taken by the caller (trainer.py:386-395 weights them by frame count anyway).

// Singular value decomposition of a 3x3 matrix (row-major H = U diag(s) V^T, s descending) by one-sided Jacobi
// rotations on the columns: stands in for np.linalg.svd (loss.py:50).  U, V are only determined up to paired column
// signs, but R = V U^T and the singular values -- all p_mpjpe uses -- are unique for non-degenerate poses.
__device__ void svd3x3(const double* H, double* U, double* sv, double* V) {
  double A[9];
  for (int i = 0; i < 9; ++i) { A[i] = H[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
      for (int i = 0; i < 3; ++i) { alpha += A[3 * i + p] * A[3 * i + p]; beta += A[3 * i + q] * A[3 * i + q]; gamma += A[3 * i + p] * A[3 * i + q]; }
      if (fabs(gamma) <= 1e-15 * sqrt(alpha * beta) || gamma == 0.0) continue;
      rotated = true;
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
      for (int i = 0; i < 3; ++i) {
        const double ap = A[3 * i + p], aq = A[3 * i + q];
        A[3 * i + p] = c * ap - sn * aq; A[3 * i + q] = sn * ap + c * aq;
        const double vp = V[3 * i + p], vq = V[3 * i + q];
        V[3 * i + p] = c * vp - sn * vq; V[3 * i + q] = sn * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  for (int k = 0; k < 3; ++k) sv[k] = sqrt(A[k] * A[k] + A[3 + k] * A[3 + k] + A[6 + k] * A[6 + k]);
  auto swap_cols = [&](int a, int b) {
    for (int i = 0; i < 3; ++i) {
      double x = A[3 * i + a]; A[3 * i + a] = A[3 * i + b]; A[3 * i + b] = x;
      x = V[3 * i + a]; V[3 * i + a] = V[3 * i + b]; V[3 * i + b] = x;
    }
    const double x = sv[a]; sv[a] = sv[b]; sv[b] = x;
  };
  if (sv[0] < sv[1]) swap_cols(0, 1);
  if (sv[1] < sv[2]) swap_cols(1, 2);
  if (sv[0] < sv[1]) swap_cols(0, 1);
  for (int k = 0; k < 2; ++k) {
    const double inv = sv[k] > 0.0 ? 1.0 / sv[k] : 0.0;
    for (int i = 0; i < 3; ++i) U[3 * i + k] = A[3 * i + k] * inv;
  }
  if (sv[2] > 1e-14 * sv[0]) {
    for (int i = 0; i < 3; ++i) U[3 * i + 2] = A[3 * i + 2] / sv[2];
  } else {  // planar pose: complete the basis; the reflection fix below makes R independent of this sign
    U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
  }
}

__global__ void eval_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, int frames, int J,
                                    const double* __restrict__ rt /* 9 + 3, or null */, double* __restrict__ acc) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= frames) return;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
  if (rt != nullptr) {
    for (int i = 0; i < 9; ++i) R[i] = rt[i];
    for (int i = 0; i < 3; ++i) T[i] = rt[9 + i];
  }
  auto world = [&](const float* src, int f, double (&o)[3]) {    // pt @ Rn2w.T + Tn2w.T
    const float* q = src + ((int64_t)f * J + lane) * 3;
    const double x = q[0], y = q[1], z = q[2];
    for (int i = 0; i < 3; ++i) o[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, R[3 * i]), __dmul_rn(y, R[3 * i + 1])), __dmul_rn(z, R[3 * i + 2])), T[i]);
  };
  const bool on = lane < J;
  double p[3] = {0, 0, 0}, t[3] = {0, 0, 0};
  if (on) { world(pred, warp, p); world(target, warp, t); }
  double e_pos = 0, s_pp = 0, s_tp = 0, e_vel = 0;
  if (on) {
    const double dx = p[0] - t[0], dy = p[1] - t[1], dz = p[2] - t[2];
    e_pos = sqrt(dx * dx + dy * dy + dz * dz);                                   // loss.py:17
    s_pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];                              // loss.py:79
    s_tp = t[0] * p[0] + t[1] * p[1] + t[2] * p[2];                              // loss.py:80
    if (warp + 1 < frames) {                                                     // loss.py:101-104
      double p1[3], t1[3];
      world(pred, warp + 1, p1); world(target, warp + 1, t1);
      const double vx = (p1[0] - p[0]) - (t1[0] - t[0]), vy = (p1[1] - p[1]) - (t1[1] - t[1]), vz = (p1[2] - p[2]) - (t1[2] - t[2]);
      e_vel = sqrt(vx * vx + vy * vy + vz * vz);
    }
  }
  const double e_root = lane == 0 ? e_pos : 0.0;
  auto wsum = [](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  const double scale = (wsum(s_tp) / J) / (wsum(s_pp) / J);                      // loss.py:79-81 (means over joints)
  double e_n = 0;
  if (on) {
    const double dx = scale * p[0] - t[0], dy = scale * p[1] - t[1], dz = scale * p[2] - t[2];
    e_n = sqrt(dx * dx + dy * dy + dz * dz);
  }
  // p_mpjpe (loss.py:30-69): X = target, Y = predicted; every lane carries one joint and the 3x3 algebra redundantly
  double e_pa = 0;
  {
    double muX[3], muY[3], X0[3], Y0[3];
    for (int i = 0; i < 3; ++i) { muX[i] = wsum(on ? t[i] : 0.0) / J; muY[i] = wsum(on ? p[i] : 0.0) / J; }
    for (int i = 0; i < 3; ++i) { X0[i] = on ? t[i] - muX[i] : 0.0; Y0[i] = on ? p[i] - muY[i] : 0.0; }
    const double normX = sqrt(wsum(X0[0] * X0[0] + X0[1] * X0[1] + X0[2] * X0[2]));
    const double normY = sqrt(wsum(Y0[0] * Y0[0] + Y0[1] * Y0[1] + Y0[2] * Y0[2]));
    for (int i = 0; i < 3; ++i) { X0[i] /= normX; Y0[i] /= normY; }
    double H[9], U[9], sv[3], V[9], R[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) H[3 * a + b] = wsum(X0[a] * Y0[b]);                 // H = X0^T Y0   (:49)
    svd3x3(H, U, sv, V);
    auto vut = [&]() {                                                            // R = V U^T     (:52, :58)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = V[3 * i] * U[3 * j] + V[3 * i + 1] * U[3 * j + 1] + V[3 * i + 2] * U[3 * j + 2];
    };
    vut();
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0.0) {                                                              // reflections   (:55-58)
      for (int i = 0; i < 3; ++i) V[3 * i + 2] = -V[3 * i + 2];
      sv[2] = -sv[2];
      vut();
    }
    const double a = (sv[0] + sv[1] + sv[2]) * normX / normY;                      // scale         (:60-62)
    if (on) {
      double d2 = 0;
      for (int j = 0; j < 3; ++j) {                                               // a * pred @ R + t - target  (:63-69)
        const double tj = muX[j] - a * (muY[0] * R[j] + muY[1] * R[3 + j] + muY[2] * R[6 + j]);
        const double al = a * (p[0] * R[j] + p[1] * R[3 + j] + p[2] * R[6 + j]) + tj;
        d2 += (al - t[j]) * (al - t[j]);
      }
      e_pa = sqrt(d2);
    }
  }
  const double a0 = wsum(e_pos), a2 = wsum(e_n), a3 = wsum(e_vel), a4 = wsum(e_pa);
  if (lane == 0) {
    atomicAdd(acc + 0, a0);
    atomicAdd(acc + 1, e_root);
    atomicAdd(acc + 2, a2);
    atomicAdd(acc + 3, a3);
    atomicAdd(acc + 4, a4);
  }
}

cudaError_t launch_eval_metrics(const float* pred, const float* target, int frames, int J, const double* rt, double* acc, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(acc, 0, 5 * sizeof(double), st);
  if (e != cudaSuccess || frames <= 0) return e;
  const int warps_per_block = 8;
  eval_metrics_kernel<<<(frames + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(pred, target, frames, J, rt, acc);
  return cudaGetLastError();
}

}  // namespace r3d
