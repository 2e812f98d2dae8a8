// Grouped FP32 GEMM with fused bias + LeakyReLU + residual epilogue (R3D_PREC_FP32 path).
//
//   out[p][m, n] = res[p][m, n] + lrelu( sum_k A[p][m, k] * W[p][n, k] + bias[p][n] )
//
// A is row-major (K contiguous), W is the BN-folded weight matrix packed K-major, so both operand
// tiles are read with 16-byte coalesced loads along K.  This is the exact-fp32 restatement of
// F.conv1d(stride == kernel) / F.linear + BatchNorm1d(eval) + LeakyReLU (+ residual) at
// lib/model/rie.py:86-99 and :122-135,159-169.  128x128x16 CTA tile, 8x8 register micro-tile,
// register-staged double buffering; it is also the on-device checker for the tcgen05 path.
#include "r3d_internal.h"

namespace r3d {

constexpr int FBM = 128, FBN = 128, FBK = 16, FTHREADS = 256;

__global__ void __launch_bounds__(FTHREADS, 2) gemm_ffma_kernel(const GemmOpDev* __restrict__ opp, int M) {
  __shared__ __align__(16) float As[2][FBK][FBM + 4];
  __shared__ __align__(16) float Ws[2][FBK][FBN + 4];

  const GemmOpDev& op = *opp;
  const int m_tiles = (M + FBM - 1) / FBM;
  int tile = blockIdx.x;
  // tile -> (problem, m tile, n tile); n fastest so CTAs sharing an A tile are co-resident
  int p = 0, n_tiles = 0;
  for (;; ++p) {
    n_tiles = (op.prob[p].n_pad + FBN - 1) / FBN;
    const int cnt = m_tiles * n_tiles;
    if (tile < cnt) break;
    tile -= cnt;
  }
  const GemmProb& pr = op.prob[p];
  const int m0 = (tile / n_tiles) * FBM, n0 = (tile % n_tiles) * FBN;
  const int K = pr.K;
  const float* __restrict__ A = reinterpret_cast<const float*>(pr.a.p0);
  const float* __restrict__ W = reinterpret_cast<const float*>(pr.w0);
  const int lda = pr.a.ld;

  const int tid = threadIdx.x;
  // global->smem staging: each thread moves 2 float4 of A and 2 of W per k-step
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = m0 + lr + 64 * i, n = n0 + lr + 64 * i;
      ra[i] = (m < M) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)m * lda + k0 + lk)) : make_float4(0, 0, 0, 0);
      rw[i] = (n < pr.n_pad) ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K + k0 + lk)) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lr + 64 * i;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y; Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
    }
  };

  const int ty = tid >> 4, tx = tid & 15;   // 16x16 threads, each 2x(4 rows) x 2x(4 cols)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / FBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * FBK);
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: bias, LeakyReLU, residual, scatter to destinations -------------------------------
  const float slope = op.slope;
  const float* __restrict__ R = reinterpret_cast<const float*>(pr.res.p0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= pr.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j] + ((n + j < pr.n_pad) ? __ldg(pr.bias + n + j) : 0.f);
        x = x > 0.f ? x : slope * x;
        v[j] = x;
      }
      if (R != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < pr.N) v[j] += __ldg(R + (int64_t)m * pr.res.ld + pr.res_col + n + j);
      }
      for (int t = 0; t < pr.ndst; ++t) {
        float* o = reinterpret_cast<float*>(pr.dst[t].m.p0) + (int64_t)m * pr.dst[t].m.ld + pr.dst[t].col + n;
        if (n + 3 < pr.N && ((pr.dst[t].m.ld | pr.dst[t].col) & 3) == 0) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < pr.N) o[j] = v[j];
        }
      }
    }
  }
}

cudaError_t launch_gemm_ffma(const GemmOpDev* d_op, const GemmOpDev& h, int M, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  const int m_tiles = (M + FBM - 1) / FBM;
  int tiles = 0;
  for (int p = 0; p < h.nprob; ++p) tiles += m_tiles * ((h.prob[p].n_pad + FBN - 1) / FBN);
  gemm_ffma_kernel<<<tiles, FTHREADS, 0, s>>>(d_op, M);
  return cudaGetLastError();
}

}  // namespace r3d
