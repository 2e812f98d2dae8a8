// Micro-benchmark: how fast can one SM push an epilogue-shaped output stream through TMA tensor stores, as a function of
// the box shape?  Emulates the first-layer GEMM's output pattern (6 problems x 82944 rows x 256 bf16 columns x 2 planes =
// 510 MB): persistent CTAs, 8 storing warps per CTA, tile = 128 rows x 256 columns, no math, staging tiles never rewritten
// (pure TMA store engine + L2/HBM write path).
//   mode 0: 32 rows x 32 cols, 64B swizzle (2 KB)   -- what the epilogue does today
//   mode 1: 32 rows x 64 cols, 128B swizzle (4 KB)
//   mode 2: 64 rows x 64 cols (8 KB, 4 warps store)   mode 3: 128 rows x 64 cols (16 KB, 4 warps... one per column group)
//   mode 4: plain st.global.v4 from registers, thread = row (64 B contiguous per thread and plane per 32-col chunk)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_store_bench scripts/tma_store_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tmap), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) store_kernel(const CUtensorMap* __restrict__ maps, uint4* p0, uint4* p1, int m_tiles_total, int rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 1024 / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, i, i, i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  uint8_t* stg = smem + warp * 16384;
  for (int tile = blockIdx.x; tile < m_tiles_total; tile += gridDim.x) {
    const int m0 = tile * 128;
    if (MODE == 0) {          // warp: rows q*32, column half h: 4 chunks x 2 planes of 32x32
      const int q = warp & 3, h = warp >> 2;
      for (int c = 0; c < 4; ++c) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          tma_store_2d(maps + 0, stg + (c & 1) * 4096, h * 128 + c * 32, m0 + q * 32);
          tma_store_2d(maps + 1, stg + (c & 1) * 4096 + 2048, h * 128 + c * 32, m0 + q * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
    } else if (MODE == 1) {   // 2 x (32 rows x 64 cols) per plane
      const int q = warp & 3, h = warp >> 2;
      for (int c = 0; c < 2; ++c) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          tma_store_2d(maps + 2, stg + (c & 1) * 8192, h * 128 + c * 64, m0 + q * 32);
          tma_store_2d(maps + 3, stg + (c & 1) * 8192 + 4096, h * 128 + c * 64, m0 + q * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
    } else if (MODE == 2) {   // 64 rows x 64 cols: warp = (row half, column quarter) , 2 planes
      const int rh = warp & 1, cq = warp >> 1;
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        tma_store_2d(maps + 4, stg, cq * 64, m0 + rh * 64);
        tma_store_2d(maps + 5, stg + 8192, cq * 64, m0 + rh * 64);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    } else if (MODE == 3) {   // 128 rows x 64 cols: warp = (plane, column quarter)
      const int pl = warp & 1, cq = warp >> 1;
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        tma_store_2d(maps + 6 + pl, stg, cq * 64, m0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    } else {                  // direct stores: thread = row
      const int q = warp & 3, h = warp >> 2;
      const int row = m0 + q * 32 + lane;
      if (row < rows) {
        for (int c = 0; c < 4; ++c) {
          const size_t o = ((size_t)row * 256 + h * 128 + c * 32) / 8;
#pragma unroll
          for (int u = 0; u < 4; ++u) { p0[o + u] = make_uint4(row, c, u, 0); p1[o + u] = make_uint4(row, c, u, 1); }
        }
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int rows = 82944 * 6, cols = 256;
  void *p0, *p1;
  CK(cudaMalloc(&p0, (size_t)rows * cols * 2));
  CK(cudaMalloc(&p1, (size_t)rows * cols * 2));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<CUtensorMap> maps(8);
  auto mk = [&](int i, void* base, int brow, int bcol, CUtensorMapSwizzle sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)bcol, (cuuint32_t)brow};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode %d failed %d\n", i, (int)r); exit(1); }
  };
  mk(0, p0, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B); mk(1, p1, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  mk(2, p0, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B); mk(3, p1, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  mk(4, p0, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B); mk(5, p1, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  mk(6, p0, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B); mk(7, p1, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap* dmaps;
  CK(cudaMalloc(&dmaps, sizeof(CUtensorMap) * 8));
  CK(cudaMemcpy(dmaps, maps.data(), sizeof(CUtensorMap) * 8, cudaMemcpyHostToDevice));
  const int smem = 128 * 1024, m_tiles = rows / 128;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = 2.0 * rows * cols * 2;
  auto run = [&](int mode, auto kern) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int grid : {148, 296}) {
      if (grid == 296) continue;      // 128 KB smem: one CTA per SM
      for (int i = 0; i < 3; ++i) kern<<<grid, 256, smem>>>(dmaps, (uint4*)p0, (uint4*)p1, m_tiles, rows);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      const int it = 20;
      for (int i = 0; i < it; ++i) kern<<<grid, 256, smem>>>(dmaps, (uint4*)p0, (uint4*)p1, m_tiles, rows);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("mode %d grid %d: %.1f us per launch, %.0f GB/s (510 MB written)\n", mode, grid, ms / it * 1e3, bytes / (ms / it * 1e-3) / 1e9);
    }
  };
  run(0, store_kernel<0>); run(1, store_kernel<1>); run(2, store_kernel<2>); run(3, store_kernel<3>); run(4, store_kernel<4>);
  return 0;
}
