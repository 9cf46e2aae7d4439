// mn_major_probe.cu -- NOT part of the product build.  Checks, on a B200, the one hardware fact the planned TMA-fed 1x1
// forward / dgrad kernel rests on (notes/next_round.md, item 6): tcgen05.mma kind::tf32 with the A operand MN-MAJOR in
// shared memory, straight from an NCHW activation tensor through TMA:
//     D[pixel][o] = sum_c X[c][pixel] * W[o][c]          (X is [C][P] per image: the GEMM-M axis, pixels, is contiguous)
// A tile = 128 pixels x 32 channels loaded by ONE cp.async.bulk.tensor.3d from the view {32 px, C, P/32} of X
// (strides 4 B, P*4 B, 128 B), box {32, 32, 4}, SWIZZLE_128B: shared memory then holds four MN atoms [32 c][32 px] of 4 KB,
// the canonical Major-MN SW128 layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) with LBO = 4096 B (next 32 pixels) and
// SBO = 1024 B (next 8 channels); K step kk of the MMA starts at base + kk*1024.  Instruction descriptor bit 15 = 1.
// The probe tries that encoding and the (LBO,SBO)-swapped one and reports which reproduces the CPU result.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../caffe_mpi_b200/csrc -o mn_major_probe mn_major_probe.cu -lcuda
//   ./mn_major_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_common.cuh"

using namespace b2c::tc;

constexpr int C = 32, P = 128, NO = 32;      // channels (K), pixels (M), output channels (N)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ w, float* __restrict__ out, uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t a_tile = sbase, b_tile = sbase + 16384, bar = sbase + 16384 + 4096, bar_mma = bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sptr + 16384 + 4096 + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 32);
  // B tile: K-major SW128, row n = 128 bytes, 16-byte chunk j stored at position j ^ (n & 7)
  for (int e = tid; e < NO * C; e += 128) {
    const int n = e / C, k = e % C;
    float* dst = reinterpret_cast<float*>(sptr + 16384 + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4);
    *dst = w[e];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384u) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(a_tile), "l"(reinterpret_cast<uint64_t>(&map_x)), "r"(bar), "r"(0), "r"(0), "r"(0) : "memory");
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(128, NO) | (1u << 15);          // A MN-major
      for (int kk = 0; kk < C / 8; ++kk)
        umma_tf32(tmem_base, desc_sw128(a_tile + kk * 1024, lbo, sbo), desc_sw128(b_tile + kk * 32, 16, 1024), idesc, kk != 0);
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < NO; ++j) out[(warp * 32 + lane) * NO + j] = v[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

int main() {
  std::vector<float> x(C * P), w(NO * C), ref(P * NO, 0.f), got(P * NO);
  srand(1701);
  for (auto& v : x) v = (float)(rand() % 17 - 8);          // small integers: exact in TF32
  for (auto& v : w) v = (float)(rand() % 9 - 4);
  for (int p = 0; p < P; ++p)
    for (int o = 0; o < NO; ++o) { float s = 0; for (int c = 0; c < C; ++c) s += x[c * P + p] * w[o * C + c]; ref[p * NO + o] = s; }
  float *dx, *dw, *dout;
  cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&dout, got.size() * 4);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
  alignas(64) CUtensorMap map;
  cuuint64_t dims[3] = {32, (cuuint64_t)C, (cuuint64_t)(P / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)P * 4, 128};
  cuuint32_t box[3] = {32, 32, 4}, es[3] = {1, 1, 1};
  CUresult r = ((Enc)f)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (is a stride of 128 B below the next one accepted?)\n", (int)r); return 2; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const uint32_t variants[2][2] = {{4096, 1024}, {1024, 4096}};
  int rc = 1;
  for (auto& v : variants) {
    cudaMemset(dout, 0, got.size() * 4);
    probe_kernel<<<1, 128, 32768>>>(map, dw, dout, v[0], v[1]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("LBO=%u SBO=%u: CUDA error %s\n", v[0], v[1], cudaGetErrorString(e)); return 3; }
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t i = 0; i < got.size(); ++i) bad += got[i] != ref[i];
    printf("A MN-major SW128, LBO=%u SBO=%u: %s (%d of %zu elements differ)\n", v[0], v[1], bad ? "MISMATCH" : "EXACT", bad, got.size());
    if (!bad) rc = 0;
  }
  return rc;
}
