// Experiment: SWIZZLE_NONE K-major descriptor over a flat array of 16-B entries (8 bf16 each):
//   A[m][k<8] = E[s0+m][k],  A[m][8+k] = E[s0+m+lbo_entries][k]   (SBO = 128 B, LBO = lbo_entries*16 B)
// B (sw128) selects column n<16 -> D[m][n] = A[m][n].
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../sbl_for_multilingual_lip_reading_b200/csrc/sblk_common.cuh"
using namespace sblk;

__device__ __forceinline__ uint64_t make_desc_none(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;  // layout type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128, 1) k(int s0, int lbo_entries, int swap, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* E = smem;                   // 4096 entries * 16 B = 64 KB
  uint8_t* B = smem + 4096 * 16;       // 64 rows * 128 B
  for (int i = threadIdx.x; i < 4096 * 8; i += 128) {
    int e = i / 8, j = i % 8;
    *reinterpret_cast<__nv_bfloat16*>(E + e * 16 + j * 2) = __float2bfloat16(float((e * 5 + j * 3) % 17 - 8));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    int n = i / 64, kk = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(B + n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2) =
        __float2bfloat16((n == kk && n < 16) ? 1.0f : 0.0f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    uint32_t a_addr = smem_u32(E) + s0 * 16;
    uint64_t da = swap ? make_desc_none(a_addr, 128, lbo_entries * 16) : make_desc_none(a_addr, lbo_entries * 16, 128);
    uint64_t db = make_desc_sw128(smem_u32(B));
    umma_bf16(tb, da, db, make_idesc_bf16(128, 64), 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 0x7702);
  tc_fence_after_sync();
  int warp = threadIdx.x >> 5;
  uint32_t v[32];
  tmem_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 64);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  static float h[128 * 16];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int s0s[4] = {0, 3, 44, 137};
  int lbos[5] = {8, 1, 44, 1000, 2068};
  for (int swap = 0; swap < 2; ++swap)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 5; ++b) {
        k<<<1, 128, 100 * 1024>>>(s0s[a], lbos[b], swap, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 16; ++n) {
            int e_ = s0s[a] + m + (n >= 8 ? lbos[b] : 0), j = n & 7;
            float ref = float((e_ * 5 + j * 3) % 17 - 8);
            if (h[m * 16 + n] != ref) ++bad;
          }
        printf("swap %d s0 %3d lbo_entries %4d: mismatches %d / 2048\n", swap, s0s[a], lbos[b], bad);
      }
  return 0;
}
