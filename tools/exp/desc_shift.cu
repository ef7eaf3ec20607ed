// Experiment: can a SWIZZLE_128B K-major UMMA descriptor start at an arbitrary 128-B row of a swizzled pixel array?
// A = 256 "pixels" x 64 ch (sw128, base 1024-aligned); B = identity 64x64 -> D[m][n] = A[shift+m][n].
// For each shift in 0..17 and base_offset mode (0: field = 0, 1: field = (start>>7)&7) print #mismatches.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../sbl_for_multilingual_lip_reading_b200/csrc/sblk_common.cuh"
using namespace sblk;

__global__ void __launch_bounds__(128, 1) k(int shift, int mode, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* A = smem;                 // 256 rows * 128 B = 32 KB
  uint8_t* B = smem + 256 * 128;     // 64 rows * 128 B
  for (int i = threadIdx.x; i < 256 * 64; i += 128) {
    int p = i / 64, kk = i % 64;
    float v = float((p * 7 + kk * 3) % 13 - 6);
    *reinterpret_cast<__nv_bfloat16*>(A + p * 128 + (((kk >> 3) ^ (p & 7)) << 4) + (kk & 7) * 2) = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    int n = i / 64, kk = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(B + n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2) =
        __float2bfloat16(n == kk ? 1.0f : 0.0f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    uint32_t a_addr = smem_u32(A) + shift * 128;
    uint64_t da = make_desc_sw128(a_addr);
    if (mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    uint64_t db = make_desc_sw128(smem_u32(B));
    for (int kq = 0; kq < 4; ++kq) umma_bf16(tb, da + 2 * kq, db + 2 * kq, make_idesc_bf16(128, 64), kq > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 0x7701);
  tc_fence_after_sync();
  int warp = threadIdx.x >> 5;
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 64);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  static float h[128 * 64];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift = 0; shift < 18; ++shift) {
      k<<<1, 128, 64 * 1024>>>(shift, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = float(((shift + m) * 7 + n * 3) % 13 - 6);
          if (h[m * 64 + n] != ref) ++bad;
        }
      printf("mode %d shift %2d: mismatches %d / 8192\n", mode, shift, bad);
    }
  return 0;
}
