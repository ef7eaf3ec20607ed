// Microbenchmark: does draining accumulators with tcgen05.ld slow a concurrent tcgen05.mma stream (and vice versa)?
// Warp 0 issues TS-form M128 x N96 x K16 MMAs into buffer 0; warps 4..4+W-1 read 32-column windows of buffer 1.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../sbl_for_multilingual_lip_reading_b200/csrc/sblk_common.cuh"
using namespace sblk;

__device__ __forceinline__ uint64_t desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(db), "r"(idesc),
               "r"(acc) : "memory");
}

// do_mma / ld_warps: which sides run.  ss: 1 = SS-form MMAs (A from smem) instead of TS.
__global__ void __launch_bounds__(384, 1) k(int iters, int do_mma, int ld_warps, int ss, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0 && do_mma) {
    if (threadIdx.x == 0) {
      uint32_t b_addr = smem_u32(smem) + 16384;
      const uint32_t idesc = make_idesc_bf16(128, 96);
      uint64_t dB[4], dA[4];
      uint32_t aT[4];
      for (int q = 0; q < 4; ++q) { aT[q] = tb + 448 + q * 8; dB[q] = desc_none(b_addr + q * 704, 7872, 128); dA[q] = make_desc_sw128(smem_u32(smem)) + 2 * q; }
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 6; ++rep)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ss) umma_bf16(tb, dA[q], dB[q], idesc, 1); else umma_ts(tb, aT[q], dB[q], idesc, 1);
          }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0, 0x7703);
      out[blockIdx.x * 2] = clock64() - t0;
    }
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + 96 + ((warp - 4) >> 2) * 32;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t v[32], w[32];
      tmem_ld_32x32b_x32(taddr, v);
      tmem_ld_32x32b_x32(taddr + 44, w);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i] ^ w[i];
    }
    if (acc == 0x12345678u) out[0] = 1;
    if (threadIdx.x == 128) out[blockIdx.x * 2 + 1] = clock64() - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

void run(int do_mma, int ld_warps, int ss, long long* d) {
  const int iters = 500;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaMemset(d, 0, 148 * 16);
  k<<<148, 384, 100 * 1024>>>(iters, do_mma, ld_warps, ss, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[296];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long m0 = 0, m1 = 0;
  for (int i = 0; i < 148; ++i) { m0 = h[2 * i] > m0 ? h[2 * i] : m0; m1 = h[2 * i + 1] > m1 ? h[2 * i + 1] : m1; }
  printf("%s mma=%d ld_warps=%d : MMA side %.1f cycles per 24 MMAs (floor 1152), LDTM side %.1f cycles per iteration (2 x32 loads per warp)\n",
         ss ? "SS" : "TS", do_mma, ld_warps, (double)m0 / iters, (double)m1 / iters);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  for (int ss = 0; ss < 2; ++ss) {
    run(1, 0, ss, d);
    run(0, 4, ss, d);
    run(0, 8, ss, d);
    run(1, 4, ss, d);
    run(1, 8, ss, d);
  }
  return 0;
}
