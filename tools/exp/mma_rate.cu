// Microbenchmark: issue rate of tcgen05.mma (M128 x N x K16, bf16, SS mode) on operands fixed in smem.
// One CTA per SM; prints cycles per MMA for N in {64,128,256} and A layout in {SW128, NONE}.
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

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int layout, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tb = tslot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    uint32_t a_addr = smem_u32(smem);             // A: 16 KB (128 rows x 128 B)
    uint32_t b_addr = smem_u32(smem) + 16384;     // B: N rows x 128 B (up to 32 KB)
    uint64_t db = make_desc_sw128(b_addr);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint64_t da = layout == 0 ? make_desc_sw128(a_addr) + 2 * q : desc_none(a_addr + q * 704, 10368, 128);
        umma_bf16(tb + (it & 1) * N, da, db + 2 * q, make_idesc_bf16(128, N), q > 0);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 0x7703);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int N>
void run(int layout, long long* d) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N><<<148, 128, 100 * 1024>>>(iters, layout, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d A-layout=%s : %.1f cycles per MMA (math floor %d)\n", N, layout ? "NONE " : "SW128",
         (double)mx / (iters * 4.0), N / 2);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  for (int rep = 0; rep < 2; ++rep) {
    run<64>(0, d); run<64>(1, d); run<128>(0, d); run<128>(1, d); run<256>(0, d); run<256>(1, d);
  }
  return 0;
}
