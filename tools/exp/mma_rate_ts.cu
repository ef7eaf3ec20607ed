// Microbenchmark: rate of tcgen05.mma M128 x N x K16 (bf16) with the A operand in TENSOR MEMORY (TS form) and the B
// operand in shared memory (SWIZZLE_NONE K-major or SW128), against the SS form; accumulating into one buffer or
// rotating over `nbuf` accumulator buffers.  One CTA per SM; prints cycles per MMA.
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

// mode 0: TS + B NONE, 1: TS + B SW128, 2: SS (A NONE, B SW128), 3: SS (A SW128, B NONE)
template <int N, int mode>
__global__ void __launch_bounds__(128, 1) k(int iters, int nbuf, long long* out, int rnd) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    // two bf16 values in [-2, 2): sign | exponent 0x3f / 0x3e..0x40 | random mantissa
    const uint32_t v = (h & 0x807f807fu) | 0x3f003f00u;
    reinterpret_cast<uint32_t*>(smem)[i] = rnd ? v : 0u;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tb = tslot;
  if (rnd) {   // random A operand in TMEM columns 448..479 (all 128 lanes)
    uint32_t w[8];
    for (int c = 0; c < 4; ++c) {
      for (int e = 0; e < 8; ++e) {
        uint32_t h = (threadIdx.x * 977u + c * 131u + e * 7919u) * 2654435761u;
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        w[e] = (h & 0x807f807fu) | 0x3f003f00u;
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"r"(tb + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 448 + c * 8), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                   "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before_sync();
  }
  __syncthreads();
  tc_fence_after_sync();
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    uint32_t a_addr = smem_u32(smem);             // A: 16 KB
    uint32_t b_addr = smem_u32(smem) + 16384;     // B: up to 64 KB
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a_tmem = tb + 448;             // 32 columns of A (4 K16 chunks)
    uint64_t dA[4], dB[4];
    uint32_t aT[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      aT[q] = a_tmem + q * 8;
      if (mode == 0) dB[q] = desc_none(b_addr + q * 704, 7872, 128);
      else if (mode == 1) dB[q] = make_desc_sw128(b_addr) + 2 * q;
      else if (mode == 2) { dA[q] = desc_none(a_addr + q * 704, 7872, 128); dB[q] = make_desc_sw128(b_addr) + 2 * q; }
      else { dA[q] = make_desc_sw128(a_addr) + 2 * q; dB[q] = desc_none(b_addr + q * 704, 7872, 128); }
    }
    t0 = clock64();
    int b = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tb + b * N;
      if (++b >= nbuf) b = 0;
#pragma unroll
      for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (mode < 2) umma_ts(d, aT[q], dB[q], idesc, 1);
          else umma_bf16(d, dA[q], dB[q], idesc, 1);
        }
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

template <int N, int mode>
void run(int nbuf, long long* d, int rnd) {
  const int iters = 2000;
  if (nbuf * N > 448) return;
  cudaFuncSetAttribute(k<N, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, mode><<<148, 128, 100 * 1024>>>(iters, nbuf, d, rnd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const char* names[4] = {"TS  B=NONE ", "TS  B=SW128", "SS A=NONE B=SW128", "SS A=SW128 B=NONE"};
  printf("%s N=%3d %-18s nbuf=%d : %.1f cycles per MMA (math floor %d)\n", rnd ? "random" : "zeros ", N, names[mode], nbuf,
         (double)mx / (iters * 16.0), N / 2);
}

int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 148 * 8);
  for (int rnd = 0; rnd < 2; ++rnd) { const int nbuf = 1;
#define ALLN(M) run<64, M>(nbuf, d, rnd); run<96, M>(nbuf, d, rnd); run<128, M>(nbuf, d, rnd); run<256, M>(nbuf, d, rnd);
    ALLN(0) ALLN(1) ALLN(2) ALLN(3)
  }
  return 0;
}
