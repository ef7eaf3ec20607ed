// sblk_attention.cuh — fused scaled-dot-product self-attention for short sequences (T <= 128).
// Reference: ScaledDotProductAttention.forward, SBL/transformer/attention.py:72-83 and the head
// split / merge around it in MultiHeadAttention.forward, attention.py:41-55:
//   attn = softmax(Q K^T / sqrt(d_k) masked_fill(key >= length, -inf)) ; out = attn V
// One CTA per (clip, head); K and V of the head live in shared memory as fp32; one warp per query row:
// each lane owns keys {lane, lane+32, ...}, max / sum are warp-shuffle reductions, everything fp32.
// 0.05 % of the encoder's FLOPs -> latency bound; no tensor cores on purpose.
#pragma once
#include "sblk_common.cuh"

namespace sblk {

struct AttnParams {
  const __nv_bfloat16* qkv;  // [N*T, 3*H*64] : q | k | v, head h at columns h*64 inside each third
  __nv_bfloat16* out;        // [N*T, H*64]   : heads concatenated (== permute(1,2,0,3).view(b, lq, -1))
  float* probs;              // optional [H*N, T, T], batch index h*N + b (attention.py:45,52), or nullptr
  const int* lengths;        // optional [N] valid key counts, or nullptr (= all T)
  int N, T, H;
  float scale;               // 1 / temperature
};

template <int KPL>  // keys per lane: supports T <= 32*KPL
__global__ void __launch_bounds__(128)
attention_kernel(const AttnParams p) {
  constexpr int D = 64;
  constexpr int LDK = D + 1;  // padded fp32 row -> conflict-free column walks
  extern __shared__ float sm[];
  float* sK = sm;                      // [T][65]
  float* sV = sK + p.T * LDK;          // [T][65]
  float* sQ = sV + p.T * LDK;          // [T][64]

  grid_dep_wait();

  const int b = blockIdx.x / p.H;
  const int h = blockIdx.x - b * p.H;
  const int T = p.T;
  const int ld = 3 * p.H * D;
  const int len = (p.lengths != nullptr) ? min(max(__ldg(p.lengths + b), 0), T) : T;

  // cooperative load: T rows x 3 x 64 bf16, 8 values (16 B) per thread-iteration
  const __nv_bfloat16* base = p.qkv + static_cast<size_t>(b) * T * ld + h * D;
  for (int i = threadIdx.x; i < T * 3 * (D / 8); i += blockDim.x) {
    const int c8 = i % (D / 8);
    int r = i / (D / 8);
    const int which = r % 3;
    const int t = r / 3;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * ld + which * p.H * D) + c8);
    float* dst = (which == 0) ? (sQ + t * D + c8 * 8) : ((which == 1 ? sK : sV) + t * LDK + c8 * 8);
    dst[0] = bf16_lo(u.x); dst[1] = bf16_hi(u.x);
    dst[2] = bf16_lo(u.y); dst[3] = bf16_hi(u.y);
    dst[4] = bf16_lo(u.z); dst[5] = bf16_hi(u.z);
    dst[6] = bf16_lo(u.w); dst[7] = bf16_hi(u.w);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int q = warp; q < T; q += (blockDim.x >> 5)) {
    const float* qrow = sQ + q * D;
    float s[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) s[i] = 0.0f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      const float qd = qrow[d];
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const int j = lane + 32 * i;
        if (j < T) s[i] = fmaf(qd, sK[j * LDK + d], s[i]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int j = lane + 32 * i;
      s[i] = (j < len) ? s[i] * p.scale : -INFINITY;
      mx = fmaxf(mx, s[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      s[i] = __expf(s[i] - mx);  // all-masked row: exp(-inf - -inf) = NaN, as in the reference softmax
      sum += s[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < KPL; ++i) s[i] *= inv;

    if (p.probs != nullptr) {
      float* pr = p.probs + (static_cast<size_t>(h) * p.N + b) * T * T + static_cast<size_t>(q) * T;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const int j = lane + 32 * i;
        if (j < T) pr[j] = s[i];
      }
    }

    float o0 = 0.0f, o1 = 0.0f;  // output dims lane and lane + 32
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int jmax = min(32, T - 32 * i);
      for (int jj = 0; jj < jmax; ++jj) {
        const float pj = __shfl_sync(0xffffffffu, s[i], jj);
        const float* vr = sV + (32 * i + jj) * LDK;
        o0 = fmaf(pj, vr[lane], o0);
        o1 = fmaf(pj, vr[lane + 32], o1);
      }
    }
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * T + q) * (p.H * D) + h * D;
    orow[lane] = __float2bfloat16_rn(o0);
    orow[lane + 32] = __float2bfloat16_rn(o1);
  }
  grid_dep_launch();
}

}  // namespace sblk
