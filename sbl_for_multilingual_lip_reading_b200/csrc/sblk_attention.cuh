// sblk_attention.cuh — fused scaled-dot-product self-attention for short sequences (T <= 128).
// Reference: ScaledDotProductAttention.forward, SBL/transformer/attention.py:72-83 and the head
// split / merge around it in MultiHeadAttention.forward, attention.py:41-55:
//   attn = softmax(Q K^T / sqrt(d_k) masked_fill(key >= length, -inf)) ; out = attn V
// One WARP per (clip, head): Q, K, V of the head are staged in (warp-private, XOR-swizzled) shared memory, S = Q K^T
// and O = P V run on warp-level tensor-core MMAs (m16n8k16 on the encoder's 16-bit operand format, fp32 accumulate; 0.05 % of the encoder's FLOPs, far
// too small for a tcgen05 tile), the softmax lives in the accumulator registers (fp32, quad-shuffle max / sum), and P is
// fed to the second MMA as a 16-bit hi + lo pair so the probabilities keep fp32-level accuracy.
#pragma once
#include "sblk_common.cuh"

namespace sblk {

struct AttnParams {
  const enc16_t* qkv;  // [N*T, 3*H*64] : q | k | v, head h at columns h*64 inside each third
  enc16_t* out;        // [N*T, H*64]   : heads concatenated (== permute(1,2,0,3).view(b, lq, -1))
  float* probs;              // optional [H*N, T, T], batch index h*N + b (attention.py:45,52), or nullptr
  const int* lengths;        // optional [N] valid key counts, or nullptr (= all T)
  int N, T, H;
  float scale;               // 1 / temperature
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_e16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(SBLK_MMA_SYNC_E16 " {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int ATTN_WARPS = 2;

// One 16-query tile of one (clip, head): S = Q K^T (tensor cores), masked fp32 softmax in the accumulator registers,
// O = P V with P as hi + lo, 16-bit store.  Q / K / V are 16-bit (enc16_t) tiles in shared memory with 128-byte rows (64
// features), 16-byte chunk c of ABSOLUTE row r stored at chunk c ^ (r & 7).  qrow0 = absolute row of this tile's first
// query, krow0 = absolute row of the clip's key 0; rows [krow0, krow0 + 8 NT) must hold finite values.
//   out_clip  : &out[(b*T) * ld_out + h*64]          (row q of the clip at + q * ld_out)
//   probs_clip: &probs[(h*N + b) * T * T] or nullptr
template <int NT>
__device__ __forceinline__ void attention_mtile(uint32_t sQ_u, uint32_t sK_u, uint32_t sV_u, int qrow0, int krow0,
                                                int q_first, int T, int len, float scale, int lane,
                                                enc16_t* out_clip, int ld_out, float* probs_clip) {
  constexpr int D = 64;
  constexpr int KT = NT / 2;             // 16-key tiles for P V
  const int g = lane >> 2;       // fragment row within an 8-row group
  const int tq = lane & 3;       // fragment column pair
  // ---- S = Q K^T for 16 query rows
  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f; }
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    uint32_t a0, a1, a2, a3;
    {
      const int row = qrow0 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = 2 * kk + (lane >> 4);
      ldmatrix_x4(sQ_u + row * 128 + ((c ^ (row & 7)) << 4), a0, a1, a2, a3);
    }
#pragma unroll
    for (int j2 = 0; j2 < NT / 2; ++j2) {
      // two key tiles per ldmatrix.x4: matrices (keys 16*j2.., chunk 2kk), (.., chunk 2kk+1), (keys +8, ..), (..)
      const int row = krow0 + j2 * 16 + (lane & 7) + (lane >> 4) * 8;
      const int c = 2 * kk + ((lane >> 3) & 1);
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(sK_u + row * 128 + ((c ^ (row & 7)) << 4), b0, b1, b2, b3);
      mma_e16_16816(s[2 * j2], a0, a1, a2, a3, b0, b1);
      mma_e16_16816(s[2 * j2 + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- masked softmax over keys (rows g and g + 8 of this tile), fp32
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = (j * 8 + 2 * tq + e) < len;
      s[j][e] = ok ? s[j][e] * scale : -INFINITY;
      s[j][2 + e] = ok ? s[j][2 + e] * scale : -INFINITY;
      mx0 = fmaxf(mx0, s[j][e]);
      mx1 = fmaxf(mx1, s[j][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      // all-masked row: exp(-inf - -inf) = NaN, as in the reference softmax
      s[j][e] = __expf(s[j][e] - mx0);
      s[j][2 + e] = __expf(s[j][2 + e] - mx1);
      sum0 += s[j][e];
      sum1 += s[j][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;
  }
  const int q0 = q_first + g, q1 = q0 + 8;
  if (probs_clip != nullptr) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = j * 8 + 2 * tq + e;
        if (key < T) {
          if (q0 < T) probs_clip[static_cast<size_t>(q0) * T + key] = s[j][e];
          if (q1 < T) probs_clip[static_cast<size_t>(q1) * T + key] = s[j][2 + e];
        }
      }
    }
  }
  // ---- O = P V with P as hi + lo
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      // A-fragment register r: (row g | g+8, keys 16kt + 2tq.. | +8) == accumulator regs of key tiles 2kt, 2kt+1
      const float x = s[2 * kt + (r >> 1)][(r & 1) * 2 + 0];
      const float y = s[2 * kt + (r >> 1)][(r & 1) * 2 + 1];
      hi[r] = pack_e16x2(x, y);
      lo[r] = pack_e16x2(x - e16_lo(hi[r]), y - e16_hi(hi[r]));
    }
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      // V^T fragments for d-chunks 2*n2 and 2*n2+1: matrices (keys 16kt.., chunk), (keys +8, chunk), (.., chunk+1), (..)
      const int row = krow0 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = 2 * n2 + (lane >> 4);
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(sV_u + row * 128 + ((c ^ (row & 7)) << 4), b0, b1, b2, b3);
      mma_e16_16816(o[2 * n2], hi[0], hi[1], hi[2], hi[3], b0, b1);
      mma_e16_16816(o[2 * n2], lo[0], lo[1], lo[2], lo[3], b0, b1);
      mma_e16_16816(o[2 * n2 + 1], hi[0], hi[1], hi[2], hi[3], b2, b3);
      mma_e16_16816(o[2 * n2 + 1], lo[0], lo[1], lo[2], lo[3], b2, b3);
    }
  }
  // ---- store: row q of the clip, features 8n + 2tq, +1
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (q0 < T)
      *reinterpret_cast<uint32_t*>(out_clip + static_cast<size_t>(q0) * ld_out + n * 8 + 2 * tq) =
          pack_e16x2(o[n][0], o[n][1]);
    if (q1 < T)
      *reinterpret_cast<uint32_t*>(out_clip + static_cast<size_t>(q1) * ld_out + n * 8 + 2 * tq) =
          pack_e16x2(o[n][2], o[n][3]);
  }
}

// NT = number of 8-key tiles (T <= 8*NT); rows padded to TP = 8*NT (multiple of 16).
template <int NT>
__global__ void __launch_bounds__(ATTN_WARPS * 32)
attention_kernel(const AttnParams p) {
  constexpr int D = 64;
  constexpr int TP = 8 * NT;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * ATTN_WARPS + warp;    // (clip, head) index
  uint8_t* sQ = attn_smem + warp * (3 * TP * 128);
  uint8_t* sK = sQ + TP * 128;
  uint8_t* sV = sK + TP * 128;

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data
  grid_dep_wait();
  if (pair >= p.N * p.H) return;

  const int b = pair / p.H;
  const int h = pair - b * p.H;
  const int T = p.T;
  const int ld = 3 * p.H * D;
  const int len = (p.lengths != nullptr) ? min(max(__ldg(p.lengths + b), 0), T) : T;

  // stage Q | K | V rows of this head (enc16_t, 128 B per row, 16-B chunk c stored at c ^ (row & 7)); pad rows are zero
  const enc16_t* base = p.qkv + static_cast<size_t>(b) * T * ld + h * D;
  for (int i = lane; i < 3 * TP * 8; i += 32) {
    const int c = i & 7;
    const int row = (i >> 3) % TP;
    const int which = (i >> 3) / TP;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (row < T)
      u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(row) * ld + which * p.H * D) + c);
    *reinterpret_cast<uint4*>(sQ + which * (TP * 128) + row * 128 + ((c ^ (row & 7)) << 4)) = u;
  }
  __syncwarp();

  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  const int mt_count = (T + 15) >> 4;
  enc16_t* out_clip = p.out + static_cast<size_t>(b) * T * (p.H * D) + h * D;
  float* probs_clip = p.probs != nullptr ? p.probs + (static_cast<size_t>(h) * p.N + b) * T * T : nullptr;
  for (int mt = 0; mt < mt_count; ++mt)
    attention_mtile<NT>(sQ_u, sK_u, sV_u, mt * 16, 0, mt * 16, T, len, p.scale, lane, out_clip, p.H * D, probs_clip);
}

}  // namespace sblk
