// sblk_gemm_ln.cuh — Linear (d_model = 512 outputs) + bias + residual + LayerNorm (+ positional encoding, + pad mask)
// in ONE kernel: y = LN(A W^T + b + residual) * gamma + beta (+ pe[t]) (* keep[b, t]).
// Reference call sites (SBL/transformer): attention.py:57-58 (fc -> dropout(identity) -> layer_norm(out + residual)),
// module.py:49-51 (w_2 -> layer_norm(out + x)), encoder.py:53-55 (layer_norm_in(linear_in(x)) + PE) and the
// `*= non_pad_mask` of encoder.py:86,89.
//
// At the BASELINE batch (928 tokens) these GEMMs are latency-bound, and the separate LayerNorm launch that followed
// each of them cost as much as the GEMM.  LayerNorm needs whole 512-wide rows, a single CTA owning 128 x 512 outputs
// would leave 140 SMs idle, so a CLUSTER of 4 CTAs owns one 128-row tile: CTA r computes columns [128 r, 128 r + 128)
// with its own TMA / tcgen05 pipeline (bf16 operands, fp32 accumulate in TMEM), and the row statistics are combined
// through distributed shared memory: every CTA writes its per-row partial sums into all four CTAs' smem, a cluster
// barrier publishes them.  Statistics are two-pass (mean, then sum of squared deviations) in fp32 like the unfused
// kernel; the pre-LN values v = acc + bias + residual are parked in TMEM (tcgen05.st) between the passes.
#pragma once
#include "sblk_common.cuh"
#include "sblk_igemm2.cuh"

namespace sblk {

struct GemmLnParams {
  int M;                  // rows (tokens)
  int K;                  // reduction length (multiple of 64)
  int T;                  // frames per clip (pe row / pad mask index t = m % T)
  const float* bias;      // [512] or nullptr
  const float* residual;  // [M, 512] fp32 or nullptr
  const float* gamma;     // [512]
  const float* beta;      // [512]
  const float* pe;        // [>= T, 512] or nullptr
  const int* lengths;     // [M / T] or nullptr
  float* out_f32;         // [M, 512] or nullptr
  enc16_t* out_bf16;  // [M, 512] or nullptr
  float eps;
};

namespace gln {
constexpr int D = 512;
constexpr int CLUSTER = 4;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = D / CLUSTER;   // 128
constexpr int A_BYTES = BLOCK_M * 128;
constexpr int B_BYTES = BLOCK_N * 128;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
constexpr int STAGES = 6;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 128;
}  // namespace gln

__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__global__ void __cluster_dims__(gln::CLUSTER, 1, 1) __launch_bounds__(gln::THREADS, 1)
gemm_ln512_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmLnParams p) {
  using namespace gln;
  constexpr uint32_t IDESC = make_idesc_e16(BLOCK_M, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float part_sum[CLUSTER][BLOCK_M];   // [source CTA][row]: written by every CTA of the cluster (DSMEM)
  __shared__ float part_sq[CLUSTER][BLOCK_M];

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = static_cast<int>(blockIdx.x / CLUSTER) * BLOCK_M;
  const int n0 = static_cast<int>(rank) * BLOCK_N;
  const int num_kb = p.K / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);
  // split cluster barrier: the matching wait sits in front of the first distributed-shared-memory store, so every
  // CTA of the cluster is known to be running (its shared memory live) before a peer writes into it
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data
  grid_dep_wait();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1u, 0x0501);
      uint8_t* a_dst = smem + stage * STAGE_BYTES;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * 64, m0);
        tma_load_2d(a_dst + A_BYTES, &tmB, &full_bar[stage], kb * 64, n0);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase, 0x0502);
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
      const uint64_t da = make_desc_sw128(a_addr);
      const uint64_t db = make_desc_sw128(a_addr + A_BYTES);
      const uint32_t da_lo = static_cast<uint32_t>(da), db_lo = static_cast<uint32_t>(db);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, desc_with_lo(da, da_lo + 2 * k), desc_with_lo(db, db_lo + 2 * k), IDESC,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(&tfull_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  }

  // ------------------------------------------------ epilogue: warps 2-5 own TMEM lane quarter (warp % 4); every
  // thread of the cluster takes part in the two cluster barriers.
  const bool epi = warp >= 2;
  const int quarter = warp & 3;
  const int row = quarter * 32 + lane;
  const int m = m0 + row;
  const bool row_ok = epi && m < p.M;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  float mean = 0.0f, rstd = 0.0f;

  __syncwarp();
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (epi) {
    mbar_wait(&tfull_bar, 0, 0x0503);
    tc_fence_after_sync();
    // pass 1: v = acc + bias + residual -> TMEM ; partial row sum
    float s = 0.0f;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
      const int col = n0 + c * 32;
      if (p.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(bp + j);
          v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
          v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
          v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
        }
      }
      if (p.residual != nullptr && row_ok) {
        const float4* rp = reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(m) * D + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r4 = __ldg(rp + j);
          v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + r4.x);
          v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + r4.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + r4.z);
          v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + r4.w);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) s += __uint_as_float(v[j]);
      tmem_st_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), v);
    }
    tmem_st_wait();
    const uint32_t slot = smem_u32(&part_sum[rank][row]);
#pragma unroll
    for (uint32_t r = 0; r < CLUSTER; ++r) st_cluster_f32(mapa_u32(slot, r), s);
  }
  __syncwarp();
  cluster_sync_all();
  if (epi) {
    mean = (part_sum[0][row] + part_sum[1][row] + part_sum[2][row] + part_sum[3][row]) * (1.0f / D);
    // pass 2: partial sum of squared deviations
    float q = 0.0f;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = __uint_as_float(v[j]) - mean;
        q += d * d;
      }
    }
    const uint32_t slot = smem_u32(&part_sq[rank][row]);
#pragma unroll
    for (uint32_t r = 0; r < CLUSTER; ++r) st_cluster_f32(mapa_u32(slot, r), q);
  }
  __syncwarp();
  cluster_sync_all();
  if (epi) {
    rstd = rsqrtf((part_sq[0][row] + part_sq[1][row] + part_sq[2][row] + part_sq[3][row]) * (1.0f / D) + p.eps);
    // (tcgen05.ld is warp-collective: every lane runs the loop, only the stores are predicated on row_ok)
    const int t = m % p.T;
    float keep = 1.0f;
    if (row_ok && p.lengths != nullptr && t >= __ldg(p.lengths + m / p.T)) keep = 0.0f;
    // pass 3: normalise, affine, + PE, pad mask, store
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
      if (row_ok) {
        const int col = n0 + c * 32;
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col) + j);
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + col) + j);
          y[4 * j + 0] = (__uint_as_float(v[4 * j + 0]) - mean) * rstd * g.x + b.x;
          y[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * g.y + b.y;
          y[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * g.z + b.z;
          y[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * g.w + b.w;
        }
        if (p.pe != nullptr) {
          const float4* ep = reinterpret_cast<const float4*>(p.pe + static_cast<size_t>(t) * D + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 e = __ldg(ep + j);
            y[4 * j + 0] += e.x; y[4 * j + 1] += e.y; y[4 * j + 2] += e.z; y[4 * j + 3] += e.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] *= keep;
        if (p.out_f32 != nullptr) {
          float4* op = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(m) * D + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        }
        if (p.out_bf16 != nullptr) {
          uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(m) * D + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_e16x2(y[8 * j + 0], y[8 * j + 1]);
            o.y = pack_e16x2(y[8 * j + 2], y[8 * j + 3]);
            o.z = pack_e16x2(y[8 * j + 4], y[8 * j + 5]);
            o.w = pack_e16x2(y[8 * j + 6], y[8 * j + 7]);
            op[j] = o;
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace sblk
