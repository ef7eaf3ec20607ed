// sblk_qkv_attn.cuh — the self-attention half of an encoder layer up to (not including) the output projection, in ONE
// kernel: per (clip group, head)  [q | k | v] = x W_h^T + b_h  (tcgen05, fp32 accumulate in TMEM)  ->  bf16 tiles in
// shared memory  ->  softmax(q k^T * scale, keys >= length masked) v  ->  head slice of the concatenated output.
// Reference: MultiHeadAttention.forward up to the head merge, SBL/transformer/attention.py:41-55, and
// ScaledDotProductAttention.forward, attention.py:72-83.
//
// Why fused: at the BASELINE batch (32 clips x 29 frames) the projection GEMM and the attention kernel are both
// latency-bound launches, and the packed QKV activation made a round trip through L2 between them.  A clip's
// attention only needs that clip's own rows, so a CTA that owns the rows of G = floor(128 / T) whole clips and the
// 192 projection columns of one head (weights packed head-major: q_h | k_h | v_h) has everything in its accumulator.
// Grid = ceil(N / G) x H CTAs (8 x 8 at the BASELINE shape).
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-9 epilogue + attention:
//   phase 1: TMEM -> + bias -> bf16 -> Q / K / V smem tiles (128-byte rows, software XOR swizzle)
//   phase 2: one warp per (clip, 16-query tile): mma.sync attention straight out of those tiles (sblk_attention.cuh)
#pragma once
#include "sblk_common.cuh"
#include "sblk_attention.cuh"

namespace sblk {

struct QkvAttnParams {
  int N;                 // clips
  int T;                 // frames per clip (<= 128)
  int H;                 // heads
  int G;                 // clips per 128-row tile = max(1, 128 / T)
  int K;                 // d_model (multiple of 64)
  const float* bias;     // [H * 192] head-major (q_h | k_h | v_h)
  const int* lengths;    // [N] valid key counts or nullptr
  enc16_t* out;    // [N*T, H*64]
  float scale;           // 1 / temperature
};

namespace qa {
constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 192;
constexpr int A_BYTES = BLOCK_M * 128;          // 16 KB
constexpr int B_BYTES = BLOCK_N * 128;          // 24 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 40 KB
constexpr int STAGES = 4;
constexpr int QKV_ROWS = 144;                   // 128 tile rows + 16 zero rows (key padding of the last clip)
constexpr int QKV_BYTES = QKV_ROWS * 128;
constexpr int OFF_QKV = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_QKV + 3 * QKV_BYTES + 1024;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 256;
}  // namespace qa

template <int NT>
__global__ void __launch_bounds__(qa::THREADS, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const QkvAttnParams p) {
  using namespace qa;
  constexpr uint32_t IDESC = make_idesc_e16(BLOCK_M, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;          // clip group
  const int h = blockIdx.y;             // head
  const int T = p.T;
  const int m0 = tile * p.G * T;        // first token row of the group
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
  if (warp >= 2) {
    // zero the 16 padding rows behind each of the Q / K / V tiles (read as masked keys of the group's last clip)
    const int etid = threadIdx.x - 64;
    for (int i = etid; i < 3 * 16 * 8; i += EPI_WARPS * 32) {
      const int which = i / 128;
      const int r = (i >> 3) & 15;
      *reinterpret_cast<uint4*>(smem + OFF_QKV + which * QKV_BYTES + (BLOCK_M + r) * 128 + ((i & 7) << 4)) =
          make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data
  grid_dep_wait();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1u, 0x0601);
      uint8_t* a_dst = smem + stage * STAGE_BYTES;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * 64, m0);
        tma_load_2d(a_dst + A_BYTES, &tmB, &full_bar[stage], kb * 64, h * BLOCK_N);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase, 0x0602);
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
  } else {
    // ------------------------------------------------ epilogue + attention (8 warps)
    const int ew = warp - 2;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int chalf = ew >> 2;               // which 96 of the 192 projection columns
    const int row = quarter * 32 + lane;     // tile row (token m0 + row)
    uint8_t* sQKV = smem + OFF_QKV;
    mbar_wait(&tfull_bar, 0, 0x0603);
    tc_fence_after_sync();
    // phase 1: accumulator -> + bias -> bf16 -> Q / K / V tiles
#pragma unroll 1
    for (int c3 = 0; c3 < 3; ++c3) {
      const int cc = chalf * 3 + c3;         // 32-column chunk 0..5: q q k k v v
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(cc * 32), v);
      tmem_ld_wait();
      const float4* bp = reinterpret_cast<const float4*>(p.bias + h * BLOCK_N + cc * 32);
      uint8_t* dst_row = sQKV + (cc >> 1) * QKV_BYTES + row * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b0 = __ldg(bp + 2 * j), b1 = __ldg(bp + 2 * j + 1);
        uint4 o;
        o.x = pack_e16x2(__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y);
        o.y = pack_e16x2(__uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w);
        o.z = pack_e16x2(__uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y);
        o.w = pack_e16x2(__uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w);
        const int chunk = (cc & 1) * 4 + j;
        *reinterpret_cast<uint4*>(dst_row + ((chunk ^ (row & 7)) << 4)) = o;
      }
    }
    tc_fence_before_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // phase 2: one warp per (clip of the group, 16-query tile)
    const uint32_t sQ_u = smem_u32(sQKV), sK_u = sQ_u + QKV_BYTES, sV_u = sK_u + QKV_BYTES;
    const int mt_count = (T + 15) >> 4;
    const int units = p.G * mt_count;
    for (int u = ew; u < units; u += EPI_WARPS) {
      const int c = u / mt_count;
      const int mt = u - c * mt_count;
      const int b = tile * p.G + c;
      if (b >= p.N) continue;
      const int len = (p.lengths != nullptr) ? min(max(__ldg(p.lengths + b), 0), T) : T;
      enc16_t* out_clip = p.out + static_cast<size_t>(b) * T * (p.H * 64) + h * 64;
      attention_mtile<NT>(sQ_u, sK_u, sV_u, c * T + mt * 16, c * T, mt * 16, T, len, p.scale, lane, out_clip,
                          p.H * 64, nullptr);
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
