// sblk_flatconv.cuh — stride-1 3x3 convolution (64 -> 64 channels) as a *flat shifted-window* implicit GEMM.
// Reference: BasicBlock conv1/bn1/relu and conv2/bn2/+=residual/relu of ResNet layer1,
//            SBL/transformer/video_frontend.py:10-12,28-41 (layer1: 4 convs, 21 % of the path's FLOPs).
//
// Why not the im2col kernel (sblk_igemm.cuh): with N = 64 a 128x64 tile re-fetches 9 shifted copies of its
// activations plus the whole 72 KB filter from L2 (216 KB per tile, measured ~9 TB/s L2->SM = the fabric limit,
// tensor pipe 20 % busy).  Here activations live in HBM in a zero-haloed flat layout
//     pixel(f, y, x) -> row  m = (f*(H+1) + 1 + y) * (W+2) + 1 + x   of a [M_total, 64] bf16 matrix
// (one zero row between frames, zero columns left/right), so a 3x3 tap (r, s) of 128 consecutive output rows is
// just the SAME matrix shifted by (r-1)*(W+2) + (s-1) rows.  A CTA stages each run of pixels ONCE (TMA, SWIZZLE_128B)
// and issues the 9 taps as UMMA descriptors whose start address is shifted by whole 128-byte rows (the swizzle is a
// function of the absolute smem address, so row-shifted descriptors read the staged pixels correctly), with the
// filter resident in shared memory.  L2->SM traffic drops from 216 KB to ~19 KB per tile and the kernel becomes
// bound by the tensor core's smem operand fetch (48 cycles per M128xN64xK16 MMA).
// Outputs that fall on halo positions are written as zeros, so the result is again a valid flat layout.
// The residual add of conv2 is done BY THE TENSOR CORE: the residual tile (128 rows x 64 ch, same flat rows as the
// output) is TMA-loaded next to the activations and accumulated with 4 extra MMAs against a 64x64 identity that is
// packed behind the 9 filter taps (bf16 x 1.0 is exact in the fp32 accumulator), so the epilogue is identical for
// both convs: + bias, ReLU, one rounding to bf16, smem-staged so that every global store is a full 512-byte warp store.
#pragma once
#include "sblk_common.cuh"

namespace sblk {

namespace fc {
constexpr int C = 64;                    // Cin == Cout
constexpr int TILE_M = 128;
constexpr int TILES_PER_STAGE = 1;  // (stage == tile)       // small stages, many in flight: L2->SM latency (~3 us under load) is what
                                         // has to be covered, not bandwidth
constexpr int BOX_PIX = 192;             // one TMA box per stage: 192 >= 128 + 2*(Wp+1) for Wp <= 31
constexpr int STAGE_PIX = BOX_PIX;
constexpr int A_BOX_BYTES = STAGE_PIX * 128;   // 24576 (multiple of 1024)
constexpr int R_BOX_BYTES = TILE_M * 128;      // 16384: residual tile
constexpr int STAGE_BYTES = A_BOX_BYTES + R_BOX_BYTES;   // 40960
constexpr int A_STAGES = 3;
constexpr int W_TAPS = 10;               // 9 filter taps + the identity used for the residual
constexpr int W_BYTES = W_TAPS * C * 128;   // 81920: ten [64 cout][64 cin] SWIZZLE_128B tiles
constexpr int OFF_W = 0;
constexpr int OFF_A = W_BYTES;           // 80 * 1024
constexpr int OFF_STG = OFF_A + A_STAGES * STAGE_BYTES;   // epilogue staging tile: 128 rows x 128 B (bf16)
constexpr int STG_BYTES = TILE_M * 128;
constexpr int SMEM_BYTES = OFF_STG + STG_BYTES + 1024;   // 222208
constexpr int ACC_STAGES = 4;
constexpr int TMEM_COLS = ACC_STAGES * C;  // 256
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
}  // namespace fc

struct FlatConvParams {
  int m_total;       // rows of the flat activation matrix = (F*(H+1) + 1) * (W+2)
  int num_tiles;     // ceil(m_total / 128)
  int H, W;          // frame size (Wp = W + 2, Hp = H + 1)
  int relu;
  const float* bias;               // [64] folded BN shift
  int has_res;                     // residual tile is loaded through tmR and accumulated by the tensor core
  __nv_bfloat16* out;              // flat layout [m_total, 64]
  int debug_mode;                  // timing experiments only: bit0 = no A loads, bit1 = epilogue only frees TMEM
};

__global__ void __launch_bounds__(fc::THREADS, 1)
flatconv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmR, const FlatConvParams p) {
  using namespace fc;
  constexpr uint32_t IDESC = make_idesc_bf16(TILE_M, C);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[A_STAGES];
  __shared__ uint64_t empty_bar[A_STAGES];
  __shared__ uint64_t tfull_bar[ACC_STAGES];
  __shared__ uint64_t tempty_bar[ACC_STAGES];
  __shared__ uint64_t weights_bar;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int Wp = p.W + 2;

  // contiguous, balanced tile range of this CTA
  const int base_cnt = p.num_tiles / gridDim.x;
  const int rem = p.num_tiles - base_cnt * gridDim.x;
  const int my_cnt = base_cnt + (static_cast<int>(blockIdx.x) < rem ? 1 : 0);
  const int tile_begin = blockIdx.x * base_cnt + min(static_cast<int>(blockIdx.x), rem);
  const int tile_end = tile_begin + my_cnt;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmR);
#pragma unroll
    for (int i = 0; i < A_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
#pragma unroll
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    mbar_init(&weights_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data
  grid_dep_wait();

  if (warp == 0) {
    // ------------------------------------------------ loader: filter once, then one stage per tile
    if (elect_one()) {
      mbar_arrive_expect_tx(&weights_bar, W_BYTES);
#pragma unroll
      for (int t = 0; t < W_TAPS; ++t) tma_load_2d(smem + OFF_W + t * (C * 128), &tmW, &weights_bar, t * 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&empty_bar[stage], phase ^ 1u, 0x0301);
      uint8_t* dst = smem + OFF_A + stage * STAGE_BYTES;
      const int px0 = tile * TILE_M - (Wp + 1);   // may be negative / run past the end: TMA zero-fills
      if (elect_one()) {
        if (p.debug_mode & 1) {
          mbar_arrive(&full_bar[stage]);
        } else {
          mbar_arrive_expect_tx(&full_bar[stage], p.has_res ? STAGE_BYTES : A_BOX_BYTES);
          tma_load_2d(dst, &tmX, &full_bar[stage], 0, px0);
          if (p.has_res) tma_load_2d(dst + A_BOX_BYTES, &tmR, &full_bar[stage], 0, tile * TILE_M);
        }
      }
      __syncwarp();
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform control flow, one elected lane issues)
    mbar_wait(&weights_bar, 0, 0x0302);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint64_t dw0 = make_desc_sw128(smem_base + OFF_W);
    const uint32_t dw0_lo = static_cast<uint32_t>(dw0);
    uint32_t tap_off[9];   // (r*Wp + s) rows of 128 B, in descriptor units of 16 B
#pragma unroll
    for (int t = 0; t < 9; ++t) tap_off[t] = static_cast<uint32_t>(((t / 3) * Wp + (t % 3)) * 8);
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&full_bar[stage], phase, 0x0303);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 0x0304);
      tc_fence_after_sync();
      const uint32_t a_stage = smem_base + OFF_A + stage * STAGE_BYTES;
      const uint64_t da0 = make_desc_sw128(a_stage);
      const uint32_t da0_lo = static_cast<uint32_t>(da0);
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * C);
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          // tap (r, s): the staged pixel matrix shifted by r*Wp + s rows (stage starts Wp+1 rows early)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, desc_with_lo(da0, da0_lo + tap_off[t] + static_cast<uint32_t>(2 * k)),
                      desc_with_lo(dw0, dw0_lo + static_cast<uint32_t>(t * (C * 128 / 16) + 2 * k)), IDESC,
                      (t > 0 || k > 0) ? 1u : 0u);
        }
        if (p.has_res) {
          // + residual: R[128 x 64] * I[64 x 64] accumulated in fp32 (exact)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, desc_with_lo(da0, da0_lo + static_cast<uint32_t>(A_BOX_BYTES / 16 + 2 * k)),
                      desc_with_lo(dw0, dw0_lo + static_cast<uint32_t>(9 * (C * 128 / 16) + 2 * k)), IDESC, 1u);
        }
        umma_commit(&tfull_bar[acc]);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------ epilogue (8 warps)
    // Phase A (TMEM layout: thread = output row, 32 channels): + bias, ReLU, bf16 -> staging tile (XOR-swizzled rows).
    // Phase B (store layout: 8 threads = one 128-byte output row): halo mask -> coalesced 512-byte warp stores.
    // A thread-per-row store would touch 32 cache lines per instruction.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chalf = ew >> 2;
    const int etid = threadIdx.x - 64;        // 0..255
    const int Hp = p.H + 1;
    int acc = 0;
    uint32_t acc_phase = 0;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = __ldg(p.bias + chalf * 32 + j);
    uint8_t* stg = smem + OFF_STG;
    const int b_chunk = etid & 7;             // phase B: 16-byte output chunk (8 channels)
    const int b_row0 = etid >> 3;             // phase B: rows b_row0 + 32*i
    const int arow = quarter * 32 + lane;     // phase A: output row of this thread
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      if (p.debug_mode & 2) {
        mbar_wait(&tfull_bar[acc], acc_phase, 0x0306);
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      mbar_wait(&tfull_bar[acc], acc_phase, 0x0305);
      tc_fence_after_sync();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * C + chalf * 32), v);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      // ---- phase A
      {
        uint8_t* rowp = stg + arow * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = __uint_as_float(v[8 * q + e]) + bias_r[8 * q + e];
            if (p.relu) f[e] = fmaxf(f[e], 0.0f);
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
          const int cc = chalf * 4 + q;
          *reinterpret_cast<uint4*>(rowp + ((cc ^ (arow & 7)) << 4)) = o;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- phase B
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = b_row0 + 32 * i;
        const int m = tile * TILE_M + row;
        const int R = m / Wp;
        const int c = m - R * Wp;
        const bool in_buf = m < p.m_total;
        const bool valid = c >= 1 && c <= p.W && R >= 1 && ((R - 1) % Hp) < p.H;
        uint4 o = *reinterpret_cast<const uint4*>(stg + row * 128 + ((b_chunk ^ (row & 7)) << 4));
        if (!valid) o = make_uint4(0u, 0u, 0u, 0u);   // halo positions stay zero for the next conv
        if (in_buf) *(reinterpret_cast<uint4*>(p.out + static_cast<size_t>(m) * C) + b_chunk) = o;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");   // staging tile is free again
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
