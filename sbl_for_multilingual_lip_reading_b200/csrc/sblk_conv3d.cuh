// sblk_conv3d.cuh — fused visual frontend stem for sm_100a:
//   Conv3d(1->64, k=(5,7,7), s=(1,2,2), p=(2,3,3), no bias) + BatchNorm3d(eval, folded) + ReLU
//   + MaxPool3d((1,3,3), s=(1,2,2), p=(0,1,1)), written as per-frame NHWC bf16 [F,22,22,64].
// Reference: Lipreading.frontend3D + the transpose/contiguous/view that follows it,
//            SBL/transformer/video_frontend.py:99-104,111-115.
//
// Implicit GEMM on tcgen05: M = conv output pixels, N = 64 channels, K = (dt, r, s8) = 5*7*8 = 280 -> 288.
// Cin = 1, so TMA im2col cannot build the A operand.  Instead one CTA owns a half frame (22/23 conv rows):
//   * loader thread: ONE 3-D TMA brings the bf16 input patch (5 frames x 51 rows x 96 cols) into smem
//   * 8 producer warps expand it into SWIZZLE_128B A tiles (128 pixels x 64 K per ring stage); each 16-B
//     chunk of a row is 8 consecutive input pixels of one (dt, r) filter row — the 8th multiplies a zero weight
//   * 1 MMA thread: 18 x tcgen05.mma (M128 N64 K16) per 128-pixel tile into a 2-deep TMEM ring
//   * 4 epilogue warps: TMEM -> +bias, ReLU -> bf16 -> swizzled smem ring of conv pixels, then the 3x3/s2
//     max-pool is taken straight out of that ring and stored coalesced.  The 7.2 MB/clip un-pooled
//     activation never reaches HBM.
#pragma once
#include "sblk_common.cuh"

namespace sblk {

namespace c3d {
constexpr int IN_HW = 88;
constexpr int HP = 94;          // padded rows of the prepped frame (3 + 88 + 3)
constexpr int WP = 96;          // padded cols (3 + 88 + 5)
constexpr int TPAD = 2;         // zero frames before/after each clip
constexpr int CONV_HW = 44;
constexpr int POOL_HW = 22;
constexpr int COUT = 64;
constexpr int KPAD = 320;       // packed weight row length (5 K-blocks of 64); taps live in [0, 288)
constexpr int PATCH_ROWS = 51;
constexpr int PATCH_FRAMES = 5;
constexpr int PATCH_BYTES = PATCH_FRAMES * PATCH_ROWS * WP * 2;  // 48960
constexpr int A_STAGES = 4;
constexpr int A_STAGE_BYTES = 128 * 128;                          // 128 pixels x 64 bf16
constexpr int B_BYTES = 5 * COUT * 128;                           // 40960
constexpr int RING_PIX = 512;
constexpr int RING_BYTES = RING_PIX * 128;
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_STAGES * A_STAGE_BYTES;           // 65536
constexpr int OFF_RING = OFF_B + B_BYTES;                         // 106496
constexpr int OFF_PATCH = OFF_RING + RING_BYTES;                  // 172032
constexpr int SMEM_BYTES = OFF_PATCH + PATCH_BYTES + 1024;        // 222016
constexpr int TILES_PER_UNIT = 8;
constexpr int NUM_PRODUCERS = 256;
constexpr int THREADS = 32 * 14;  // loader, mma, 4 epilogue, 8 producer warps
constexpr int TMEM_COLS = 128;    // 2 accumulator stages x 64 columns
}  // namespace c3d

struct Conv3dParams {
  int frames;                 // F = N*T
  int T;                      // frames per clip
  const float* bias;          // [64] folded BN shift
  __nv_bfloat16* out;         // [F,22,22,64]
};

__global__ void __launch_bounds__(c3d::THREADS, 1)
conv3d_bn_relu_pool_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                           const Conv3dParams p) {
  using namespace c3d;
  constexpr uint32_t IDESC = make_idesc_bf16(128, COUT);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[A_STAGES];
  __shared__ uint64_t empty_bar[A_STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint64_t patch_full_bar;
  __shared__ uint64_t patch_free_bar;
  __shared__ uint64_t weights_bar;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = p.frames * 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
#pragma unroll
    for (int i = 0; i < A_STAGES; ++i) {
      mbar_init(&full_bar[i], NUM_PRODUCERS);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 4);
    mbar_init(&tempty_bar[1], 4);
    mbar_init(&patch_full_bar, 1);
    mbar_init(&patch_free_bar, NUM_PRODUCERS);
    mbar_init(&weights_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  grid_dep_wait();

  if (warp == 0) {
    // ------------------------------------------------ loader: weights once, then one patch per unit
    if (lane == 0) {
      mbar_arrive_expect_tx(&weights_bar, B_BYTES);
#pragma unroll
      for (int j = 0; j < 5; ++j) tma_load_2d(smem + OFF_B + j * (COUT * 128), &tmW, &weights_bar, j * 64, 0);
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int f = u >> 1;
        const int half = u & 1;
        const int n = f / p.T;
        const int t = f - n * p.T;
        mbar_wait(&patch_free_bar, phase ^ 1u, 0x0201);
        mbar_arrive_expect_tx(&patch_full_bar, PATCH_BYTES);
        // padded-time frames t .. t+4 hold real frames t-2 .. t+2; rows start at 0 (half 0) or 42 (half 1)
        tma_load_3d(smem + OFF_PATCH, &tmX, &patch_full_bar, 0, half ? 42 : 0, n * (p.T + 2 * TPAD) + t);
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(&weights_bar, 0, 0x0202);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t db0 = make_desc_sw128(smem_base + OFF_B);
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        for (int tile = 0; tile < TILES_PER_UNIT; ++tile) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 0x0203);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * COUT);
#pragma unroll 1
          for (int j = 0; j < 5; ++j) {
            mbar_wait(&full_bar[stage], phase, 0x0204);
            tc_fence_after_sync();
            const uint64_t da = make_desc_sw128(smem_base + OFF_A + stage * A_STAGE_BYTES);
            const uint64_t db = db0 + static_cast<uint64_t>((j * COUT * 128) >> 4);
            const int nk = (j < 4) ? 4 : 2;  // K = 288 = 4*64 + 32
            for (int k = 0; k < nk; ++k) {
              umma_bf16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), IDESC,
                        (j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&tfull_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------ epilogue + fused max-pool (warps 2..5)
    const int quarter = warp & 3;
    const int ew = warp - 2;           // 0..3, pooling work split
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    float bias_r[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) bias_r[j] = __ldg(p.bias + j);
    uint8_t* ring = smem + OFF_RING;

    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int f = u >> 1;
      const int half = u & 1;
      const int y_base = half ? 21 : 0;
      const int nrows = half ? 23 : 22;
      const int py_end = half ? 22 : 11;
      int py_next = half ? 11 : 0;
      for (int tile = 0; tile < TILES_PER_UNIT; ++tile) {
        mbar_wait(&tfull_bar[acc], acc_phase, 0x0205);
        tc_fence_after_sync();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * COUT);
        const int pm = tile * 128 + row;
        const int slot = pm & (RING_PIX - 1);
        uint8_t* dst_row = ring + slot * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float g[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              g[e] = fmaxf(__uint_as_float(v[q * 8 + e]) + bias_r[c * 32 + q * 8 + e], 0.0f);
            uint4 o;
            o.x = pack_bf16x2(g[0], g[1]);
            o.y = pack_bf16x2(g[2], g[3]);
            o.z = pack_bf16x2(g[4], g[5]);
            o.w = pack_bf16x2(g[6], g[7]);
            const int chunk = c * 4 + q;
            *reinterpret_cast<uint4*>(dst_row + ((chunk ^ (slot & 7)) << 4)) = o;
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }

        // conv pixels of this tile are now in the ring; pool every pooled row that just completed
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int rows_done = (tile == TILES_PER_UNIT - 1) ? nrows : min(nrows, (128 * (tile + 1)) / CONV_HW);
        while (py_next < py_end) {
          const int gy_lo = max(2 * py_next - 1, 0);
          const int gy_hi = min(2 * py_next + 1, CONV_HW - 1);
          if (gy_hi - y_base >= rows_done) break;
          for (int px = ew; px < POOL_HW; px += 4) {
            const int gx_lo = max(2 * px - 1, 0);
            const int gx_hi = min(2 * px + 1, CONV_HW - 1);
            __nv_bfloat162 best = __floats2bfloat162_rn(0.0f, 0.0f);  // post-ReLU values are >= 0
            for (int gy = gy_lo; gy <= gy_hi; ++gy) {
              for (int gx = gx_lo; gx <= gx_hi; ++gx) {
                const int s2 = ((gy - y_base) * CONV_HW + gx) & (RING_PIX - 1);
                const uint32_t w = *reinterpret_cast<const uint32_t*>(
                    ring + s2 * 128 + (((lane >> 2) ^ (s2 & 7)) << 4) + ((lane & 3) << 2));
                best = __hmax2(best, *reinterpret_cast<const __nv_bfloat162*>(&w));
              }
            }
            __nv_bfloat16* op = p.out + ((static_cast<size_t>(f) * POOL_HW + py_next) * POOL_HW + px) * COUT;
            reinterpret_cast<__nv_bfloat162*>(op)[lane] = best;
          }
          ++py_next;
        }
      }
    }
  } else {
    // ------------------------------------------------ A-operand producers (warps 6..13, 256 threads)
    const int ptid = threadIdx.x - 6 * 32;
    const int m = ptid & 127;
    const int hsel = ptid >> 7;
    const uint8_t* patch = smem + OFF_PATCH;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t patch_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      mbar_wait(&patch_full_bar, patch_phase, 0x0206);
      patch_phase ^= 1u;
      for (int tile = 0; tile < TILES_PER_UNIT; ++tile) {
        const int pm = tile * 128 + m;
        int yl = pm / CONV_HW;
        const int x = pm - yl * CONV_HW;
        yl = min(yl, 22);  // rows past the unit are never pooled; keep their reads inside the patch
        const uint8_t* src_px = patch + ((2 * yl) * WP + 2 * x) * 2;
        uint8_t* dst_row = smem + OFF_A + m * 128;
        const int sw = m & 7;
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 0x0207);
          uint8_t* dst = dst_row + stage * A_STAGE_BYTES;
          if (j < 4 || hsel == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int cc = hsel * 4 + q;   // chunk inside this 64-wide K block
              const int c = j * 8 + cc;      // global chunk = dt*7 + r
              uint4 val = make_uint4(0u, 0u, 0u, 0u);
              if (c < 35) {
                const int dt = c / 7;
                const int r = c - dt * 7;
                const uint32_t* sp =
                    reinterpret_cast<const uint32_t*>(src_px + ((dt * PATCH_ROWS + r) * WP) * 2);
                val.x = sp[0]; val.y = sp[1]; val.z = sp[2]; val.w = sp[3];
              }
              *reinterpret_cast<uint4*>(dst + ((cc ^ sw) << 4)) = val;
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[stage]);
          if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      mbar_arrive(&patch_free_bar);  // this thread no longer reads the patch
    }
  }

  grid_dep_launch();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace sblk
