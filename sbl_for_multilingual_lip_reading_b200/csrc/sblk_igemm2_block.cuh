// sblk_igemm2_block.cuh — a whole BasicBlock of ResNet layer 3 (256 channels, 6x6 maps) in ONE launch of the CTA-pair
// TMA-im2col implicit GEMM: conv1 (+BN+ReLU) -> conv2 (+BN, + downsample branch as K-extension or + residual, ReLU).
// Reference: BasicBlock.forward, SBL/transformer/video_frontend.py:28-41, downsample :68-72.
//
// Why: with 256-wide pair tiles the two convs of a layer-3 block are launches of 131 pair tiles on 74 pairs: 19 us of
// k-loop and ~8 us of launch boundary each (the next kernel cannot touch its input before the last CTA of the previous
// one has retired and its own pipeline has refilled).  But conv2 of an output pixel only reads conv1 outputs of the
// SAME FRAME (zero padding is per frame), so a pair tile made of whole frames — 7 frames x 36 pixels = 252 of the 256
// rows — can go through conv1 AND conv2 with no dependency on any other tile: no grid-wide synchronisation, no second
// launch.  Every CTA pair walks its tiles twice,
//     conv1(tile a), conv1(tile b), ... , conv2(tile a), conv2(tile b), ...
// so the epilogue of conv1(a) (bias, ReLU, bf16 y1 rows to global memory, fence, pair-wide mbarrier) is long done when
// the producer warp asks for conv2(a)'s im2col tiles of y1, and the TMEM accumulator ring keeps the tensor pipe busy
// across the phase change.  Arithmetic per output element is exactly that of the two separate launches (same k order,
// same roundings): bit-identical.
//
// Everything else — roles (warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, 8 epilogue warps), barrier protocol,
// staged coalesced stores, PDL — is sblk_igemm2.cuh's igemm2_kernel<256, true>.
#pragma once
#include "sblk_common.cuh"
#include "sblk_igemm2.cuh"

namespace sblk {

struct BlockConvParams {
  int M;               // F * P * Q output rows of both convs
  int P, Q;            // output map of both convs
  int rows_per_tile;   // whole frames per pair tile: (256 / (P*Q)) * P*Q
  int num_tiles;       // ceil(M / rows_per_tile)
  int c1_cblocks;      // Cin / 64 of conv1 (3x3, pad 1, stride c1_stride over x)
  int c1_stride;
  int N;               // output channels of both convs: 256 (layer 3) or 512 (layer 4)
  int n_tiles;         // N / 256 column tiles per frame group; work unit = (frame-group tile, column tile)
  int c2_cblocks;      // N / 64 of conv2 (3x3, pad 1, stride 1 over y1)
  int ext_cblocks;     // k-blocks of the 1x1 / stride ext_stride downsample branch over x behind conv2's own (0 = none)
  int ext_stride;
  const float* bias1;              // [256] folded bn1 shift
  const float* bias2;              // [256] folded bn2 shift (+ folded downsample shift when ext_cblocks > 0)
  const __nv_bfloat16* residual;   // [M, N] identity residual of conv2 (blocks without a downsample branch) or nullptr
  __nv_bfloat16* y1;               // [M, N] workspace: relu(bn1(conv1 x))
  __nv_bfloat16* out;              // [M, N]
  // n_tiles > 1: conv2 of a unit needs conv1 of ALL column tiles of its frame group, computed by neighbouring CTA pairs
  // (unit indices are adjacent, so they run at the same time).  flags[tile] counts the epilogue warps that have published
  // their y1 rows (n_tiles x 2 CTAs x 8 warps), flags[num_tiles + tile] the producers that have seen it; the last one
  // resets both, so the counters are zero again when the kernel ends (zero-initialised once by the caller).
  unsigned int* flags;
};

constexpr int BLK_MAX_TILES = 32;   // units one CTA pair may own (y1-ready barriers); more -> the caller launches conv by conv

__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Igemm2Cfg<256>::THREADS, 1)
igemm2_block_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB3,
                    const BlockConvParams p) {
  using Cfg = Igemm2Cfg<256>;
  constexpr int BLOCK_N = 256;
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t IDESC = make_idesc_bf16(Cfg::PAIR_M, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar[Cfg::ACC_STAGES];
  __shared__ uint64_t tempty_bar[Cfg::ACC_STAGES];
  __shared__ uint64_t y1_ready[BLK_MAX_TILES];   // both CTAs: conv1 output rows of my j-th tile are in global memory
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[BLOCK_N];   // bias1, replaced by bias2 at the phase change (227 KB budget)

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_units = p.num_tiles * p.n_tiles;     // unit = (frame-group tile, column tile), column tile fastest
  const int n_my = pair_id < num_units ? (num_units - pair_id + num_pairs - 1) / num_pairs : 0;   // my units
  const int n_items = 2 * n_my;                      // item i: phase i / n_my (0 conv1, 1 conv2), my unit i % n_my
  const int nkb1 = 9 * p.c1_cblocks;
  const int nkb2_main = 9 * p.c2_cblocks;
  const int nkb2 = nkb2_main + p.ext_cblocks;
  const int nstages = STAGES - 1;                    // the last ring stage holds the epilogue's staging tiles
  const int pq = p.P * p.Q;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
    if (p.ext_cblocks > 0) {
      tma_prefetch_desc(&tmA3);
      tma_prefetch_desc(&tmB3);
    }
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
#pragma unroll
    for (int i = 0; i < Cfg::ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * Cfg::EPI_WARPS);
    }
#pragma unroll
    for (int i = 0; i < BLK_MAX_TILES; ++i) mbar_init(&y1_ready[i], 2 * Cfg::EPI_WARPS);   // epilogue warps of both CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(&tmem_base_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);

  grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs): own A rows, own half of the B tile
    int stage = 0;
    uint32_t phase = 0;
    grid_dep_wait();
    for (int it = 0; it < n_items; ++it) {
      const int ph2 = it >= n_my ? 1 : 0;
      const int j = it - ph2 * n_my;
      const int unit = pair_id + j * num_pairs;
      const int tile = unit / p.n_tiles;
      const int n0 = (unit - tile * p.n_tiles) * BLOCK_N + static_cast<int>(rank) * Cfg::BH;
      const int m0 = tile * p.rows_per_tile + static_cast<int>(rank) * Cfg::BLOCK_M;
      const int img = m0 / pq;
      const int rem = m0 - img * pq;
      const int oh = rem / p.Q;
      const int ow = rem - oh * p.Q;
      if (ph2 && p.n_tiles > 1) {
        // conv1's rows of this frame group — all column tiles, i.e. the epilogue warps of n_tiles CTA pairs — are in
        // global memory once the group's counter has reached n_tiles x 2 CTAs x 8 warps
        if (lane == 0) {
          const unsigned int target = static_cast<unsigned int>(p.n_tiles * 2 * Cfg::EPI_WARPS);
          unsigned int seen, polls = 0;
          unsigned long long t0 = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.flags + tile) : "memory");
            if (seen >= target) break;
            if ((++polls & 255u) == 0u) {
              unsigned long long now;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
              if (t0 == 0) t0 = now;
              if (now - t0 > SBLK_WATCHDOG_NS) {
                unsigned int* wd = g_sblk_watchdog_ptr;
                if (wd != nullptr) {
                  atomicCAS_system(wd, 0u, 0x0a07u | 0x80000000u);
                  __threadfence_system();
                }
                __trap();
              }
            }
          } while (true);
          // every producer of the group (n_tiles pairs x 2 CTAs) passes here exactly once: the last one resets the counters
          if (atomicAdd(p.flags + p.num_tiles + tile, 1u) == static_cast<unsigned int>(2 * p.n_tiles - 1)) {
            p.flags[tile] = 0u;
            p.flags[p.num_tiles + tile] = 0u;
          }
        }
        __syncwarp();
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
      } else if (ph2) {
        // conv1's rows of this tile (written by the epilogue warps of BOTH CTAs) are in global memory
        mbar_wait(&y1_ready[j], 0, 0x0a01);
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy global writes -> visible to the TMA reads below
      }
      const int nkb = ph2 ? nkb2 : nkb1;
      const int cblocks = ph2 ? p.c2_cblocks : p.c1_cblocks;
      const int stride = ph2 ? 1 : p.c1_stride;
      const int base_w = ow * stride - 1, base_h = oh * stride - 1;
      int cb = 0, r = 0, s = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, 0x0a02);
        uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* b_dst = a_dst + Cfg::A_BYTES;
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * (Cfg::A_BYTES + Cfg::B_BYTES));
          if (!ph2) {
            tma2_load_im2col_4d(a_dst, &tmA1, bar, cb * 64, base_w, base_h, img, static_cast<uint16_t>(s),
                                static_cast<uint16_t>(r));
            tma2_load_2d(b_dst, &tmB1, bar, kb * 64, n0);
          } else if (kb < nkb2_main) {
            tma2_load_im2col_4d(a_dst, &tmA2, bar, cb * 64, base_w, base_h, img, static_cast<uint16_t>(s),
                                static_cast<uint16_t>(r));
            tma2_load_2d(b_dst, &tmB2, bar, kb * 64, n0);
          } else {
            tma2_load_im2col_4d(a_dst, &tmA3, bar, (kb - nkb2_main) * 64, ow * p.ext_stride, oh * p.ext_stride, img, 0, 0);
            tma2_load_2d(b_dst, &tmB3, bar, (kb - nkb2_main) * 64, n0);
          }
        }
        __syncwarp();
        if (++cb == cblocks) {
          cb = 0;
          if (++s == 3) { s = 0; ++r; }
        }
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
    for (int i = 0; i < nstages; ++i) {   // drain (see sblk_igemm2.cuh)
      mbar_wait(&empty_bar[stage], phase ^ 1u, 0x0a03);
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: leader CTA only
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0; it < n_items; ++it) {
        const int nkb = it >= n_my ? nkb2 : nkb1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 0x0a04);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 0x0a05);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = make_desc_sw128(a_addr);
          const uint64_t db = make_desc_sw128(a_addr + Cfg::A_BYTES);
          const uint32_t da_lo = static_cast<uint32_t>(da), db_lo = static_cast<uint32_t>(db);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < Cfg::BLOCK_K / 16; ++k)
              umma2_bf16(d_tmem, desc_with_lo(da, da_lo + 2 * k), desc_with_lo(db, db_lo + 2 * k), IDESC,
                         (kb > 0 || k > 0) ? 1u : 0u);
            umma2_commit_mc(&empty_bar[stage]);
            if (kb == nkb - 1) umma2_commit_mc(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4), both CTAs
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int row = quarter * 32 + lane;
    constexpr int NCH = BLOCK_N / 32 / 2;          // 32-column chunks per warp
    const int c_begin = half * NCH;
    uint8_t* const stg = smem + Cfg::OFF_STG + ew * Cfg::STG_WARP_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_bias = -1;
    grid_dep_wait();
    for (int it = 0; it < n_items; ++it) {
      const int ph2 = it >= n_my ? 1 : 0;
      const int j = it - ph2 * n_my;
      const int unit = pair_id + j * num_pairs;
      const int tile = unit / p.n_tiles;
      const int n_blk = unit - tile * p.n_tiles;
      const int row_in_tile = static_cast<int>(rank) * Cfg::BLOCK_M + row;
      const int m = tile * p.rows_per_tile + row_in_tile;
      const bool row_ok = row_in_tile < p.rows_per_tile && m < p.M;   // rows past the tile's whole frames belong to the next tile
      const int ncol0 = n_blk * BLOCK_N + c_begin * 32;   // first output channel of this warp
      const int bkey = ph2 * p.n_tiles + n_blk;
      if (bkey != cur_bias) {   // phase / column-tile change (block-uniform): the eight epilogue warps swap the bias slice
        asm volatile("bar.sync 1, 256;" ::: "memory");
        bias_s[threadIdx.x - 64] = __ldg((ph2 ? p.bias2 : p.bias1) + n_blk * BLOCK_N + (threadIdx.x - 64));
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_bias = bkey;
      }
      const float* bs = bias_s;
      uint4 res[NCH * 4];
      if (ph2 && p.residual != nullptr && row_ok) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(m) * p.N + ncol0);
#pragma unroll
        for (int q = 0; q < NCH * 4; ++q) res[q] = __ldg(rp + q);
      } else {
#pragma unroll
        for (int q = 0; q < NCH * 4; ++q) res[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&tfull_bar[acc], acc_phase, 0x0a06);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * BLOCK_N + c_begin * 32);
      __nv_bfloat16* const dst = ph2 ? p.out : p.y1;
      const unsigned long long row_off_b = static_cast<unsigned long long>(m) * p.N * 2ull;
#pragma unroll
      for (int c2 = 0; c2 < NCH / 2; ++c2) {
        uint32_t v2[2][32];
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c2 * 64), v2[0]);
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c2 * 64 + 32), v2[1]);
        tmem_ld_wait();
        uint4 o[8];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cl = (c_begin + c2 * 2 + hh) * 32;   // column inside the 256-wide tile
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = *reinterpret_cast<const float4*>(&bs[cl + 8 * q]);
            const float4 b1 = *reinterpret_cast<const float4*>(&bs[cl + 8 * q + 4]);
            const uint4 r4 = res[(c2 * 2 + hh) * 4 + q];
            float f[8];
            f[0] = __uint_as_float(v2[hh][8 * q + 0]) + b0.x + bf16_lo(r4.x);
            f[1] = __uint_as_float(v2[hh][8 * q + 1]) + b0.y + bf16_hi(r4.x);
            f[2] = __uint_as_float(v2[hh][8 * q + 2]) + b0.z + bf16_lo(r4.y);
            f[3] = __uint_as_float(v2[hh][8 * q + 3]) + b0.w + bf16_hi(r4.y);
            f[4] = __uint_as_float(v2[hh][8 * q + 4]) + b1.x + bf16_lo(r4.z);
            f[5] = __uint_as_float(v2[hh][8 * q + 5]) + b1.y + bf16_hi(r4.z);
            f[6] = __uint_as_float(v2[hh][8 * q + 6]) + b1.z + bf16_lo(r4.w);
            f[7] = __uint_as_float(v2[hh][8 * q + 7]) + b1.w + bf16_hi(r4.w);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);   // both convs of a block end in ReLU
            o[hh * 4 + q].x = pack_bf16x2(f[0], f[1]);
            o[hh * 4 + q].y = pack_bf16x2(f[2], f[3]);
            o[hh * 4 + q].z = pack_bf16x2(f[4], f[5]);
            o[hh * 4 + q].w = pack_bf16x2(f[6], f[7]);
          }
        }
        warp_store_rows128_idx(stg, o, reinterpret_cast<uint8_t*>(dst + ncol0 + c2 * 64), row_off_b, row_ok, lane);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // the leader's barrier
      if (!ph2) {
        // publish this warp's y1 rows: gpu-scope fence, cross-proxy fence (they are read back by TMA), then one arrive per
        // warp on the y1-ready barrier of BOTH CTAs of the pair (a frame may straddle the two CTAs' row ranges)
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (p.n_tiles > 1) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + tile) : "memory");
          } else {
            mbar_arrive_release_cluster(mapa_u32(smem_u32(&y1_ready[j]), 0));
            mbar_arrive_release_cluster(mapa_u32(smem_u32(&y1_ready[j]), 1));
          }
        }
      }
      if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace sblk
