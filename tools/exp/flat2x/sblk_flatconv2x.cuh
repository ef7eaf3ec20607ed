// sblk_flatconv2x.cuh — stride-1 3x3 convolution 64 -> 64 over the zero-haloed flat layout with TWO output pixels per
// accumulator row: the layer-1 form of sblk_flatconv2.cuh with 27 % fewer shared-memory operand bytes per FLOP.
// Reference: BasicBlock conv1/bn1/relu and conv2/bn2/+=residual/relu of ResNet layer1 (64 ch, 22x22),
//            SBL/transformer/video_frontend.py:10-12,28-41.
//
// Why: flatconv2_kernel<1> is bound by the shared-memory fetch of its MMA operands (per-tile stamps: the tile period is
// 91 % of the fetch time of its 36 N = 64 SS-form MMAs; every M128 x N64 x K16 MMA reads 4 KB of pixels and 1 KB of
// filter for 32 cycles of math).  The pixels are the expensive operand and each fetch feeds only 64 output columns.
// Here accumulator row j of a CTA holds output pixels m0 = row0 + 2j (columns 0-63) AND m0 + 1 (columns 64-127):
//     out[m0 + q] = sum_{r,s} X[m0 + q + (r-1)*Wp + (s-1)] W[r,s]      q = 0, 1
//                 = sum_{r,s'} X[m0 + (r-1)*Wp + (s'-1)] W[r, s'-q]    s' = s + q = 0 .. 3
// so for every (r, s') ONE pixel operand — flat rows row0 + 2j + delta, delta = (r-1)*Wp + s' - 1, i.e. every other row
// of the flat matrix — multiplies the stacked filter [W[r,s'] | W[r,s'-1]] (N = 128; the two ends s' = 0 / 3 have only
// one half, N = 64 into columns 0-63 / 64-127).  Rows of one parity are staged as their own run (TMA over the
// [rows/2, 2, 64] view of the flat matrix), so "every other row, shifted by delta" is again a plain row-shifted UMMA
// descriptor: parity = delta & 1, shift = (delta - parity) / 2.  Per 256 output rows of a CTA: 6 N=64 + 6 N=128 MMAs
// per K16 step (5 and 6 KB of operands) instead of 18 N=64 MMAs — 132 KB instead of 180 KB per 128 output rows, and the
// N = 128 MMAs run at their math floor.
//
// Bit-identical to flatconv2_kernel<1>: every output element accumulates its taps in the same (r, s, k16) order — s'
// ascends, so columns 0-63 see s = 0,1,2 and columns 64-127 see s = 0,1,2 as well.  The one MMA that would have to
// initialise columns 64-127 while accumulating into columns 0-63 (r = 0, s' = 1, first K16 step) is issued as two
// N = 64 MMAs (hence the extra W[0,1] filter slot split like the single-tap slots).
//
// Filter slots per CTA (cta_group::2: CTA `rank` supplies B rows [rank*N/2, (rank+1)*N/2) of every MMA):
//   (r, s'=0): W[r,0] rows rank*32..+32            (r, s'=3): W[r,2] rows rank*32..+32         4 KB each
//   (r, s'=1): rank 0: W[r,1], rank 1: W[r,0]      (r, s'=2): rank 0: W[r,2], rank 1: W[r,1]   8 KB each (all 64 rows)
//   extra    : W[0,1] rows rank*32..+32                                                        4 KB      -> 76 KB resident
// Everything else (roles, barrier protocol, residual / store-staging tiles, TMA stores, PDL) follows flatconv2_kernel.
#pragma once
#include "sblk_common.cuh"
#include "sblk_igemm2.cuh"
#include "sblk_flatconv2.cuh"

namespace sblk {

struct Fc2xCfg {
  static constexpr int C = 64;
  static constexpr int TILE_ROWS = 256;               // output rows per CTA and tile (512 per pair)
  static constexpr int KOFF_MAX = 17;                  // (Wp + 2) / 2 + 1 for Wp <= 31
  static constexpr int BOX_H = 168;                    // rows of one parity run: 128 + 2 * KOFF_MAX, rounded up to 8
  static constexpr int A_BOX_BYTES = BOX_H * 128;      // 21 KB
  static constexpr int A_STAGE_BYTES = 2 * A_BOX_BYTES;
  static constexpr int A_STAGES = 2;
  static constexpr int R_BYTES = TILE_ROWS * 128;      // residual / store-staging tile: 32 KB
  static constexpr int R_BUFS = 2;                     // one per epilogue group
  static constexpr int B_R_BYTES = 24 * 1024;          // slots of one filter row r: 4 + 8 + 8 + 4 KB
  static constexpr int B_BYTES = 3 * B_R_BYTES + 4096; // + the split W[0,1] slot
  static constexpr int OFF_B = 0;
  static constexpr int OFF_A = B_BYTES;                // 76 KB (1024-aligned)
  static constexpr int OFF_R = OFF_A + A_STAGES * A_STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_R + R_BUFS * R_BYTES + 1024;
  static constexpr int ACC_COLS = 128;
  static constexpr int ACC_STAGES = 4;
  static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;
  static constexpr int EPI_GROUPS = 2;
  static constexpr int EPI_WARPS = 4 * EPI_GROUPS;
  static constexpr int THREADS = 64 + EPI_WARPS * 32 + 64;
  static_assert(OFF_A % 1024 == 0 && A_BOX_BYTES % 1024 == 0 && OFF_R % 1024 == 0, "swizzle atoms");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct FlatConv2xParams {
  int m_total;       // rows of the flat activation matrix (even)
  int num_tiles;     // pair tiles = ceil(m_total / 512)
  int H, W;
  int relu;
  int has_res;
  const float* bias;
  unsigned long long* dbg;     // profiling aid (SBLK_DEBUG builds, SBLK_FLAT_STAMPS=1): per-tile clock64 stamps of CTA 0
  int debug_mode;              // SBLK_DEBUG builds, timing experiments (wrong results): 1 = epilogue releases the
                               // accumulator without reading it, 2 = no MMAs are issued, 4 = no residual read
};

__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* d, uint32_t bar_cluster_addr, int c0,
                                             int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Fc2xCfg::THREADS, 1)
flatconv2x_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                  const FlatConv2xParams p) {
  using Cfg = Fc2xCfg;
  constexpr int A_STAGES = Cfg::A_STAGES;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr uint32_t IDESC128 = make_idesc_bf16(256, 128);
  constexpr uint32_t IDESC64 = make_idesc_bf16(256, 64);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[A_STAGES];      // leader: both parity runs of both CTAs landed
  __shared__ uint64_t a_empty[A_STAGES];     // both CTAs: stage released by the MMAs (multicast commit)
  __shared__ uint64_t b_full;                // leader: resident filter slots of both CTAs landed
  __shared__ uint64_t r_full[Cfg::R_BUFS];   // local: staging tile usable (residual landed / previous store read out)
  __shared__ uint64_t s_ready[Cfg::R_BUFS];  // local: the 4 warps of an epilogue group finished writing the tile
  __shared__ uint64_t tfull_bar[ACC_STAGES];
  __shared__ uint64_t tempty_bar[ACC_STAGES];   // leader: 4 epilogue warps (one group) x 2 CTAs
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[Cfg::C];

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int Wp = p.W + 2;
  const int koff = (Wp + 2) / 2 + 1;         // staged runs start koff half-rows before the tile

  // contiguous, balanced range of pair tiles for this CTA pair
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int base_cnt = p.num_tiles / num_pairs;
  const int rem = p.num_tiles - base_cnt * num_pairs;
  const int my_cnt = base_cnt + (pair_id < rem ? 1 : 0);
  const int tile_begin = pair_id * base_cnt + min(pair_id, rem);
  const int tile_end = tile_begin + my_cnt;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmO);
#pragma unroll
    for (int i = 0; i < A_STAGES; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(&b_full, 1);
#pragma unroll
    for (int i = 0; i < Cfg::R_BUFS; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&s_ready[i], 4);
    }
#pragma unroll
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);   // the 4 warps of one epilogue group x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(&tmem_base_slot, Cfg::TMEM_COLS);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + Cfg::C) bias_s[threadIdx.x - 64] = __ldg(p.bias + (threadIdx.x - 64));
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data

  if (warp == 0) {
    // ------------------------------------------------ activation loader (both CTAs): the two parity runs of a tile
    grid_dep_wait();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int row0 = tile * 512 + static_cast<int>(rank) * Cfg::TILE_ROWS;   // first output row of this CTA (even)
      mbar_wait(&a_empty[stage], phase ^ 1u, 0x0901);
      uint8_t* a_dst = smem + Cfg::OFF_A + stage * Cfg::A_STAGE_BYTES;
      const uint32_t bar = mapa_u32(smem_u32(&a_full[stage]), 0);
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(&a_full[stage], 2u * Cfg::A_STAGE_BYTES);
        // half-rows k0 .. k0 + BOX_H of parity 0 / 1 (flat rows 2k + parity); out-of-range rows are zero-filled
        tma2_load_3d(a_dst, &tmX, bar, 0, 0, row0 / 2 - koff);
        tma2_load_3d(a_dst + Cfg::A_BOX_BYTES, &tmX, bar, 0, 1, row0 / 2 - koff);
      }
      __syncwarp();
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
    for (int i = 0; i < A_STAGES; ++i) {   // drain (see sblk_igemm2.cuh)
      mbar_wait(&a_empty[stage], phase ^ 1u, 0x0903);
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 3 + Cfg::EPI_WARPS) {
    // ------------------------------------------------ store + residual warp (both CTAs), as in flatconv2_kernel
    grid_dep_wait();
    if (lane == 0) {
      constexpr int R_BUFS = Cfg::R_BUFS;
      auto recycle = [&](int j_next) {         // staging tile j_next % R_BUFS is free again: prepare it for tile j_next
        if (j_next >= my_cnt) return;
        const int rb = j_next % R_BUFS;
        if (p.has_res) {
          const int row0 = (tile_begin + j_next) * 512 + static_cast<int>(rank) * Cfg::TILE_ROWS;
          mbar_arrive_expect_tx(&r_full[rb], Cfg::R_BYTES);
          tma_load_2d(smem + Cfg::OFF_R + rb * Cfg::R_BYTES, &tmR, &r_full[rb], 0, row0);
        } else {
          mbar_arrive(&r_full[rb]);
        }
      };
      if (p.has_res)
        for (int j = 0; j < R_BUFS; ++j) recycle(j);   // first round: the tiles are free, only the residuals are missing
      for (int j = 0; j < my_cnt; ++j) {
        const int rb = j % R_BUFS;
        const int row0 = (tile_begin + j) * 512 + static_cast<int>(rank) * Cfg::TILE_ROWS;
        mbar_wait(&s_ready[rb], static_cast<uint32_t>(j / R_BUFS) & 1u, 0x0902);
        tma_store_2d(&tmO, smem + Cfg::OFF_R + rb * Cfg::R_BYTES, 0, row0);   // rows past the end are clipped by TMA
        bulk_commit_group();
        bulk_wait_group_read0();
        recycle(j + R_BUFS);
      }
      bulk_wait_group0();                      // all output bytes are in global memory before the CTA retires
    }
    __syncwarp();
  } else if (warp == 2 + Cfg::EPI_WARPS) {
    // ------------------------------------------------ filter loader (both CTAs): resident slots, see the header
    const uint32_t bar = mapa_u32(smem_u32(&b_full), 0);
    if (elect_one()) {
      if (leader) mbar_arrive_expect_tx(&b_full, 2u * Cfg::B_BYTES);
      const int half = static_cast<int>(rank) * 32;
      const int rk = static_cast<int>(rank);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        uint8_t* b = smem + Cfg::OFF_B + r * Cfg::B_R_BYTES;
        tma2_load_2d(b, &tmW, bar, (3 * r + 0) * 64, half);                  // s' = 0: W[r,0], own half of the rows
        tma2_load_2d(b + 4096, &tmW, bar, (3 * r + 1 - rk) * 64, 0);         // s' = 1: W[r,1] (rank 0) / W[r,0] (rank 1)
        tma2_load_2d(b + 8192, &tmW, bar, (3 * r + 1 - rk) * 64, 32);
        tma2_load_2d(b + 12288, &tmW, bar, (3 * r + 2 - rk) * 64, 0);        // s' = 2: W[r,2] / W[r,1]
        tma2_load_2d(b + 16384, &tmW, bar, (3 * r + 2 - rk) * 64, 32);
        tma2_load_2d(b + 20480, &tmW, bar, (3 * r + 2) * 64, half);          // s' = 3: W[r,2], own half of the rows
      }
      tma2_load_2d(smem + Cfg::OFF_B + 3 * Cfg::B_R_BYTES, &tmW, bar, 1 * 64, half);   // W[0,1] split like a single tap
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: leader CTA only (one elected lane issues)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t db0 = make_desc_sw128(smem_base + Cfg::OFF_B);
      const uint32_t db0_lo = static_cast<uint32_t>(db0);
      // per (r, s'): pixel-operand offset (parity run + row shift) and filter slot, in descriptor units of 16 B
      uint32_t a_off[12], b_off[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int sp = 0; sp < 4; ++sp) {
          const int delta = (r - 1) * Wp + sp - 1;
          const int par = delta & 1;
          const int shift = ((delta - par) >> 1) + koff;   // >= 0
          a_off[r * 4 + sp] = static_cast<uint32_t>(par * (Cfg::A_BOX_BYTES / 16) + shift * 8);
          const int slot = sp == 0 ? 0 : sp == 1 ? 4096 : sp == 2 ? 12288 : 20480;
          b_off[r * 4 + sp] = static_cast<uint32_t>((r * Cfg::B_R_BYTES + slot) / 16);
        }
      }
      const uint32_t b_extra = static_cast<uint32_t>(3 * Cfg::B_R_BYTES / 16);
      mbar_wait(&b_full, 0, 0x0906);
      unsigned long long* const dbg = (p.dbg != nullptr && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        if (dbg) dbg[(tile - tile_begin) * 16 + 0] = clock64();
        mbar_wait(&a_full[stage], phase, 0x0907);
        if (dbg) dbg[(tile - tile_begin) * 16 + 1] = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 0x0908);
        if (dbg) dbg[(tile - tile_begin) * 16 + 2] = clock64();
        tc_fence_after_sync();
        const uint64_t da0 = make_desc_sw128(smem_base + Cfg::OFF_A + stage * Cfg::A_STAGE_BYTES);
        const uint32_t da0_lo = static_cast<uint32_t>(da0);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * Cfg::ACC_COLS);
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            if (p.debug_mode & 2) break;
#pragma unroll
            for (int sp = 0; sp < 4; ++sp) {
              const uint32_t a_lo = da0_lo + a_off[r * 4 + sp];
              const uint32_t b_lo = db0_lo + b_off[r * 4 + sp];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = desc_with_lo(da0, a_lo + static_cast<uint32_t>(2 * k));
                if (sp == 0) {           // W[r,0] -> columns 0-63
                  umma2_bf16(d_tmem, da, desc_with_lo(db0, b_lo + static_cast<uint32_t>(2 * k)), IDESC64,
                             (r > 0 || k > 0) ? 1u : 0u);
                } else if (sp == 3) {    // W[r,2] -> columns 64-127
                  umma2_bf16(d_tmem + 64u, da, desc_with_lo(db0, b_lo + static_cast<uint32_t>(2 * k)), IDESC64, 1u);
                } else if (r == 0 && sp == 1 && k == 0) {
                  // first contribution to columns 64-127 (W[0,0], initialises them) next to an accumulating one into
                  // columns 0-63 (W[0,1]): two N = 64 MMAs out of the single-tap slots
                  umma2_bf16(d_tmem + 64u, da, desc_with_lo(db0, db0_lo + b_off[0]), IDESC64, 0u);
                  umma2_bf16(d_tmem, da, desc_with_lo(db0, db0_lo + b_extra), IDESC64, 1u);
                } else {                 // [W[r,s'] | W[r,s'-1]] -> columns 0-127
                  umma2_bf16(d_tmem, da, desc_with_lo(db0, b_lo + static_cast<uint32_t>(2 * k)), IDESC128, 1u);
                }
              }
            }
          }
          umma2_commit_mc(&tfull_bar[acc]);
          umma2_commit_mc(&a_empty[stage]);
        }
        __syncwarp();
        if (dbg) dbg[(tile - tile_begin) * 16 + 3] = clock64();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (both CTAs): two groups of 4 warps, alternate tiles.
    // Thread = accumulator row j = two output rows 2j (columns 0-63) and 2j + 1 (columns 64-127): + bias (+ residual,
    // read from the staging tile), ReLU, halo rows -> 0, bf16 written back in place in the layout TMA SWIZZLE_128B
    // gave the residual and expects for the store (16-byte chunk c of row rho at c ^ (rho & 7)).  Lanes j and j + 4 of
    // a quarter-warp hold rows with the same rho & 7, so odd groups of four lanes take each pair of chunks in the
    // opposite order: every 16-byte access of a quarter-warp hits eight different bank groups.
    grid_dep_wait();
    const int ew = warp - 2;
    const int grp = ew >> 2;                   // epilogue group = staging buffer
    const int quarter = warp & 3;
    const int arow = quarter * 32 + lane;      // accumulator row of this thread
    const int swap = (lane >> 2) & 1;
    const int Hp = p.H + 1;
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    for (int j = grp; j < my_cnt; j += Cfg::EPI_GROUPS) {
      const int tile = tile_begin + j;
      const int row0 = tile * 512 + static_cast<int>(rank) * Cfg::TILE_ROWS;
      const int acc = j & (ACC_STAGES - 1);
      const uint32_t acc_phase = static_cast<uint32_t>(j >> 2) & 1u;
      const int rb = j % Cfg::R_BUFS;
      const uint32_t rphase = static_cast<uint32_t>(j / Cfg::R_BUFS) & 1u;
      uint8_t* stg = smem + Cfg::OFF_R + rb * Cfg::R_BYTES;
      unsigned long long* const dbg =
          (p.dbg != nullptr && blockIdx.x == 0 && (ew & 3) == 0 && lane == 0) ? p.dbg + j * 16 : nullptr;
      if (dbg) dbg[5] = clock64();
      mbar_wait(&tfull_bar[acc], acc_phase, 0x090a);
      tc_fence_after_sync();
      if (dbg) dbg[6] = clock64();
      mbar_wait(&r_full[rb], p.has_res ? rphase : (rphase ^ 1u), 0x090b);
      if (dbg) dbg[7] = clock64();
      const bool use_res = p.has_res && !(p.debug_mode & 4);
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        if (p.debug_mode & 1) break;
        const int rho = 2 * arow + q;          // row inside the CTA's 256-row tile
        const int m = row0 + rho;
        const int R = m / Wp;
        const int cpos = m - R * Wp;
        const bool valid = cpos >= 1 && cpos <= p.W && R >= 1 && ((R - 1) % Hp) < p.H;   // else: halo row -> zeros
        uint32_t v2[2][32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * Cfg::ACC_COLS + q * 64);
        tmem_ld_32x32b_x32(taddr, v2[0]);
        tmem_ld_32x32b_x32(taddr + 32u, v2[1]);
        tmem_ld_wait();
        uint8_t* rowp = stg + rho * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t (&v)[32] = v2[hh];
#pragma unroll
          for (int qq = 0; qq < 4; qq += 2) {
            const int cc = hh * 4 + qq;        // even chunk of the pair (cc, cc + 1)
            uint4* slot_a = reinterpret_cast<uint4*>(rowp + (((cc ^ swap) ^ (rho & 7)) << 4));       // first access
            uint4* slot_b = reinterpret_cast<uint4*>(rowp + (((cc ^ swap ^ 1) ^ (rho & 7)) << 4));   // second access
            uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;   // residual chunks cc, cc + 1
            if (use_res) {
              const uint4 ra = *slot_a;
              const uint4 rb4 = *slot_b;
              r0 = swap ? rb4 : ra;
              r1 = swap ? ra : rb4;
            }
            uint4 o[2];
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int col0 = hh * 32 + (qq + e2) * 8;      // column inside this row's 64 channels
              const uint4 r4 = e2 == 0 ? r0 : r1;
              const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[col0]);
              const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[col0 + 4]);
              const int vb = (qq + e2) * 8;
              float f[8];
              f[0] = __uint_as_float(v[vb + 0]) + b0.x; f[1] = __uint_as_float(v[vb + 1]) + b0.y;
              f[2] = __uint_as_float(v[vb + 2]) + b0.z; f[3] = __uint_as_float(v[vb + 3]) + b0.w;
              f[4] = __uint_as_float(v[vb + 4]) + b1.x; f[5] = __uint_as_float(v[vb + 5]) + b1.y;
              f[6] = __uint_as_float(v[vb + 6]) + b1.z; f[7] = __uint_as_float(v[vb + 7]) + b1.w;
              if (use_res) {
                f[0] += bf16_lo(r4.x); f[1] += bf16_hi(r4.x); f[2] += bf16_lo(r4.y); f[3] += bf16_hi(r4.y);
                f[4] += bf16_lo(r4.z); f[5] += bf16_hi(r4.z); f[6] += bf16_lo(r4.w); f[7] += bf16_hi(r4.w);
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
              }
              o[e2].x = pack_bf16x2(f[0], f[1]);
              o[e2].y = pack_bf16x2(f[2], f[3]);
              o[e2].z = pack_bf16x2(f[4], f[5]);
              o[e2].w = pack_bf16x2(f[6], f[7]);
              if (!valid) o[e2] = make_uint4(0u, 0u, 0u, 0u);   // halo positions stay zero for the next conv
            }
            *slot_a = swap ? o[1] : o[0];
            *slot_b = swap ? o[0] : o[1];
          }
        }
      }
      // the accumulator is consumed: hand it back to the MMA issuer (leader's barrier)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + static_cast<uint32_t>(acc * 8));
      if (dbg) dbg[8] = clock64();
      fence_proxy_async_smem();               // generic-proxy tile writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_ready[rb]);   // 4 warps -> the store warp issues the tile's TMA store
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
