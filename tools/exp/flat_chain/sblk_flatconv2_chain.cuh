// sblk_flatconv2.cuh — stride-1 3x3 convolution C -> C (C = 64 or 128) over the zero-haloed flat layout, as a
// CTA-PAIR (tcgen05 cta_group::2) shifted-window implicit GEMM.
// Reference: BasicBlock conv1/bn1/relu and conv2/bn2/+=residual/relu of ResNet layer1 (64 ch, 22x22) and layer2
//            (128 ch, 11x11), SBL/transformer/video_frontend.py:10-12,28-41.
//
// Layout and trick are those of sblk_flatconv.cuh: pixel (f, y, x) lives at row (f*(H+1) + 1 + y)*(W+2) + 1 + x of a
// [rows, C] bf16 matrix whose other rows are zero, so tap (r, s) of 128 consecutive output rows is the SAME staged
// pixel run shifted by (r-1)*(W+2) + (s-1) rows; each run is staged once (TMA, SWIZZLE_128B) and the 9 taps are UMMA
// descriptors with row-shifted start addresses.  What changes here:
//   * two CTAs (one TPC) compute 256 rows x C with ONE tcgen05.mma.cta_group::2 stream: each CTA stages its own 128-row
//     pixel run but only HALF of the filter rows, so the shared-memory operand bytes per MMA drop from A+B to A+B/2
//     (the N = 64 MMA of layer1 is bound by exactly that fetch: 6 KB -> 5 KB per M128xN64xK16) and layer2 gets the
//     9x cut in L2->SM activation traffic the im2col kernel cannot give (ncu: 418 MB per launch at the L2 fabric limit).
//   * C = 128: the filter (288 KB) does not fit next to the activations, so its 64-row halves stream through a small
//     ring, one 8 KB tile per (tap, channel block); C = 64: the 36 KB half filter stays resident.
//   * the residual tile is TMA-loaded into the buffer the epilogue uses as its store-staging tile: + bias, + residual,
//     ReLU and the halo mask happen in fp32 in place (one rounding to bf16), and the tile leaves with a TMA store.
//   * the epilogue is two independent groups of 4 warps working on alternate tiles (own staging buffer each), so the
//     TMEM-load -> math -> store latency chain of one tile overlaps the next tile's.
// CHAINS (levels > 1).  One launch runs up to FC2_MAX_LEVELS dependent convs of the same shape (a ResNet stage:
// conv1, conv2 + residual, conv1, ...): level l reads the output of level l - 1.  Launched one by one these convs are
// bound by HBM, not by the tensor pipe: every 64-channel 22x22 conv of the BASELINE batch reads 57 MB, writes 57 MB
// (+ 57 MB of residual) in ~25-40 us, and a whole tensor written by one launch has left the L2 before the next one
// reads it; a kernel boundary adds ~6-8 us (drain, launch, cold pipeline fill).  Here
//   * every CTA pair keeps its contiguous tile range through all levels; a pair tile of level l needs only the pair
//     tiles t-1, t, t+1 of level l - 1 (the 3x3 reach is W + 3 <= 256 rows).  Dependencies are tracked per tile by
//     completion COUNTERS in global memory (`flags`, never reset: both CTAs of the pair add 1 once their TMA store of
//     the tile has completed; a launch waits for its own epoch's value, derived from the counter of one of its own
//     tiles at kernel start);
//   * the tile order is DEPTH-FIRST (ChainIter): chunks of <= 8 tiles go through all levels before the next chunk
//     starts, level l lagging l tiles behind level l - 1 (its rightmost tile needs the next tile of the level below), so
//     a tile is consumed ~7 steps after it was produced — out of the L2 — and the residuals are L2 hits as well.  The
//     tiles within l of a range end depend on the neighbour pairs' last tiles and are done at the end, level by level;
//   * the filter changes with the level: C = 64 keeps two filter buffers (this run's and the next run's, swapped tap by
//     tap behind the MMAs); C = 128 streams its filters anyway.
// Deadlock-free for any CTA scheduling: a CTA's own order is a topological order of its own tiles, and what it needs
// from a neighbour is never later in the neighbour's order than what the neighbour needs from it.
// Roles: warp 0 activation loader, warp 1 MMA issuer (leader CTA) + TMEM owner, 8 (C = 128) or 12 (C = 64) epilogue
// warps, then the filter loader warp, the store + residual warp and (chains) the publisher warp.  Barrier protocol as in sblk_igemm2.cuh (full barriers in the leader).
#pragma once
#include "sblk_common.cuh"
#include "sblk_igemm2.cuh"

namespace sblk {

constexpr int FC2_MAX_LEVELS = 4;

template <int CB>
struct Fc2Cfg {
  static constexpr int C = 64 * CB;                 // Cin == Cout
  static constexpr int TILE_M = 128;                // rows per CTA (256 per pair)
  static constexpr int BOX_PIX = CB == 1 ? 192 : 160;   // staged pixel run: 128 + 2*(Wp+1) rows (Wp <= 31 / 15)
  static constexpr int A_BOX_BYTES = BOX_PIX * 128;
  static constexpr int A_STAGE_BYTES = CB * A_BOX_BYTES;          // 24 KB / 40 KB
  static constexpr int A_STAGES = CB == 1 ? 3 : 2;    // a stage lasts a whole tile (>= one L2 round trip)
  static constexpr int R_BOX_BYTES = TILE_M * 128;                // one 64-channel block of the residual / output tile
  static constexpr int R_BYTES = CB * R_BOX_BYTES;                // 16 KB / 32 KB
  // residual / store-staging tiles: C = 64 has room for two per epilogue group, so the residual of the group's NEXT
  // tile is already in flight while the current one is still being stored (else: load latency on the critical path)
  static constexpr int R_BUFS = CB == 1 ? 4 : 2;
  static constexpr int BH = C / 2;                                // filter rows staged per CTA
  static constexpr int B_TILE_BYTES = BH * 128;                   // one (tap, channel block) k-block: 4 KB / 8 KB
  static constexpr int B_TILES = 9 * CB;                          // k-blocks per tile
  static constexpr bool B_RESIDENT = CB == 1;
  static constexpr int B_SLOTS = B_RESIDENT ? 2 * B_TILES : 10;   // resident: two filters (one slot per tap); streamed: one slot lasts 4 MMAs
  static constexpr int OFF_B = 0;
  static constexpr int OFF_A = ((B_SLOTS * B_TILE_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_R = OFF_A + A_STAGES * A_STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_R + R_BUFS * R_BYTES + 1024;
  static constexpr int ACC_STAGES = 4;
  static constexpr int TMEM_COLS = ACC_STAGES * C;                // 256 / 512
  // epilogue groups of 4 warps (one TMEM lane quarter each) working on successive tiles.  Two are enough: the C = 64
  // kernel is SHARED-MEMORY-BANDWIDTH bound (per 128-row tile the 36 N=64 MMAs fetch 180 KB of operands, the
  // residual / staging / TMA traffic adds 88 KB: ~2100 cycles at 128 B/clk vs ~2550 measured), so a third group
  // (measured: 36.4 vs 35.5 us) only adds contention
  static constexpr int EPI_GROUPS = 2;
  // convs per launch (chains): C = 128 has 2 KB of static shared memory left next to its 225 KB of tiles
  static constexpr int MAX_LEVELS = CB == 1 ? FC2_MAX_LEVELS : 3;
  static constexpr int EPI_WARPS = 4 * EPI_GROUPS;
  static constexpr int THREADS = 64 + EPI_WARPS * 32 + 96;        // + filter loader warp + store / residual warp + publisher warp
};

// shared -> global TMA store of one box (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* d, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(d)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group2() { asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group6() { asm volatile("cp.async.bulk.wait_group 6;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group20() { asm volatile("cp.async.bulk.wait_group 20;" ::: "memory"); }
// generic-proxy acquire -> async-proxy (TMA) reads of global memory issued afterwards see the acquired data
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until the completion counter has reached `target` (wrap-safe); wall-clock watchdog like mbar_wait
__device__ __forceinline__ void flag_wait(const int* flag, int target, uint32_t code) {
  if (ld_acquire_gpu(flag) - target >= 0) return;
  uint32_t polls = 0;
  uint64_t t0 = 0;
  while (ld_acquire_gpu(flag) - target < 0) {
    if ((++polls & 255u) == 0u) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > SBLK_WATCHDOG_NS) {
        unsigned int* wd = g_sblk_watchdog_ptr;
        if (wd != nullptr) {
          atomicCAS_system(wd, 0u, code | 0x80000000u);
          __threadfence_system();
        }
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}


// Depth-first order of the (level, tile) steps of one CTA pair with n tiles (header comment).  Chunks of K tiles; in
// chunk c level l does the tiles (c*K - l .. (c+1)*K - 1 - l) inside its interior [l, n - 1 - l].  In the LAST chunk
// the interior run of level l >= 1 is followed by the level's margins [0, l) and [n - l, n): they need the neighbour
// pairs' last tiles of level l - 1, which exist by then, and are one whole run older than the margins of level l + 1
// that need them.  levels == 1: tiles 0 .. n - 1.
// (Python model + exhaustive check of the order: tests/test_host_cpu.py::test_chain_schedule_model.)
struct ChainIter {
  int n, levels, K, nc;
  int c, lp, side;       // chunk, level, 0 = interior run / 1 = low margin / 2 = high margin
  int t, t_end;
  int emitted;
  __device__ __forceinline__ void load_run() {
    const int m = lp;
    if (side == 0) {
      int hi = min((c + 1) * K - 1, n - 1) - lp;
      int lo = c == 0 ? m : min(c * K - 1, n - 1) - lp + 1;
      lo = max(lo, m);
      hi = min(hi, n - 1 - m);
      t = lo;
      t_end = hi >= lo ? hi + 1 : lo;
    } else if (n - 1 - m < m) {      // no interior at this level: everything is margin
      t = 0;
      t_end = side == 1 ? n : 0;
    } else {
      t = side == 1 ? 0 : n - m;
      t_end = side == 1 ? m : n;
    }
  }
  __device__ __forceinline__ void init(int n_, int levels_, int kmax) {
    n = n_; levels = levels_;
    nc = max(1, (n + kmax - 1) / kmax);   // balanced chunks of at most kmax tiles
    K = (n + nc - 1) / nc;
    c = 0; lp = 0; side = 0; emitted = 0;
    load_run();
  }
  // next step: level l, local tile tl (0 .. n - 1); false once all levels * n steps have been handed out
  __device__ __forceinline__ bool next(int& l, int& tl) {
    if (emitted == levels * n) return false;
    while (t >= t_end) {
      if (c == nc - 1 && lp >= 1 && side < 2) {
        ++side;
      } else {
        side = 0;
        if (++lp == levels) { lp = 0; ++c; }
      }
      load_run();
    }
    l = lp;
    tl = t++;
    ++emitted;
    return true;
  }
};

// tensor maps of every level: input (staged pixel runs), filter, residual, output
struct FlatChainMaps {
  CUtensorMap x[FC2_MAX_LEVELS], w[FC2_MAX_LEVELS], r[FC2_MAX_LEVELS], o[FC2_MAX_LEVELS];
};

struct FlatConv2Params {
  int m_total;       // rows of the flat activation matrix = (F*(H+1) + 1) * (W+2)
  int num_tiles;     // pair tiles = ceil(m_total / 256)
  int H, W;
  int levels;                        // convs in this launch (1 = a plain conv)
  int relu[FC2_MAX_LEVELS];
  int has_res[FC2_MAX_LEVELS];
  int res_level[FC2_MAX_LEVELS];     // level whose OUTPUT is this level's residual, or -1 (a tensor written before the launch)
  const float* bias[FC2_MAX_LEVELS]; // [C] folded BN shift
  int* flags;                        // [levels][num_tiles] tile completion counters (levels > 1), else nullptr
  int chunk_tiles;                   // chains: tiles per depth-first chunk (ChainIter)
  unsigned long long* dbg;     // profiling aid (SBLK_FLAT_STAMPS=1): per-tile clock64 stamps of CTA 0, or nullptr
  int debug_mode;              // SBLK_DEBUG builds, timing experiments (wrong results): 16 = issue the MMAs with N = 32
};

template <int CB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Fc2Cfg<CB>::THREADS, 1)
flatconv2_kernel(const __grid_constant__ FlatChainMaps tm, const __grid_constant__ FlatConv2Params p) {
  using Cfg = Fc2Cfg<CB>;
  constexpr int C = Cfg::C;
  constexpr int A_STAGES = Cfg::A_STAGES;
  constexpr int B_SLOTS = Cfg::B_SLOTS;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr uint32_t IDESC = make_idesc_bf16(256, C);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[A_STAGES];      // leader: pixel runs of both CTAs landed
  __shared__ uint64_t a_empty[A_STAGES];     // both CTAs: slot released by the MMAs (multicast commit)
  __shared__ uint64_t b_full[B_SLOTS];       // leader: filter k-block halves of both CTAs landed (resident: one per filter buffer)
  __shared__ uint64_t b_empty[B_SLOTS];      // both CTAs (resident, chains: filter buffer released by the last tile of a filter run)
  __shared__ uint64_t r_full[Cfg::R_BUFS];   // local: staging tile may be used (residual landed / previous store read out)
  __shared__ uint64_t s_ready[Cfg::R_BUFS];  // local: the 4 warps of an epilogue group finished writing the tile
  __shared__ uint64_t tfull_bar[ACC_STAGES];
  __shared__ uint64_t tempty_bar[ACC_STAGES];   // leader: 4 epilogue warps (one group) x 2 CTAs
  __shared__ uint64_t k_bar;                 // chains: this launch's counter target is known (k_s)
  __shared__ int k_s;
  __shared__ int complete_s;                 // chains: stores [0, complete_s) of this CTA have completed (store warp -> publisher)
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[Cfg::MAX_LEVELS][C];

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int Wp = p.W + 2;
  const int levels = p.levels;

  // contiguous, balanced range of pair tiles for this CTA pair (the same range at every level); every role walks the
  // same (level, tile) step sequence with its own ChainIter
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int base_cnt = p.num_tiles / num_pairs;
  const int rem = p.num_tiles - base_cnt * num_pairs;
  const int my_cnt = base_cnt + (pair_id < rem ? 1 : 0);
  const int tile_begin = pair_id * base_cnt + min(pair_id, rem);
  const int total_steps = levels * my_cnt;

  if (threadIdx.x == 0) {
    for (int l = 0; l < levels; ++l) {
      tma_prefetch_desc(&tm.x[l]);
      tma_prefetch_desc(&tm.w[l]);
      tma_prefetch_desc(&tm.r[l]);
      tma_prefetch_desc(&tm.o[l]);
    }
#pragma unroll
    for (int i = 0; i < A_STAGES; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
#pragma unroll
    for (int i = 0; i < B_SLOTS; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
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
    mbar_init(&k_bar, 1);
    complete_s = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(&tmem_base_slot, Cfg::TMEM_COLS);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C) {
    for (int l = 0; l < levels; ++l) bias_s[l][threadIdx.x - 64] = __ldg(p.bias[l] + (threadIdx.x - 64));
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data
  // (the filter loader below starts before grid_dep_wait: weights do not depend on the previous kernel)

  if (warp == 0) {
    // ------------------------------------------------ activation loader (both CTAs)
    grid_dep_wait();
    int target = 0;
    if (p.flags != nullptr) {
      // every counter stands at 2 * (launches so far) — both CTAs of a pair add 1 per tile and launch — and no tile of
      // this launch is complete yet, whatever the previous launch's tile partition was
      if (lane == 0) {
        target = ld_acquire_gpu(p.flags + tile_begin) + 2;
        k_s = target;
        mbar_arrive(&k_bar);
      }
      target = __shfl_sync(0xffffffffu, target, 0);
    }
    ChainIter it;
    it.init(my_cnt, levels, p.chunk_tiles);
    int stage = 0;
    uint32_t phase = 0;
    int ready_upto = 0;     // chains: the inputs of steps < ready_upto are known to be complete
    for (int i = 0; i < total_steps; ++i) {
      int l, tl;
      it.next(l, tl);
      const int tile = tile_begin + tl;
      const int row0 = tile * 256 + static_cast<int>(rank) * Cfg::TILE_M;   // first output row of this CTA
      if (l > 0 && i >= ready_upto) {
        // the staged run reaches into the pair tiles tile - 1 .. tile + 1 of the previous level.  One round trip to L2
        // checks the counters of the next 10 steps (lane = (step, neighbour)); the loader spins only when the step it
        // needs next is not ready
        const int ds = lane / 3;
        ChainIter la = it;
        int l2 = l, tl2 = tl;
        bool exists = true;
        for (int k = 0; k < ds && exists; ++k) exists = la.next(l2, tl2);
        const int t2 = tile_begin + tl2 - 1 + (lane - 3 * ds);
        bool ok = true;
        if (lane < 30 && exists && l2 > 0 && t2 >= 0 && t2 < p.num_tiles)
          ok = ld_acquire_gpu(p.flags + (l2 - 1) * p.num_tiles + t2) - target >= 0;
        const uint32_t bad = __ballot_sync(0xffffffffu, !ok);
        int n_ready = bad == 0u ? 10 : (__ffs(static_cast<int>(bad)) - 1) / 3;   // leading steps with all neighbours ready
        if (n_ready == 0) {
          const int t1 = tile - 1 + lane;
          if (lane < 3 && t1 >= 0 && t1 < p.num_tiles) flag_wait(p.flags + (l - 1) * p.num_tiles + t1, target, 0x070c);
          n_ready = 1;
        }
        __syncwarp();
        ready_upto = i + n_ready;
      }
      mbar_wait(&a_empty[stage], phase ^ 1u, 0x0701);
      uint8_t* a_dst = smem + Cfg::OFF_A + stage * Cfg::A_STAGE_BYTES;
      const uint32_t bar = mapa_u32(smem_u32(&a_full[stage]), 0);
      if (elect_one()) {
        if (l > 0) fence_proxy_async_global();
        if (leader) mbar_arrive_expect_tx(&a_full[stage], 2u * Cfg::A_STAGE_BYTES);
#pragma unroll
        for (int cb = 0; cb < CB; ++cb)   // may start before row 0 / run past the end: TMA zero-fills
          tma2_load_2d(a_dst + cb * Cfg::A_BOX_BYTES, &tm.x[l], bar, cb * 64, row0 - (Wp + 1));
      }
      __syncwarp();
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
    for (int i = 0; i < A_STAGES; ++i) {   // drain (see sblk_igemm2.cuh)
      mbar_wait(&a_empty[stage], phase ^ 1u, 0x0703);
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 3 + Cfg::EPI_WARPS) {
    // ------------------------------------------------ store + residual warp (both CTAs).  One lane owns every TMA store
    // of the CTA (bulk async-groups are per thread) and recycles the staging tiles: as soon as a store has been READ
    // out of shared memory the tile is handed back — with the residual of its next user already requested (has_res) or
    // as a plain arrive.  Nothing in the epilogue warps ever waits for a store, and a residual load is in flight two
    // (C = 64) tiles ahead of its use.  (Clock stamps before: the MMA issuer idled ~40 % on a_full because residual
    // waits sat in the activation loader's in-order loop, and the epilogue groups idled on their own store reads.)
    // Chains: the number of COMPLETED stores is handed to the publisher warp.
    grid_dep_wait();
    if (lane == 0) {
      constexpr int R_BUFS = Cfg::R_BUFS;
      constexpr int LAG = R_BUFS / 2 - 1;      // stores allowed to be still reading when the next one is issued
      int target = 0;
      if (p.flags != nullptr) {
        mbar_wait(&k_bar, 0, 0x070d);
        target = k_s;
      }
      ChainIter it_rec, it_st;
      it_rec.init(my_cnt, levels, p.chunk_tiles);
      it_st.init(my_cnt, levels, p.chunk_tiles);
      int rec = 0;          // next step whose staging tile has to be prepared (in step order)
      int rec_l = 0, rec_tl = 0;
      bool rec_have = it_rec.next(rec_l, rec_tl);
      int read_done = 0;    // stores [0, read_done) have been read out of shared memory
      int complete = 0;     // stores [0, complete) have completed (their bytes are in global memory)
      int pub = 0;          // value last handed to the publisher
      // Prepare the staging tiles of as many steps as possible: tile rec % R_BUFS is free once the store of step
      // rec - R_BUFS has been read out; a residual that this launch itself produces (res_level >= 0) can be loaded once
      // its tile's completion counter says so (checked without blocking: retried after the next store).
      auto recycle = [&]() {
        while (rec_have && rec < read_done + R_BUFS) {
          const int rb = rec % R_BUFS;
          const int tile = tile_begin + rec_tl;
          if (p.has_res[rec_l]) {
            const int rl = p.res_level[rec_l];
            if (rl >= 0) {
              if (ld_acquire_gpu(p.flags + rl * p.num_tiles + tile) - target < 0) break;
              fence_proxy_async_global();
            }
            const int row0 = tile * 256 + static_cast<int>(rank) * Cfg::TILE_M;
            uint8_t* r_dst = smem + Cfg::OFF_R + rb * Cfg::R_BYTES;
            mbar_arrive_expect_tx(&r_full[rb], Cfg::R_BYTES);
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
              tma_load_2d(r_dst + cb * Cfg::R_BOX_BYTES, &tm.r[rec_l], &r_full[rb], cb * 64, row0);
          } else {
            mbar_arrive(&r_full[rb]);
          }
          ++rec;
          rec_have = it_rec.next(rec_l, rec_tl);
        }
      };
      // completed stores are handed to the publisher warp (the release fence in front of a counter update takes
      // microseconds when HBM is saturated; this warp must keep the stores and the staging tiles moving)
      auto publish = [&]() {
        if (p.flags != nullptr && complete > pub) {
          pub = complete;
          asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&complete_s)), "r"(complete) : "memory");
        }
      };
      // chains: stores allowed to be still in flight when the next one is issued.  A tile is consumed ~7 steps after it
      // was produced (depth-first order), so completion has to be tracked closely
      const int pub_lag = my_cnt >= 5 ? 1 : 0;
      recycle();                               // first round: the tiles are free, only the residuals are missing
      for (int j = 0; j < total_steps; ++j) {
        const int rb = j % R_BUFS;
        int l, tl;
        it_st.next(l, tl);
        const int row0 = (tile_begin + tl) * 256 + static_cast<int>(rank) * Cfg::TILE_M;
        const uint32_t sparity = static_cast<uint32_t>(j / R_BUFS) & 1u;
        if (levels > 1 && complete < j && !mbar_try_wait(&s_ready[rb], sparity)) {
          // Nothing to store yet.  The tile this warp is about to wait for may depend on stores it has not reported
          // yet (completion is tracked with a lag, below): after a short spin — tiles arrive every ~0.8 us in steady
          // state — finish and report everything stored so far.
          const long long t_idle = clock64();
          bool ready = false;
          while (!(ready = mbar_try_wait(&s_ready[rb], sparity)) && clock64() - t_idle < 3000) {}
          if (!ready) {
            bulk_wait_group0();
            read_done = complete = j;
            publish();
            recycle();
          }
        }
        mbar_wait(&s_ready[rb], sparity, 0x0702);
        const uint8_t* stg = smem + Cfg::OFF_R + rb * Cfg::R_BYTES;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb)        // rows past the end of the tensor are clipped by TMA
          tma_store_2d(&tm.o[l], stg + cb * Cfg::R_BOX_BYTES, cb * 64, row0);
        bulk_commit_group();
        if (LAG == 1) bulk_wait_group_read1(); else bulk_wait_group_read0();
        read_done = max(read_done, j + 1 - LAG);
        recycle();
        if (levels > 1 && j >= pub_lag) {
          if (pub_lag == 1) bulk_wait_group1(); else bulk_wait_group0();
          complete = max(complete, j + 1 - pub_lag);
          publish();
          recycle();
        }
      }
      bulk_wait_group0();                      // all output bytes are in global memory before the CTA retires
      complete = total_steps;
      publish();
    }
    __syncwarp();
  } else if (warp == 4 + Cfg::EPI_WARPS) {
    // ------------------------------------------------ publisher (chains): bumps the completion counter of every step
    // whose store the store warp has seen complete.  One gpu-scope release fence covers all steps found at a poll.
    if (lane == 0 && p.flags != nullptr) {
      ChainIter it;
      it.init(my_cnt, levels, p.chunk_tiles);
      int pub = 0;
      while (pub < total_steps) {
        int c;
        asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(c) : "r"(smem_u32(&complete_s)) : "memory");
        if (c > pub) {
          __threadfence();
          for (; pub < c; ++pub) {
            int l, tl;
            it.next(l, tl);
            if (l < levels - 1 || l == 0)
              asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p.flags + l * p.num_tiles + tile_begin + tl), "r"(1) : "memory");
          }
        } else {
          __nanosleep(32);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2 + Cfg::EPI_WARPS) {
    // ------------------------------------------------ filter loader (both CTAs): own half of the output channels
    const int n0 = static_cast<int>(rank) * Cfg::BH;
    ChainIter it;
    it.init(my_cnt, levels, p.chunk_tiles);
    if (Cfg::B_RESIDENT) {
      // a filter run = consecutive steps of one level.  Run r uses filter buffer r & 1 (b_full / b_empty [r & 1]): its
      // filter is requested as soon as the last tile of run r - 2 has released the buffer, a whole run ahead of its use
      // (one commit per run: nine multicast commits behind a tile's MMAs delayed the next tile by ~1500 cycles)
      int cur = -1, frun = -1;
      for (int i = 0; i < total_steps; ++i) {
        int l, tl;
        it.next(l, tl);
        if (l == cur) continue;
        cur = l;
        ++frun;
        const int buf = frun & 1;
        const int s0 = buf * Cfg::B_TILES;
        if (frun >= 2) mbar_wait(&b_empty[buf], static_cast<uint32_t>((frun - 2) >> 1) & 1u, 0x0704);
        const uint32_t bar = mapa_u32(smem_u32(&b_full[buf]), 0);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&b_full[buf], 2u * Cfg::B_TILES * Cfg::B_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < Cfg::B_TILES; ++t)
            tma2_load_2d(smem + Cfg::OFF_B + (s0 + t) * Cfg::B_TILE_BYTES, &tm.w[l], bar, t * 64, n0);
        }
        __syncwarp();
      }
    } else {
      int slot = 0;
      uint32_t phase = 0;
      for (int i = 0; i < total_steps; ++i) {
        int l, tl;
        it.next(l, tl);
        for (int kb = 0; kb < Cfg::B_TILES; ++kb) {   // k-block kb = (channel block, tap) in MMA order
          const int cb = kb / 9;
          const int t = kb - cb * 9;
          mbar_wait(&b_empty[slot], phase ^ 1u, 0x0704);
          const uint32_t bar = mapa_u32(smem_u32(&b_full[slot]), 0);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&b_full[slot], 2u * Cfg::B_TILE_BYTES);
            tma2_load_2d(smem + Cfg::OFF_B + slot * Cfg::B_TILE_BYTES, &tm.w[l], bar, (t * CB + cb) * 64, n0);
          }
          __syncwarp();
          if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
        }
      }
      for (int i = 0; i < B_SLOTS; ++i) {   // drain
        mbar_wait(&b_empty[slot], phase ^ 1u, 0x0705);
        if (++slot == B_SLOTS) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: leader CTA only (one elected lane issues)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int slot = 0;
      uint32_t bphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t db0 = make_desc_sw128(smem_base + Cfg::OFF_B);
      const uint32_t db0_lo = static_cast<uint32_t>(db0);
      const uint32_t idesc = (p.debug_mode & 16) ? make_idesc_bf16(256, 32) : IDESC;
      uint32_t tap_off[9];   // (r*Wp + s) rows of 128 B, in descriptor units of 16 B
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_off[t] = static_cast<uint32_t>(((t / 3) * Wp + (t % 3)) * 8);
      unsigned long long* const dbg = (p.dbg != nullptr && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr;
      ChainIter it;
      it.init(my_cnt, levels, p.chunk_tiles);
      int l = 0, tl = 0, cur = -1, frun = -1;
      it.next(l, tl);
      for (int i = 0; i < total_steps; ++i) {
        int l_next = -1, tl_next = 0;
        const bool have = it.next(l_next, tl_next);        // one step of look-ahead: does the filter run end here?
        const bool first = l != cur;                       // first step of a filter run
        const bool last = levels > 1 && (!have || l_next != l);   // last step of a filter run
        if (first) { cur = l; ++frun; }
        if (dbg) dbg[i * 16 + 0] = clock64();
        mbar_wait(&a_full[stage], phase, 0x0707);
        if (dbg) dbg[i * 16 + 1] = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 0x0708);
        if (dbg) dbg[i * 16 + 2] = clock64();
        tc_fence_after_sync();
        const uint64_t da0 = make_desc_sw128(smem_base + Cfg::OFF_A + stage * Cfg::A_STAGE_BYTES);
        const uint32_t da0_lo = static_cast<uint32_t>(da0);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * C);
        if (Cfg::B_RESIDENT) {
          // resident filters (C = 64): one straight run of 36 MMAs per tile; the first tile of a filter run waits for its
          // taps (requested a whole run earlier), the last tile of a chained run hands the tap slots to the run after next
          const int s0 = (frun & 1) * Cfg::B_TILES;
          const uint32_t fb_lo = db0_lo + static_cast<uint32_t>(s0 * (Cfg::B_TILE_BYTES / 16));
          if (first) {
            mbar_wait(&b_full[frun & 1], static_cast<uint32_t>(frun >> 1) & 1u, 0x0706);
            tc_fence_after_sync();
          }
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t b_lo = fb_lo + static_cast<uint32_t>(t * (Cfg::B_TILE_BYTES / 16));
              const uint32_t a_lo = da0_lo + tap_off[t];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_bf16(d_tmem, desc_with_lo(da0, a_lo + static_cast<uint32_t>(2 * k)),
                           desc_with_lo(db0, b_lo + static_cast<uint32_t>(2 * k)), idesc, (t > 0 || k > 0) ? 1u : 0u);
            }
            if (last) umma2_commit_mc(&b_empty[frun & 1]);
          }
          __syncwarp();
        } else {
#pragma unroll 1
          for (int cb = 0; cb < CB; ++cb) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              mbar_wait(&b_full[slot], bphase, 0x0709);
              tc_fence_after_sync();
              const uint32_t b_lo = db0_lo + static_cast<uint32_t>(slot * (Cfg::B_TILE_BYTES / 16));
              const uint32_t a_lo = da0_lo + static_cast<uint32_t>(cb * (Cfg::A_BOX_BYTES / 16)) + tap_off[t];
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma2_bf16(d_tmem, desc_with_lo(da0, a_lo + static_cast<uint32_t>(2 * k)),
                             desc_with_lo(db0, b_lo + static_cast<uint32_t>(2 * k)), idesc,
                             (cb > 0 || t > 0 || k > 0) ? 1u : 0u);
                umma2_commit_mc(&b_empty[slot]);
              }
              __syncwarp();
              if (++slot == B_SLOTS) { slot = 0; bphase ^= 1u; }
            }
          }
        }
        if (elect_one()) {
          umma2_commit_mc(&tfull_bar[acc]);
          umma2_commit_mc(&a_empty[stage]);
        }
        __syncwarp();
        if (dbg) dbg[i * 16 + 3] = clock64();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
        l = l_next;
        tl = tl_next;
      }
    }
  } else {
    // ------------------------------------------------ epilogue (both CTAs): two groups of 4 warps, alternate steps
    // thread = output row (TMEM lane), all C channels: + bias (+ residual, read from the staging tile the loader
    // filled), ReLU, halo rows -> 0, bf16 written back IN PLACE (16-byte chunk c of row r at c ^ (r & 7): the layout TMA
    // SWIZZLE_128B gave the residual and expects for the store), then one thread issues the TMA store of the tile.
    grid_dep_wait();
    const int ew = warp - 2;
    const int grp = ew >> 2;                   // epilogue group = staging buffer
    const int quarter = warp & 3;
    const int arow = quarter * 32 + lane;      // output row of this thread inside the CTA's 128-row tile
    const int Hp = p.H + 1;
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    ChainIter it;
    it.init(my_cnt, levels, p.chunk_tiles);
    {
      int l0, t0;
      for (int k = 0; k < grp; ++k) it.next(l0, t0);   // this group's first step
    }
    for (int j = grp; j < total_steps; j += Cfg::EPI_GROUPS) {
      int l, tl;
      it.next(l, tl);
      {
        int l0, t0;
        for (int k = 1; k < Cfg::EPI_GROUPS; ++k) it.next(l0, t0);   // the other groups' steps
      }
      const int tile = tile_begin + tl;
      const bool has_res = p.has_res[l] != 0;
      const bool relu = p.relu[l] != 0;
      const float* const bias_l = bias_s[l];
      const int row0 = tile * 256 + static_cast<int>(rank) * Cfg::TILE_M;
      const int acc = j & (ACC_STAGES - 1);
      const uint32_t acc_phase = static_cast<uint32_t>(j >> 2) & 1u;
      const int rb = j % Cfg::R_BUFS;
      const uint32_t rphase = static_cast<uint32_t>(j / Cfg::R_BUFS) & 1u;
      uint8_t* stg = smem + Cfg::OFF_R + rb * Cfg::R_BYTES;
      unsigned long long* const dbg =
          (p.dbg != nullptr && blockIdx.x == 0 && (ew & 3) == 0 && lane == 0) ? p.dbg + j * 16 : nullptr;
      if (dbg) dbg[5] = clock64();
      mbar_wait(&tfull_bar[acc], acc_phase, 0x070a);
      tc_fence_after_sync();
      if (dbg) dbg[6] = clock64();
      // the staging tile is usable: its residual has landed, or its previous store has been read out
      mbar_wait(&r_full[rb], rphase, 0x070b);
      if (dbg) { dbg[7] = clock64(); dbg[11] = static_cast<unsigned long long>(l * 1000 + tl); }
      const int m = row0 + arow;
      const int R = m / Wp;
      const int cpos = m - R * Wp;
      const bool valid = cpos >= 1 && cpos <= p.W && R >= 1 && ((R - 1) % Hp) < p.H;   // else: halo row -> zeros
#pragma unroll 1
      for (int c64 = 0; c64 < C / 64; ++c64) {
        // both 32-column halves of the 64-channel block are requested before the first is used
        uint32_t v2[2][32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * C + c64 * 64);
        tmem_ld_32x32b_x32(taddr, v2[0]);
        tmem_ld_32x32b_x32(taddr + 32u, v2[1]);
        tmem_ld_wait();
        uint8_t* rowp = stg + c64 * Cfg::R_BOX_BYTES + arow * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t (&v)[32] = v2[hh];
          const int col0 = c64 * 64 + hh * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = *reinterpret_cast<const float4*>(&bias_l[col0 + 8 * q]);
            const float4 b1 = *reinterpret_cast<const float4*>(&bias_l[col0 + 8 * q + 4]);
            float f[8];
            f[0] = __uint_as_float(v[8 * q + 0]) + b0.x; f[1] = __uint_as_float(v[8 * q + 1]) + b0.y;
            f[2] = __uint_as_float(v[8 * q + 2]) + b0.z; f[3] = __uint_as_float(v[8 * q + 3]) + b0.w;
            f[4] = __uint_as_float(v[8 * q + 4]) + b1.x; f[5] = __uint_as_float(v[8 * q + 5]) + b1.y;
            f[6] = __uint_as_float(v[8 * q + 6]) + b1.z; f[7] = __uint_as_float(v[8 * q + 7]) + b1.w;
            const int cc = hh * 4 + q;                               // 16-byte chunk inside the 128-byte row block
            uint4* slot = reinterpret_cast<uint4*>(rowp + ((cc ^ (arow & 7)) << 4));
            if (has_res) {
              const uint4 r4 = *slot;
              f[0] += bf16_lo(r4.x); f[1] += bf16_hi(r4.x); f[2] += bf16_lo(r4.y); f[3] += bf16_hi(r4.y);
              f[4] += bf16_lo(r4.z); f[5] += bf16_hi(r4.z); f[6] += bf16_lo(r4.w); f[7] += bf16_hi(r4.w);
            }
            if (relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            if (!valid) o = make_uint4(0u, 0u, 0u, 0u);   // halo positions stay zero for the next conv
            *slot = o;
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
      if (dbg) dbg[9] = clock64();
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
