// sblk_encoder_stack.cuh — the WHOLE transformer encoder stack in one launch.
// Reference: Encoder.forward, SBL/transformer/encoder.py:36-67 (linear_in -> LayerNorm -> + positional encoding, then
// n_layers x EncoderLayer.forward, encoder.py:83-91 = MultiHeadAttention.forward attention.py:32-60 +
// ScaledDotProductAttention.forward attention.py:72-83 + PositionwiseFeedForward.forward module.py:47-52, with the
// `*= non_pad_mask` of encoder.py:86,89 and the masks of utils.py:98-113,140-147).
//
// Why one kernel: at the BASELINE batch (32 clips x 29 frames = 928 tokens) the stack is 25 dependent GEMM-shaped
// steps of ~1.5 GFLOP each; as separate launches every step costs ~10 us of launch / drain / fill latency, 280 us
// for 36 GFLOP.  But clips never interact inside the stack (attention is per clip, LayerNorm per token), so a
// group of floor(128 / T) whole clips (<= 128 token rows = one tcgen05 M tile) can run through all layers without ever
// synchronising with another group.  One thread-block CLUSTER of CL CTAs owns one group and splits every step along
// the OUTPUT FEATURES, so each CTA streams only 1/CL of the weights (L2-resident, 6.3 MB per layer):
//   stage      A operand (all 128 rows)   B operand (this CTA's weight rows)           epilogue
//   IN         x_in  [128, d_in]          w_in   rows [rank*NS, +NS)   NS = 512/CL      + bias -> LayerNorm -> + PE
//   QKV+attn   x16   [128, 512]           w_heads rows of head rank/PARTS (q|k|v, 192)  + bias -> smem Q/K/V -> softmax(QK^T)V
//   FC         att16 [128, 512]           w_fc   rows [rank*NS, +NS)                    + bias + residual -> LayerNorm -> mask
//   W1         x16   [128, 512]           w_1    rows [rank*NW, +NW)   NW = d_inner/CL  + bias -> ReLU
//   W2         h16   [128, d_inner]       w_2    rows [rank*NS, +NS)                    + bias + residual -> LayerNorm -> mask
// Activations travel between stages through small L2-resident global buffers (x16 / att16 / h16, bf16; the fp32
// residual stream lives in the output buffer) published with cluster barriers (release / acquire, plus proxy fences
// because the next stage reads them with TMA); LayerNorm row statistics are combined across the cluster through
// distributed shared memory (per-CTA mean and M2, Chan's parallel-variance merge: one exchange, two-pass accuracy).
// With MC the A tiles are loaded ONCE per cluster: CTA r fetches rows [r*128/CL, ...) of every k-block and multicasts
// them into all CL CTAs' rings (ring slots are then released cluster-wide by a multicast tcgen05.commit).
//
// Two clip groups per cluster (EncStackParams::gpc == 2).  The chain of one group is latency-bound: per layer ~10 us of
// operand delivery + MMAs, ~7 us of epilogues and ~6 us of cluster barriers, each waiting for the previous one.  With
// gpc == 2 a cluster runs groups A and B through the same stages alternately — virtual stage v = 2 * stage + group —
// and the producer and MMA warps work ONE virtual stage ahead of the barrier sequence: while the epilogue warps
// finish (stage, A) and wait in its barriers, the operands of (stage, B) stream in and its MMAs run (second TMEM
// accumulator), and vice versa.  Every (stage, group) keeps its own cluster barriers in the same global order, the
// LayerNorm epilogues of A / B live in the two halves of the epilogue warps (each keeps ITS group's residual-stream
// slice in registers), and all arithmetic is per group exactly as with gpc == 1: results are bit-identical, the
// cluster just serves two groups in little more than the time of one.
//
// Roles (192 threads): warp 0 TMA producer (prefetches the next stage's weight tiles BEFORE the barrier that publishes
// its activations), warp 1 MMA issuer / TMEM owner, warps 2-5 epilogue (one token row per thread) + attention.
#pragma once
#include "sblk_common.cuh"
#include "sblk_igemm2.cuh"
#include "sblk_gemm_ln.cuh"
#include "sblk_attention.cuh"

namespace sblk {

struct EncStackParams {
  int N, T, G, L;            // clips, frames per clip, clips per group (cluster), layers
  int M;                     // N * T token rows
  int d_in;                  // input feature width (K of linear_in), multiple of 128
  int d_inner;               // FFN width, multiple of 16 * CL
  const float* b_in;         // [512]
  const float* g_in;         // [512] layer_norm_in weight
  const float* be_in;        // [512] layer_norm_in bias
  const float* pe;           // [>= T, 512]
  const float* b_heads;      // [L][8*192] head-major q|k|v biases
  const float* b_fc;         // [L][512]
  const float* g1;           // [L][512] slf_attn.layer_norm
  const float* be1;
  const float* b_w1;         // [L][d_inner]
  const float* b_w2;         // [L][512]
  const float* g2;           // [L][512] pos_ffn.layer_norm
  const float* be2;
  const int* lengths;        // [N] or nullptr (all T)
  float* out;                // [M, 512] fp32: residual stream while the stack runs, enc_output at the end
  enc16_t* x16;        // [M, 512]      workspace: bf16 copy of the residual stream (GEMM operand)
  enc16_t* att16;      // [M, 512]      workspace: concatenated attention heads
  enc16_t* h16;        // [M, d_inner]  workspace: relu(w_1 x)
  float scale;               // 1 / temperature
  float eps;                 // LayerNorm eps (all three LayerNorms of the reference use the default 1e-5)
  unsigned long long* dbg;   // optional [stages][8] clock64 stamps of cluster 0 / CTA 0 (profiling aid), or nullptr
  unsigned int* resident;    // optional: every CTA adds 1 when it starts running (sblk_gate_wait: co-scheduling hint)
  int gpc;                   // clip groups per cluster: 1, or 2 (interleaved, see above)
};

template <int CL>
struct EncCfg {
  static constexpr int D = 512;
  static constexpr int H = 8;
  static constexpr int PARTS = CL / H;                 // CTAs sharing one head in the QKV + attention stage
  static constexpr int NS = D / CL;                    // LayerNorm-stage output columns per CTA
  static constexpr int QKV_N = 192;
  static constexpr int A_BYTES = 128 * 128;            // one k-block (64 bf16) of the 128-row A tile
  static constexpr int B_ROWS_MAX = CL == 16 ? 192 : 256;
  static constexpr int SLOT_BYTES = A_BYTES + B_ROWS_MAX * 128;
  static constexpr int STAGES = CL == 16 ? 4 : 3;
  static constexpr int QKV_ROWS = 144;                 // 128 tile rows + 16 zero rows (key padding of the last clip)
  static constexpr int QKV_BYTES = QKV_ROWS * 128;
  static constexpr int OFF_QKV = STAGES * SLOT_BYTES;
  static constexpr int SMEM_BYTES = OFF_QKV + 3 * QKV_BYTES + 1024;
  static constexpr int EPI_WARPS = 8;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int VEC_FLOATS = QKV_N + B_ROWS_MAX + 6 * NS;   // per-layer bias / gamma / beta slices of this CTA
  static constexpr int TMEM_COLS = 512;                // one 256-column accumulator per interleaved clip group
  static constexpr int MC_ROWS = 128 / CL;             // A rows each CTA fetches and multicasts (8-row swizzle atoms)
  static_assert(2 * (A_BYTES + NS * 128) <= SLOT_BYTES, "two LayerNorm-stage k-blocks must fit one ring slot");
  static_assert(16 * 128 * 8 <= QKV_BYTES, "statistics exchange area aliases the Q tile");
};

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive without release semantics (threads that published nothing: producer, MMA issuer, idle epilogue warps); the
// releasing form compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive
__device__ __forceinline__ void cl_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// generic-proxy global writes -> visible to async-proxy (TMA) reads issued after the next acquire
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_v2f32(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
// 2-D tiled TMA load multicast to every CTA of `mask` (same smem offset and same mbarrier offset in each of them)
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* d, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the same-offset mbarrier of every CTA in `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// Coalesced row stores through a per-warp staging tile: lane l holds one whole row segment in registers (its token
// row), but 32 lanes x 16 B at a row pitch of 1-4 KB is 32 cache lines per store instruction (LSU-bound: ncu /
// clock stamps showed 1.9 us for a 48 KB tile).  The segment goes to shared memory (16-byte chunks XOR-swizzled,
// conflict-free both ways) and comes back transposed, so one instruction writes 4 (8) complete 128 (64) byte rows.
__device__ __forceinline__ void warp_store_rows128(uint8_t* stg, const uint4 (&d)[8], uint8_t* gbase, size_t pitch,
                                                   int valid_rows, int lane) {
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = d[c];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3), c = lane & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
    if (r < valid_rows) *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(r) * pitch + c * 16) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_store_rows64(uint8_t* stg, const uint4 (&d)[4], uint8_t* gbase, size_t pitch,
                                                  int valid_rows, int lane) {
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = d[c];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), c = lane & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    if (r < valid_rows) *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(r) * pitch + c * 16) = v;
  }
  __syncwarp();
}

struct EncStage {
  const CUtensorMap* tmA;
  const CUtensorMap* tmB;
  int b_row;      // first weight row of this CTA's B tile
  int n;          // B tile rows = output columns of this CTA
  int num_kb;     // K / 64
  int kbps;       // k-blocks per ring slot (2 for the narrow LayerNorm stages)
  int kind;       // 0 IN, 1 QKV, 2 FC, 3 W1, 4 W2
  int layer;
  int end_barriers;   // cluster barriers every thread executes at the end of this stage
};

template <int CL, int NT, bool MC>
__global__ void __launch_bounds__(EncCfg<CL>::THREADS, 1)
encoder_stack_kernel(const __grid_constant__ CUtensorMap tmXin, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ CUtensorMap tmAtt, const __grid_constant__ CUtensorMap tmH,
                     const __grid_constant__ CUtensorMap tmWin, const __grid_constant__ CUtensorMap tmWh,
                     const __grid_constant__ CUtensorMap tmWfc, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const EncStackParams p) {
  using Cfg = EncCfg<CL>;
  constexpr int D = Cfg::D;
  constexpr int NS = Cfg::NS;
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint16_t ALL = static_cast<uint16_t>((1u << CL) - 1u);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar[2];   // per interleaved group
  __shared__ uint32_t tmem_base_slot;
  __shared__ float vec_layer[2][Cfg::VEC_FLOATS];   // double-buffered per layer: b_heads | b_w1 | b_fc g1 be1 | b_w2 g2 be2
  __shared__ float vec_in[3 * NS];                  // b_in | layer_norm_in gamma | beta

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQKV = smem + Cfg::OFF_QKV;
  float2* part = reinterpret_cast<float2*>(sQKV);   // [16][128] (mean, M2) per 32-column slice; aliases the Q tile

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int T = p.T;
  const int total_groups = (p.N + p.G - 1) / p.G;
  const int g_first = (static_cast<int>(blockIdx.x) / CL) * p.gpc;   // first clip group of this cluster
  const int ng = min(p.gpc, total_groups - g_first);                  // groups this cluster interleaves (1 or 2)
  const int NW = p.d_inner / CL;
  const int head = rank / Cfg::PARTS;
  const int num_stages = 1 + 4 * p.L;
  const int NV = num_stages * ng;                                     // virtual stages: v = stage * ng + group slot
  unsigned long long* const dbg = (p.dbg != nullptr && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr;
  auto stamp = [&](int v, int which) {   // stamps follow the cluster's first group only
    if (dbg != nullptr && v % ng == 0) dbg[(v / ng) * 8 + which] = static_cast<unsigned long long>(clock64());
  };

  auto stage_desc = [&](int s) -> EncStage {
    EncStage d;
    if (s == 0) {
      d.tmA = &tmXin; d.tmB = &tmWin; d.b_row = rank * NS; d.n = NS; d.num_kb = p.d_in / 64; d.kbps = 2; d.kind = 0;
      d.layer = 0; d.end_barriers = 2;
      return d;
    }
    const int l = (s - 1) >> 2;
    const int k = ((s - 1) & 3) + 1;
    d.kind = k; d.layer = l;
    if (k == 1) {
      d.tmA = &tmX; d.tmB = &tmWh; d.b_row = l * (Cfg::H * Cfg::QKV_N) + head * Cfg::QKV_N; d.n = Cfg::QKV_N;
      d.num_kb = D / 64; d.kbps = 1; d.end_barriers = 1;
    } else if (k == 2) {
      d.tmA = &tmAtt; d.tmB = &tmWfc; d.b_row = l * D + rank * NS; d.n = NS; d.num_kb = D / 64; d.kbps = 2;
      d.end_barriers = 2;
    } else if (k == 3) {
      d.tmA = &tmX; d.tmB = &tmW1; d.b_row = l * p.d_inner + rank * NW; d.n = NW; d.num_kb = D / 64; d.kbps = 1;
      d.end_barriers = 1;
    } else {
      d.tmA = &tmH; d.tmB = &tmW2; d.b_row = l * D + rank * NS; d.n = NS; d.num_kb = p.d_inner / 64; d.kbps = 2;
      d.end_barriers = 2;
    }
    return d;
  };

  if (threadIdx.x == 0) {
    // a running CTA means its whole cluster has been placed; a gate kernel on another stream (sblk_gate_wait) holds a
    // co-running kernel chain back until every cluster of this launch owns its SMs
    if (p.resident != nullptr) atomicAdd(p.resident, 1u);
    tma_prefetch_desc(&tmXin); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmAtt); tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmWin); tma_prefetch_desc(&tmWh); tma_prefetch_desc(&tmWfc); tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], MC ? CL : 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
  if (warp >= 2) {
    // zero the 16 padding rows behind each of the Q / K / V tiles (read as masked keys of the group's last clip)
    const int etid = threadIdx.x - 64;
    for (int i = etid; i < 3 * 16 * 8; i += Cfg::EPI_THREADS) {
      const int which = i / 128;
      const int r = (i >> 3) & 15;
      *reinterpret_cast<uint4*>(sQKV + which * Cfg::QKV_BYTES + (128 + r) * 128 + ((i & 7) << 4)) =
          make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  // every CTA of the cluster is running with its barriers initialised before any peer multicasts / stores into it
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_slot, 0);

  grid_dep_launch();  // PDL: let the next kernel start its prologue now; its own wait orders the data
  grid_dep_wait();
  if (p.dbg != nullptr && rank == 0 && threadIdx.x == 0) {   // per-cluster start time (ns) behind the stage stamps
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    p.dbg[num_stages * 8 + g_first * 2] = now;
  }

  // Cluster-barrier participation of the producer and MMA warps: every thread of the cluster has to arrive at every
  // cluster barrier, but these two warps publish nothing, so each ARRIVES at barrier k + 1 right after its wait on
  // barrier k (the earliest the ISA allows) and WAITS only when it needs the guarantee (need(k): k barriers completed).
  // The epilogue warps of one clip group then do not wait for this warp's ring-paced load / issue loops of the other
  // group.  (Servicing the barriers from inside those polling loops as well — a "barrier pump" — was measured: no gain,
  // the remaining serialisation is the epilogue chain itself and the memory-system contention between the groups.)
  int bar_total = 0;
  for (int v = 0; v < NV; ++v) bar_total += stage_desc(v / ng).end_barriers;
  int bar_done = 0;
  auto pump_start = [&]() { if (bar_total > 0) cl_arrive_relaxed(); };
  auto need = [&](int k) {
    while (bar_done < k) {
      cl_wait();
      if (++bar_done < bar_total) cl_arrive_relaxed();
    }
  };

  if (warp == 0) {
    // ================================================================= TMA producer
    int slot = 0;
    uint32_t phase = 0;
    // reserve up to STAGES ring slots for stage d and start their weight-tile loads; returns the slots reserved
    auto prefetch_b = [&](const EncStage& d) -> int {
      const int units = d.num_kb / d.kbps;
      const int pre = units < STAGES ? units : STAGES;
      const uint32_t b_bytes = static_cast<uint32_t>(d.n) * 128u;
      int s = slot;
      uint32_t ph = phase;
      for (int u = 0; u < pre; ++u) {
        mbar_wait(&empty_bar[s], ph ^ 1u, 0x0701);
        uint8_t* base = smem + s * Cfg::SLOT_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(d.kbps) * (Cfg::A_BYTES + b_bytes));
          for (int i = 0; i < d.kbps; ++i)
            tma_load_2d(base + d.kbps * Cfg::A_BYTES + i * b_bytes, d.tmB, &full_bar[s], (u * d.kbps + i) * 64,
                        d.b_row);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      return pre;
    };
    auto load_a = [&](const EncStage& d, uint8_t* base, uint64_t* bar, int u, int m0) {
      for (int i = 0; i < d.kbps; ++i) {
        const int kc = (u * d.kbps + i) * 64;
        if (MC) {
          tma_load_2d_mc(base + i * Cfg::A_BYTES + rank * (Cfg::MC_ROWS * 128), d.tmA, bar, kc,
                         m0 + rank * Cfg::MC_ROWS, ALL);
        } else {
          tma_load_2d(base + i * Cfg::A_BYTES, d.tmA, bar, kc, m0);
        }
      }
    };
    auto finish_loads = [&](const EncStage& d, int pre, int m0) {
      const int units = d.num_kb / d.kbps;
      const uint32_t b_bytes = static_cast<uint32_t>(d.n) * 128u;
      for (int u = 0; u < units; ++u) {
        uint8_t* base = smem + slot * Cfg::SLOT_BYTES;
        if (u >= pre) mbar_wait(&empty_bar[slot], phase ^ 1u, 0x0702);
        if (elect_one()) {
          if (u >= pre) {
            mbar_arrive_expect_tx(&full_bar[slot], static_cast<uint32_t>(d.kbps) * (Cfg::A_BYTES + b_bytes));
            for (int i = 0; i < d.kbps; ++i)
              tma_load_2d(base + d.kbps * Cfg::A_BYTES + i * b_bytes, d.tmB, &full_bar[slot],
                          (u * d.kbps + i) * 64, d.b_row);
          }
          load_a(d, base, &full_bar[slot], u, m0);
        }
        __syncwarp();
        if (++slot == STAGES) { slot = 0; phase ^= 1u; }
      }
    };

    // virtual stage v = stage * ng + group slot; the loads of v + ng (same group, next stage) are issued right after
    // the barriers of v, i.e. (ng == 2) one virtual stage AHEAD of the barrier sequence
    auto v_m0 = [&](int v) { return (g_first + v % ng) * p.G * T; };
    for (int v = 0; v < ng; ++v) {
      const EncStage d0 = stage_desc(0);
      const int pre0 = prefetch_b(d0);
      stamp(v, 0);
      finish_loads(d0, pre0, v_m0(v));
      stamp(v, 5);
    }
    pump_start();
    int cum = 0;   // barriers up to and including virtual stage v
    for (int v = 0; v < NV; ++v) {
      cum += stage_desc(v / ng).end_barriers;
      if (v + ng >= NV) continue;
      const EncStage nxt = stage_desc((v + ng) / ng);
      const int pre = prefetch_b(nxt);          // weight tiles do not depend on the barrier
      need(cum);                                // the activations of (stage, group) are published
      fence_proxy_async_all();
      stamp(v + ng, 0);
      finish_loads(nxt, pre, v_m0(v + ng));
      stamp(v + ng, 5);
    }
    need(bar_total);
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    int slot = 0;
    uint32_t phase = 0;
    auto issue = [&](int v) {
      const EncStage d = stage_desc(v / ng);
      const int gi = v % ng;
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(gi * 256);
      const int units = d.num_kb / d.kbps;
      const uint32_t idesc = make_idesc_e16(128, d.n);
      const uint32_t b_bytes = static_cast<uint32_t>(d.n) * 128u;
      tc_fence_after_sync();
      for (int u = 0; u < units; ++u) {
        mbar_wait(&full_bar[slot], phase, 0x0703);
        tc_fence_after_sync();
        const uint32_t base = smem_u32(smem + slot * Cfg::SLOT_BYTES);
        const uint64_t da = make_desc_sw128(base);
        const uint64_t db = make_desc_sw128(base + d.kbps * Cfg::A_BYTES);
        const uint32_t da_lo = static_cast<uint32_t>(da), db_lo = static_cast<uint32_t>(db);
        const uint32_t b_step = b_bytes >> 4;
        const bool last = u == units - 1;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, desc_with_lo(da, da_lo + 2 * k), desc_with_lo(db, db_lo + 2 * k), idesc,
                      (u > 0 || k > 0) ? 1u : 0u);
          if (d.kbps == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, desc_with_lo(da, da_lo + (Cfg::A_BYTES >> 4) + 2 * k),
                        desc_with_lo(db, db_lo + b_step + 2 * k), idesc, 1u);
          }
          if (MC) umma_commit_mc(&empty_bar[slot], ALL); else umma_commit(&empty_bar[slot]);
          if (last) umma_commit(&tfull_bar[gi]);
        }
        __syncwarp();
        if (++slot == STAGES) { slot = 0; phase ^= 1u; }
      }
      stamp(v, 1);
    };
    // like the producer: the MMAs of v + ng are issued after the barriers of v (which also guarantee that the epilogue
    // has drained that group's accumulator)
    for (int v = 0; v < ng; ++v) issue(v);
    pump_start();
    int cum = 0;
    for (int v = 0; v < NV; ++v) {
      cum += stage_desc(v / ng).end_barriers;
      if (v + ng >= NV) continue;
      need(cum);   // the epilogue has drained this group's accumulator
      issue(v + ng);
    }
    need(bar_total);
  } else {
    // ================================================================= epilogue warps
    // 8 warps: warp pairs (ew, ew + 4) share a TMEM lane quarter = 32 token rows.  The LayerNorm stages run on four of
    // them (one row per thread, NS columns in registers): the first four, or — two interleaved groups — the first four
    // for group A and the other four for group B, so that every thread keeps exactly ONE group's residual-stream
    // slice in registers; the wide stages (QKV tiles, attention, W1) use all 8.
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int t = row % T;
    const int etid = static_cast<int>(threadIdx.x) - 64;
    uint32_t tph0 = 0, tph1 = 0;   // accumulator-full phase per interleaved group
    const uint32_t my_part = smem_u32(&part[rank * (NS / 32) * 128 + row]);   // 32-column partial rank*NS/32 (+ sub)
    const int col0 = rank * NS;
    uint8_t* const stg_ln = sQKV + 16 * 128 * 8 + (ew & 3) * 4096;       // behind the statistics area (LayerNorm stages)
    uint8_t* const stg_w1 = sQKV + ew * 4096;                            // W1 stage: the Q/K/V tiles are idle
    float xres[NS];   // this thread's slice of the fp32 residual stream: lives in registers across all stages
#pragma unroll
    for (int j = 0; j < NS; ++j) xres[j] = 0.0f;

    // per-column vectors of one layer for this CTA's column slices, staged in shared memory one layer ahead
    auto load_layer_vectors = [&](int l, float* dst) {
      const int total = Cfg::QKV_N + NW + 6 * NS;
      for (int e = etid; e < total; e += Cfg::EPI_THREADS) {
        float val;
        if (e < Cfg::QKV_N) {
          val = __ldg(p.b_heads + l * (Cfg::H * Cfg::QKV_N) + head * Cfg::QKV_N + e);
        } else if (e < Cfg::QKV_N + NW) {
          val = __ldg(p.b_w1 + static_cast<size_t>(l) * p.d_inner + rank * NW + (e - Cfg::QKV_N));
        } else {
          const int k = (e - Cfg::QKV_N - NW) / NS;
          const int j = (e - Cfg::QKV_N - NW) - k * NS;
          const float* src = k == 0 ? p.b_fc : k == 1 ? p.g1 : k == 2 ? p.be1 : k == 3 ? p.b_w2 : k == 4 ? p.g2 : p.be2;
          val = __ldg(src + l * D + col0 + j);
        }
        dst[e] = val;
      }
    };
    if (etid < 3 * NS) {
      const int k = etid / NS, j = etid - k * NS;
      vec_in[etid] = __ldg((k == 0 ? p.b_in : k == 1 ? p.g_in : p.be_in) + col0 + j);
    }
    if (p.L > 0) load_layer_vectors(0, vec_layer[0]);
    asm volatile("bar.sync 1, 256;" ::: "memory");

    for (int vs = 0; vs < NV; ++vs) {
      const int s = vs / ng;
      const int gi = vs - s * ng;
      const EncStage d = stage_desc(s);
      const float* lv = vec_layer[d.layer & 1];
      // this virtual stage's clip group
      const int group = g_first + gi;
      const int m0 = group * p.G * T;
      const int nclips = min(p.G, p.N - group * p.G);
      const int rows_valid = nclips * T;
      const bool row_ok = row < rows_valid;
      const int clip = group * p.G + row / T;
      const int warp_valid = min(max(rows_valid - quarter * 32, 0), 32);   // valid token rows of this warp's lane quarter
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(gi * 256);
      uint64_t* const tfull = &tfull_bar[gi];
      const uint32_t tph = gi == 0 ? tph0 : tph1;
      if (d.kind == 0 || d.kind == 2 || d.kind == 4) {
        // ------------------------------------------------ + bias (+ residual) -> LayerNorm (+ PE | * pad mask)
        if (half == (ng == 2 ? gi : 0)) {
          float keep = 1.0f;
          if (row_ok && p.lengths != nullptr && t >= __ldg(p.lengths + clip)) keep = 0.0f;
          const float* bias = d.kind == 0 ? vec_in : lv + Cfg::QKV_N + NW + (d.kind == 2 ? 0 : 3 * NS);
          const float* gamma = bias + NS;
          const float* beta = bias + 2 * NS;
          float v[NS];
#pragma unroll
          for (int j = 0; j < NS; ++j) v[j] = d.kind != 0 ? xres[j] : 0.0f;
          mbar_wait(tfull, tph, 0x0704);
          tc_fence_after_sync();
          if (ew == 0) stamp(vs, 2);
#pragma unroll
          for (int c = 0; c < NS / 32; ++c) {
            uint32_t u[32];
            tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] += __uint_as_float(u[j]) + bias[c * 32 + j];
          }
          tc_fence_before_sync();
          // row statistics in 32-column partials (one per sub-slice), whatever the cluster size: the merged mean / M2
          // are then bit-identical for clusters of 8 and 16, so a clip's output does not depend on the batch it is in
#pragma unroll
          for (int sub = 0; sub < NS / 32; ++sub) {
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += v[sub * 32 + j];
            const float mean_c = sum * (1.0f / 32.0f);
            float m2_c = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float dv = v[sub * 32 + j] - mean_c;
              m2_c += dv * dv;
            }
#pragma unroll
            for (uint32_t r = 0; r < static_cast<uint32_t>(CL); ++r)
              st_cluster_v2f32(mapa_u32(my_part + static_cast<uint32_t>(sub * 128 * 8), r), mean_c, m2_c);
          }
          if (ew == 0) stamp(vs, 4);
          cl_arrive();
          float pev[NS];   // positional encoding row (IN stage only): in flight while the barrier completes
          if (d.kind == 0) {
            const float4* ep = reinterpret_cast<const float4*>(p.pe + static_cast<size_t>(t) * D + col0);
#pragma unroll
            for (int j = 0; j < NS / 4; ++j) {
              const float4 e = __ldg(ep + j);
              pev[4 * j + 0] = e.x; pev[4 * j + 1] = e.y; pev[4 * j + 2] = e.z; pev[4 * j + 3] = e.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) pev[j] = 0.0f;
          }
          cl_wait();
          if (ew == 0) stamp(vs, 6);
          float mean = 0.0f;
#pragma unroll
          for (int r = 0; r < 16; ++r) mean += part[r * 128 + row].x;
          mean *= (1.0f / 16.0f);
          float m2 = 0.0f;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float2 pr = part[r * 128 + row];
            const float dm = pr.x - mean;
            m2 += pr.y + 32.0f * dm * dm;
          }
          const float rstd = rsqrtf(m2 * (1.0f / D) + p.eps);
          const float mul = d.kind == 0 ? 1.0f : keep;   // encoder.py:53-55 applies no pad mask after layer_norm_in
#pragma unroll
          for (int j = 0; j < NS; ++j) v[j] = (((v[j] - mean) * rstd) * gamma[j] + beta[j] + pev[j]) * mul;
#pragma unroll
          for (int j = 0; j < NS; ++j) xres[j] = v[j];
          if (s != num_stages - 1) {
            // bf16 copy of the stream for the next GEMM
            uint8_t* gb = reinterpret_cast<uint8_t*>(p.x16 + static_cast<size_t>(m0 + quarter * 32) * D + col0);
            if (NS == 32) {
              uint4 o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                o[j].x = pack_e16x2(v[8 * j + 0], v[8 * j + 1]);
                o[j].y = pack_e16x2(v[8 * j + 2], v[8 * j + 3]);
                o[j].z = pack_e16x2(v[8 * j + 4], v[8 * j + 5]);
                o[j].w = pack_e16x2(v[8 * j + 6], v[8 * j + 7]);
              }
              warp_store_rows64(stg_ln, o, gb, D * 2, warp_valid, lane);
            } else {
              uint4 o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                o[j].x = pack_e16x2(v[(8 * j + 0) % NS], v[(8 * j + 1) % NS]);
                o[j].y = pack_e16x2(v[(8 * j + 2) % NS], v[(8 * j + 3) % NS]);
                o[j].z = pack_e16x2(v[(8 * j + 4) % NS], v[(8 * j + 5) % NS]);
                o[j].w = pack_e16x2(v[(8 * j + 6) % NS], v[(8 * j + 7) % NS]);
              }
              warp_store_rows128(stg_ln, o, gb, D * 2, warp_valid, lane);
            }
          } else {
            // last stage: the fp32 stream is the encoder output
            uint8_t* gb = reinterpret_cast<uint8_t*>(p.out + static_cast<size_t>(m0 + quarter * 32) * D + col0);
#pragma unroll
            for (int h2 = 0; h2 < NS / 32; ++h2) {
              uint4 o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                o[j] = make_uint4(__float_as_uint(v[h2 * 32 + 4 * j]), __float_as_uint(v[h2 * 32 + 4 * j + 1]),
                                  __float_as_uint(v[h2 * 32 + 4 * j + 2]), __float_as_uint(v[h2 * 32 + 4 * j + 3]));
              warp_store_rows128(stg_ln, o, gb + h2 * 128, D * 4, warp_valid, lane);
            }
          }
          if (ew == 0) stamp(vs, 7);
          fence_proxy_async_all();
          cl_arrive();
          cl_wait();
        } else {
          cl_arrive_relaxed(); cl_wait();
          cl_arrive_relaxed(); cl_wait();
        }
      } else if (d.kind == 1) {
        // ------------------------------------------------ + bias -> bf16 Q / K / V tiles -> attention of this head
        // (next layer's per-column vectors are staged now, while the projection GEMM is still running)
        if (gi == 0 && d.layer + 1 < p.L) load_layer_vectors(d.layer + 1, vec_layer[(d.layer + 1) & 1]);
        mbar_wait(tfull, tph, 0x0705);
        tc_fence_after_sync();
        if (ew == 0) stamp(vs, 2);
#pragma unroll 1
        for (int c3 = 0; c3 < 3; ++c3) {
          const int cc = half * 3 + c3;        // 32-column chunk 0..5: q q k k v v
          uint32_t u[32];
          tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(cc * 32), u);
          tmem_ld_wait();
          const float* bp = lv + cc * 32;
          uint8_t* dst_row = sQKV + (cc >> 1) * Cfg::QKV_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);   // rows past the group's clips hold zeros (finite masked keys)
            if (row_ok) {
              o.x = pack_e16x2(__uint_as_float(u[8 * j + 0]) + bp[8 * j + 0], __uint_as_float(u[8 * j + 1]) + bp[8 * j + 1]);
              o.y = pack_e16x2(__uint_as_float(u[8 * j + 2]) + bp[8 * j + 2], __uint_as_float(u[8 * j + 3]) + bp[8 * j + 3]);
              o.z = pack_e16x2(__uint_as_float(u[8 * j + 4]) + bp[8 * j + 4], __uint_as_float(u[8 * j + 5]) + bp[8 * j + 5]);
              o.w = pack_e16x2(__uint_as_float(u[8 * j + 6]) + bp[8 * j + 6], __uint_as_float(u[8 * j + 7]) + bp[8 * j + 7]);
            }
            const int chunk = (cc & 1) * 4 + j;
            *reinterpret_cast<uint4*>(dst_row + ((chunk ^ (row & 7)) << 4)) = o;
          }
        }
        tc_fence_before_sync();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const uint32_t sQ_u = smem_u32(sQKV), sK_u = sQ_u + Cfg::QKV_BYTES, sV_u = sK_u + Cfg::QKV_BYTES;
        const int mt_count = (T + 15) >> 4;
        const int units = nclips * mt_count;
        for (int un = (rank % Cfg::PARTS) + Cfg::PARTS * ew; un < units; un += Cfg::PARTS * Cfg::EPI_WARPS) {
          const int c = un / mt_count;
          const int mt = un - c * mt_count;
          const int b = group * p.G + c;
          const int len = (p.lengths != nullptr) ? min(max(__ldg(p.lengths + b), 0), T) : T;
          enc16_t* out_clip = p.att16 + static_cast<size_t>(b) * T * D + head * 64;
          attention_mtile<NT>(sQ_u, sK_u, sV_u, c * T + mt * 16, c * T, mt * 16, T, len, p.scale, lane, out_clip, D,
                              nullptr);
        }
        // the Q tile doubles as the statistics exchange area of the next stage: every warp is done reading it
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0) stamp(vs, 4);
        fence_proxy_async_all();
        cl_arrive();
        cl_wait();
      } else {
        // ------------------------------------------------ W1: + bias -> ReLU -> bf16 h
        const float* bias = lv + Cfg::QKV_N;
        uint8_t* hbase = reinterpret_cast<uint8_t*>(p.h16 + static_cast<size_t>(m0 + quarter * 32) * p.d_inner +
                                                    rank * NW);
        mbar_wait(tfull, tph, 0x0706);
        tc_fence_after_sync();
        if (ew == 0) stamp(vs, 2);
#pragma unroll 1
        for (int blk = half; blk < NW / 64; blk += 2) {   // 64-column blocks, alternating between the two warps
          uint4 o[8];
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t u[32];
            tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(blk * 64 + c2 * 32), u);
            tmem_ld_wait();
            const float* bp = bias + blk * 64 + c2 * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[c2 * 4 + j].x = pack_e16x2(fmaxf(__uint_as_float(u[8 * j + 0]) + bp[8 * j + 0], 0.0f),
                                            fmaxf(__uint_as_float(u[8 * j + 1]) + bp[8 * j + 1], 0.0f));
              o[c2 * 4 + j].y = pack_e16x2(fmaxf(__uint_as_float(u[8 * j + 2]) + bp[8 * j + 2], 0.0f),
                                            fmaxf(__uint_as_float(u[8 * j + 3]) + bp[8 * j + 3], 0.0f));
              o[c2 * 4 + j].z = pack_e16x2(fmaxf(__uint_as_float(u[8 * j + 4]) + bp[8 * j + 4], 0.0f),
                                            fmaxf(__uint_as_float(u[8 * j + 5]) + bp[8 * j + 5], 0.0f));
              o[c2 * 4 + j].w = pack_e16x2(fmaxf(__uint_as_float(u[8 * j + 6]) + bp[8 * j + 6], 0.0f),
                                            fmaxf(__uint_as_float(u[8 * j + 7]) + bp[8 * j + 7], 0.0f));
            }
          }
          warp_store_rows128(stg_w1, o, hbase + blk * 128, static_cast<size_t>(p.d_inner) * 2, warp_valid, lane);
        }
        tc_fence_before_sync();
        if (ew == 0) stamp(vs, 4);
        fence_proxy_async_all();
        cl_arrive();
        cl_wait();
      }
      if (gi == 0) tph0 ^= 1u; else tph1 ^= 1u;
      if (ew == 0) stamp(vs, 3);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (p.dbg != nullptr && rank == 0 && threadIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    p.dbg[num_stages * 8 + g_first * 2 + 1] = now;
  }
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace sblk
