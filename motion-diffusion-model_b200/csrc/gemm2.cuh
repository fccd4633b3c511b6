// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM:  D[M,N] = A[M,K] * W[N,K]^T, 256 x 256 tiles.
//
// Why pairs: at d_model = 512 the operands never leave the L2, and the L2 -> shared-memory fabric (~7 TB/s on this
// part, ~24 B/clk/SM) -- not the tensor pipe, not HBM -- bounds a 128 x 256 single-CTA tile (48 KB of operands per
// 512 MMA cycles = 96 B/clk).  With cta_group::2 the two SMs of a cluster execute ONE 256 x 256 x 16 UMMA per issue:
// each CTA stages its own 128 rows of A and only HALF of the W tile (128 of the 256 rows); the tensor cores read the
// other half from the peer's shared memory.  Operand traffic per CTA drops to 32 KB per 512 MMA cycles (2/3).
//
// Roles per CTA (192 threads, identical code in both CTAs of the pair):
//   warp 0     : TMA producer for ITS OWN shared memory (A rows of this CTA, its half of W); completion bytes of both
//                CTAs are reported to the LEADER's "full" barrier.
//   warp 1     : TMEM allocator (pair-wide allocation); in the leader also the single-thread MMA issuer.  tcgen05.commit
//                multicasts "slot free" / "accumulator ready" to the barriers of both CTAs.
//   warps 2..9 : epilogue of this CTA's 128 accumulator rows (same functors as gemm.cuh); accumulator release is
//                reported to the leader's barrier (remote mbarrier arrive).
#pragma once
#include "gemm.cuh"
#include "ptx.cuh"

namespace b200 {

// debug: clock64 stamps of cluster 0 (MMA issuer / first epilogue warp).  Compiled in only with -DB200_TRACE (the
// tools/trace_*.py build: B200MDM_TRACE=1 python -m b200mdm.build -> lib/libb200mdm_trace.so): in a production build the
// pointer test was a dependent global load at the head of every tile / item of every hot kernel (3.8 % + 2.9 % of the
// LayerNorm-GEMM's stall samples, profiles/r02_e_ncu_gemm_resid_ln_cluster.txt).
__device__ long long* g_gemm2_trace = nullptr;
#ifdef B200_TRACE
#define B200_TRACE_PTR(cond, ptr) ((b200::g_gemm2_trace != nullptr && (cond)) ? (ptr) : static_cast<long long*>(nullptr))
#else
#define B200_TRACE_PTR(cond, ptr) (static_cast<long long*>(nullptr))
#endif

#ifndef B200_GEMM2_EPI_WARPS
#define B200_GEMM2_EPI_WARPS 8
#endif
constexpr int GEMM2_EPI_WARPS = B200_GEMM2_EPI_WARPS;   // 8 or 16 (4 lane quarters x 2 or 4 column parts)
constexpr int GEMM2_THREADS = 64 + 32 * GEMM2_EPI_WARPS;
constexpr int GEMM2_BLOCK_N = 256;   // per pair: 256 x 256 output tile; per CTA: 128 rows x 256 columns of accumulator
constexpr int GEMM2_TILE_M = 256;

template <class Epi>
struct Gemm2Smem {
  static constexpr int A_BYTES = 128 * GEMM_BLOCK_K * 2;  // this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * GEMM_BLOCK_K * 2;  // this CTA's half (128 rows) of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
  static constexpr int EPI_BYTES = GEMM2_EPI_WARPS * Epi::SMEM_PER_WARP;
  static constexpr int budget = 227 * 1024 - 1024 - EPI_BYTES - GEMM_BIAS_BYTES - GEMM_BAR_BYTES;
  static constexpr int STAGES = (budget / STAGE_BYTES) > 6 ? 6 : (budget / STAGE_BYTES);
  static constexpr int TOTAL = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + GEMM_BIAS_BYTES + GEMM_BAR_BYTES;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm2_f16_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_c, int M, int N, int K,
                  const __grid_constant__ typename Epi::Params ep) {
  using SM = Gemm2Smem<Epi>;
  constexpr int STAGES = SM::STAGES;
  constexpr uint32_t ACC_STRIDE = 256;
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* epi_smem = smem + STAGES * SM::STAGE_BYTES;
  float* bias_all = reinterpret_cast<float*>(epi_smem + SM::EPI_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + SM::EPI_BYTES + GEMM_BIAS_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  (leader's copy is the live one)
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  (each CTA waits on its own)
  uint64_t* acc_full = bars + 2 * STAGES;       // [2]       (each CTA waits on its own)
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]       (leader's copy is the live one)
  uint64_t* epi_bars = bars + 2 * STAGES + 4;   // [GEMM2_EPI_WARPS][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bars + GEMM2_EPI_WARPS * 4);
  static_assert((2 * 6 + 4 + GEMM2_EPI_WARPS * 4) * 8 + 8 <= GEMM_BAR_BYTES, "barrier area too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (M + GEMM2_TILE_M - 1) / GEMM2_TILE_M;
  const int tiles_n = (N + GEMM2_BLOCK_N - 1) / GEMM2_BLOCK_N;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  pdl_launch_dependents();
  Epi::preload(ep, bias_all, N, threadIdx.x, blockDim.x);   // visible to the epilogue warps after the barrier below
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's expect_tx arrival; the peer only contributes transaction bytes
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);                     // one multicast commit
      mbar_init(&acc_empty[s], 2 * GEMM2_EPI_WARPS);   // the epilogue warps of both CTAs
    }
    for (int s = 0; s < GEMM2_EPI_WARPS * 4; ++s) mbar_init(&epi_bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated in both SMs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        const int a_row = m_blk * GEMM2_TILE_M + static_cast<int>(rank) * 128;
        const int b_row = n_blk * GEMM2_BLOCK_N + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * SM::STAGE_BYTES;
          uint8_t* sb = sa + SM::A_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          // The leader arms its barrier for the bytes of BOTH CTAs.  The peer never arrives: its bytes may even land
          // first (the transaction count simply goes negative until the leader's expect_tx is posted).
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * SM::STAGE_BYTES);
          tma_load_2d_2cta(sa, &map_a, leader_full, kb * GEMM_BLOCK_K, a_row);
          tma_load_2d_2cta(sb, &map_b, leader_full, kb * GEMM_BLOCK_K, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(256, GEMM2_BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        long long* tr = B200_TRACE_PTR(cluster_id == 0 && it < 8, g_gemm2_trace + it * 8);
        if (tr) tr[0] = clock64();
        mbar_wait_cluster(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        if (tr) tr[1] = clock64();
        const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * SM::STAGE_BYTES);
          const uint32_t sb = sa + SM::A_BYTES;
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_f16_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2cta_mc(&empty_bar[stage], 0b11);   // slot free in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta_mc(&acc_full[as], 0b11);         // accumulator ready in both CTAs
        if (tr) tr[2] = clock64();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9), both CTAs
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    EpiCtx ctx;
    ctx.smem = epi_smem + (warp - 2) * Epi::SMEM_PER_WARP;
    ctx.bars = epi_bars + (warp - 2) * 4;
    ctx.map_c = &map_c;
    ctx.bias_all = bias_all;
    ctx.lane = lane;
    ctx.M = M;
    ctx.N = N;
    ctx.seq = 0;
    ctx.primed = false;
    ctx.trace = nullptr;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * GEMM2_TILE_M + static_cast<int>(rank) * 128 + q * 32;
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      long long* tr = B200_TRACE_PTR(blockIdx.x == 0 && warp == 2 && lane == 0 && it < 8, g_gemm2_trace + it * 8);
      if (tr) tr[4] = clock64();
      if (tr) { mbar_wait(&acc_full[as], aphase); tr[5] = clock64(); }
      ctx.trace = (tr && it >= 2 && it < 6) ? g_gemm2_trace + 64 + (it - 2) * 16 : nullptr;
      epilogue_tile<GEMM2_BLOCK_N, Epi, GEMM2_EPI_WARPS / 4>(ctx, ep, taddr, row0, n_blk * GEMM2_BLOCK_N, part, &acc_full[as], aphase, [&]() {
        if (leader) mbar_arrive(&acc_empty[as]);
        else mbar_arrive_remote_relaxed(mapa_shared(smem_u32(&acc_empty[as]), 0));
      });
      if (tr) tr[6] = clock64();
    }
    Epi::finish(ctx);
  }

  // No CTA may exit (or free TMEM) while its peer can still read its shared memory or signal its barriers.
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

}  // namespace b200
