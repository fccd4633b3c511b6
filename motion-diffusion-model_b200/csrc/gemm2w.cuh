// W-resident variant of the CTA-pair GEMM (gemm2.cuh):  out16[M,N] = fp16(act(A[M,K] * W[N,K]^T + bias)),  K <= 512.
//
// Why: the layer GEMMs of this model never leave the L2, and what bounds gemm2_f16_tcgen05 at K = 512 is the
// L2 -> shared-memory fabric, not the tensor pipe: every 256 x 256 pair tile makes each CTA pull 128 KB of A and
// 128 KB of W (its half of the W tile) for 4096 cycles of MMA -- 62 B/clk against the ~24-32 B/clk an SM gets when
// all 148 stream at once (FFN up-projection at B = 64: 203 MB of operands per launch in 35.6 us = 5.7 TB/s, 11.7 k
// cycles per tile; profiles/r02_e_ncu_gemm2_f16_tcgen05.txt).  At K = 512 one CTA's half of a W tile is exactly
// 128 KB, so it can simply STAY in shared memory: a cluster is bound to one column block n_blk for the whole launch,
// loads its W half once (k-block by k-block, so the first tile still pipelines like the streaming kernel) and then
// streams only A.  Operand traffic of the FFN up-projection drops from 203 MB to 120 MB per launch, 128 KB per tile
// and CTA (32 B/clk for 4096 MMA cycles).
//
// Tile order: cluster c owns column block c % tiles_n and, among the cnt clusters of that block, every cnt-th row
// block starting at c / tiles_n.  The launcher only picks this kernel when that costs no extra round of tiles
// compared with the strided order of gemm2_f16_tcgen05 (engine.cu: gemm2w_pays).
//
// Shared memory (227 KB): 128 KB resident W half | 4 x 16 KB A stages | 8 x 4 KB epilogue slabs | 1 KB bias of the
// cluster's 256 columns | barriers.  Roles, barriers and the epilogue are those of gemm2.cuh; the extra barriers
// w_full[kb] (leader's copy live, armed once) gate the MMAs of the first tile only.
#pragma once
#include "epilogues.cuh"
#include "gemm2.cuh"

namespace b200 {

constexpr int GEMM2W_KB_MAX = 8;    // resident k-blocks of 64: K <= 512
constexpr int GEMM2W_STAGES = 4;
constexpr int GEMM2W_BAR_BYTES = 512;
constexpr int GEMM2W_BIAS_BYTES = GEMM2_BLOCK_N * 4;

struct Gemm2wSmem {
  static constexpr int A_BYTES = 128 * GEMM_BLOCK_K * 2;      // this CTA's 128 rows of A, one k-block
  static constexpr int W_KB_BYTES = 128 * GEMM_BLOCK_K * 2;   // this CTA's half (128 rows) of the W tile, one k-block
  static constexpr int W_BYTES = GEMM2W_KB_MAX * W_KB_BYTES;  // 128 KB
  static constexpr int EPI_BYTES = GEMM2_EPI_WARPS * EpiBiasF16<true>::SMEM_PER_WARP;
  static constexpr int TOTAL = 1024 + W_BYTES + GEMM2W_STAGES * A_BYTES + EPI_BYTES + GEMM2W_BIAS_BYTES + GEMM2W_BAR_BYTES;
  static_assert(TOTAL <= 227 * 1024, "W-resident pair GEMM does not fit in shared memory");
  static_assert(EpiBiasF16<true>::SMEM_PER_WARP == EpiBiasF16<false>::SMEM_PER_WARP, "one layout for both activations");
};

template <bool GELU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm2w_f16_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_c, int M, int N, int K,
                   const __grid_constant__ typename EpiBiasF16<GELU>::Params ep) {
  using Epi = EpiBiasF16<GELU>;
  using SM = Gemm2wSmem;
  constexpr int STAGES = GEMM2W_STAGES;
  constexpr uint32_t ACC_STRIDE = 256;
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;                                   // [KB_MAX][128 rows x 128 B], SWIZZLE_128B
  uint8_t* tiles = smem + SM::W_BYTES;                    // [STAGES][128 rows x 128 B]
  uint8_t* epi_smem = tiles + STAGES * SM::A_BYTES;
  float* bias_blk = reinterpret_cast<float*>(epi_smem + SM::EPI_BYTES);   // bias of columns [n_blk*256, +256)
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + SM::EPI_BYTES + GEMM2W_BIAS_BYTES);
  uint64_t* full_bar = bars;                              // [STAGES]  (leader's copy is the live one)
  uint64_t* empty_bar = bars + STAGES;                    // [STAGES]  (each CTA waits on its own)
  uint64_t* acc_full = bars + 2 * STAGES;                 // [2]       (each CTA waits on its own)
  uint64_t* acc_empty = bars + 2 * STAGES + 2;            // [2]       (leader's copy is the live one)
  uint64_t* w_full = bars + 2 * STAGES + 4;               // [KB_MAX]  (leader's copy is the live one)
  uint64_t* epi_bars = w_full + GEMM2W_KB_MAX;            // [GEMM2_EPI_WARPS][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bars + GEMM2_EPI_WARPS * 4);
  static_assert((2 * GEMM2W_STAGES + 4 + GEMM2W_KB_MAX + GEMM2_EPI_WARPS * 4) * 8 + 8 <= GEMM2W_BAR_BYTES, "barrier area too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (M + GEMM2_TILE_M - 1) / GEMM2_TILE_M;
  const int tiles_n = (N + GEMM2_BLOCK_N - 1) / GEMM2_BLOCK_N;   // launcher guarantees tiles_n <= num_clusters
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;      // launcher guarantees num_kb <= KB_MAX
  const int n_blk = cluster_id % tiles_n;
  const int m_first = cluster_id / tiles_n;
  const int m_step = (num_clusters - n_blk + tiles_n - 1) / tiles_n;   // clusters that own this column block

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < GEMM2_BLOCK_N; i += blockDim.x) {      // weights: not written by the previous kernel
    const int col = n_blk * GEMM2_BLOCK_N + i;
    bias_blk[i] = col < N ? ep.bias[col] : 0.f;
  }
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
    for (int s = 0; s < GEMM2W_KB_MAX; ++s) mbar_init(&w_full[s], 1);
    for (int s = 0; s < GEMM2_EPI_WARPS * 4; ++s) mbar_init(&epi_bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated in both SMs, bias staged
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      const int b_row = n_blk * GEMM2_BLOCK_N + static_cast<int>(rank) * 128;
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int m_blk = m_first; m_blk < tiles_m; m_blk += m_step) {
        const int a_row = m_blk * GEMM2_TILE_M + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (first) {
            // resident W k-block: its own buffer, loaded once, no slot to wait for
            const uint32_t leader_wfull = mapa_shared(smem_u32(&w_full[kb]), 0);
            if (leader) mbar_expect_tx(&w_full[kb], 2 * SM::W_KB_BYTES);
            tma_load_2d_2cta(wres + kb * SM::W_KB_BYTES, &map_b, leader_wfull, kb * GEMM_BLOCK_K, b_row);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * SM::A_BYTES);
          tma_load_2d_2cta(tiles + stage * SM::A_BYTES, &map_a, leader_full, kb * GEMM_BLOCK_K, a_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(256, GEMM2_BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int m_blk = m_first; m_blk < tiles_m; m_blk += m_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_cluster(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (it == 0) mbar_wait_cluster(&w_full[kb], 0);   // W k-block of both CTAs has landed (first tile only)
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(tiles + stage * SM::A_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(wres + kb * SM::W_KB_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_f16_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2cta_mc(&empty_bar[stage], 0b11);   // A slot free in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta_mc(&acc_full[as], 0b11);         // accumulator ready in both CTAs
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
    ctx.bias_all = bias_blk - n_blk * GEMM2_BLOCK_N;   // Epi::chunk indexes it with the absolute column
    ctx.lane = lane;
    ctx.M = M;
    ctx.N = N;
    ctx.seq = 0;
    ctx.primed = false;
    ctx.trace = nullptr;
    int it = 0;
    for (int m_blk = m_first; m_blk < tiles_m; m_blk += m_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * GEMM2_TILE_M + static_cast<int>(rank) * 128 + q * 32;
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
      epilogue_tile<GEMM2_BLOCK_N, Epi, GEMM2_EPI_WARPS / 4>(ctx, ep, taddr, row0, n_blk * GEMM2_BLOCK_N, part, &acc_full[as], aphase, [&]() {
        if (leader) mbar_arrive(&acc_empty[as]);
        else mbar_arrive_remote_relaxed(mapa_shared(smem_u32(&acc_empty[as]), 0));
      });
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
