// CTA-pair tcgen05 GEMM with the residual add AND the LayerNorm fused into the epilogue (post-norm encoder layer,
// reference: nn.TransformerEncoderLayer built at model/mdm.py:77-84):
//
//     h <- LayerNorm( h + A W^T + bias ; gamma, beta, eps )          N = d_model = 512, in place on h32, plus fp16 copy
//
// LayerNorm needs complete rows, so a CTA keeps BOTH 256-column halves of its 128 rows in tensor memory (all 512
// TMEM columns) before the epilogue starts; the pair still shares W through cta_group::2 (each CTA stages half of every
// 256-row W tile).  Epilogue, thread = row, two warps per TMEM lane quarter (one per 256-column half):
//   pass 1   v = acc + bias + residual (residual slabs by TMA load), per-row sum / sum of squares, v written back to
//            TMEM over the accumulator (tcgen05.st)
//   combine  the two warps of a quarter exchange their partial sums through shared memory (named barrier, 64 threads)
//   pass 2   y = (v - mean) * rstd * gamma + beta -> fp32 slab -> TMA store to h32 ; fp16 copy -> st.global to h16
// This removes the separate LayerNorm kernel (one full read + two full writes of the residual stream per LayerNorm)
// and the fp32 read-modify-write of the residual epilogue.  Cost: no MMA / epilogue overlap inside a CTA (TMEM full).
#pragma once
#include "epilogues.cuh"
#include "gemm2.cuh"

namespace b200 {

constexpr int LN2_EPI_WARPS = 8;    // two warps per TMEM lane quarter, one per column half
constexpr int LN2_THREADS = 64 + 32 * LN2_EPI_WARPS;

constexpr int LN_D = 512;
constexpr int LN_EPI_SMEM_PER_WARP = 12 * 1024;   // 3 rotating fp32 slabs (residual in, pass 1 / normalised out, pass 2)
constexpr int LN_PARAM_BYTES = 3 * LN_D * 4;      // bias | gamma | beta
constexpr int LN_STATS_BYTES = 4 * 2 * 32 * 8;    // [quarter][part][lane] (sum, sumsq)

struct Gemm2LnSmem {
  static constexpr int STAGE_BYTES = 32 * 1024;
  static constexpr int EPI_BYTES = LN2_EPI_WARPS * LN_EPI_SMEM_PER_WARP;
  static constexpr int AUX_BYTES = LN_PARAM_BYTES + LN_STATS_BYTES + GEMM_BAR_BYTES;
  static constexpr int budget = 227 * 1024 - 1024 - EPI_BYTES - AUX_BYTES;
  static constexpr int STAGES = (budget / STAGE_BYTES) > 6 ? 6 : (budget / STAGE_BYTES);
  static constexpr int TOTAL = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + AUX_BYTES;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

struct LnParams {
  const float* bias;    // [512] projection bias
  const float* gamma;   // [512]
  const float* beta;    // [512]
  float eps;
  long long* trace;     // optional: per-phase clock64 stamps of (block 0, first epilogue warp), 8 per tile
};

// map_a: A [M, K] fp16 (box 128 rows); map_b: W [512, K] fp16 (box 128 rows); map_h32: h32 [M, 512] fp32 (box 32 rows x
// 32 cols), loaded and stored; h16: [M, 512] fp16, written directly (64 contiguous bytes per thread per chunk).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN2_THREADS, 1)
gemm2_resid_ln_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_h32, __half* __restrict__ h16, int M, int K,
                       const LnParams lp) {
  using SM = Gemm2LnSmem;
  constexpr int STAGES = SM::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* epi_smem = smem + STAGES * SM::STAGE_BYTES;
  float* prm = reinterpret_cast<float*>(epi_smem + SM::EPI_BYTES);          // bias | gamma | beta
  float2* stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(prm) + LN_PARAM_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + LN_STATS_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  (leader's copy is the live one)
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;       // [1]
  uint64_t* acc_empty = bars + 2 * STAGES + 1;  // [1]       (leader's copy is the live one)
  uint64_t* epi_bars = bars + 2 * STAGES + 2;   // [LN2_EPI_WARPS][3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bars + LN2_EPI_WARPS * 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = (M + GEMM2_TILE_M - 1) / GEMM2_TILE_M;   // one tile = 256 rows x 512 columns per pair
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  for (int i = threadIdx.x; i < LN_D; i += blockDim.x) {
    prm[i] = lp.bias[i];
    prm[LN_D + i] = lp.gamma[i];
    prm[2 * LN_D + i] = lp.beta[i];
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_h32);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * LN2_EPI_WARPS);
    for (int s = 0; s < LN2_EPI_WARPS * 3; ++s) mbar_init(&epi_bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int a_row = tile * GEMM2_TILE_M + static_cast<int>(rank) * 128;
        for (int n_blk = 0; n_blk < 2; ++n_blk) {
          const int b_row = n_blk * GEMM2_BLOCK_N + static_cast<int>(rank) * 128;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = tiles + stage * SM::STAGE_BYTES;
            uint8_t* sb = sa + 16384;
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * SM::STAGE_BYTES);
            tma_load_2d_2cta(sa, &map_a, leader_full, kb * GEMM_BLOCK_K, a_row);
            tma_load_2d_2cta(sb, &map_b, leader_full, kb * GEMM_BLOCK_K, b_row);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
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
        mbar_wait_cluster(acc_empty, (it & 1) ^ 1);   // epilogues of both CTAs have drained the previous tile
        tc_fence_after();
        for (int n_blk = 0; n_blk < 2; ++n_blk) {
          const uint32_t tmem_d = tmem_base + n_blk * 256;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_cluster(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(tiles + stage * SM::STAGE_BYTES);
            const uint64_t da = umma_desc_k_sw128(sa);
            const uint64_t db = umma_desc_k_sw128(sa + 16384);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
              umma_f16_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2cta_mc(&empty_bar[stage], 0b11);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit_2cta_mc(acc_full, 0b11);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9), both CTAs
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;       // which 256-column half of the rows this warp handles
    uint8_t* wsm = epi_smem + (warp - 2) * LN_EPI_SMEM_PER_WARP;
    uint8_t* s32[3] = {wsm, wsm + 4096, wsm + 8192};
    uint64_t* rbar = epi_bars + (warp - 2) * 3;
    const float* bias_s = prm + part * 256;
    const float* gamma_s = prm + LN_D + part * 256;
    const float* beta_s = prm + 2 * LN_D + part * 256;
    uint32_t rseq = 0;   // residual slabs consumed so far (buffer = rseq % 3, parity = (rseq / 3) & 1)
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int row0 = tile * GEMM2_TILE_M + static_cast<int>(rank) * 128 + q * 32;
      const bool live = row0 < M;
      const int colw = part * 256;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + colw;
      auto load_resid = [&](uint32_t seq, int c) {
        if (lane == 0) {
          bulk_wait_group_read<0>();   // any earlier store out of these buffers has been read
          mbar_expect_tx(&rbar[seq % 3], 4096);
          tma_load_2d(s32[seq % 3], &map_h32, &rbar[seq % 3], colw + 32 * c, row0);
        }
      };
      const bool tr = lp.trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && it < 4;
      if (tr) lp.trace[it * 8 + 0] = clock64();
      if (live) {   // the three residual buffers are filled while the tensor core is still working on this tile
        load_resid(rseq, 0);
        load_resid(rseq + 1, 1);
        load_resid(rseq + 2, 2);
      }
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      if (tr) lp.trace[it * 8 + 1] = clock64();
      float sum = 0.f, sumsq = 0.f;
      if (live) {
        // ---- pass 1: v = acc + bias + residual; statistics; v back to TMEM
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t raw[32];
          tmem_ld_32x32(taddr + 32 * c, raw);
          mbar_wait(&rbar[rseq % 3], (rseq / 3) & 1);
          tmem_ld_wait();
          const uint8_t* slab = s32[rseq % 3];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = *reinterpret_cast<const float4*>(slab + slab_off(lane, j));
            const float4 b = *reinterpret_cast<const float4*>(bias_s + 32 * c + 4 * j);
            const float v0 = r.x + (__uint_as_float(raw[4 * j + 0]) + b.x);
            const float v1 = r.y + (__uint_as_float(raw[4 * j + 1]) + b.y);
            const float v2 = r.z + (__uint_as_float(raw[4 * j + 2]) + b.z);
            const float v3 = r.w + (__uint_as_float(raw[4 * j + 3]) + b.w);
            sum += (v0 + v1) + (v2 + v3);
            sumsq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, sumsq))));
            raw[4 * j + 0] = __float_as_uint(v0);
            raw[4 * j + 1] = __float_as_uint(v1);
            raw[4 * j + 2] = __float_as_uint(v2);
            raw[4 * j + 3] = __float_as_uint(v3);
          }
          tmem_st_32x32(taddr + 32 * c, raw);
          __syncwarp();   // every lane is done with this slab: refill it three chunks ahead
          if (c + 3 < 8) load_resid(rseq + 3, c + 3);
          ++rseq;
        }
        tmem_st_wait();
      }
      if (tr) lp.trace[it * 8 + 2] = clock64();
      // ---- combine the two column halves of each row (warps w and w+4 own the same 32 rows)
      stats[(q * 2 + part) * 32 + lane] = make_float2(sum, sumsq);
      named_bar_sync(1 + q, 64);
      const float2 other = stats[(q * 2 + (part ^ 1)) * 32 + lane];
      named_bar_sync(1 + q, 64);   // both warps have read before the next tile overwrites
      const float mean = (sum + other.x) * (1.f / LN_D);
      const float var = fmaxf((sumsq + other.y) * (1.f / LN_D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + lp.eps);
      if (tr) lp.trace[it * 8 + 3] = clock64();
      if (live) {
        // ---- pass 2: normalise, write h32 (fp32) and h16 (fp16)
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t raw[32];
          tmem_ld_32x32(taddr + 32 * c, raw);
          tmem_ld_wait();
          uint8_t* o32 = s32[c % 3];
          if (c >= 3) {            // this slab was stored three chunks ago: two younger stores may stay in flight
            if (lane == 0) bulk_wait_group_read<2>();
            __syncwarp();
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g = *reinterpret_cast<const float4*>(gamma_s + 32 * c + 4 * j);
            const float4 be = *reinterpret_cast<const float4*>(beta_s + 32 * c + 4 * j);
            float4 y;
            y.x = (__uint_as_float(raw[4 * j + 0]) - mean) * rstd * g.x + be.x;
            y.y = (__uint_as_float(raw[4 * j + 1]) - mean) * rstd * g.y + be.y;
            y.z = (__uint_as_float(raw[4 * j + 2]) - mean) * rstd * g.z + be.z;
            y.w = (__uint_as_float(raw[4 * j + 3]) - mean) * rstd * g.w + be.w;
            *reinterpret_cast<float4*>(o32 + slab_off(lane, j)) = y;
            pk[2 * j] = pack_half2(y.x, y.y);
            pk[2 * j + 1] = pack_half2(y.z, y.w);
          }
          if (row0 + lane < M) {   // fp16 copy straight from registers: 64 contiguous bytes of this thread's row
            uint4* dst = reinterpret_cast<uint4*>(h16 + static_cast<size_t>(row0 + lane) * LN_D + colw + 32 * c);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_h32, o32, colw + 32 * c, row0);
            bulk_commit_group();
          }
        }
      }
      if (tr) lp.trace[it * 8 + 4] = clock64();
      // accumulator (and the v it was overwritten with) fully consumed: hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(acc_empty);
        else mbar_arrive_remote(mapa_shared(smem_u32(acc_empty), 0));
      }
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace b200
