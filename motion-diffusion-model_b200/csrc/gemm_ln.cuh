// tcgen05 GEMM with residual add + LayerNorm fused into the epilogue, N = d_model = 512 split over a 2-CTA cluster:
//
//     h <- LayerNorm( h + A W^T + bias ; gamma, beta, eps )        in place on the residual stream (fp16 [hi | lo])
//     (post-norm nn.TransformerEncoderLayer of the reference, built at model/mdm.py:77-84)
//
// Both CTAs of a cluster work on the SAME 128 rows; CTA r owns columns [256 r, 256 r + 256) (its own 256 rows of W,
// cta_group::1 MMAs, 128 x 256 accumulator).  Each CTA therefore needs only 256 TMEM columns per tile and keeps TWO
// accumulator stages, so the two-pass LayerNorm epilogue of tile i overlaps the MMAs of tile i+1 -- and the epilogue
// traffic (one read + 1.5 writes of the residual stream per LayerNorm) is spread over the whole kernel instead of
// arriving in chip-wide bursts.
//
// Epilogue, thread = row, warps (q = lane quarter, part = 128-column half of the CTA's 256 columns):
//   pass 1    v = acc + bias + residual (TMA-loaded slabs); per-row partial sum / sum of squares; v -> TMEM (in place)
//   exchange  partials of the two parts meet in shared memory (named barrier); the part-0 warp pushes the CTA's partial
//             into the PEER CTA's shared memory with st.async (data + mbarrier complete_tx in one message, no release
//             fence: round 1's `mbarrier.arrive.release.cluster` showed up as stall_membar 0.5 per issue); both CTAs now
//             own the statistics of the complete 512-wide rows
//   pass 2    y = (v - mean) rstd gamma + beta -> [hi | lo] fp16 slab -> TMA store
// History (profiles/): an fp32 stream + separate fp16 copy cost 6 B/element of stores and 41.6 / 49.2 us per launch;
// TMA slabs beat direct 256-bit loads/stores (43.8 / 51.8 us vs 52.0 / 59.7 us).  Round 2: keeping v in 128 registers
// (one TMEM read instead of read + write-back + read; 12 warps with setmaxnreg) did not shorten the tile loop -- the
// kernel is bound by L2 -> shared-memory bytes (operands + residual, ~49 B/clk/SM), not by tcgen05.ld -- and cost 5 % of
// the sampling loop (A/B on one box: 81.7 vs 77.9 ms), so it was not kept; st.async alone gave 78.1 -> 76.9 ms.
#pragma once
#include "epilogues.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"   // g_gemm2_trace (debug stamps)

namespace b200 {

constexpr int GLN_EPI_WARPS = 8;    // two warps per TMEM lane quarter, one per column half
constexpr int GLN_THREADS = 64 + 32 * GLN_EPI_WARPS;

constexpr int GLN_D = 512;
constexpr int GLN_BN = 256;                       // columns per CTA
constexpr int GLN_WARP_SMEM = 8 * 1024;           // two fp32 slabs per epilogue warp (residual in / normalised out)
constexpr int GLN_STAGE_BYTES = (128 + GLN_BN) * GEMM_BLOCK_K * 2;   // 48 KB
constexpr int GLN_AUX_BYTES = 3 * GLN_BN * 4 /*bias gamma beta*/ + 3 * 2048 /*local, remote, total stats*/ + 1024 /*barriers*/;

struct GemmLnSmem {
  static constexpr int EPI_BYTES = GLN_EPI_WARPS * GLN_WARP_SMEM;
  static constexpr int budget = 227 * 1024 - 1024 - EPI_BYTES - GLN_AUX_BYTES;
  static constexpr int STAGES = (budget / GLN_STAGE_BYTES) > 6 ? 6 : (budget / GLN_STAGE_BYTES);
  static constexpr int TOTAL = 1024 + STAGES * GLN_STAGE_BYTES + EPI_BYTES + GLN_AUX_BYTES;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

struct GemmLnParams {
  const float* bias;    // [512]
  const float* gamma;   // [512]
  const float* beta;    // [512]
  float eps;
};

// map_a: A [M, K] fp16, box 128 rows; map_b: W [512, K] fp16, box 256 rows;
// map_res: the residual stream h, fp16 [M, 1024] = [hi | lo] (hi + lo carries ~22 bits), box {32 cols, 32 rows} with
// 64-byte rows (SWIZZLE_64B): a 4 KB slab = the hi half-slab of 32 columns + the lo half-slab.  Loaded and stored in place.
// The hi half doubles as the fp16 A operand of the next GEMM (the trans_dec engine feeds both halves, K = 1024).
// Versus an fp32 stream + fp16 copy this writes 4 instead of 6 bytes per element: the epilogue is bound by the SM's
// store bandwidth (measured ~13 B/clk/SM), not by the MMAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GLN_THREADS, 1)
gemm_resid_ln_cluster(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_res, int M, int K, const GemmLnParams lp) {
  using SM = GemmLnSmem;
  constexpr int STAGES = SM::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* epi_smem = smem + STAGES * GLN_STAGE_BYTES;
  float* prm = reinterpret_cast<float*>(epi_smem + SM::EPI_BYTES);                 // bias | gamma | beta  (this CTA's 256 cols)
  float2* st_local = reinterpret_cast<float2*>(prm + 3 * GLN_BN);                  // [stage 2][q 4][part 2][lane 32] -> 4 KB? no: see below
  // layout of the 6 KB statistics area: local [2][4][2][32] float2 would be 4 KB; we keep [2 stages][4 q][32] per array
  // and let part p of a quarter use array p (local0 / local1), plus remote and total:
  float2* st_part0 = st_local;                     // [2][4][32]
  float2* st_part1 = st_local + 256;               // [2][4][32]
  float2* st_remote = st_local + 512;              // [2][4][32]  written by the PEER CTA
  // (total is recomputed by each warp from part0 + part1 + remote, no extra array needed)
  uint64_t* bars = reinterpret_cast<uint64_t*>(st_local + 768);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;          // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;     // [2]
  uint64_t* xbar = bars + 2 * STAGES + 4;          // [2 stages][4 q]  peer's statistics have landed
  uint64_t* rbar = xbar + 8;                       // [GLN_EPI_WARPS][2] residual slabs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + GLN_EPI_WARPS * 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;   // one tile per cluster = 128 rows x 512 columns
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int col_cta = static_cast<int>(rank) * GLN_BN;

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < GLN_BN; i += blockDim.x) {
    prm[i] = lp.bias[col_cta + i];
    prm[GLN_BN + i] = lp.gamma[col_cta + i];
    prm[2 * GLN_BN + i] = lp.beta[col_cta + i];
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_res);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], GLN_EPI_WARPS);
    }
    for (int s = 0; s < 8; ++s) mbar_init(&xbar[s], 1);    // one expect_tx arrival; the peer's 32 lanes deliver 256 bytes with st.async
    for (int s = 0; s < GLN_EPI_WARPS * 2; ++s) mbar_init(&rbar[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anybody signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * GLN_STAGE_BYTES;
          uint8_t* sb = sa + 16384;
          mbar_expect_tx(&full_bar[stage], GLN_STAGE_BYTES);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * GEMM_BLOCK_K, tile * GEMM_BLOCK_M);
          tma_load_2d(sb, &map_b, &full_bar[stage], kb * GEMM_BLOCK_K, col_cta);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(GEMM_BLOCK_M, GLN_BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * GLN_STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + 16384);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9)
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;                 // 128-column half of this CTA's 256 columns
    uint8_t* wsm = epi_smem + (warp - 2) * GLN_WARP_SMEM;
    uint8_t* slab[2] = {wsm, wsm + 4096};
    uint64_t* rb = rbar + (warp - 2) * 2;
    uint32_t rseq = 0;                                // residual slabs consumed (buffer = rseq & 1, parity = (rseq >> 1) & 1)
    const int colw = part * 128;                      // first column of this warp inside the CTA's 256
    const float* bias_s = prm + colw;
    const float* gamma_s = prm + GLN_BN + colw;
    const float* beta_s = prm + 2 * GLN_BN + colw;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = tile * GEMM_BLOCK_M + q * 32;
      const bool live = row0 < M;
      const int gcol = col_cta + colw;                // first global column of this warp
      const uint32_t taddr = tmem_base + as * 256 + (static_cast<uint32_t>(q * 32) << 16) + colw;
      auto load_resid = [&](uint32_t seq, int c) {
        if (lane == 0) {
          bulk_wait_group_read<0>();                  // stores out of these slabs (previous tile, pass 2) have been read
          mbar_expect_tx(&rb[seq & 1], 4096);
          tma_load_2d(slab[seq & 1], &map_res, &rb[seq & 1], gcol + 32 * c, row0);
          tma_load_2d(slab[seq & 1] + 2048, &map_res, &rb[seq & 1], GLN_D + gcol + 32 * c, row0);
        }
      };
      if (live) {
        load_resid(rseq, 0);
        load_resid(rseq + 1, 1);
      }
      if (part == 0 && lane == 0) mbar_expect_tx(&xbar[(it & 1) * 4 + q], 256);
      long long* tr = B200_TRACE_PTR(blockIdx.x == 0 && warp == 2 && lane == 0 && it >= 1 && it < 3, g_gemm2_trace + (it - 1) * 16);
      int tri = 0;
      if (tr) tr[tri++] = clock64();
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      if (tr) tr[tri++] = clock64();
      float sum = 0.f, sumsq = 0.f;
      if (live) {
        // ---- pass 1 (64 accumulator columns per tcgen05.ld / tcgen05.st: the TMEM port arbitrates per instruction while
        // the next tile's MMAs run, see ptx.cuh)
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t raw[64];
          tmem_ld_32x64(taddr + 64 * cc, raw);
          tmem_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int c = 2 * cc + hh;
            mbar_wait(&rb[rseq & 1], (rseq >> 1) & 1);
            const uint8_t* sl = slab[rseq & 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // 8 columns: 16 bytes of hi + 16 bytes of lo
              float r[8];
              join_hi_lo8(*reinterpret_cast<const uint4*>(sl + slab64_off(lane, j)),
                          *reinterpret_cast<const uint4*>(sl + 2048 + slab64_off(lane, j)), r);
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 32 * c + 8 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 32 * c + 8 * j + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float v = r[i] + (__uint_as_float(raw[32 * hh + 8 * j + i]) + bb[i]);
                sum += v;
                sumsq = fmaf(v, v, sumsq);
                raw[32 * hh + 8 * j + i] = __float_as_uint(v);
              }
            }
            __syncwarp();
            if (c + 2 < 4) load_resid(rseq + 2, c + 2);
            ++rseq;
            if (tr) tr[tri++] = clock64();
          }
          tmem_st_32x64(taddr + 64 * cc, raw);
        }
        tmem_st_wait();
      }
      // ---- statistics of the full 512-wide rows: part 0 + part 1 of this CTA + the peer CTA's 256 columns
      const int sidx = (as * 4 + q) * 32 + lane;
      (part == 0 ? st_part0 : st_part1)[sidx] = make_float2(sum, sumsq);
      named_bar_sync(1 + q, 64);
      if (part == 0) {
        const float2 a = st_part0[sidx], b = st_part1[sidx];
        st_async_f32x2(mapa_shared(smem_u32(&st_remote[sidx]), rank ^ 1), a.x + b.x, a.y + b.y,
                       mapa_shared(smem_u32(&xbar[as * 4 + q]), rank ^ 1));
      }
      mbar_wait_cluster(&xbar[as * 4 + q], aphase);   // the peer's 32 lanes have delivered their partials
      const float2 p0 = st_part0[sidx], p1 = st_part1[sidx], pr = st_remote[sidx];
      const float mean = ((p0.x + p1.x) + pr.x) * (1.f / GLN_D);
      const float var = fmaxf(((p0.y + p1.y) + pr.y) * (1.f / GLN_D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + lp.eps);
      if (tr) tr[tri++] = clock64();
      if (live) {
        // ---- pass 2
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t raw64[64];
          tmem_ld_32x64(taddr + 64 * cc, raw64);
          tmem_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
          const int c = 2 * cc + hh;
          const uint32_t* raw = raw64 + 32 * hh;
          uint8_t* o32 = slab[c & 1];
          if (c >= 2) {
            if (lane == 0) bulk_wait_group_read<1>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g0 = *reinterpret_cast<const float4*>(gamma_s + 32 * c + 8 * j);
            const float4 g1 = *reinterpret_cast<const float4*>(gamma_s + 32 * c + 8 * j + 4);
            const float4 e0 = *reinterpret_cast<const float4*>(beta_s + 32 * c + 8 * j);
            const float4 e1 = *reinterpret_cast<const float4*>(beta_s + 32 * c + 8 * j + 4);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = (__uint_as_float(raw[8 * j + i]) - mean) * rstd * gg[i] + ee[i];
            uint32_t hi[4], lo[4];
            split_hi_lo8(y, hi, lo);
            *reinterpret_cast<uint4*>(o32 + slab64_off(lane, j)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(o32 + 2048 + slab64_off(lane, j)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_res, o32, gcol + 32 * c, row0);
            tma_store_2d(&map_res, o32 + 2048, GLN_D + gcol + 32 * c, row0);
            bulk_commit_group();
          }
          if (tr) tr[tri++] = clock64();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }

  // the peer may still be writing into this CTA's shared memory / signalling its barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200
