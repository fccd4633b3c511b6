// Fused QKV projection + softmax attention of one transformer layer (north-star kernel #1):
//
//     att[b, s, h*128 : h*128+128] = softmax( (h_b Wq_h^T + bq)(h_b Wk_h^T + bk)^T / sqrt(128) + key_mask_b ) (h_b Wv_h^T + bv)
//
// i.e. nn.MultiheadAttention up to (not including) its output projection, as used by nn.TransformerEncoderLayer in the
// reference (model/mdm.py:77-84; key_padding_mask of model/mdm.py:241-247 = a per-sample prefix mask `kvlen`).  The
// qkv tensor ([M, 1536] fp16, 77 MB per layer at the benchmark size) never exists: Q, K and V of one (sample, head) are
// produced into shared memory in the tensor-core operand layouts and consumed in place.
//
// Work item = (sample, head), processed by a CTA PAIR (cluster of 2, cta_group::2 MMAs, M = 256 = all <= 256 tokens of
// the sample: CTA r owns tokens [128 r, 128 r + 128)).  Persistent: items are strided over the clusters.
//   1. projection  acc[256 x 384] = h_b[256 x 512] [Wq_h; Wk_h; Wv_h]^T    8 k-blocks through a 3-stage TMA ring;
//        two UMMAs per k-step: N = 256 (Q | K) and N = 128 (V); W is split across the pair by the hardware (CTA 0 stages
//        Wq_h and half of Wv_h, CTA 1 Wk_h and the other half) -- each CTA loads 16 KB of tokens + 24 KB of weights per k-block
//   2. epilogue    + bias, -> fp16, written as UMMA operands:  Q_r [128 x 128] and K_r [128 keys x 128] K-major (K_r is
//        exactly this CTA's half of the B operand of step 3), V as [256 keys x 64 dh] MN-major: CTA c holds dh
//        [64 c, 64 c + 64) of ALL keys, so each CTA sends one 64-wide half of its V rows to its peer through distributed
//        shared memory (16 KB per item) -- the only data the two CTAs exchange
//   3. S = Q K^T   [256 x 256 keys] fp32 in TMEM (8 UMMAs)
//   4. softmax     thread = query row, two warps per row (128 keys each, max / sum exchanged through shared memory);
//        P -> TMEM as fp16
//   5. O = P V     A = P from TMEM, B = V (MN-major), 16 UMMAs;  6. O / rowsum -> fp16 -> swizzled slabs -> TMA store
// TMEM (512 columns per CTA): [0,256) Q|K accumulator, then S (it may be overwritten as soon as Q and K have been read,
// while V is still being converted), then P as fp16 over the first half of each warp's own S columns; [256,384) V accumulator; O in [384,512) -- outside the projection accumulator, so the next item's projection
// MMAs are issued right behind this item's P V and run under its output epilogue.
// What bounds the CUDA-core side (measured, profiles/r02_a_trace_qkv_attn_v1.txt): tcgen05.ld moves 16 B/clk per lane
// quarter, i.e. a [128 x N] fp32 accumulator costs 8 N cycles per read -- 3072 for the projection, 2048 for S, 1024 for
// O.  Every TMEM value is therefore read exactly once (S is held in 128 registers between the max and the exp pass).
#pragma once
#include <cuda_fp16.h>

#include "attention_tc.cuh"   // tmem_st_32x16
#include "gemm2.cuh"          // g_gemm2_trace (debug stamps)
#include "epilogues.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int QA_THREADS = 384;              // warpgroup 0: warp 0 TMA, warp 1 MMA / TMEM, warps 2-3 idle; warpgroups 1-2: epilogue
// Registers: 12 warps = 3 per SM sub-partition = 168 registers per thread at launch.  The epilogue warps hold a whole
// 128-key row of S (and the LayerNorm-style single-read epilogues 128 accumulator values) in registers, so warpgroup 0
// hands its share back (setmaxnreg.dec) and the epilogue warpgroups grow: CTRL + 2 x EPI <= 512 per lane.  The two setmaxnreg
// must sit INSIDE the role branches: after a join ptxas compiles everything for the smaller count.
constexpr int QA_REGS_CTRL = 88, QA_REGS_EPI = 208;   // 88 + 2 x 208 = 504 = 3 x 168: the pool is what the CTA got at launch
constexpr int QA_STAGES = 3;
constexpr int QA_STAGE_BYTES = 16384 + 16384 + 8192;   // A [128 x 64] | W (q or k) [128 x 64] | W (v half) [64 x 64]
constexpr int QA_TILE = 32768;               // [128 x 128] fp16
struct QkvAttnSmem {
  static constexpr int RING = QA_STAGES * QA_STAGE_BYTES;
  static constexpr int Q_OFF = RING, K_OFF = RING + QA_TILE, V_OFF = RING + 2 * QA_TILE;
  static constexpr int BIAS_OFF = RING + 3 * QA_TILE;          // 1536 floats
  static constexpr int XCH_OFF = BIAS_OFF + 1536 * 4;          // max[2][128], sum[2][128] floats
  static constexpr int BAR_OFF = XCH_OFF + 2048;
  static constexpr int TOTAL = 1024 + BAR_OFF + 256;
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// map_h : residual stream hi half viewed [n_samples][S][512] (row pitch `ld`), box {64, 128, 1}
// map_w128 / map_w64 : in_proj_weight fp16 [1536, 512], boxes of 128 / 64 rows
// map_o : att16 viewed [n_samples][S][512], box {64, 32, 1}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(QA_THREADS, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w128,
                     const __grid_constant__ CUtensorMap map_w64, const __grid_constant__ CUtensorMap map_o,
                     const float* __restrict__ bqkv, const int* __restrict__ kvlen, int n_samples, int S, float scale_log2) {
  using SM = QkvAttnSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* sQ = smem + SM::Q_OFF;
  uint8_t* sK = smem + SM::K_OFF;
  uint8_t* sV = smem + SM::V_OFF;
  float* bias_s = reinterpret_cast<float*>(smem + SM::BIAS_OFF);
  float* xmax = reinterpret_cast<float*>(smem + SM::XCH_OFF);   // [2][128]
  float* xsum = xmax + 256;                                     // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* full_bar = bars;                 // [3]  leader's copy is live (expect_tx of both CTAs' bytes)
  uint64_t* empty_bar = bars + 3;            // [3]  multicast commit
  uint64_t* proj_done = bars + 6;            //      multicast commit: accumulator complete
  uint64_t* qkv_ready = bars + 7;            //      leader: 16 warp arrivals (both CTAs): Q and K are in shared memory
  uint64_t* s_done = bars + 8;               //      multicast commit: S complete
  uint64_t* p_ready = bars + 9;              //      leader: 16 warp arrivals: P is in TMEM, V (incl. the peer's half) in shared memory
  uint64_t* o_done = bars + 10;              //      multicast commit: O complete
  uint64_t* tmem_free = bars + 11;           //      leader: 16 warp arrivals: O has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_items = n_samples * 4;

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 1536; i += blockDim.x) bias_s[i] = bqkv[i];   // weights: not produced by the previous kernel
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_h);
    tma_prefetch_desc(&map_w128);
    tma_prefetch_desc(&map_w64);
    tma_prefetch_desc(&map_o);
    for (int s = 0; s < QA_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(proj_done, 1);
    mbar_init(qkv_ready, 16);
    mbar_init(s_done, 1);
    mbar_init(p_ready, 16);
    mbar_init(o_done, 1);
    mbar_init(tmem_free, 16);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(QA_REGS_EPI));
    // ---------------------------------------------------------------- epilogue / softmax warps (4..11), both CTAs
    const int q = warp & 3;                    // TMEM lane quarter
    const int part = (warp - 4) >> 2;          // 0: Q columns + dh [0,64) of V + keys [0,128); 1: K + dh [64,128) + keys [128,256)
    const int row = 32 * q + lane;             // token row inside this CTA's tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
    const uint32_t qkv_ready_leader = mapa_shared(smem_u32(qkv_ready), 0);
    const uint32_t p_ready_leader = mapa_shared(smem_u32(p_ready), 0);
    const uint32_t tmem_free_leader = mapa_shared(smem_u32(tmem_free), 0);
    uint8_t* qk_tile = part == 0 ? sQ : sK;
    // V destination: CTA `part` owns dh [64 part, +64) of every key; this thread's key row there is 128 rank + row
    const int key_row = 128 * static_cast<int>(rank) + row;
    const uint32_t v_dst = mapa_shared(smem_u32(sV) + key_row * 128, static_cast<uint32_t>(part));
    uint8_t* o_slab = sQ + part * 16384 + q * 4096;   // rows [32q, 32q+32) of dh atom `part` of the (dead) Q tile
    auto ex2 = [](float x) {
      float y;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      return y;
    };
    int it = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
      const uint32_t par = it & 1;
      const int smp = item >> 2, h = item & 3;
      const int kvl = min(kvlen[smp], S);
      long long* tr = B200_TRACE_PTR(blockIdx.x == 0 && warp == 4 && lane == 0 && it >= 1 && it < 3, g_gemm2_trace + (it - 1) * 16);
      if (tr) tr[0] = clock64();
      mbar_wait(proj_done, par);
      tc_fence_after();
      if (tr) tr[1] = clock64();
      // the previous item's output slabs live in the Q tile: their TMA stores must have finished reading
      if (lane == 0) bulk_wait_group_read<0>();
      named_bar_sync(1 + q, 64);
      // ---- projection epilogue: Q or K (4 chunks of 32 columns), then one 64-wide half of V (2 chunks)
      {
        const float* bqk = bias_s + part * 512 + h * 128;
        const float* bv = bias_s + 1024 + h * 128 + 64 * part;
        uint32_t ra[32], rb[32];
        auto store_qk = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = *reinterpret_cast<const float4*>(bqk + 32 * c + 8 * j);
            const float4 b1 = *reinterpret_cast<const float4*>(bqk + 32 * c + 8 * j + 4);
            const uint32_t w0 = pack_half2(__uint_as_float(r[8 * j + 0]) + b0.x, __uint_as_float(r[8 * j + 1]) + b0.y);
            const uint32_t w1 = pack_half2(__uint_as_float(r[8 * j + 2]) + b0.z, __uint_as_float(r[8 * j + 3]) + b0.w);
            const uint32_t w2 = pack_half2(__uint_as_float(r[8 * j + 4]) + b1.x, __uint_as_float(r[8 * j + 5]) + b1.y);
            const uint32_t w3 = pack_half2(__uint_as_float(r[8 * j + 6]) + b1.z, __uint_as_float(r[8 * j + 7]) + b1.w);
            const int chunk = (c & 1) * 4 + j;       // 16-byte chunk inside the 128-byte row of dh atom (c >> 1)
            *reinterpret_cast<uint4*>(qk_tile + (c >> 1) * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(w0, w1, w2, w3);
          }
        };
        auto store_v = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = *reinterpret_cast<const float4*>(bv + 32 * c + 8 * j);
            const float4 b1 = *reinterpret_cast<const float4*>(bv + 32 * c + 8 * j + 4);
            const uint32_t w0 = pack_half2(__uint_as_float(r[8 * j + 0]) + b0.x, __uint_as_float(r[8 * j + 1]) + b0.y);
            const uint32_t w1 = pack_half2(__uint_as_float(r[8 * j + 2]) + b0.z, __uint_as_float(r[8 * j + 3]) + b0.w);
            const uint32_t w2 = pack_half2(__uint_as_float(r[8 * j + 4]) + b1.x, __uint_as_float(r[8 * j + 5]) + b1.y);
            const uint32_t w3 = pack_half2(__uint_as_float(r[8 * j + 6]) + b1.z, __uint_as_float(r[8 * j + 7]) + b1.w);
            const int chunk = c * 4 + j;
            st_shared_cluster_v4(v_dst + ((chunk ^ (key_row & 7)) << 4), w0, w1, w2, w3);
          }
        };
        const uint32_t tqk = trow + part * 128, tv = trow + 256 + part * 64;
        tmem_ld_32x32(tqk, ra);
        tmem_ld_wait();
        tmem_ld_32x32(tqk + 32, rb);
        store_qk(ra, 0);
        tmem_ld_wait();
        tmem_ld_32x32(tqk + 64, ra);
        store_qk(rb, 1);
        tmem_ld_wait();
        tmem_ld_32x32(tqk + 96, rb);
        store_qk(ra, 2);
        tmem_ld_wait();
        tmem_ld_32x32(tv, ra);
        store_qk(rb, 3);
        // Q / K are complete: let S = Q K^T start while V is still being written.  Only LOCAL stores precede this
        // release, so it does not have to wait for distributed-shared-memory traffic.
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(qkv_ready_leader);   // release.cluster
        if (tr) tr[2] = clock64();
        tmem_ld_wait();
        tmem_ld_32x32(tv + 32, rb);
        store_v(ra, 0);
        tmem_ld_wait();
        store_v(rb, 1);
        // (the V half that went to the peer is published together with P, below)
      }
      tc_fence_before();
      if (tr) tr[3] = clock64();
      // ---- softmax: the valid 16-key blocks of S are split evenly between the two warps of a row (at S = 197: 7 + 6 blocks
      // instead of 8 + 5: the S read and the exponentials are the critical path of the item); ONE read of S into
      // registers, row max / sum shared with the partner warp through shared memory
      mbar_wait(s_done, par);
      tc_fence_after();
      if (tr) tr[4] = clock64();
      const int nblk = max(1, (kvl + 15) >> 4);            // 16-key blocks holding a valid key (P V stops there too)
      const int split_blk = (nblk + 1) >> 1;               // part 0: blocks [0, split), part 1: [split, nblk)
      const int key0 = part == 0 ? 0 : 16 * split_blk;     // first key (= first S column) of this warp
      const int nkeys = part == 0 ? 16 * split_blk : 16 * (nblk - split_blk);   // <= 128, multiple of 16
      const int n32 = nkeys >> 5;
      const bool tail16 = (nkeys & 16) != 0;
      float sum = 0.f;
      {
        const uint32_t tS = trow + key0;
        const uint32_t tP = trow + key0;               // P (fp16 pairs) over the first half of this warp's own S columns
        uint32_t r0[32], r1[32], r2[32], r3[32], rt[16];
        if (n32 > 0) tmem_ld_32x32(tS, r0);
        if (n32 > 1) tmem_ld_32x32(tS + 32, r1);
        if (n32 > 2) tmem_ld_32x32(tS + 64, r2);
        if (n32 > 3) tmem_ld_32x32(tS + 96, r3);
        if (tail16) tmem_ld_32x16(tS + 32 * n32, rt);
        tmem_ld_wait();
        float mx = -INFINITY;
        auto max32 = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + 32 * c + j < kvl) mx = fmaxf(mx, __uint_as_float(r[j]));
        };
        if (n32 > 0) max32(r0, 0);
        if (n32 > 1) max32(r1, 1);
        if (n32 > 2) max32(r2, 2);
        if (n32 > 3) max32(r3, 3);
        if (tail16) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (key0 + 32 * n32 + j < kvl) mx = fmaxf(mx, __uint_as_float(rt[j]));
        }
        if (tr) tr[5] = clock64();
        xmax[part * 128 + row] = mx;
        named_bar_sync(1 + q, 64);
        mx = fmaxf(xmax[row], xmax[128 + row]);
        const float off = (mx == -INFINITY) ? 0.f : mx * scale_log2;
        auto softmax32 = [&](const uint32_t (&r)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = ex2(fmaf(__uint_as_float(r[j]), scale_log2, -off));
            const float e1 = ex2(fmaf(__uint_as_float(r[j + 1]), scale_log2, -off));
            const float p0 = (key0 + 32 * c + j < kvl) ? e0 : 0.f;
            const float p1 = (key0 + 32 * c + j + 1 < kvl) ? e1 : 0.f;
            sum += p0 + p1;
            pk[j >> 1] = pack_half2(p0, p1);
          }
          tmem_st_32x16(tP + 16 * c, pk);
        };
        if (n32 > 0) softmax32(r0, 0);
        if (n32 > 1) softmax32(r1, 1);
        if (n32 > 2) softmax32(r2, 2);
        if (n32 > 3) softmax32(r3, 3);
        if (tail16) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float e0 = ex2(fmaf(__uint_as_float(rt[j]), scale_log2, -off));
            const float e1 = ex2(fmaf(__uint_as_float(rt[j + 1]), scale_log2, -off));
            const float p0 = (key0 + 32 * n32 + j < kvl) ? e0 : 0.f;
            const float p1 = (key0 + 32 * n32 + j + 1 < kvl) ? e1 : 0.f;
            sum += p0 + p1;
            pk[j >> 1] = pack_half2(p0, p1);
          }
          tmem_st_32x8(tP + 16 * n32, pk);
        }
      }
      xsum[part * 128 + row] = sum;
      tmem_st_wait();
      fence_proxy_async_all();      // this warp's V rows (local or in the peer's shared memory) -> visible to the tensor cores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(p_ready_leader);   // release.cluster
      if (tr) tr[6] = clock64();
      // ---- O / rowsum -> fp16 -> slab -> TMA store (this warp: rows [32q, +32), dh [64 part, +64))
      mbar_wait(o_done, par);
      tc_fence_after();
      if (tr) tr[7] = clock64();
      named_bar_sync(1 + q, 64);
      const float tot = xsum[row] + xsum[128 + row];
      const float inv = tot > 0.f ? 1.f / tot : 0.f;
      {
        const uint32_t tO = trow + 384 + 64 * part;
        uint32_t ro[64];
        tmem_ld_32x64(tO, ro);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tmem_free_leader);   // the accumulator columns may be overwritten
        auto store_o = [&](int c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t* r = ro + 32 * c + 8 * j;
            const uint32_t w0 = pack_half2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
            const uint32_t w1 = pack_half2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
            const uint32_t w2 = pack_half2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
            const uint32_t w3 = pack_half2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
            *reinterpret_cast<uint4*>(o_slab + slab_off(lane, c * 4 + j)) = make_uint4(w0, w1, w2, w3);
          }
        };
        store_o(0);
        store_o(1);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int row0 = 128 * static_cast<int>(rank) + 32 * q;
        if (row0 < S) {
          tma_store_3d(&map_o, o_slab, h * 128 + 64 * part, row0, smp);
          bulk_commit_group();
        }
      }
      __syncwarp();
      if (tr) tr[8] = clock64();
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  
  } else {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(QA_REGS_CTRL));
  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int smp = item >> 2, h = item & 3;
        const int wrow = (leader ? 0 : 512) + h * 128;          // Wq_h rows (CTA 0) / Wk_h rows (CTA 1)
        const int vrow = 1024 + h * 128 + 64 * static_cast<int>(rank);
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * QA_STAGE_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * QA_STAGE_BYTES);
          tma_load_3d_2cta(sa, &map_h, leader_full, kb * 64, 128 * static_cast<int>(rank), smp);
          tma_load_2d_2cta(sa + 16384, &map_w128, leader_full, kb * 64, wrow);
          tma_load_2d_2cta(sa + 32768, &map_w64, leader_full, kb * 64, vrow);
          if (++stage == QA_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(256, 256);
      constexpr uint32_t idesc_v = umma_idesc_f16(256, 128);
      constexpr uint32_t idesc_s = umma_idesc_f16(256, 256);
      constexpr uint32_t idesc_o = umma_idesc_f16(256, 128, 0, 1);   // B (= V) is MN-major
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
        const uint32_t par = it & 1;
        const int nk16 = max(1, (min(kvlen[item >> 2], S) + 15) >> 4);   // 16-key steps of P V that hold a valid key
        const int split_blk = (nk16 + 1) >> 1;                          // P of blocks [0, split) at column 8 kk, the rest at 16 split + 8 (kk - split)
        long long* tr = B200_TRACE_PTR(cluster_id == 0 && it >= 1 && it < 3, g_gemm2_trace + 32 + (it - 1) * 8);
        if (tr) tr[0] = clock64();
        // projection: [0,384) was last read by the previous item's projection epilogue (Q|K, V accumulators: the previous
        // qkv_ready / p_ready were waited for) and by the previous P V MMA (P in [128,256): MMAs execute in issue order)
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * QA_STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(sa), db1 = umma_desc_k_sw128(sa + 16384), db2 = umma_desc_k_sw128(sa + 32768);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ss_2cta(tmem_base, da + 2 * k, db1 + 2 * k, idesc_qk, (kb | k) != 0);
            umma_f16_ss_2cta(tmem_base + 256, da + 2 * k, db2 + 2 * k, idesc_v, (kb | k) != 0);
          }
          umma_commit_2cta_mc(&empty_bar[stage], 0b11);
          if (++stage == QA_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta_mc(proj_done, 0b11);
        if (tr) tr[1] = clock64();
        // S = Q K^T over the (read) Q|K accumulator columns [0,256)
        mbar_wait_cluster(qkv_ready, par);
        tc_fence_after();
        if (tr) tr[2] = clock64();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_f16_ss_2cta(tmem_base, umma_desc_k_sw128(smem_u32(sQ) + off), umma_desc_k_sw128(smem_u32(sK) + off),
                           idesc_s, ks != 0);
        }
        umma_commit_2cta_mc(s_done, 0b11);
        // O = P V into [384,512): the previous item's O has been read by the output epilogues of both CTAs
        mbar_wait_cluster(p_ready, par);
        if (it > 0) mbar_wait_cluster(tmem_free, par ^ 1);
        tc_fence_after();
        if (tr) tr[3] = clock64();
#pragma unroll 4
        for (int kk = 0; kk < nk16; ++kk)
          umma_f16_ts_2cta(tmem_base + 384, tmem_base + (kk < split_blk ? 8 * kk : 16 * split_blk + 8 * (kk - split_blk)),
                           umma_desc_mn_sw128(smem_u32(sV) + kk * 2048, 0, 1024), idesc_o, kk != 0);
        umma_commit_2cta_mc(o_done, 0b11);
        if (tr) tr[4] = clock64();
      }
    }
  }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace b200
