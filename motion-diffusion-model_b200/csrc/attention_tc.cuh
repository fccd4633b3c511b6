// Self-attention core on the 5th-generation tensor cores: one CTA per (sample, head), two 128-row query tiles.
//   S = Q K^T        tcgen05.mma SS   (Q, K: K-major 128B-swizzled shared-memory tiles brought by TMA)
//   P = softmax(S)   single pass over all keys (<= 256): fp32 in registers, scale folded into exp2, prefix key mask;
//                    P is written back to TMEM as packed fp16 *over* the S columns it came from (tcgen05.st)
//   O = P V          tcgen05.mma TS   (A = P from TMEM, B = V as an MN-major shared-memory operand: V is [key, dh]
//                    with dh contiguous, exactly what the QKV projection wrote -- no transpose anywhere)
//   O / rowsum -> fp16 -> swizzled slabs (re-using the dead Q tile) -> TMA store
// (reference: nn.MultiheadAttention inside nn.TransformerEncoderLayer, built at model/mdm.py:77-84; the
//  key_padding_mask of model/mdm.py:241-247 is a prefix mask => per-sample valid-key count `kvlen`.)
//
// Warp roles (320 threads): warp 0 TMA loader, warp 1 TMEM allocator + MMA issuer, warps 2-5 softmax/epilogue of
// query tile 0, warps 6-9 of query tile 1 (warp w owns TMEM lanes 32*(w%4)..+31; thread = query row).
// TMEM (512 columns): tile i uses columns [256 i, 256 i + 256):  S at +0..+keys, P (fp16 pairs) at +0..+keys/2,
// O at +128..+256 (written only after the softmax has consumed S).
// While tile 0 is in its softmax, the tensor core runs QK^T of tile 1; PV of tile 0 overlaps the softmax of tile 1.
#pragma once
#include <cuda_fp16.h>

#include "epilogues.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int ATC_THREADS = 320;
constexpr int ATC_DH = 128;
constexpr int ATC_MAX_KEYS = 256;

struct AttnTcSmem {
  // all regions are multiples of 1024 bytes (128B-swizzle atoms); key-dependent sizes are computed at run time
  static __host__ __device__ constexpr int q_bytes() { return 2 * 2 * 128 * 128; }  // 2 tiles x 2 dh-atoms x [128 x 128 B]
  static __host__ __device__ int kv_atom_bytes(int keys) { return keys * 128; }       // one dh-atom of K or V
  static __host__ __device__ int total(int keys) { return 1024 + q_bytes() + 4 * kv_atom_bytes(keys) + 256; }
};

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// map_q : qkv16 viewed [n_samples][S][3d], box {64, 128, 1}        (Q tiles)
// map_kv: same view, box {64, keys, 1}                              (K / V atoms; keys = round_up(S, 16) <= 256)
// map_o : att16 viewed [n_samples][S][d], box {64, 32, 1}            (per-warp output slabs)
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_o, const int* __restrict__ kvlen, int S, int d, int keys,
                    float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_atom = keys * 128;
  uint8_t* sQ = smem;                         // [tile][atom][128 x 128 B]
  uint8_t* sK = sQ + AttnTcSmem::q_bytes();   // [atom][keys x 128 B]
  uint8_t* sV = sK + 2 * kv_atom;             // [atom][keys x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kv_atom);
  uint64_t* bar_qk = bars;        // [2]  Q tile i (+ K for i = 0) landed
  uint64_t* bar_v = bars + 2;     //      V landed
  uint64_t* bar_s = bars + 3;     // [2]  S_i = Q_i K^T complete
  uint64_t* bar_p = bars + 5;     // [2]  P_i written to TMEM (4 warp arrivals)
  uint64_t* bar_o = bars + 7;     // [2]  O_i = P_i V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, smp = blockIdx.y;
  const int n_tiles = (S > 128) ? 2 : 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    tma_prefetch_desc(&map_o);
    mbar_init(&bar_qk[0], 1);
    mbar_init(&bar_qk[1], 1);
    mbar_init(bar_v, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 4);
      mbar_init(&bar_o[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader
    if (elect_one()) {
      const int cq = h * ATC_DH, ck = d + h * ATC_DH, cv = 2 * d + h * ATC_DH;
      mbar_expect_tx(&bar_qk[0], 2 * 128 * 128 + 2 * kv_atom);
      for (int a = 0; a < 2; ++a) tma_load_3d(sQ + a * 16384, &map_q, &bar_qk[0], cq + 64 * a, 0, smp);
      for (int a = 0; a < 2; ++a) tma_load_3d(sK + a * kv_atom, &map_kv, &bar_qk[0], ck + 64 * a, 0, smp);
      if (n_tiles == 2) {
        mbar_expect_tx(&bar_qk[1], 2 * 128 * 128);
        for (int a = 0; a < 2; ++a) tma_load_3d(sQ + 32768 + a * 16384, &map_q, &bar_qk[1], cq + 64 * a, 128, smp);
      }
      mbar_expect_tx(bar_v, 2 * kv_atom);
      for (int a = 0; a < 2; ++a) tma_load_3d(sV + a * kv_atom, &map_kv, bar_v, cv + 64 * a, 0, smp);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_f16(128, keys);
      const uint32_t idesc_o = umma_idesc_f16(128, ATC_DH, 0, 1);  // B (= V) is MN-major
      // S_i = Q_i K^T for both tiles back to back
      for (int i = 0; i < n_tiles; ++i) {
        mbar_wait(&bar_qk[i], 0);
        if (i == 1) mbar_wait(&bar_qk[0], 0);
        tc_fence_after();
        const uint32_t tS = tmem_base + i * 256;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + i * 32768 + (ks >> 2) * 16384) + (ks & 3) * 32);
          const uint64_t db = umma_desc_k_sw128(smem_u32(sK + (ks >> 2) * kv_atom) + (ks & 3) * 32);
          umma_f16_ss(tS, da, db, idesc_s, ks != 0);
        }
        umma_commit(&bar_s[i]);
      }
      // O_i = P_i V
      mbar_wait(bar_v, 0);
      for (int i = 0; i < n_tiles; ++i) {
        mbar_wait(&bar_p[i], 0);
        tc_fence_after();
        const uint32_t tP = tmem_base + i * 256;
        const uint32_t tO = tmem_base + i * 256 + 128;
        const int nk = keys >> 4;
        for (int kk = 0; kk < nk; ++kk) {
          // V operand: N (= dh) spans the two 64-wide atoms (LBO = atom size), K (= keys) advances 16 rows = 2048 B
          const uint64_t db = umma_desc_mn_sw128(smem_u32(sV) + kk * 2048, kv_atom, 1024);
          umma_f16_ts(tO, tP + kk * 8, db, idesc_o, kk != 0);
        }
        umma_commit(&bar_o[i]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output warps
    const int tile = (warp - 2) >> 2;   // 0 or 1
    const int q = warp & 3;             // TMEM lane quarter
    if (tile < n_tiles) {
      const int kvl = min(kvlen[smp], S);
      const uint32_t tS = tmem_base + tile * 256 + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t tO = tS + 128;
      mbar_wait(&bar_s[tile], 0);
      tc_fence_after();
      const int n32 = keys >> 5, tail16 = keys & 16;
      // ---- pass 1: row maximum over the valid keys
      float mx = -INFINITY;
      for (int c = 0; c < n32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tS + 32 * c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 * c + j < kvl) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      if (tail16) {
        uint32_t r[16];
        tmem_ld_32x16(tS + 32 * n32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (32 * n32 + j < kvl) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      const float off = (mx == -INFINITY) ? 0.f : mx * scale_log2;
      // ---- pass 2: p = exp2(s*scale - max*scale); P (fp16 pairs) overwrites the S columns it trails
      float sum = 0.f;
      for (int c = 0; c < n32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tS + 32 * c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = (32 * c + j < kvl) ? exp2f(fmaf(__uint_as_float(r[j]), scale_log2, -off)) : 0.f;
          float p1 = (32 * c + j + 1 < kvl) ? exp2f(fmaf(__uint_as_float(r[j + 1]), scale_log2, -off)) : 0.f;
          sum += p0 + p1;
          pk[j >> 1] = pack_half2(p0, p1);
        }
        tmem_st_32x16(tS + 16 * c, pk);
      }
      if (tail16) {
        uint32_t r[16];
        tmem_ld_32x16(tS + 32 * n32, r);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p0 = (32 * n32 + j < kvl) ? exp2f(fmaf(__uint_as_float(r[j]), scale_log2, -off)) : 0.f;
          float p1 = (32 * n32 + j + 1 < kvl) ? exp2f(fmaf(__uint_as_float(r[j + 1]), scale_log2, -off)) : 0.f;
          sum += p0 + p1;
          pk[j >> 1] = pack_half2(p0, p1);
        }
        tmem_st_32x8(tS + 16 * n32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[tile]);
      // ---- O / rowsum -> fp16 slabs (the Q tile is dead once bar_s fired) -> TMA store
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      mbar_wait(&bar_o[tile], 0);
      tc_fence_after();
      uint8_t* slab0 = sQ + tile * 32768 + q * 4096;   // rows [32q, 32q+32) of dh-atom 0
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tO + 32 * c, r);
        tmem_ld_wait();
        uint8_t* slab = slab0 + (c >> 1) * 16384;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v4;
          v4.x = pack_half2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
          v4.y = pack_half2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
          v4.z = pack_half2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
          v4.w = pack_half2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(slab + slab_off(lane, (c & 1) * 4 + j)) = v4;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int row0 = tile * 128 + q * 32;
        if (row0 < S) {
          tma_store_3d(&map_o, slab0, h * ATC_DH, row0, smp);
          tma_store_3d(&map_o, slab0 + 16384, h * ATC_DH + 64, row0, smp);
          bulk_commit_group();
          bulk_wait_group<0>();
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200
