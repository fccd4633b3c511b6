// Self-attention core on the 5th-generation tensor cores: one CTA per (sample, head, 128-row query tile), two CTAs
// resident per SM so that one CTA's loads / softmax overlap the other's MMAs.
//   S = Q K^T        tcgen05.mma SS   (Q, K: K-major 128B-swizzled shared-memory tiles brought by TMA)
//   P = softmax(S)   single pass over all keys (<= 256): fp32 in registers, scale folded into exp2, prefix key mask;
//                    P is written back to TMEM as packed fp16 *over* the S columns it came from (tcgen05.st)
//   O = P V          tcgen05.mma TS   (A = P from TMEM, B = V as an MN-major shared-memory operand: V is [key, dh]
//                    with dh contiguous, exactly what the QKV projection wrote -- no transpose anywhere).  V is loaded
//                    into the shared memory K occupied, as soon as QK^T has consumed K: its latency hides behind the
//                    softmax, and the CTA needs only 32 KB (Q) + 128 B x keys x 2 (K, then V) of shared memory.
//   O / rowsum -> fp16 -> swizzled slabs (re-using the dead Q tile) -> TMA store
// (reference: nn.MultiheadAttention inside nn.TransformerEncoderLayer, built at model/mdm.py:77-84; the
//  key_padding_mask of model/mdm.py:241-247 is a prefix mask => per-sample valid-key count `kvlen`.)
//
// Warp roles (192 threads): warp 0 TMA loader, warp 1 TMEM allocator + MMA issuer, warps 2-5 softmax / output
// (warp w owns TMEM lanes 32*(w%4)..+31; thread = query row).
// TMEM (256 columns per CTA): S at +0..+keys, P (fp16 pairs) at +0..+keys/2, O at +128..+256 (written only after
// the softmax has consumed S).
#pragma once
#include <cuda_fp16.h>

#include "epilogues.cuh"
#include "gemm2.cuh"   // g_gemm2_trace (debug stamps)
#include "ptx.cuh"

namespace b200 {

constexpr int ATC_THREADS = 192;
constexpr int ATC_DH = 128;
constexpr int ATC_MAX_KEYS = 256;

struct AttnTcSmem {
  static __host__ __device__ constexpr int q_bytes() { return 2 * 128 * 128; }    // 2 dh-atoms x [128 rows x 128 B]
  static __host__ __device__ int kv_atom_bytes(int keys) { return keys * 128; }   // one dh-atom of K (later V)
  static __host__ __device__ int total(int keys) { return 1024 + q_bytes() + 2 * kv_atom_bytes(keys) + 128; }
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

// map_q : qkv16 viewed [n_samples][S][3d], box {64, 128, 1}        (Q tile)
// map_kv: same view, box {64, keys, 1}                              (K / V atoms; keys = round_up(S, 16) <= 256)
// map_o : att16 viewed [n_samples][S][d], box {64, 32, 1}            (per-warp output slabs)
//         WIDE: [n_samples][S][2d] = [hi | lo] with hi + lo = O to ~22 bits (trans_dec engine)
// grid = (heads, n_samples, ceil(S / 128))
template <bool WIDE>
__global__ void __launch_bounds__(ATC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_o, const int* __restrict__ kvlen, int S, int d, int keys,
                    float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_atom = keys * 128;
  uint8_t* sQ = smem;                         // [atom][128 x 128 B]; later the output slabs
  uint8_t* sKV = sQ + AttnTcSmem::q_bytes();  // [atom][keys x 128 B]: K, then V
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + 2 * kv_atom);
  uint64_t* bar_qk = bars;        // Q + K landed
  uint64_t* bar_v = bars + 1;     // V landed (in K's place)
  uint64_t* bar_s = bars + 2;     // S = Q K^T complete (K and Q are dead)
  uint64_t* bar_p = bars + 3;     // P written to TMEM (4 warp arrivals)
  uint64_t* bar_o = bars + 4;     // O = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, smp = blockIdx.y, tile = blockIdx.z;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    tma_prefetch_desc(&map_o);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA loader
    if (elect_one()) {
      const int cq = h * ATC_DH, ck = d + h * ATC_DH, cv = 2 * d + h * ATC_DH;
      mbar_expect_tx(bar_qk, 2 * 128 * 128 + 2 * kv_atom);
      for (int a = 0; a < 2; ++a) tma_load_3d(sQ + a * 16384, &map_q, bar_qk, cq + 64 * a, tile * 128, smp);
      for (int a = 0; a < 2; ++a) tma_load_3d(sKV + a * kv_atom, &map_kv, bar_qk, ck + 64 * a, 0, smp);
      // V replaces K as soon as the tensor core has finished reading K
      mbar_wait(bar_s, 0);
      mbar_expect_tx(bar_v, 2 * kv_atom);
      for (int a = 0; a < 2; ++a) tma_load_3d(sKV + a * kv_atom, &map_kv, bar_v, cv + 64 * a, 0, smp);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_f16(128, keys);
      const uint32_t idesc_o = umma_idesc_f16(128, ATC_DH, 0, 1);  // B (= V) is MN-major
      mbar_wait(bar_qk, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + (ks >> 2) * 16384) + (ks & 3) * 32);
        const uint64_t db = umma_desc_k_sw128(smem_u32(sKV + (ks >> 2) * kv_atom) + (ks & 3) * 32);
        umma_f16_ss(tmem_base, da, db, idesc_s, ks != 0);
      }
      umma_commit(bar_s);
      // O = P V
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const int nk = keys >> 4;
      for (int kk = 0; kk < nk; ++kk) {
        // V operand: N (= dh) spans the two 64-wide atoms (LBO = atom size), K (= keys) advances 16 rows = 2048 B
        const uint64_t db = umma_desc_mn_sw128(smem_u32(sKV) + kk * 2048, kv_atom, 1024);
        umma_f16_ts(tmem_base + 128, tmem_base + kk * 8, db, idesc_o, kk != 0);
      }
      umma_commit(bar_o);
    }
  } else {
    // ------------------------------------------------------------------ softmax + output warps
    const int q = warp & 3;             // TMEM lane quarter
    long long* tr = B200_TRACE_PTR(blockIdx.x == 1 && blockIdx.y == 37 && blockIdx.z == 0 && warp == 2 && lane == 0, g_gemm2_trace);
    if (tr) tr[0] = clock64();
    const int kvl = min(kvlen[smp], S);
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tO = tS + 128;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (tr) tr[1] = clock64();
    const int n32 = keys >> 5, tail16 = keys & 16;
    // ---- pass 1: row maximum over the valid keys (next chunk's tcgen05.ld in flight while this one is reduced)
    float mx = -INFINITY;
    {
      uint32_t ra[32], rb[32];
      uint32_t rt[16];
      auto max32 = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 * c + j < kvl) mx = fmaxf(mx, __uint_as_float(r[j]));
      };
      if (n32 > 0) tmem_ld_32x32(tS, ra);
#pragma unroll 1
      for (int c = 0; c < n32; c += 2) {
        tmem_ld_wait();
        if (c + 1 < n32) tmem_ld_32x32(tS + 32 * (c + 1), rb);
        else if (tail16) tmem_ld_32x16(tS + 32 * n32, rt);
        max32(ra, c);
        if (c + 1 < n32) {
          tmem_ld_wait();
          if (c + 2 < n32) tmem_ld_32x32(tS + 32 * (c + 2), ra);
          else if (tail16) tmem_ld_32x16(tS + 32 * n32, rt);
          max32(rb, c + 1);
        }
      }
      if (tail16) {
        if (n32 == 0) tmem_ld_32x16(tS, rt);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (32 * n32 + j < kvl) mx = fmaxf(mx, __uint_as_float(rt[j]));
      }
    }
    const float off = (mx == -INFINITY) ? 0.f : mx * scale_log2;
    if (tr) tr[2] = clock64();
    // ---- pass 2: p = exp2(s*scale - max*scale); P (fp16 pairs) overwrites the S columns it trails.  The load of
    // chunk c+1 is issued BEFORE the store of chunk c (register double buffer): a tcgen05.ld queued behind a
    // tcgen05.st of the same thread otherwise waits for the store (measured: 1600 cycles per chunk vs 290 in pass 1).
    float sum = 0.f;
    auto ex2 = [](float x) {
      float y;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      return y;
    };
    auto softmax32 = [&](uint32_t (&r)[32], int c) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float e0 = ex2(fmaf(__uint_as_float(r[j]), scale_log2, -off));
        const float e1 = ex2(fmaf(__uint_as_float(r[j + 1]), scale_log2, -off));
        const float p0 = (32 * c + j < kvl) ? e0 : 0.f;
        const float p1 = (32 * c + j + 1 < kvl) ? e1 : 0.f;
        sum += p0 + p1;
        pk[j >> 1] = pack_half2(p0, p1);
      }
      tmem_st_32x16(tS + 16 * c, pk);
    };
    {
      uint32_t ra[32], rb[32];
      uint32_t rt[16];
      if (n32 > 0) tmem_ld_32x32(tS, ra);
#pragma unroll 1
      for (int c = 0; c < n32; c += 2) {
        tmem_ld_wait();
        if (c + 1 < n32) tmem_ld_32x32(tS + 32 * (c + 1), rb);
        else if (tail16) tmem_ld_32x16(tS + 32 * n32, rt);
        softmax32(ra, c);
        if (c + 1 < n32) {
          tmem_ld_wait();
          if (c + 2 < n32) tmem_ld_32x32(tS + 32 * (c + 2), ra);
          else if (tail16) tmem_ld_32x16(tS + 32 * n32, rt);
          softmax32(rb, c + 1);
        }
      }
      if (tail16) {
        if (n32 == 0) tmem_ld_32x16(tS, rt);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float e0 = ex2(fmaf(__uint_as_float(rt[j]), scale_log2, -off));
          const float e1 = ex2(fmaf(__uint_as_float(rt[j + 1]), scale_log2, -off));
          const float p0 = (32 * n32 + j < kvl) ? e0 : 0.f;
          const float p1 = (32 * n32 + j + 1 < kvl) ? e1 : 0.f;
          sum += p0 + p1;
          pk[j >> 1] = pack_half2(p0, p1);
        }
        tmem_st_32x8(tS + 16 * n32, pk);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    if (tr) tr[3] = clock64();
    // ---- O / rowsum -> fp16 slabs (the Q tile is dead once bar_s fired) -> TMA store
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    mbar_wait(bar_o, 0);
    tc_fence_after();
    if (tr) tr[4] = clock64();
    uint8_t* slab0 = sQ + q * 4096;   // rows [32q, 32q+32) of dh-atom 0
#pragma unroll 1
    for (int part = 0; part < (WIDE ? 2 : 1); ++part) {
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tO + 32 * c, r);
        tmem_ld_wait();
        uint8_t* slab = slab0 + (c >> 1) * 16384;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x0 = __uint_as_float(r[8 * j + 2 * i]) * inv, x1 = __uint_as_float(r[8 * j + 2 * i + 1]) * inv;
            w[i] = pack_half2(x0, x1);
            if (WIDE && part == 1) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
              w[i] = pack_half2(x0 - f.x, x1 - f.y);
            }
          }
          *reinterpret_cast<uint4*>(slab + slab_off(lane, (c & 1) * 4 + j)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if (tr && part == 0) tr[5] = clock64();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int row0 = tile * 128 + q * 32;
        if (row0 < S) {
          const int colbase = h * ATC_DH + part * d;
          tma_store_3d(&map_o, slab0, colbase, row0, smp);
          tma_store_3d(&map_o, slab0 + 16384, colbase + 64, row0, smp);
          bulk_commit_group();
          bulk_wait_group<0>();
        }
      }
      __syncwarp();
    }
    __syncwarp();
    if (tr) tr[6] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace b200
