// Fused GEMM epilogues (see gemm.cuh for the calling protocol).  Thread = accumulator row.
//
// Row-major outputs are written as 128-byte-per-row slabs (32 rows of the warp x 64 fp16 or 32 fp32 columns = 4 KB)
// staged in warp-private shared memory in the TMA 128-byte swizzle and shipped with cp.async.bulk.tensor stores;
// residual inputs arrive the same way with TMA loads.  No global load/store instruction is issued by these
// epilogues except broadcast bias reads, and out-of-range rows / columns are clipped by the TMA unit.
// Feature-major outputs ([B, J, T], T contiguous) are written straight from registers because consecutive
// accumulator rows are consecutive frames (coalesced along T).
#pragma once
#include <cuda_fp16.h>

#include "gemm.cuh"
#include "ptx.cuh"

#ifndef B200_WIDE_PACKED
#define B200_WIDE_PACKED 1   // 0: scalar GELU + conversion-based hi/lo split in the [hi | lo] epilogue of the trans_dec FFN (A/B builds)
#endif
#ifndef B200_GELU_PACKED
#define B200_GELU_PACKED 1   // 0: the scalar GELU epilogue (A/B builds)
#endif

namespace b200 {

// byte offset of 16-byte chunk j (0..7) of row r inside a [32 x 128 B] slab with the TMA SWIZZLE_128B pattern
__device__ __forceinline__ uint32_t slab_off(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }
// byte offset of 16-byte chunk j (0..3) of row r inside a [32 x 64 B] half-slab with the TMA SWIZZLE_64B pattern
// (address bits 4-5 xor bits 7-8).  A residual-stream slab is two of these: hi halves at +0, lo halves at +2048.
__device__ __forceinline__ uint32_t slab64_off(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

// 8 fp32 values -> 4 packed half2 "hi" words + 4 packed half2 "lo" words with hi + lo = v to ~22 bits
__device__ __forceinline__ void split_hi_lo8(const float (&v)[8], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
}
// One pair: hi = fp16(v), lo = fp16(v - float(hi)) with the subtraction as a mixed-precision add (add.f32.f16 -> FHADD
// with the negation folded into the operand: no separate fp16 -> fp32 conversion).  Bit-identical to split_hi_lo8.
__device__ __forceinline__ void split_hi_lo2(float2 v, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v.x, v.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  float2 r;
  asm("{\n\t.reg .b16 l, u;\n\t.reg .b32 n;\n\tneg.f16x2 n, %2;\n\tmov.b32 {l, u}, n;\n\t"
      "add.f32.f16 %0, l, %3;\n\tadd.f32.f16 %1, u, %4;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "r"(hi), "f"(v.x), "f"(v.y));
  const __half2 l2 = __floats2half2_rn(r.x, r.y);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}
// the inverse for one 16-byte group of hi and one of lo: 8 fp32 values
__device__ __forceinline__ void join_hi_lo8(const uint4& h, const uint4& l, float (&v)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
    v[2 * i] = a.x + b.x;
    v[2 * i + 1] = a.y + b.y;
  }
}

// exact (erf) GELU, torch F.gelu default (reference model/mdm.py:80 activation="gelu"):  gelu(x) = x * Phi(x).
// Phi(-t) = 2^q(t) with q a degree-6 minimax fit of log2(0.5*erfc(t/sqrt 2)) on [0, 5.5] (|Phi error| < 1.5e-7,
// |gelu error| < 8e-7 over all x, checked against torch fp64 in tests); Phi(t) = 1 - Phi(-t).  11 FP32 ops + one
// MUFU.EX2 instead of the ~25 of erff() -- the epilogue of the FFN up-projection runs 26 M of these per layer.
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fminf(fabsf(x), 5.5f);
  float q = 1.9175331544829533e-05f;
  q = fmaf(q, t, -0.0006586098461411893f);
  q = fmaf(q, t, 0.007754423655569553f);
  q = fmaf(q, t, -0.05296541005373001f);
  q = fmaf(q, t, -0.4590602517127991f);
  q = fmaf(q, t, -1.1511220932006836f);
  q = fmaf(q, t, -0.9999997019767761f);
  float a;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(q));
  const float phi = (x >= 0.f) ? (1.0f - a) : a;
  return x * phi;
}

// The same arithmetic (bit for bit) on 16 values at once, two per 64-bit register: Blackwell's packed fp32 instructions
// (fma.rn.f32x2 & co: FFMA2 / FADD2 / FMUL2 in SASS) halve the instruction count of the Horner chain -- 9 instead of 13.5
// instructions per element for the whole bias + GELU + fp16 epilogue (ptxas keeps two or three of the eight independent
// chains in flight whatever the source order or `asm volatile` says).  Measured (profiles/r02_m_*, r02_n_*): warp
// instructions of the FFN up-projection 12.6 M -> 9.8 M per launch, kernel 34.3 -> 33.3 us, loop 0.5-0.8 % shorter with an
// identical checksum: once its operands are resident (gemm2w.cuh) that GEMM is paced by the TMEM port (accumulator
// read-modify-writes of the running MMAs against the epilogue's tcgen05.ld: DESIGN.md section 4.1), not by ALU issue.
__device__ __forceinline__ void gelu_erf_x16(float2 (&x)[8]) {
  float2 t[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = make_float2(fminf(fabsf(x[i].x), 5.5f), fminf(fabsf(x[i].y), 5.5f));
  const float c6 = 1.9175331544829533e-05f, c5 = -0.0006586098461411893f, c4 = 0.007754423655569553f,
              c3 = -0.05296541005373001f, c2 = -0.4590602517127991f, c1 = -1.1511220932006836f, c0 = -0.9999997019767761f;
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(make_float2(c6, c6), t[i], make_float2(c5, c5));
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(q[i], t[i], make_float2(c4, c4));
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(q[i], t[i], make_float2(c3, c3));
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(q[i], t[i], make_float2(c2, c2));
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(q[i], t[i], make_float2(c1, c1));
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(q[i], t[i], make_float2(c0, c0));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q[i].x) : "f"(q[i].x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q[i].y) : "f"(q[i].y));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 om = __fadd2_rn(make_float2(1.0f, 1.0f), make_float2(-q[i].x, -q[i].y));   // 1 - a
    const float2 phi = make_float2(x[i].x >= 0.f ? om.x : q[i].x, x[i].y >= 0.f ? om.y : q[i].y);
    x[i] = __fmul2_rn(x[i], phi);
  }
}

// Stage the bias of the tile in flight (up to 256 columns) in warp-private shared memory so that the row-owning
// threads read it with broadcast LDS instead of dependent global loads inside the chunk loop.
__device__ __forceinline__ void stage_bias(float* bs, const float* bias, int col_base, int N, int lane) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = i * 128 + lane * 4;
    const int col = col_base + idx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col + 3 < N) {
      v = __ldg(reinterpret_cast<const float4*>(bias + col));
    } else {
      if (col + 0 < N) v.x = bias[col + 0];
      if (col + 1 < N) v.y = bias[col + 1];
      if (col + 2 < N) v.z = bias[col + 2];
    }
    *reinterpret_cast<float4*>(bs + idx) = v;
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------
// out16[row, col] = fp16( act(acc + bias[col]) )          (QKV projection, FFN up-projection)
// Two 32-column chunks fill one 64-column (128-byte) slab; slabs are double-buffered.  map_c: fp16 [M, N],
// box {64 cols, 32 rows}, SWIZZLE_128B.
template <bool GELU>
struct EpiBiasF16 {
  // one output slab + the tile's bias: with 8 epilogue warps the wait for the previous slab's TMA store overlaps the
  // other warps' work, and the 4 KB saved per warp buy one more operand pipeline stage (the layer GEMMs are bound
  // by operand bytes in flight)
  static constexpr int NSLAB = 1;
  static constexpr int SMEM_PER_WARP = NSLAB * 4096;
  static constexpr bool RELEASE_EARLY = true;
  struct Params {
    const float* bias;
  };
  // the bias of ALL N columns is staged once per CTA (a per-tile global load sat on the epilogue's critical path:
  // 1-2.7 k cycles between "accumulator ready" and the first chunk, measured with clock64 stamps)
  static __device__ __forceinline__ void preload(const Params& p, float* dst, int N, int tid, int nthreads) {
    for (int i = tid; i < N; i += nthreads) dst[i] = p.bias[i];
  }
  static __device__ __forceinline__ void tile_begin(EpiCtx&, const Params&, int, int) {}
  static __device__ __forceinline__ void chunk(EpiCtx& ctx, const Params&, uint32_t (&raw)[32], int row0, int col0, int) {
    chunk_bias(ctx, ctx.bias_all + col0, raw, row0, col0);   // chunk() is only called for col0 < N, N % 32 == 0 here
  }
  static __device__ __forceinline__ void chunk_bias(EpiCtx& ctx, const float* bs, uint32_t (&raw)[32], int row0, int col0) {
    const int half = (col0 >> 5) & 1;
    uint8_t* slab = ctx.smem + (ctx.seq % NSLAB) * 4096;
    if (half == 0) {
      // the store that last read this buffer (NSLAB slabs ago) must have finished reading
      if (ctx.lane == 0) bulk_wait_group_read<NSLAB - 1>();
      __syncwarp();
    }
    uint32_t pk[16];
#if B200_GELU_PACKED
    if (GELU) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {       // 16 columns at a time: eight packed pairs (see gelu_erf_x16)
        float2 x[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = *reinterpret_cast<const float4*>(bs + 16 * g + 4 * i);
          x[2 * i] = __fadd2_rn(make_float2(__uint_as_float(raw[16 * g + 4 * i]), __uint_as_float(raw[16 * g + 4 * i + 1])), make_float2(b.x, b.y));
          x[2 * i + 1] = __fadd2_rn(make_float2(__uint_as_float(raw[16 * g + 4 * i + 2]), __uint_as_float(raw[16 * g + 4 * i + 3])), make_float2(b.z, b.w));
        }
        gelu_erf_x16(x);
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[8 * g + i] = pack_half2(x[i].x, x[i].y);
      }
    } else
#endif
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bs + j);
      float x0 = __uint_as_float(raw[j + 0]) + b.x, x1 = __uint_as_float(raw[j + 1]) + b.y;
      float x2 = __uint_as_float(raw[j + 2]) + b.z, x3 = __uint_as_float(raw[j + 3]) + b.w;
      if (GELU) {
        x0 = gelu_erf(x0); x1 = gelu_erf(x1); x2 = gelu_erf(x2); x3 = gelu_erf(x3);
      }
      pk[j / 2] = pack_half2(x0, x1);
      pk[j / 2 + 1] = pack_half2(x2, x3);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(slab + slab_off(ctx.lane, half * 4 + j)) =
          make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    if (half == 1 || col0 + 32 >= ctx.N) {
      fence_proxy_async_smem();
      __syncwarp();
      if (ctx.lane == 0) {
        tma_store_2d(ctx.map_c, slab, col0 - 32 * half, row0);
        bulk_commit_group();
      }
      ctx.seq++;
    }
  }
  static __device__ __forceinline__ void tile_end(EpiCtx&, const Params&, int, int, uint32_t) {}
  static __device__ __forceinline__ void finish(EpiCtx& ctx) {
    if (ctx.lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }
};

// EpiBiasF16 for a GEMM wider than the staged-vector limit (the cross-attention K/V projection of all decoder layers
// at once, N = L * 2d = 8192): the bias is read from global memory per chunk (one broadcast float4 per 4 columns).
struct EpiBiasF16Global : EpiBiasF16<false> {
  static constexpr bool UNSTAGED = true;
  static __device__ __forceinline__ void preload(const Params&, float*, int, int, int) {}
  static __device__ __forceinline__ void chunk(EpiCtx& ctx, const Params& p, uint32_t (&raw)[32], int row0, int col0, int) {
    chunk_bias(ctx, p.bias + col0, raw, row0, col0);
  }
};

// ---------------------------------------------------------------------------------------------------------
// EpiBiasF16 with the result kept as a [hi | lo] fp16 pair: out16[row, col] = hi, out16[row, lo_col + col] = lo with
// hi + lo = act(acc + bias) to ~22 bits (trans_dec engine, where guidance 7.5 amplifies activation rounding; the
// consumer GEMM runs over K = 2N against [W | W]).  map_c: fp16 [M, 2N], box {64 cols, 32 rows}, SWIZZLE_128B.
template <bool GELU>
struct EpiBiasF16Wide {
  static constexpr int SMEM_PER_WARP = 2 * 4096;   // one hi slab + one lo slab
  static constexpr bool RELEASE_EARLY = true;
  struct Params {
    const float* bias;
    int lo_col;
  };
  static __device__ __forceinline__ void preload(const Params& p, float* dst, int N, int tid, int nthreads) {
    for (int i = tid; i < N; i += nthreads) dst[i] = p.bias[i];
  }
  static __device__ __forceinline__ void tile_begin(EpiCtx&, const Params&, int, int) {}
  static __device__ __forceinline__ void chunk(EpiCtx& ctx, const Params& p, uint32_t (&raw)[32], int row0, int col0,
                                               int) {
    const int half = (col0 >> 5) & 1;
    uint8_t* slab_hi = ctx.smem;
    uint8_t* slab_lo = ctx.smem + 4096;
    if (half == 0) {
      if (ctx.lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
    }
    const float* bs = ctx.bias_all + col0;
#if B200_GELU_PACKED && B200_WIDE_PACKED
    if (GELU) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {       // 16 columns at a time (see gelu_erf_x16)
        float2 x[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = *reinterpret_cast<const float4*>(bs + 16 * g + 4 * i);
          x[2 * i] = __fadd2_rn(make_float2(__uint_as_float(raw[16 * g + 4 * i]), __uint_as_float(raw[16 * g + 4 * i + 1])), make_float2(b.x, b.y));
          x[2 * i + 1] = __fadd2_rn(make_float2(__uint_as_float(raw[16 * g + 4 * i + 2]), __uint_as_float(raw[16 * g + 4 * i + 3])), make_float2(b.z, b.w));
        }
        gelu_erf_x16(x);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_hi_lo2(x[i], hi[i], lo[i]);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * g + jj;
          *reinterpret_cast<uint4*>(slab_hi + slab_off(ctx.lane, half * 4 + j)) = make_uint4(hi[4 * jj], hi[4 * jj + 1], hi[4 * jj + 2], hi[4 * jj + 3]);
          *reinterpret_cast<uint4*>(slab_lo + slab_off(ctx.lane, half * 4 + j)) = make_uint4(lo[4 * jj], lo[4 * jj + 1], lo[4 * jj + 2], lo[4 * jj + 3]);
        }
      }
    } else
#endif
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x0 = __uint_as_float(raw[8 * j + 2 * i]) + bs[8 * j + 2 * i];
        float x1 = __uint_as_float(raw[8 * j + 2 * i + 1]) + bs[8 * j + 2 * i + 1];
        if (GELU) {
          x0 = gelu_erf(x0); x1 = gelu_erf(x1);
        }
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
      }
      *reinterpret_cast<uint4*>(slab_hi + slab_off(ctx.lane, half * 4 + j)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(slab_lo + slab_off(ctx.lane, half * 4 + j)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (half == 1 || col0 + 32 >= ctx.N) {
      fence_proxy_async_smem();
      __syncwarp();
      if (ctx.lane == 0) {
        tma_store_2d(ctx.map_c, slab_hi, col0 - 32 * half, row0);
        tma_store_2d(ctx.map_c, slab_lo, p.lo_col + col0 - 32 * half, row0);
        bulk_commit_group();
      }
      ctx.seq++;
    }
  }
  static __device__ __forceinline__ void tile_end(EpiCtx&, const Params&, int, int, uint32_t) {}
  static __device__ __forceinline__ void finish(EpiCtx& ctx) {
    if (ctx.lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------------------
// InputProcess + positional encoding (reference model/mdm.py:238,252,343-349):
//   GEMM rows are (b, s) over B*S, s >= 1 is frame s-1:  h = acc + (bias + pe[s]).  The frame rows are identical for
//   the cond / uncond halves of the packed CFG batch, so every slab is TMA-stored twice (one tensor map per half --
//   the per-half maps also clip the rows of the last M tile that belong to the other half).  Row s == 0 (the
//   conditioning token, mdm.py:251) is produced by tok0_rows_kernel right after this GEMM.
//   Output: the residual stream as fp16 [hi | lo] (hi half = the next GEMM's A operand).
struct EpiEmbed {
  static constexpr int SMEM_PER_WARP = 2 * 4096;  // two [hi | lo] slabs
  static constexpr bool RELEASE_EARLY = true;
  template <class P> static __device__ __forceinline__ void preload(const P&, float*, int, int, int) {}
  struct Params {
    CUtensorMap res_c, res_u;                 // residual stream [rows, 2d] = [hi | lo] of the two CFG halves,
                                              // box {32 cols, 32 rows} (64-byte rows, SWIZZLE_64B)
    const float* pe_bias;                     // [S, d] = pe[s] + bias
    int S, d, halves;
  };
  static __device__ __forceinline__ void tile_begin(EpiCtx&, const Params&, int, int) {}
  static __device__ __forceinline__ void chunk(EpiCtx& ctx, const Params& p, uint32_t (&raw)[32], int row0, int col0,
                                               int) {
    uint8_t* slab = ctx.smem + (ctx.seq & 1) * 4096;
    if (ctx.lane == 0) bulk_wait_group_read<1>();  // the group that last read this slab (two chunks ago) is done
    __syncwarp();
    const int row = row0 + ctx.lane;
    const int s = (row < ctx.M) ? row % p.S : 0;
    const float* pb = p.pe_bias + static_cast<size_t>(s) * p.d + col0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(pb + 8 * j));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(pb + 8 * j + 4));
      const float o[8] = {__uint_as_float(raw[8 * j + 0]) + a0.x, __uint_as_float(raw[8 * j + 1]) + a0.y,
                          __uint_as_float(raw[8 * j + 2]) + a0.z, __uint_as_float(raw[8 * j + 3]) + a0.w,
                          __uint_as_float(raw[8 * j + 4]) + a1.x, __uint_as_float(raw[8 * j + 5]) + a1.y,
                          __uint_as_float(raw[8 * j + 6]) + a1.z, __uint_as_float(raw[8 * j + 7]) + a1.w};
      uint32_t hi[4], lo[4];
      split_hi_lo8(o, hi, lo);
      *reinterpret_cast<uint4*>(slab + slab64_off(ctx.lane, j)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(slab + 2048 + slab64_off(ctx.lane, j)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (ctx.lane == 0) {
      tma_store_2d(&p.res_c, slab, col0, row0);
      tma_store_2d(&p.res_c, slab + 2048, p.d + col0, row0);
      if (p.halves == 2) {
        tma_store_2d(&p.res_u, slab, col0, row0);
        tma_store_2d(&p.res_u, slab + 2048, p.d + col0, row0);
      }
      bulk_commit_group();
    }
    ctx.seq++;
  }
  static __device__ __forceinline__ void tile_end(EpiCtx&, const Params&, int, int, uint32_t) {}
  static __device__ __forceinline__ void finish(EpiCtx& ctx) {
    if (ctx.lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------------------
// OutputProcess + (inpainting) + sampler arithmetic fused (reference model/mdm.py:372-386,
// diffusion/gaussian_diffusion.py:300-304, 254-257, 525-540, 757-778).  GEMM rows are (b, s) over B*S; row s>=1
// is frame t = s-1; column j < J is a feature.  All tensors are the reference layout [B, J*F, T].
//   mode 0: out = x0                       (model forward only)
//   mode 1: DDPM   x_{t-1} = c1*x0 + c2*x_t + (nz*sigma)*eps
//   mode 2: DDIM   eps_hat = (sr*x_t - x0)/srm1 ; x_{t-1} = x0*sqrt_abp + coef*eps_hat + (nz*sigma)*eps
// Per-step scalars come from a device table indexed by the device-side step state, so the very same launch
// (and CUDA graph) serves every step of the loop.
constexpr int SCHED_STRIDE = 8;  // floats per schedule row: c1 c2 sig_ddpm sr srm1 sqrt_abp coef_eps sig_ddim
struct StepState {
  int done;      // steps completed so far (indexes the noise tape)
  int cur;       // schedule index i of the step in flight
  int start;     // schedule index of the first step (num_timesteps - 1 - skip)
  int pad;
  const float* noise;            // loop mode: base of the noise tape (set per loop, so the step graph is reusable)
  long long noise_step_stride;   // loop mode: elements between consecutive steps of the tape
  unsigned long long seed;       // in-engine Philox noise (philox_normal_kernel): stream seed ...
  long long sample_base;         // ... and the global index of this workspace's sample 0
};

struct EpiOutStep {
  static constexpr int SMEM_PER_WARP = 1024;  // unused
  static constexpr bool RELEASE_EARLY = true;
  template <class P> static __device__ __forceinline__ void preload(const P&, float*, int, int, int) {}
  struct Params {
    const float* bias;        // [J]
    const float* x_t;         // [B, J, T]
    const float* noise;       // explicit eps for one step; nullptr => tape described by *state (loop mode)
    float* x_out;             // [B, J, T]
    float* pred_xstart;       // nullable
    const unsigned char* inpaint_mask;  // nullable, bool [B, J, T]
    const float* inpaint_motion;        // [B, J, T]
    const float* sched;       // [n_steps, SCHED_STRIDE]
    const StepState* state;
    long long noise_batch_stride;  // J*T normally, 0 for const_noise
    int B, S, T, J, mode;
    int s_off;                // rows s < s_off of a sequence are not frames of x (cond token / DiP prefix); t = s - s_off
    int clip_denoised;        // clamp x0 to [-1, 1] after the inpainting blend (gaussian_diffusion.py:348-352)
  };
  static __device__ __forceinline__ void tile_begin(EpiCtx&, const Params&, int, int) {}
  static __device__ __forceinline__ void chunk(EpiCtx& ctx, const Params& p, uint32_t (&raw)[32], int row0, int col0,
                                               int) {
    const int row = row0 + ctx.lane;
    if (row >= ctx.M) return;
    const int b = row / p.S, s = row - b * p.S;
    if (s < p.s_off) return;
    const int t = s - p.s_off;
    float c1 = 0.f, c2 = 0.f, sg = 0.f, sr = 0.f, srm1 = 1.f, sq = 0.f, ce = 0.f;
    const float* nz = nullptr;
    if (p.mode != 0) {
      const StepState st = *p.state;
      const float* row_s = p.sched + static_cast<size_t>(st.cur) * SCHED_STRIDE;
      c1 = row_s[0]; c2 = row_s[1]; sr = row_s[3]; srm1 = row_s[4]; sq = row_s[5]; ce = row_s[6];
      sg = (p.mode == 1) ? row_s[2] : row_s[7];
      nz = (p.noise != nullptr ? p.noise : st.noise + static_cast<long long>(st.done) * st.noise_step_stride) +
           static_cast<long long>(b) * p.noise_batch_stride;
    }
    // x_out may alias x_t (in-place loop): batch every load of the chunk before the first store so that they are
    // all in flight together (consecutive lanes = consecutive frames => each load/store is one coalesced line)
    const size_t base = static_cast<size_t>(b) * p.J * p.T + t;
#pragma unroll
    for (int h = 0; h < 32; h += 16) {
      float xv[16], nv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = col0 + h + j;
        xv[j] = (p.mode != 0 && col < p.J) ? p.x_t[base + static_cast<size_t>(col) * p.T] : 0.f;
        nv[j] = (p.mode != 0 && col < p.J) ? nz[static_cast<size_t>(col) * p.T + t] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = col0 + h + j;
        if (col < p.J) {
          const size_t idx = base + static_cast<size_t>(col) * p.T;
          float x0 = __uint_as_float(raw[h + j]) + __ldg(p.bias + col);
          if (p.inpaint_mask != nullptr && p.inpaint_mask[idx]) x0 = p.inpaint_motion[idx];
          if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
          if (p.pred_xstart != nullptr) p.pred_xstart[idx] = x0;
          float o = x0;
          if (p.mode == 1) {
            const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv[j]));
            o = __fadd_rn(mean, __fmul_rn(sg, nv[j]));
          } else if (p.mode == 2) {
            const float eh = __fdiv_rn(__fsub_rn(__fmul_rn(sr, xv[j]), x0), srm1);
            const float mean = __fadd_rn(__fmul_rn(x0, sq), __fmul_rn(ce, eh));
            o = __fadd_rn(mean, __fmul_rn(sg, nv[j]));
          }
          p.x_out[idx] = o;
        }
      }
    }
  }
  static __device__ __forceinline__ void tile_end(EpiCtx&, const Params&, int, int, uint32_t) {}
  static __device__ __forceinline__ void finish(EpiCtx&) {}
};

}  // namespace b200
