// Self-attention core for one (sample, head) per CTA:  O = softmax(Q K^T / sqrt(dh) + key_mask) V
// (reference: nn.MultiheadAttention inside nn.TransformerEncoderLayer, built at model/mdm.py:77-84;
//  key_padding_mask from model/mdm.py:241-247 is always a prefix mask => a per-sample valid-key count).
//
// Revision 1 of this kernel keeps the whole K and V of the head resident in shared memory (S <= ~440 tokens;
// the reference's sequences are <= 197) with a 16-byte-chunk XOR swizzle, and runs the two matmuls on the
// legacy warp-level tensor path (mma.sync m16n8k16, fp16 in / fp32 accumulate) with an online softmax over
// 32-key blocks.  The tcgen05/TMEM version (fused with the QKV projection) replaces it in a later revision;
// the interface (qkv16 in, att16 out) is already the one that kernel needs.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace b200 {

constexpr int ATT_DH = 128;
constexpr int ATT_WARPS = 7;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int ATT_KB = 32;  // keys per online-softmax block

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// smem byte offset of 16-B chunk `chunk` (0..15) of row `row` in a [rows][128] fp16 tile
__device__ __forceinline__ uint32_t att_swz(int row, int chunk) { return row * 256 + ((chunk ^ (row & 7)) << 4); }

// qkv : [n_samples * S, 3*d] fp16 (q | k | v, head h at columns h*128 of each block)
// out : [n_samples * S, d]   fp16
// kvlen[sample] = number of valid keys (prefix mask); grid = (heads, n_samples)
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, const int* __restrict__ kvlen, int S,
                     int d, float scale_log2) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int h = blockIdx.x, smp = blockIdx.y;
  const int S_pad = (S + 15) & ~15;
  uint8_t* sK = att_smem;
  uint8_t* sV = att_smem + static_cast<size_t>(S_pad) * 256;
  const uint32_t sK_u = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t sV_u = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ld = static_cast<size_t>(3) * d;
  const __half* base = qkv + static_cast<size_t>(smp) * S * ld + h * ATT_DH;

  // ---- stage K and V of this head (zero rows beyond S)
  for (int idx = tid; idx < S_pad * 16; idx += ATT_THREADS) {
    const int row = idx >> 4, chunk = idx & 15;
    const uint32_t off = att_swz(row, chunk);
    if (row < S) {
      const __half* src = base + static_cast<size_t>(row) * ld + chunk * 8;
      cp_async16(sK_u + off, src + d);
      cp_async16(sV_u + off, src + 2 * d);
    } else {
      *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int kvl = min(kvlen[smp], S);
  const int g = lane >> 2, t = lane & 3;
  const int n_mt = S_pad >> 4;

  for (int mt = warp; mt < n_mt; mt += ATT_WARPS) {
    const int r0 = mt * 16;
    // ---- Q fragments (A operand), straight from global
    uint32_t qa[8][4];
    {
      const int ra = r0 + g, rb = r0 + g + 8;
      const __half* qa_p = base + static_cast<size_t>(ra) * ld + 2 * t;
      const __half* qb_p = base + static_cast<size_t>(rb) * ld + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        qa[ks][0] = ra < S ? *reinterpret_cast<const uint32_t*>(qa_p + ks * 16) : 0u;
        qa[ks][1] = rb < S ? *reinterpret_cast<const uint32_t*>(qb_p + ks * 16) : 0u;
        qa[ks][2] = ra < S ? *reinterpret_cast<const uint32_t*>(qa_p + ks * 16 + 8) : 0u;
        qa[ks][3] = rb < S ? *reinterpret_cast<const uint32_t*>(qb_p + ks * 16 + 8) : 0u;
      }
    }
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

    for (int k0 = 0; k0 < S_pad; k0 += ATT_KB) {
      const int ngrp = min(2, (S_pad - k0) >> 4);  // 16-key groups present in this block
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      // ---- S = Q K^T
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
        for (int gp = 0; gp < 2; ++gp) {
          if (gp < ngrp) {
            const int key = k0 + gp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
            const int chunk = ks * 2 + ((lane >> 3) & 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sK_u + att_swz(key, chunk), b0, b1, b2, b3);
            mma_16816(s[2 * gp], qa[ks], b0, b1);
            mma_16816(s[2 * gp + 1], qa[ks], b2, b3);
          }
        }
      }
      // ---- mask + online softmax (rows g and g+8 of the tile)
      float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key = k0 + nt * 8 + 2 * t + (c & 1);
          float val = s[nt][c] * scale_log2;
          if (key >= kvl) val = -INFINITY;
          s[nt][c] = val;
          if (c < 2) mx_a = fmaxf(mx_a, val); else mx_b = fmaxf(mx_b, val);
        }
      }
      mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
      mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
      mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
      mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
      const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
      const float ms_a = (mn_a == -INFINITY) ? 0.f : mn_a, ms_b = (mn_b == -INFINITY) ? 0.f : mn_b;
      const float al_a = exp2f(m_a - ms_a), al_b = exp2f(m_b - ms_b);
      m_a = mn_a; m_b = mn_b;
      float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - ms_a); s[nt][1] = exp2f(s[nt][1] - ms_a);
        s[nt][2] = exp2f(s[nt][2] - ms_b); s[nt][3] = exp2f(s[nt][3] - ms_b);
        sum_a += s[nt][0] + s[nt][1];
        sum_b += s[nt][2] + s[nt][3];
      }
      l_a = l_a * al_a + sum_a;
      l_b = l_b * al_b + sum_b;
#pragma unroll
      for (int i = 0; i < 16; ++i) { o[i][0] *= al_a; o[i][1] *= al_a; o[i][2] *= al_b; o[i][3] *= al_b; }
      // ---- O += P V
#pragma unroll
      for (int gp = 0; gp < 2; ++gp) {
        if (gp < ngrp) {
          uint32_t pa[4];
          pa[0] = pack_h2(s[2 * gp][0], s[2 * gp][1]);
          pa[1] = pack_h2(s[2 * gp][2], s[2 * gp][3]);
          pa[2] = pack_h2(s[2 * gp + 1][0], s[2 * gp + 1][1]);
          pa[3] = pack_h2(s[2 * gp + 1][2], s[2 * gp + 1][3]);
          const int key = k0 + gp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
          for (int dp = 0; dp < 8; ++dp) {
            const int chunk = 2 * dp + (lane >> 4);
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(sV_u + att_swz(key, chunk), b0, b1, b2, b3);
            mma_16816(o[2 * dp], pa, b0, b1);
            mma_16816(o[2 * dp + 1], pa, b2, b3);
          }
        }
      }
    }
    // ---- normalise and store
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
    const float inv_a = 1.f / l_a, inv_b = 1.f / l_b;
    const int ra = r0 + g, rb = r0 + g + 8;
    __half* oa = out + (static_cast<size_t>(smp) * S + ra) * d + h * ATT_DH + 2 * t;
    __half* ob = out + (static_cast<size_t>(smp) * S + rb) * d + h * ATT_DH + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      if (ra < S) *reinterpret_cast<__half2*>(oa + nt * 8) = __floats2half2_rn(o[nt][0] * inv_a, o[nt][1] * inv_a);
      if (rb < S) *reinterpret_cast<__half2*>(ob + nt * 8) = __floats2half2_rn(o[nt][2] * inv_b, o[nt][3] * inv_b);
    }
  }
}

}  // namespace b200
