// Small HBM/L2-bound kernels around the GEMMs: input packing (transpose + fp16 hi/lo split), per-step
// conditioning token, LayerNorm rows, CFG blend of the hidden rows, weight repacking, table set-up.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "epilogues.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------
// x [B, JF, T] fp32 (reference layout, T contiguous)  ->  xin16 [B*S, ld] fp16 rows (b, s = 1 + t_off + t):
//   columns [0,Kp) = hi, [Kp,2Kp) = lo, [2Kp,3Kp) = hi    (A' of the 3-pass split GEMM  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo)
// Row s = 0 (conditioning token slot) and the pad columns stay zero from allocation time.
__global__ void pack_input_kernel(const float* __restrict__ x, __half* __restrict__ xin, int B, int JF, int T, int S,
                                  int Kp, int ld, int row_off) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int j0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int j = j0 + i, t = t0 + tx;
    tile[i][tx] = (j < JF && t < T) ? x[(static_cast<size_t>(b) * JF + j) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, j = j0 + tx;
    if (t < T && j < JF) {
      const float v = tile[tx][i];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      __half* dst = xin + (static_cast<size_t>(b) * S + row_off + t) * ld + j;
      dst[0] = hi;
      dst[Kp] = lo;
      dst[2 * Kp] = hi;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Conditioning-token rows of the sequence (reference model/mdm.py:195,218-220,251-252):
//   h[b', s=0, :] = (condproj[b', :] + temb_table[t(b'), :]) + pe[0, :]
//   t(b') = tvec[b' % B] when tvec != nullptr (model called with explicit timesteps), else timestep_map[state->cur]
// Runs right after the embedding GEMM (which leaves placeholder values in these rows).
// The residual stream is an fp16 [hi | lo] pair per element (row = 2d halves, hi + lo carries ~22 bits).
__global__ void tok0_rows_kernel(__half* __restrict__ hres, const float* __restrict__ condproj,
                                 const float* __restrict__ temb_table, const float* __restrict__ pe,
                                 const int* __restrict__ tvec, const int* __restrict__ tmap,
                                 const StepState* __restrict__ state, int B, int S, int d, int temb_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int bp = blockIdx.x;
  int t = (tvec != nullptr) ? tvec[bp % B] : tmap[state->cur];
  t = min(max(t, 0), temb_rows - 1);
  const size_t row = static_cast<size_t>(bp) * S;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = (condproj[static_cast<size_t>(bp) * d + c] + temb_table[static_cast<size_t>(t) * d + c]) + pe[c];
    const __half hi = __float2half_rn(v);
    hres[row * 2 * d + c] = hi;
    hres[row * 2 * d + d + c] = __float2half_rn(v - __half2float(hi));
  }
}

// pe_bias[s, c] = pe[s, c] + bias[c]  (per (B,T) workspace table for the embedding epilogue)
__global__ void pe_bias_kernel(float* __restrict__ out, const float* __restrict__ pe, const float* __restrict__ bias,
                               int S, int d) {
  const int s = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) out[static_cast<size_t>(s) * d + c] = bias[c] + pe[static_cast<size_t>(s) * d + c];
}

__global__ void step_advance_kernel(StepState* state) {
  pdl_launch_dependents();
  pdl_wait();
  state->done += 1;
  state->cur -= 1;
}
__global__ void step_set_kernel(StepState* state, int done, int cur, const float* noise, long long noise_step_stride,
                                unsigned long long seed, long long sample_base) {
  state->done = done;
  state->cur = cur;
  state->start = cur;
  state->noise = noise;
  state->noise_step_stride = noise_step_stride;
  state->seed = seed;
  state->sample_base = sample_base;
}

// ---------------------------------------------------------------------------------------------------------
// The engine's own noise stream (B200MDM_FLAG_PHILOX_NOISE / b200mdm_philox_normal): replaces the reference's
// th.randn_like(x) per step (diffusion/gaussian_diffusion.py:525, :770) when the caller asks for a stream that does not
// depend on how the batch is split over GPUs or on how many steps are drawn at once.
//   Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
//   counter = (q, step_id, g_lo, g_hi ^ 0x4d444d42)   q = element index / 4 inside the sample, g = global sample index
//   the 4 output words w0..w3 -> u_k = ((w_k >> 8) + 0.5) * 2^-24 in (0, 1);
//   elements 4q..4q+3 = r0 cos(2 pi u1), r0 sin(2 pi u1), r1 cos(2 pi u3), r1 sin(2 pi u3),  r0 = sqrt(-2 ln u0), r1 = sqrt(-2 ln u2)
// step_id = schedule index of the step that consumes the eps (state->cur when `state` != nullptr); x_T uses 0xffffffff.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__global__ void philox_normal_kernel(float* __restrict__ out, int B, long long n, unsigned long long seed,
                                     long long sample_base, uint32_t step_id, const StepState* __restrict__ state) {
  pdl_launch_dependents();
  pdl_wait();
  if (state != nullptr) {
    seed = state->seed;
    sample_base = state->sample_base;
    step_id = static_cast<uint32_t>(state->cur);
  }
  const long long qn = (n + 3) / 4, total = qn * B;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / qn, q = i - b * qn;
    const unsigned long long g = static_cast<unsigned long long>(sample_base + b);
    uint32_t c[4] = {static_cast<uint32_t>(q), step_id, static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32) ^ 0x4d444d42u};
    philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u0 = (static_cast<float>(c[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-08f;
      const float u1 = (static_cast<float>(c[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-08f;
      const float r = sqrtf(-2.0f * logf(u0));
      float sn, cs;
      sincospif(2.0f * u1, &sn, &cs);
      z[2 * h] = r * cs;
      z[2 * h + 1] = r * sn;
    }
    float* dst = out + b * n + 4 * q;
    if (4 * q + 3 < n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < n) dst[j] = z[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// CFG blend on the hidden rows + fp16 hi/lo split for the 3-pass output GEMM.
//   v = h_u + scale[b] * (h_c - h_u)   (same expression as utils/sampler_util.py:34, applied before the linear
//   OutputProcess: W(h_u + s(h_c-h_u)) + b == out_u + s(out_c - out_u) exactly in real arithmetic)
//   halves == 1: v = h.       g16 row layout: [hi | lo | hi], ld = 3*d.
__global__ void blend_split_kernel(const __half* __restrict__ hres, __half* __restrict__ g16,
                                   const float* __restrict__ scale, int B, int S, int T, int s_off, int d, int halves) {
  pdl_launch_dependents();
  pdl_wait();
  // one warp per FRAME row: the rows s < s_off of a sequence (condition token / DiP prefix) never reach x, so g16
  // holds B*T rows only (12544 = 98 tiles of 128 at B=64, T=196 -- 294 output tiles, two full waves of 148 CTAs)
  const int orow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (orow >= B * T) return;
  const int b = orow / T;
  const int row = b * S + s_off + (orow - b * T);
  const __half* hc = hres + static_cast<size_t>(row) * 2 * d;                         // [hi | lo] rows
  const __half* hu = hres + (static_cast<size_t>(B) * S + row) * 2 * d;
  const float sc = (halves == 2) ? scale[b] : 0.f;
  __half* dst = g16 + static_cast<size_t>(orow) * 3 * d;
  for (int c = lane * 2; c < d; c += 64) {
    const float2 ah = __half22float2(*reinterpret_cast<const __half2*>(hc + c));
    const float2 al = __half22float2(*reinterpret_cast<const __half2*>(hc + d + c));
    float2 a = make_float2(ah.x + al.x, ah.y + al.y);
    if (halves == 2) {
      const float2 uh = __half22float2(*reinterpret_cast<const __half2*>(hu + c));
      const float2 ul = __half22float2(*reinterpret_cast<const __half2*>(hu + d + c));
      const float2 u = make_float2(uh.x + ul.x, uh.y + ul.y);
      a.x = __fadd_rn(u.x, __fmul_rn(sc, __fsub_rn(a.x, u.x)));
      a.y = __fadd_rn(u.y, __fmul_rn(sc, __fsub_rn(a.y, u.y)));
    }
    const __half2 hi = __floats2half2_rn(a.x, a.y);
    const float2 hif = __half22float2(hi);
    const __half2 lo = __floats2half2_rn(a.x - hif.x, a.y - hif.y);
    *reinterpret_cast<__half2*>(dst + c) = hi;
    *reinterpret_cast<__half2*>(dst + d + c) = lo;
    *reinterpret_cast<__half2*>(dst + 2 * d + c) = hi;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Weight repacking (one-time, at load).
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}
// W [N, K] fp32 -> [W16 | W16] fp16 [N, 2K]: partner of activations stored as [hi | lo] along K (the trans_dec engine keeps
// its fp16 activations to ~22 mantissa bits this way; the product A_hi W + A_lo W accumulates in fp32 on the tensor core).
__global__ void f32_to_f16_dup_kernel(const float* __restrict__ src, __half* __restrict__ dst, int N, int K) {
  const size_t n = static_cast<size_t>(N) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / K, k = i % K;
    const __half h = __float2half_rn(src[i]);
    dst[r * 2 * K + k] = h;
    dst[r * 2 * K + K + k] = h;
  }
}
// W [N, K] fp32 -> W' [Npad, 3*Kp] fp16 = [hi | hi | lo] (zero padding), partner of the [hi | lo | hi] activations.
__global__ void split_weight_kernel(const float* __restrict__ w, __half* __restrict__ out, int N, int K, int Kp) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = w[static_cast<size_t>(n) * K + k];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    __half* dst = out + static_cast<size_t>(n) * 3 * Kp + k;
    dst[0] = hi;
    dst[Kp] = hi;
    dst[2 * Kp] = lo;
  }
}

// y[r, c] = act( sum_k x[r, k] * w[c, k] + b[c] ), fp32, one warp per output element (tiny set-up GEMVs:
// timestep-embedding MLP for every model timestep, text projection once per loop)
template <int ACT>  // 0 none, 1 SiLU
__global__ void small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ b, float* __restrict__ y, int R, int C, int K,
                                    int x_ld) {
  const size_t widx = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (widx >= static_cast<size_t>(R) * C) return;
  const int r = static_cast<int>(widx / C), c = static_cast<int>(widx % C);
  const float* xr = x + static_cast<size_t>(r) * x_ld;
  const float* wr = w + static_cast<size_t>(c) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(xr[k], wr[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    acc += (b != nullptr) ? b[c] : 0.f;
    if (ACT == 1) acc = acc / (1.f + expf(-acc));
    y[static_cast<size_t>(r) * C + c] = acc;
  }
}

// condproj rows for the packed batch: first B rows conditional, next B rows unconditional.
//   text  : cond = (W clip + b) already in proj[B, d];  uncond = bias          (mask_cond zeros => bias only)
//   action: cond = action_embedding[a[b]];              uncond = 0             (model/mdm.py:225-227)
//   none  : 0
__global__ void condproj_fill_kernel(float* __restrict__ condproj, const float* __restrict__ proj,
                                     const float* __restrict__ bias, const float* __restrict__ action_emb,
                                     const int* __restrict__ action, int B, int d, int rows, int first_uncond,
                                     int cond_mode) {
  const int bp = blockIdx.x;
  if (bp >= rows) return;
  const bool unc = first_uncond ? true : (bp >= B);
  const int b = bp % B;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = 0.f;
    if (cond_mode == 1) v = unc ? bias[c] : proj[static_cast<size_t>(b) * d + c];
    else if (cond_mode == 2) v = unc ? 0.f : action_emb[static_cast<size_t>(action[b]) * d + c];
    condproj[static_cast<size_t>(bp) * d + c] = v;
  }
}

// x_t = sqrt_ac * x0 + sqrt_1mac * noise   (q_sample, diffusion/gaussian_diffusion.py:226-244)
__global__ void q_sample_kernel(float* __restrict__ out, const float* __restrict__ x0, const float* __restrict__ noise,
                                float a, float b, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float s = (x0 != nullptr) ? x0[i] : 0.f;
    out[i] = __fadd_rn(__fmul_rn(a, s), __fmul_rn(b, noise[i]));
  }
}

}  // namespace b200

// ===================================================================================================================
// trans_dec (DiP) helpers -- reference model/mdm.py:255-270 and torch nn.TransformerDecoderLayer cross-attention
namespace b200 {

// mem16[b', m, :] = [hi | lo] fp16 of ( memproj[b', m, :] + temb_table[t(b'), :] )   (emb = text_emb + time_emb,
// mdm.py:218-220; the time embedding is broadcast over the text tokens).  Rows are 2d wide.  grid = (Mt, Bp)
__global__ void mem_build_kernel(__half* __restrict__ mem16, const float* __restrict__ memproj,
                                 const float* __restrict__ temb_table, const int* __restrict__ tvec,
                                 const int* __restrict__ tmap, const StepState* __restrict__ state, int B, int Mt, int d,
                                 int temb_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, bp = blockIdx.y;
  int t = (tvec != nullptr) ? tvec[bp % B] : tmap[state->cur];
  t = min(max(t, 0), temb_rows - 1);
  const size_t row = static_cast<size_t>(bp) * Mt + m;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = memproj[row * d + c] + temb_table[static_cast<size_t>(t) * d + c];
    const __half hi = __float2half_rn(v);
    mem16[row * 2 * d + c] = hi;
    mem16[row * 2 * d + d + c] = __float2half_rn(v - __half2float(hi));
  }
}

// memproj rows of the packed batch: cond half = W enc + b (already in proj [B*Mt, d], row (b, m)), uncond half = b.
__global__ void memproj_fill_kernel(float* __restrict__ memproj, const float* __restrict__ proj,
                                    const float* __restrict__ bias, int B, int Mt, int d, int rows_bp, int first_uncond) {
  const int m = blockIdx.x, bp = blockIdx.y;
  if (bp >= rows_bp) return;
  const bool unc = first_uncond ? true : (bp >= B);
  const int b = bp % B;
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    memproj[(static_cast<size_t>(bp) * Mt + m) * d + c] = unc ? bias[c] : proj[(static_cast<size_t>(b) * Mt + m) * d + c];
}

// enc_text [Mt, B, C] (reference layout, model/mdm.py:185) -> [B*Mt, C] rows (b, m) so that one small GEMM projects it
__global__ void permute_mbc_kernel(const float* __restrict__ src, float* __restrict__ dst, int Mt, int B, int C) {
  const int m = blockIdx.x, b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    dst[(static_cast<size_t>(b) * Mt + m) * C + c] = src[(static_cast<size_t>(m) * B + b) * C + c];
}

// Cross-attention core: softmax(q k^T / sqrt(128) + mask) v with a handful of memory tokens (Mt <= 64).
//   q16 [n_samples*S, d] (head h at columns h*128), kv16 [n_samples*Mt, 2d] (k | v), mask [n_samples, Mt] (1 = ignore),
//   out16 [n_samples*S, 2d] = [hi | lo].  One warp per query row, lane owns 4 of the 128 head dimensions.  ~0.5 GFLOP per layer at the
//   DiP configuration (60 x 16 tokens): CUDA cores are enough, the projections around it run on the tensor cores.
// grid = (heads, n_samples), block = 128
__global__ void cross_attention_kernel(const __half* __restrict__ q16, const __half* __restrict__ kv16,
                                       const unsigned char* __restrict__ mask, __half* __restrict__ out16, int S, int Mt,
                                       int d, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __half ca_smem[];   // [2][Mt][128]
  const int h = blockIdx.x, smp = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  __half* sK = ca_smem;
  __half* sV = ca_smem + Mt * 128;
  for (int i = threadIdx.x; i < Mt * 16; i += blockDim.x) {   // 16-byte chunks
    const int m = i >> 4, ch = i & 15;
    const __half* src = kv16 + (static_cast<size_t>(smp) * Mt + m) * 2 * d + h * 128 + ch * 8;
    *reinterpret_cast<uint4*>(sK + m * 128 + ch * 8) = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(sV + m * 128 + ch * 8) = *reinterpret_cast<const uint4*>(src + d);
  }
  __syncthreads();
  const unsigned char* mk = mask + static_cast<size_t>(smp) * Mt;
  for (int s = warp; s < S; s += nwarp) {
    const size_t row = static_cast<size_t>(smp) * S + s;
    const __half2* qp = reinterpret_cast<const __half2*>(q16 + row * d + h * 128 + lane * 4);
    const float2 qa = __half22float2(qp[0]), qb = __half22float2(qp[1]);
    float sc[64];
    float mx = -INFINITY;
#pragma unroll 4
    for (int m = 0; m < Mt; ++m) {
      const __half2* kp = reinterpret_cast<const __half2*>(sK + m * 128 + lane * 4);
      const float2 ka = __half22float2(kp[0]), kb = __half22float2(kp[1]);
      float dot = qa.x * ka.x + qa.y * ka.y + qb.x * kb.x + qb.y * kb.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      dot = mk[m] ? -INFINITY : dot * scale;
      sc[m] = dot;
      mx = fmaxf(mx, dot);
    }
    float sum = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    for (int m = 0; m < Mt; ++m) {
      const float p = (mx == -INFINITY) ? 0.f : __expf(sc[m] - mx);
      sum += p;
      const __half2* vp = reinterpret_cast<const __half2*>(sV + m * 128 + lane * 4);
      const float2 va = __half22float2(vp[0]), vb = __half22float2(vp[1]);
      o0 = fmaf(p, va.x, o0); o1 = fmaf(p, va.y, o1); o2 = fmaf(p, vb.x, o2); o3 = fmaf(p, vb.y, o3);
    }
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    o0 *= inv; o1 *= inv; o2 *= inv; o3 *= inv;
    __half2* op = reinterpret_cast<__half2*>(out16 + row * 2 * d + h * 128 + lane * 4);
    const __half2 h01 = __floats2half2_rn(o0, o1), h23 = __floats2half2_rn(o2, o3);
    op[0] = h01;
    op[1] = h23;
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    __half2* lp = op + d / 2;
    lp[0] = __floats2half2_rn(o0 - f01.x, o1 - f01.y);
    lp[1] = __floats2half2_rn(o2 - f23.x, o3 - f23.y);
  }
}

}  // namespace b200
