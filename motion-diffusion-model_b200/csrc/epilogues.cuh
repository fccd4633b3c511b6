// Fused GEMM epilogues.  Each functor receives one 32x32 accumulator chunk: thread `lane` of the warp holds
// accumulator row (row0 + lane), columns [col0, col0 + 32) in v[].  Row-major outputs go through a warp-private
// 32x33 shared-memory transpose so that global accesses are row-contiguous (one 128-B line per instruction);
// feature-major outputs ([B, J, T], T contiguous) are written straight from registers because consecutive
// accumulator rows are consecutive frames.
#pragma once
#include <cuda_fp16.h>

namespace b200 {

__device__ __forceinline__ void chunk_to_cols(const float (&v)[32], float* stg, int lane) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
  __syncwarp();
}

// exact-erf GELU (torch F.gelu default, reference model/mdm.py:80 activation="gelu")
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------------------------
// out16[row, col] = fp16( act(acc + bias[col]) )          (QKV projection, FFN up-projection)
// Column phase: lane owns 8 adjacent columns (16 B of fp16) of one row; 4 lanes cover the 32 columns of a row,
// so one warp-wide store writes 8 complete 64-B row segments and 4 passes cover the chunk.
template <bool GELU>
struct EpiBiasF16 {
  struct Params {
    __half* out;
    const float* bias;
    int ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, float (&v)[32], float* stg, int row0, int col0,
                                               int lane, int M, int N) {
    chunk_to_cols(v, stg, lane);
    const int c = (lane & 3) * 8;
    const int col = col0 + c;
    const bool col_ok = col + 7 < N;
    float b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (col + j < N) ? p.bias[col + j] : 0.f;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int r = pass * 8 + (lane >> 2);
      const int row = row0 + r;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = stg[r * 33 + c + j] + b[j];
        if (GELU) x[j] = gelu_erf(x[j]);
      }
      if (row < M) {
        __half* dst = p.out + static_cast<size_t>(row) * p.ld + col;
        if (col_ok) {
          uint4 pk;
          __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
          __half2 h2 = __floats2half2_rn(x[4], x[5]), h3 = __floats2half2_rn(x[6], x[7]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst) = pk;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (col + j < N) dst[j] = __float2half_rn(x[j]);
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// h32[row, col] += acc + bias[col]      (attention out-projection / FFN down-projection + residual; the
// LayerNorm that follows is a separate row kernel in this revision).  Column phase: lane owns 4 adjacent columns
// (one float4) of one row, 8 lanes cover a 128-B row segment, 8 passes cover the chunk; all 8 residual loads are
// issued before the first use so that they are in flight together.
struct EpiResidualF32 {
  struct Params {
    float* h32;
    const float* bias;
    int ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, float (&v)[32], float* stg, int row0, int col0,
                                               int lane, int M, int N) {
    const int c = (lane & 7) * 4;
    const int col = col0 + c;       // N % 4 == 0 for every caller (N = 512)
    const int rsub = lane >> 3;
    float4 res[8];
    const bool col_ok = col + 3 < N;
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int row = row0 + pass * 4 + rsub;
      res[pass] = (row < M && col_ok) ? *reinterpret_cast<const float4*>(p.h32 + static_cast<size_t>(row) * p.ld + col)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    chunk_to_cols(v, stg, lane);
    const float4 b = col_ok ? *reinterpret_cast<const float4*>(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int r = pass * 4 + rsub;
      const int row = row0 + r;
      float4 o;
      o.x = res[pass].x + (stg[r * 33 + c + 0] + b.x);
      o.y = res[pass].y + (stg[r * 33 + c + 1] + b.y);
      o.z = res[pass].z + (stg[r * 33 + c + 2] + b.z);
      o.w = res[pass].w + (stg[r * 33 + c + 3] + b.w);
      if (row < M && col_ok) *reinterpret_cast<float4*>(p.h32 + static_cast<size_t>(row) * p.ld + col) = o;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// InputProcess + cond-token concat + positional encoding (reference model/mdm.py:238,251-252,343-349):
//   GEMM rows are (b, s) over B*S; s == 0 is the conditioning token (taken from tok0, not from the GEMM),
//   s >= 1 is frame s-1:  h = acc + bias + pe[s].  The frame rows are identical for the cond / uncond halves of
//   the packed CFG batch, so each row is written `halves` times.
struct EpiEmbed {
  struct Params {
    float* h32;
    __half* h16;
    const float* bias;   // [d]
    const float* pe;     // [max_len, d]
    const float* tok0;   // [halves*B, d]  = cond projection + timestep embedding (per step)
    int B, S, d, halves;
  };
  static __device__ __forceinline__ void apply(const Params& p, float (&v)[32], float* stg, int row0, int col0,
                                               int lane, int M, int N) {
    chunk_to_cols(v, stg, lane);
    const int col = col0 + lane;
    if (col >= N) return;
    const float b = p.bias[col];
#pragma unroll 2
    for (int rr = 0; rr < 32; ++rr) {
      const int row = row0 + rr;
      if (row >= M) break;
      const int bi = row / p.S, s = row - bi * p.S;
      const float pe = p.pe[static_cast<size_t>(s) * p.d + col];
      const float frame = stg[rr * 33 + lane] + b + pe;
      for (int hf = 0; hf < p.halves; ++hf) {
        const size_t orow = static_cast<size_t>(hf * p.B + bi) * p.S + s;
        float val = frame;
        if (s == 0) val = p.tok0[static_cast<size_t>(hf * p.B + bi) * p.d + col] + pe;
        p.h32[orow * p.d + col] = val;
        p.h16[orow * p.d + col] = __float2half_rn(val);
      }
    }
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
};

struct EpiOutStep {
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
    int clip_denoised;        // clamp x0 to [-1, 1] after the inpainting blend (gaussian_diffusion.py:348-352)
  };
  static __device__ __forceinline__ void apply(const Params& p, float (&v)[32], float* stg, int row0, int col0,
                                               int lane, int M, int N) {
    const int row = row0 + lane;
    if (row >= M) return;
    const int b = row / p.S, s = row - b * p.S;
    if (s == 0) return;
    const int t = s - 1;
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
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = col0 + j;
      if (col < p.J) {
        const size_t idx = (static_cast<size_t>(b) * p.J + col) * p.T + t;
        float x0 = v[j] + p.bias[col];
        if (p.inpaint_mask != nullptr && p.inpaint_mask[idx]) x0 = p.inpaint_motion[idx];
        if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        if (p.pred_xstart != nullptr) p.pred_xstart[idx] = x0;
        float o = x0;
        if (p.mode == 1) {
          const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, p.x_t[idx]));
          o = __fadd_rn(mean, __fmul_rn(sg, nz[static_cast<size_t>(col) * p.T + t]));
        } else if (p.mode == 2) {
          const float eh = __fdiv_rn(__fsub_rn(__fmul_rn(sr, p.x_t[idx]), x0), srm1);
          const float mean = __fadd_rn(__fmul_rn(x0, sq), __fmul_rn(ce, eh));
          o = __fadd_rn(mean, __fmul_rn(sg, nz[static_cast<size_t>(col) * p.T + t]));
        }
        p.x_out[idx] = o;
      }
    }
  }
};

}  // namespace b200
