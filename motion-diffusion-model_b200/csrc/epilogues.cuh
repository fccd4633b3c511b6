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
    // two rows per pass: lanes 0-15 -> row rr, lanes 16-31 -> row rr+1; each lane owns 2 adjacent columns
    const int c = (lane & 15) * 2;
    const int col = col0 + c;
    const bool col_ok = col + 1 < N;  // N is even for every caller
    float b0 = 0.f, b1 = 0.f;
    if (col_ok) {
      b0 = p.bias[col];
      b1 = p.bias[col + 1];
    }
#pragma unroll 4
    for (int rr = 0; rr < 32; rr += 2) {
      const int r = rr + (lane >> 4);
      const int row = row0 + r;
      float x0 = stg[r * 33 + c] + b0;
      float x1 = stg[r * 33 + c + 1] + b1;
      if (GELU) {
        x0 = gelu_erf(x0);
        x1 = gelu_erf(x1);
      }
      if (row < M && col_ok)
        *reinterpret_cast<__half2*>(p.out + static_cast<size_t>(row) * p.ld + col) = __floats2half2_rn(x0, x1);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// h32[row, col] += acc + bias[col]      (attention out-projection / FFN down-projection + residual; the
// LayerNorm that follows is a separate row kernel in this revision)
struct EpiResidualF32 {
  struct Params {
    float* h32;
    const float* bias;
    int ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, float (&v)[32], float* stg, int row0, int col0,
                                               int lane, int M, int N) {
    chunk_to_cols(v, stg, lane);
    const int col = col0 + lane;
    if (col >= N) return;
    const float b = p.bias[col];
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const int row = row0 + rr;
      if (row < M) {
        float* dst = p.h32 + static_cast<size_t>(row) * p.ld + col;
        *dst = *dst + (stg[rr * 33 + lane] + b);
      }
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
};

struct EpiOutStep {
  struct Params {
    const float* bias;        // [J]
    const float* x_t;         // [B, J, T]
    const float* noise;       // base of the tape; step k at noise + k*noise_step_stride
    float* x_out;             // [B, J, T]
    float* pred_xstart;       // nullable
    const unsigned char* inpaint_mask;  // nullable, bool [B, J, T]
    const float* inpaint_motion;        // [B, J, T]
    const float* sched;       // [n_steps, SCHED_STRIDE]
    const StepState* state;
    long long noise_step_stride;   // elements between consecutive steps of the tape (0 => one buffer)
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
      nz = p.noise + static_cast<long long>(st.done) * p.noise_step_stride +
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
