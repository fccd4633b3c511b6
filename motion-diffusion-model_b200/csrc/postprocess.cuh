// Post-loop step of the reference's sample/generate.py:161-166 for data_rep = 'hml_vec':
//   inv_transform (x * std + mean, data_loaders/humanml/data/dataset.py:309-310)
//   recover_from_ric (data_loaders/humanml/scripts/motion_process.py:366-385, 437-452): yaw = exclusive prefix sum of the
//   root yaw velocity, root XZ = prefix sum of the root-frame velocity rotated into the world frame, joints rotated by the
//   same yaw and translated by the root XZ
//   permute to [B, n_joints, 3, T]
// HBM-bound: 263 floats in, 66 out per frame.  One CTA per motion; the two scans are sequential fp64 accumulations
// (torch.cumsum on the CPU accumulates fp32 in fp64 and rounds every prefix to fp32 -- reproduced exactly; at T <= 196
// that is < 1 us).  Arithmetic follows the reference's operation order with explicit round-to-nearest mul / add (no FMA
// contraction), so only cosf / sinf can differ from the CPU result (<= 2 ulp).
#pragma once
#include <cuda_runtime.h>

namespace b200 {

struct RicArgs {
  const float* x;            // element (b, feature f, frame t) at x[b * xb + f * xf + t * xt]
  long long xb, xf, xt;
  const float* mean;         // [features] or nullptr (input already de-normalised)
  const float* std;
  float* out;                // element (b, frame t, joint j, axis c) at out[b * ob + t * ot + (3 j + c) * oc]
  long long ob, ot, oc;
  int T, joints;
};

__device__ __forceinline__ float ric_feat(const RicArgs& a, const float* xb_ptr, int f, int t) {
  const float v = xb_ptr[f * a.xf + t * a.xt];
  return a.mean ? __fadd_rn(__fmul_rn(v, a.std[f]), a.mean[f]) : v;
}
// qrot(qinv((c, 0, s, 0)), (vx, *, vz)) -- quaternion.py:16-20, 56-75 with qvec = (0, -s, 0)
__device__ __forceinline__ void ric_rot(float c, float s, float vx, float vz, float* ox, float* oz) {
  const float qy = -s;
  const float uvx = __fmul_rn(qy, vz), uvz = -__fmul_rn(qy, vx);
  const float uuvx = __fmul_rn(qy, uvz), uuvz = -__fmul_rn(qy, uvx);
  *ox = __fadd_rn(vx, __fmul_rn(2.f, __fadd_rn(__fmul_rn(c, uvx), uuvx)));
  *oz = __fadd_rn(vz, __fmul_rn(2.f, __fadd_rn(__fmul_rn(c, uvz), uuvz)));
}

// grid = batch, block = 256, dynamic shared memory = 7 * T floats
__global__ void recover_from_ric_kernel(const RicArgs a) {
  extern __shared__ float ric_smem[];
  const int T = a.T;
  float* rv = ric_smem;          // root yaw velocity, later cos(yaw)
  float* vx = rv + T;            // root-frame velocity x, later world-frame
  float* vz = vx + T;
  float* ang = vz + T;           // yaw, later sin(yaw)
  float* px = ang + T;           // root position x / z
  float* pz = px + T;
  float* cs = pz + T;
  const float* xb = a.x + blockIdx.x * a.xb;
  float* ob = a.out + blockIdx.x * a.ob;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    rv[t] = ric_feat(a, xb, 0, t);
    vx[t] = ric_feat(a, xb, 1, t);
    vz[t] = ric_feat(a, xb, 2, t);
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // yaw[t] = sum_{u < t} rot_vel[u]                         motion_process.py:369-371
    double acc = 0.0;
    ang[0] = 0.f;
    for (int t = 1; t < T; ++t) {
      acc += static_cast<double>(rv[t - 1]);
      ang[t] = static_cast<float>(acc);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float c = cosf(ang[t]), s = sinf(ang[t]);
    float wx = 0.f, wz = 0.f;
    if (t > 0) ric_rot(c, s, vx[t - 1], vz[t - 1], &wx, &wz);   // r_pos[1:, [0, 2]] = data[:-1, 1:3], rotated  :377-380
    cs[t] = c;
    px[t] = wx;   // staged: the in-place overwrite of vx / vz must wait for every thread's read of vx[t - 1]
    pz[t] = wz;
    rv[t] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0 || threadIdx.x == 32) {   // root XZ = inclusive prefix sum of the world-frame velocity  :382
    float* p = threadIdx.x == 0 ? px : pz;
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
      acc += static_cast<double>(p[t]);
      p[t] = static_cast<float>(acc);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {   // root joint: (x, height, z)                              :384, :450
    float* o = ob + t * a.ot;
    o[0] = px[t];
    o[a.oc] = ric_feat(a, xb, 3, t);
    o[2 * a.oc] = pz[t];
  }
  const int nj = a.joints - 1;
  for (int i = threadIdx.x; i < nj * T; i += blockDim.x) {   // the other joints                                  :439-447
    const int j = i / T, t = i - j * T;
    const float x = ric_feat(a, xb, 4 + 3 * j, t), y = ric_feat(a, xb, 5 + 3 * j, t), z = ric_feat(a, xb, 6 + 3 * j, t);
    float wx, wz;
    ric_rot(cs[t], rv[t], x, z, &wx, &wz);
    float* o = ob + t * a.ot + 3 * (j + 1) * a.oc;
    o[0] = __fadd_rn(wx, px[t]);
    o[a.oc] = y;
    o[2 * a.oc] = __fadd_rn(wz, pz[t]);
  }
}

}  // namespace b200
