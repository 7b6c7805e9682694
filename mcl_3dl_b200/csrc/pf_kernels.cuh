// pf_kernels.cuh — kernels of the resident particle set (scope row f3): one thread per particle around the
// host-verified per-particle functions of pf_funcs.cuh.  Written after round 1's GPU budget was spent: the functions are
// checked on the host bit for bit against the oracle, these kernels and their plumbing have not run on hardware yet.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcl3dl_b200.h"
#include "pf_funcs.cuh"

namespace mcl3dl
{
static_assert(sizeof(PfState) == sizeof(mcl3dl_state), "mcl3dl_state is PfState");

// MotionPredictionModelDifferentialDrive::predict for every particle (pf::ParticleFilter::predict, pf.h:239-245)
__global__ void pf_predict_kernel(PfState* __restrict__ states, uint32_t n, MotionDev m)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    pf_predict(states[i], m);
}

// What the measurement kernels read: pos + RAW rot (they normalise it themselves, state_6dof.h:217), and the
// odometry-error factor of the node's lambda, NormalLikelihood(sigma)(|odom_err_integ_lin_|) (src/mcl_3dl.cpp:400,422-424,
// include/mcl_3dl/nd.h:45-53: a = float(1 / sqrt(2 pi s^2)), sq2 = float(2 s^2), a * expf(-x * x / sq2)).
__global__ void pf_pack_kernel(const PfState* __restrict__ states, uint32_t n, float nd_a, float nd_sq2, mcl3dl_pose* __restrict__ poses,
                               float* __restrict__ extra /* nullptr: no odometry-error factor */)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const PfState s = states[i];
  mcl3dl_pose p;
  p.px = s.pos[0];
  p.py = s.pos[1];
  p.pz = s.pos[2];
  p._pad = 0.0f;
  p.qx = s.rot[0];
  p.qy = s.rot[1];
  p.qz = s.rot[2];
  p.qw = s.rot[3];
  poses[i] = p;
  if (extra)
  {
    F3 l;
    l.x = s.lin[0];
    l.y = s.lin[1];
    l.z = s.lin[2];
    const float x = __fsqrt_rn(dot3(l, l));  // Vec3::norm, vec3.h:148-151
    extra[i] = fmul(nd_a, expf(fdiv(fmul(-x, x), nd_sq2)));
  }
}

// The sequential float prefix sum of pf.h:189-194 and pstep of :197 — one thread, in order: a parallel scan rounds
// differently and would move picks at the interval boundaries (65 536 dependent adds ~ 0.15 ms).
__global__ void pf_accum_kernel(const float* __restrict__ probs, uint32_t n, float* __restrict__ accum, float* __restrict__ pstep)
{
  if (blockIdx.x != 0 || threadIdx.x != 0)
    return;
  float a = 0.0f;
  for (uint32_t i = 0; i < n; ++i)
  {
    a = fadd(a, probs[i]);
    accum[i] = a;
  }
  *pstep = fdiv(a, static_cast<float>(n));
}

struct Sigma6
{
  float v[6];
};

// pf::ParticleFilter::resample (pf.h:182-225): systematic pick, noise on duplicates, probability 1 / n
__global__ void pf_resample_kernel(const PfState* __restrict__ in, const float* __restrict__ accum, const float* __restrict__ pstep,
                                   uint32_t n, float initial_frac, uint64_t seed, uint32_t call, Sigma6 sigma,
                                   PfState* __restrict__ out, float* __restrict__ out_probs)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float ps = *pstep;
  bool dup;
  const uint32_t src = pf_pick(accum, n, ps, fmul(initial_frac, ps), i, dup);
  PfState s = in[src];
  if (dup)
  {
    float org[6];
    pf_noise6(seed, i, call, sigma.v, org);
    s = pf_add_noise(s, org);
  }
  out[i] = s;
  out_probs[i] = __double2float_rn(ddiv(1.0, static_cast<double>(n)));
}

}  // namespace mcl3dl
