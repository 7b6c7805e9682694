// scan_kernels.cuh — scan preprocessing on the device (scope row f4, scan half): what the node does to the accumulated
// raw cloud before every measurement (src/mcl_3dl.cpp:363-383):
//   1. pcl::VoxelGrid downsample (leaf = params.downsample_x/y/z, :363-367).  PCL is a third-party dependency that is
//      absent from the reference tree (PCL >= 1.8, version unpinned); restated from its published algorithm
//      (pcl/filters/impl/voxel_grid.hpp: applyFilter): leaf index = floor(coord * float(1 / leaf)) - floor(min * ...),
//      linear index x fastest, output in ascending voxel index, one centroid per occupied voxel (float sums; the label is
//      the majority label of the voxel, lowest on ties, as PCL's CentroidPoint / AccumulatorLabel does).  PCL orders the
//      points of one voxel with an unstable std::sort, so its float centroid sums are defined only up to summation order:
//      the device sums in input order.  "Parity unpinned" for this piece (no reference test pins VoxelGrid output).
//   2. the models' filter(): clip by planar range^2 and z window, order kept
//      (src/lidar_measurement_model_likelihood.cpp:79-103 == _beam.cpp:98-122) — exact, checked against the oracle;
//   3. PointCloudUniformSampler::sample (point_cloud_random_samplers/point_cloud_uniform_sampler.h:56-74): `num` draws with
//      replacement, uniform over the clipped points.  The reference seeds std::default_random_engine from
//      std::random_device, i.e. its draws are not reproducible by construction; the device draws from Philox-4x32-10
//      keyed by (seed, stream, draw index).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcl3dl_b200.h"
#include "pf_funcs.cuh"

namespace mcl3dl
{
struct VoxelGridDev
{
  float inv_leaf[3];  // float(1 / leaf), Eigen::Array4f inverse_leaf_size_
  int min_b[3];       // floor(min * inv_leaf)
  int div_b[3];
};

__device__ __forceinline__ uint32_t voxel_key(const VoxelGridDev& g, float x, float y, float z)
{
  const int i = __float2int_rd(fmul(x, g.inv_leaf[0])) - g.min_b[0];
  const int j = __float2int_rd(fmul(y, g.inv_leaf[1])) - g.min_b[1];
  const int k = __float2int_rd(fmul(z, g.inv_leaf[2])) - g.min_b[2];
  return static_cast<uint32_t>(i) + static_cast<uint32_t>(j) * g.div_b[0] + static_cast<uint32_t>(k) * g.div_b[0] * g.div_b[1];
}

__global__ void scan_key_kernel(const mcl3dl_point* __restrict__ pts, uint32_t n, VoxelGridDev g, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const mcl3dl_point p = pts[i];
  keys[i] = voxel_key(g, p.x, p.y, p.z);
  vals[i] = i;
}

// heads of the runs of equal keys (sorted): flag[i] = 1 where a voxel starts
__global__ void scan_heads_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ flags)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// one thread per voxel head: centroid of the run (input order: the sort is stable), majority label
__global__ void scan_centroid_kernel(const mcl3dl_point* __restrict__ pts, const uint32_t* __restrict__ keys,
                                     const uint32_t* __restrict__ order, const uint32_t* __restrict__ flags,
                                     const uint32_t* __restrict__ pos /* exclusive scan of flags */, uint32_t n,
                                     mcl3dl_point* __restrict__ out, uint32_t* __restrict__ n_out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  if (i == n - 1)
    *n_out = pos[i] + flags[i];
  if (!flags[i])
    return;
  const uint32_t key = keys[i];
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  uint32_t cnt = 0;
  for (uint32_t j = i; j < n && keys[j] == key; ++j, ++cnt)
  {
    const mcl3dl_point p = pts[order[j]];
    sx = fadd(sx, p.x);
    sy = fadd(sy, p.y);
    sz = fadd(sz, p.z);
  }
  // majority label, lowest label on ties (AccumulatorLabel: a std::map walked in key order, first maximum)
  uint32_t best_label = 0xffffffffu, best_cnt = 0;
  for (uint32_t j = i; j < i + cnt; ++j)
  {
    const uint32_t l = pts[order[j]].label;
    uint32_t c = 0;
    for (uint32_t k = i; k < i + cnt; ++k) c += pts[order[k]].label == l;
    if (c > best_cnt || (c == best_cnt && l < best_label))
    {
      best_cnt = c;
      best_label = l;
    }
  }
  mcl3dl_point o;
  const float fc = static_cast<float>(cnt);
  o.x = fdiv(sx, fc);
  o.y = fdiv(sy, fc);
  o.z = fdiv(sz, fc);
  o.label = best_label;
  out[pos[i]] = o;
}

struct ClipDev
{
  float near_sq, far_sq, z_min, z_max;  // clip_near_sq_ = clip_near * clip_near (float), likelihood.cpp:59-60
};

// local_points_filter of filter(), likelihood.cpp:83-93 / beam.cpp:102-112: true = REMOVE
__device__ __forceinline__ bool clip_removes(const ClipDev& c, const mcl3dl_point& p)
{
  const float r2 = fadd(fmul(p.x, p.x), fmul(p.y, p.y));
  if (r2 > c.far_sq)
    return true;
  if (r2 < c.near_sq)
    return true;
  if (p.z < c.z_min || c.z_max < p.z)
    return true;
  return false;
}

__global__ void scan_clip_flags_kernel(const mcl3dl_point* __restrict__ pts, const uint32_t* __restrict__ n_ptr, ClipDev a, ClipDev b,
                                       uint32_t* __restrict__ flag_a, uint32_t* __restrict__ flag_b, uint32_t cap)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap)
    return;
  const bool live = i < *n_ptr;
  const mcl3dl_point p = live ? pts[i] : mcl3dl_point{0, 0, 0, 0};
  flag_a[i] = (live && !clip_removes(a, p)) ? 1u : 0u;
  flag_b[i] = (live && !clip_removes(b, p)) ? 1u : 0u;
}

__global__ void scan_compact_kernel(const mcl3dl_point* __restrict__ pts, const uint32_t* __restrict__ flags,
                                    const uint32_t* __restrict__ pos, uint32_t cap, mcl3dl_point* __restrict__ out,
                                    uint32_t* __restrict__ n_out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap)
    return;
  if (i == cap - 1)
    *n_out = pos[i] + flags[i];
  if (flags[i])
    out[pos[i]] = pts[i];
}

// PointCloudUniformSampler::sample: draw t takes clipped point floor(u * n), u from Philox (seed, stream, t)
__global__ void scan_sample_kernel(const mcl3dl_point* __restrict__ clipped, const uint32_t* __restrict__ n_ptr, uint32_t num,
                                   uint64_t seed, uint32_t stream, mcl3dl_point* __restrict__ out, uint32_t* __restrict__ n_out)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = *n_ptr;
  if (t == 0)
    *n_out = n ? num : 0u;  // an empty clipped cloud gives an empty sample (:62-63)
  if (t >= num || n == 0)
    return;
  uint32_t c[4] = {t, stream, 0x5ca11ab1u, 0u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint32_t idx = static_cast<uint32_t>((static_cast<uint64_t>(c[0]) * n) >> 32);
  out[t] = clipped[idx];
}

}  // namespace mcl3dl
