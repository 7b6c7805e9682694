// device_funcs.cuh — the per-thread device functions of the measurement update (no warp or CTA cooperation):
// grid descriptors, the exact 1-NN searches, the DDA and KD-tree ray casts.  Kept free of anything but scalar
// intrinsics so that tests/hostsim can compile this very file for the host (with tests/hostsim/cuda_shim.h) and
// check every function against the oracle without a GPU.  The kernels that call them live in kernels.cuh.
#pragma once
#ifndef MCL3DL_HOSTSIM
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>

#include "device_math.cuh"

// MCL3DL_NEAR_BITS=1 (default) compiles the near-field screens below into the searches; -DMCL3DL_NEAR_BITS=0 builds
// the kernels without them (A/B: profiles/r01y_ab_variants.txt — results byte-identical, likelihood kernel -6 %,
// KD-tree raycaster -16 %).
#ifndef MCL3DL_NEAR_BITS
#define MCL3DL_NEAR_BITS 1
#endif
// MCL3DL_KD_SKIP=1 adds a second, coarser field to the KD-tree raycaster that proves several marching steps clear at
// once (cast_ray_kd).  Host-verified (tests/hostsim), not yet measured on a GPU: off by default.
#ifndef MCL3DL_KD_SKIP
#define MCL3DL_KD_SKIP 0
#endif
// MCL3DL_LIK_CHUNKS=1 also builds the warp-chunk likelihood kernel (lik_kernel_wc), selected at run time with
// MCL3DL_LIK=chunk.  Host-verified lane by lane (tests/hostsim), not yet measured on a GPU: off by default.
#ifndef MCL3DL_LIK_CHUNKS
#define MCL3DL_LIK_CHUNKS 0
#endif
#if MCL3DL_KD_SKIP && !MCL3DL_NEAR_BITS
#error "MCL3DL_KD_SKIP needs MCL3DL_NEAR_BITS"
#endif

namespace mcl3dl
{
// ---- near field: one bit per FINE cubic cell (edge ef = 1.01 * r / k) of the rescaled space, set when some map
// point's fine cell lies within k cells (Chebyshev) of it.  A query whose bit is clear has no map point within r:
// |p - q| < r  =>  per axis |p_a - q_a| < r = 0.99 k ef  =>  the two cell indices differ by at most k (the 1 % margin
// covers the float rounding of the shared cell function below for up to ~65 k cells per axis).  The field is a
// conservative screen in front of the exact searches: a set bit decides nothing, a clear bit skips the search.
struct NearBitsDev
{
  const uint32_t* bits;  // nullptr = no field: every query is "maybe near"
  int nx, ny, nz;        // fine cells per axis
  int pitch;             // 32-bit words per x-row
  float ox, oy, oz;      // origin in the rescaled space: k + 0.5 cells below the map's bounding box
  float inv_cell;
};

__device__ __forceinline__ int near_cell(float v, float o, float inv_cell)
{
  return __float2int_rd(fmul(fsub(v, o), inv_cell));
}

// false only if no map point can be within the radius the field was built for (queries outside the field's extent
// are at least k cells away from every point's cell; a NaN coordinate lands in cell 0 and at worst runs the search)
__device__ __forceinline__ bool near_maybe(const NearBitsDev& f, float qx, float qy, float qz)
{
  if (!f.bits)
    return true;
  const int cx = near_cell(qx, f.ox, f.inv_cell), cy = near_cell(qy, f.oy, f.inv_cell), cz = near_cell(qz, f.oz, f.inv_cell);
  if (static_cast<unsigned>(cx) >= static_cast<unsigned>(f.nx) || static_cast<unsigned>(cy) >= static_cast<unsigned>(f.ny) ||
      static_cast<unsigned>(cz) >= static_cast<unsigned>(f.nz))
    return false;
  const uint32_t w = __ldg(f.bits + (static_cast<size_t>(cz) * f.ny + cy) * f.pitch + (cx >> 5));
  return (w >> (cx & 31)) & 1u;
}

// Build side, per map point (rescaled coordinates): set the bits of every fine cell within k cells of the point's.
// x is the bit index inside a row, so one (dy, dz) row is a run of <= 2k+1 <= 31 bits = at most two word updates.
__device__ __forceinline__ void near_mark_point(const NearBitsDev& f, uint32_t* bits, int k, float sx, float sy, float sz)
{
  const int cx = near_cell(sx, f.ox, f.inv_cell), cy = near_cell(sy, f.oy, f.inv_cell), cz = near_cell(sz, f.oz, f.inv_cell);
  const int x0 = max(cx - k, 0), x1 = min(cx + k, f.nx - 1);
  if (x0 > x1)
    return;
  const int w0 = x0 >> 5, w1 = x1 >> 5;
  const uint32_t lo = 0xffffffffu << (x0 & 31), hi = 0xffffffffu >> (31 - (x1 & 31));
  for (int iz = max(cz - k, 0); iz <= min(cz + k, f.nz - 1); ++iz)
    for (int iy = max(cy - k, 0); iy <= min(cy + k, f.ny - 1); ++iy)
    {
      uint32_t* row = bits + (static_cast<size_t>(iz) * f.ny + iy) * f.pitch;
      if (w0 == w1)
        atomicOr(row + w0, lo & hi);
      else
      {
        atomicOr(row + w0, lo);
        atomicOr(row + w1, hi);
      }
    }
}

// Host side: extent of a field for search radius r (> 0) and dilation k (1..15) over the rescaled bounding box.
// The cell edge is doubled until the field fits max_bytes and 65 536 cells per axis (coarser is still conservative).
inline bool near_layout(NearBitsDev& f, float r, int k, const float sc_min[3], const float sc_max[3], size_t max_bytes)
{
  if (!(r > 0.0f) || k < 1 || k > 15)
    return false;
  float ef = 1.01f * r / static_cast<float>(k);
  for (int tries = 0; tries < 40; ++tries, ef *= 2.0f)
  {
    double dims[3];
    float org[3];
    bool ok = true;
    for (int a = 0; a < 3; ++a)
    {
      org[a] = sc_min[a] - (static_cast<float>(k) + 0.5f) * ef;
      dims[a] = floor((static_cast<double>(sc_max[a]) - org[a]) / ef) + k + 2;
      ok = ok && dims[a] >= 1.0 && dims[a] <= 65536.0;
    }
    if (!ok)
      continue;
    const double pitch = floor((dims[0] + 31.0) / 32.0);
    if (pitch * dims[1] * dims[2] * 4.0 > static_cast<double>(max_bytes))
      continue;
    f.bits = nullptr;
    f.nx = static_cast<int>(dims[0]);
    f.ny = static_cast<int>(dims[1]);
    f.nz = static_cast<int>(dims[2]);
    f.pitch = static_cast<int>(pitch);
    f.ox = org[0];
    f.oy = org[1];
    f.oz = org[2];
    f.inv_cell = 1.0f / ef;
    return true;
  }
  return false;
}

// ---- likelihood search grid: cubic cells over the RESCALED map points (p * dist_weight), CSR of
// cell -> contiguous run in `pts` (x fastest, so an x-row of cells is one contiguous run).
struct NnGridDev
{
  const uint32_t* cell_start;  // [nx*ny*nz + 1]
  const float4* pts;           // rescaled xyz, w = original map index (bits)
  // Window table for the likelihood kernel: entry (x, y, z) at ((z*nx + x)*nyp + y) = {start of cell (x,y,z),
  // packed point counts of the 1 / 2 / 3 cells starting there along x (10 | 11 | 11 bits, all-ones = overflow)}.
  // y is the fastest index, so the three y-rows of a query window are adjacent and one eval fetches its whole
  // 3x3 window with 6 aligned 16-byte loads instead of 18 scattered 4-byte ones (L1 wavefronts).
  const uint2* row3;
  int nyp;  // padded (even) y pitch of row3
  int nx, ny, nz;
  float ox, oy, oz;  // grid origin in the rescaled space
  float inv_cell;
  float wx, wy, wz;  // dist_weight
#if MCL3DL_NEAR_BITS
  NearBitsDev near;  // built for the likelihood radius (LikDev::rpad)
#endif
};

struct LikDev
{
  float match_dist_min;   // R
  float match_dist_flat;  // F
  float match_weight;     // W
  float r2;               // float(double(R)*double(R)): what pcl::KdTreeFLANN::radiusSearch hands to FLANN
  float rpad;             // window half-width, R plus a rounding guard
};

// ---- DDA grid: exactly RaycastUsingDDA's lattice (min_p_, dda_grid_size_, map_size_), occupancy as
// one bit per cell + CSR of per-cell points in map order.
struct DdaGridDev
{
  const uint32_t* occ;         // bit c of word c>>5
  const uint32_t* cell_start;  // [cells + 1]
  const float4* pts;           // raw xyz, w = label bits; sorted by cell, map order inside a cell
  int nx, ny, nz;
  float min_x, min_y, min_z;
  float max_x, max_y, max_z;
  double grid;             // dda_grid_size_
  double ray_angle_half;   // ray_angle_half_
  double min_dist_thr_sq;  // min_dist_thr_sq_
  float hit_tolerance;     // float(hit_tolerance_): Vec3 * double narrows to float (vec3.h:119)
  float hit_range_sq;
  float sin_total_ref;
  float beam_likelihood;
  float beam_likelihood_min;
  uint32_t filter_label_max;
  int short_only;
};

// ---- KD-tree raycaster (RaycastUsingKDTree, the node's default caster): marches over the likelihood
// search grid; the colliding map point's raw coordinates / label are fetched by original index.
struct KdRayDev
{
  const float4* raw_pts;  // map points as handed to set_map: xyz + label bits, original order
  float grid_min;         // map_grid_min_ (raycast_using_kdtree.h:50)
  float hit_tolerance;    // hit_tolerance_ (:52)
  float r1, r1_sq, r1_pad;  // radius sqrt(2)*grid_max/2 narrowed to float (:83), float(double r * double r), window half-width
  float r2, r2_sq, r2_pad;  // radius grid_min*2 + sqrt(2)*grid_max/2 (:95)
  double sin_den;           // map_grid_min_ * 2.0 (:98)
#if MCL3DL_NEAR_BITS
  NearBitsDev near;  // built for the marching search radius (r1_pad)
#endif
#if MCL3DL_KD_SKIP
  NearBitsDev far;   // built for a larger radius R2: a clear bit proves the next few marching steps clear as well
  float far_margin;  // 0.98 * (R2 - r1_pad): how far (rescaled metric) the march may move and still be covered
#endif
};

enum
{
  ST_SHORT = 0,
  ST_HIT = 1,
  ST_LONG = 2,
  ST_TOTAL_REFLECTION = 3
};

// --------------------------------------------------------------------------------------------
// Exact nearest-neighbour distance^2 (rescaled metric) within the radius; returns r2 if none.
__device__ __forceinline__ float nn_dist2(const NnGridDev& g, const LikDev& lp, float qx, float qy, float qz,
                                          uint32_t& n_rows, uint32_t& n_pts)
{
#if MCL3DL_NEAR_BITS
  if (!near_maybe(g.near, qx, qy, qz))
    return lp.r2;
#endif
  // Same cell function as the build kernel (monotone in its argument), applied to q -/+ rpad.
  int lx = __float2int_rd(fmul(fsub(fsub(qx, lp.rpad), g.ox), g.inv_cell));
  int ly = __float2int_rd(fmul(fsub(fsub(qy, lp.rpad), g.oy), g.inv_cell));
  int lz = __float2int_rd(fmul(fsub(fsub(qz, lp.rpad), g.oz), g.inv_cell));
  int hx = __float2int_rd(fmul(fsub(fadd(qx, lp.rpad), g.ox), g.inv_cell));
  int hy = __float2int_rd(fmul(fsub(fadd(qy, lp.rpad), g.oy), g.inv_cell));
  int hz = __float2int_rd(fmul(fsub(fadd(qz, lp.rpad), g.oz), g.inv_cell));
  lx = max(lx, 0);
  ly = max(ly, 0);
  lz = max(lz, 0);
  hx = min(hx, g.nx - 1);
  hy = min(hy, g.ny - 1);
  hz = min(hz, g.nz - 1);
  float best = lp.r2;
  if (lx > hx || ly > hy || lz > hz)
    return best;
  n_rows += static_cast<uint32_t>((hz - lz + 1) * (hy - ly + 1));
  for (int iz = lz; iz <= hz; ++iz)
  {
    for (int iy = ly; iy <= hy; ++iy)
    {
      const int row = (iz * g.ny + iy) * g.nx;
      const uint32_t s0 = __ldg(g.cell_start + row + lx);
      const uint32_t s1 = __ldg(g.cell_start + row + hx + 1);
      n_pts += s1 - s0;
      for (uint32_t s = s0; s < s1; ++s)
      {
        const float4 m = __ldg(g.pts + s);
        // flann::L2_Simple: sequential float accumulate of squared differences
        const float dx = fsub(qx, m.x);
        const float dy = fsub(qy, m.y);
        const float dz = fsub(qz, m.z);
        const float d = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
        best = fminf(best, d);  // KNNRadiusResultSet: keep d < worst
      }
    }
  }
  return best;
}

// --------------------------------------------------------------------------------------------
// Warp-chunk likelihood kernel (lik_kernel_wc, kernels.cuh; -DMCL3DL_LIK_CHUNKS=1): the per-lane pieces.  A warp works
// on 32 evals per round; the map points their windows contain are cut into CHUNKS of <= 4 consecutive points and the
// chunks are dealt to the lanes, so every lane of phase 2 has the same amount of work (the run-per-lane kernel
// lik_kernel_wi leaves half the lanes idle: runs are 1..30 points long, ncu r01x: 13.5 of 32 lanes on the load line).
// Written as functions of (lane, shared slab) so that tests/hostsim can run the rounds lane by lane on the host.
constexpr int kWcMaxRows = 9;
constexpr int kWcMaxDesc = 640;       // >= 9 * 32: one descriptor per run always fits (the overflow fallback)
constexpr int kWcMaxRunChunks = 127;  // chunk index field of a descriptor: 7 bits
constexpr int kWcOverflow = 1023;     // a lane reports this chunk count to force the whole-run fallback for its round

struct LikChunkSmem
{
  uint2 rows[kWcMaxRows][32];  // non-empty [start, end) runs of each lane's eval, compacted
  float qx[32], qy[32], qz[32];
  uint32_t best[32];           // float bits of the running min d^2 (non-negative floats order as uints)
  uint16_t desc[kWcMaxDesc];   // run k (4 bits) | eval lane (5 bits) << 4 | chunk index (7 bits) << 9
};

// Phase 1, lane = eval: the runs of the <= 3x3 window of q, from the window table (same fetch as lik_kernel_wi).
// Returns the number of runs stored to sm.rows[.][lane]; n_chunks = their total number of 4-point chunks, or
// kWcOverflow if one of the runs is too long for the descriptor's chunk index (very dense cells).
__device__ __forceinline__ int wc_window(const NnGridDev& g, const LikDev& lp, float qx, float qy, float qz, int lane,
                                         LikChunkSmem& sm, uint32_t& st_rows, uint32_t& st_pts, int& n_chunks)
{
  int nr = 0;
  n_chunks = 0;
  int lx = __float2int_rd(fmul(fsub(fsub(qx, lp.rpad), g.ox), g.inv_cell));
  int ly = __float2int_rd(fmul(fsub(fsub(qy, lp.rpad), g.oy), g.inv_cell));
  int lz = __float2int_rd(fmul(fsub(fsub(qz, lp.rpad), g.oz), g.inv_cell));
  int hx = __float2int_rd(fmul(fsub(fadd(qx, lp.rpad), g.ox), g.inv_cell));
  int hy = __float2int_rd(fmul(fsub(fadd(qy, lp.rpad), g.oy), g.inv_cell));
  int hz = __float2int_rd(fmul(fsub(fadd(qz, lp.rpad), g.oz), g.inv_cell));
  lx = max(lx, 0);
  ly = max(ly, 0);
  lz = max(lz, 0);
  hx = min(hx, g.nx - 1);
  hy = min(hy, min(g.ny - 1, ly + 2));
  hz = min(hz, min(g.nz - 1, lz + 2));
  if (lx > hx || ly > hy || lz > hz)
    return 0;
#if MCL3DL_NEAR_BITS
  if (!near_maybe(g.near, qx, qy, qz))
    return 0;
#endif
  st_rows += static_cast<uint32_t>((hz - lz + 1) * (hy - ly + 1));
  // the whole 3x3 window from the y-fastest window table: 2 aligned 16-byte loads per z layer, all issued first
  const int width = hx - lx + 1;  // 1..3 cells along x
  const int yb = ly & ~1;
  const int odd = ly & 1;
  uint4 ea[3], eb[3];
#pragma unroll
  for (int dz = 0; dz < 3; ++dz)
  {
    const int iz = min(lz + dz, hz);
    const uint4* src = reinterpret_cast<const uint4*>(g.row3 + (static_cast<size_t>(iz) * g.nx + lx) * g.nyp + yb);
    ea[dz] = __ldg(src);
    eb[dz] = __ldg(src + 1);
  }
#pragma unroll
  for (int dz = 0; dz < 3; ++dz)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
    {
      const int iy = ly + dy, iz = lz + dz;
      if (iy <= hy && iz <= hz)
      {
        // entry (iy - yb) of the four fetched ones
        uint32_t start, packed;
        if (dy == 0)
        {
          start = odd ? ea[dz].z : ea[dz].x;
          packed = odd ? ea[dz].w : ea[dz].y;
        }
        else if (dy == 1)
        {
          start = odd ? eb[dz].x : ea[dz].z;
          packed = odd ? eb[dz].y : ea[dz].w;
        }
        else
        {
          start = odd ? eb[dz].z : eb[dz].x;
          packed = odd ? eb[dz].w : eb[dz].y;
        }
        uint32_t cnt = width == 3 ? (packed >> 21) : (width == 2 ? ((packed >> 10) & 0x7ffu) : (packed & 0x3ffu));
        const uint32_t sat = width == 1 ? 0x3ffu : 0x7ffu;
        if (cnt == sat)
        {
          // the count did not fit the packed field (very dense cells): read the CSR bounds themselves
          const int row = (iz * g.ny + iy) * g.nx;
          start = __ldg(g.cell_start + row + lx);
          cnt = __ldg(g.cell_start + row + hx + 1) - start;
        }
        if (cnt)
        {
          sm.rows[nr][lane] = make_uint2(start, start + cnt);
          ++nr;
          st_pts += cnt;
          const uint32_t c4 = (cnt + 3u) >> 2;
          n_chunks = (c4 > static_cast<uint32_t>(kWcMaxRunChunks) || n_chunks >= kWcOverflow) ? kWcOverflow :
                                                                                                 n_chunks + static_cast<int>(c4);
        }
      }
    }
  return nr;
}

// After the warp's prefix sum: this lane's descriptors go to sm.desc[offset ...).  `whole` (warp-uniform): the chunk
// list would not fit (very dense maps), so every run becomes one descriptor and offset comes from the prefix sum of nr.
__device__ __forceinline__ void wc_write_descs(LikChunkSmem& sm, int lane, int nr, int offset, bool whole)
{
  int pos = offset;
  for (int k = 0; k < nr; ++k)
  {
    const uint32_t id = static_cast<uint32_t>(k) | (static_cast<uint32_t>(lane) << 4);
    if (whole)
    {
      sm.desc[pos++] = static_cast<uint16_t>(id);
      continue;
    }
    const uint2 run = sm.rows[k][lane];
    const uint32_t nc = (run.y - run.x + 3u) >> 2;
    for (uint32_t c = 0; c < nc; ++c) sm.desc[pos++] = static_cast<uint16_t>(id | (c << 9));
  }
}

// Phase 2, lane = descriptor: min d^2 over the chunk's (or, in a `whole` round, the run's) map points, merged into
// the owner eval's slot.
__device__ __forceinline__ void wc_process(LikChunkSmem& sm, uint32_t d, bool whole, const NnGridDev& g, const LikDev& lp)
{
  const int k = static_cast<int>(d & 15u), e = static_cast<int>((d >> 4) & 31u);
  const uint2 run = sm.rows[k][e];
  const float qx = sm.qx[e], qy = sm.qy[e], qz = sm.qz[e];
  uint32_t s0 = run.x, s1 = run.y;
  if (!whole)
  {
    s0 = run.x + 4u * (d >> 9);
    s1 = (s0 + 4u < run.y) ? s0 + 4u : run.y;
  }
  float best = lp.r2;
  for (uint32_t s = s0; s < s1; s += 4)
  {
    float4 mp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (s + u < s1)
        mp[u] = __ldg(g.pts + s + u);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (s + u < s1)
      {
        // flann::L2_Simple: sequential float accumulate of squared differences
        const float dx = fsub(qx, mp[u].x);
        const float dy = fsub(qy, mp[u].y);
        const float dz = fsub(qz, mp[u].z);
        best = fminf(best, fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));  // keep d < worst
      }
  }
  if (best < lp.r2)
    atomicMin(&sm.best[e], __float_as_uint(best));
}

// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int dda_to_index(float v, float mn, double grid)
{
  // toIndex, raycast_using_dda.h:205-210: float difference, double division, truncation
  return __double2int_rz(ddiv(static_cast<double>(fsub(v, mn)), grid));
}

// One ray: RaycastUsingDDA::setRay + getNextCastResult loop + getBeamStatus's decision.
//
// Measured choices (profiles/r01g_variants.txt, r01h_variants.txt; all variants bit-exact):
//   * the axis choice is evaluated branch-free (selects): 200 -> 177 us on c3, the 3-way branch
//     diverges whenever the lanes of a warp cross different faces;
//   * t_max needs float(|index - begin|): kept as a float counter (+1.0f per step) instead of an
//     int->float conversion per step;
//   * the occupancy word is re-read only when the cell leaves the current 32-cell word;
//   * replacing the twelve fp64 divisions of the set-up by guarded reciprocal multiplications, caching
//     the per-sensor ray origin, or screening the cone test in float changed nothing measurable
//     (the kernel is latency/occupancy bound, not fp64 bound) and were dropped again;
//   * occupancy matters most: __launch_bounds__(256, 4) (<= 64 registers) 200 -> 165 us.
__device__ __forceinline__ int cast_ray(const DdaGridDev& g, const F3& b, const F3& e, uint32_t& n_steps,
                                        uint32_t& n_occ, uint32_t& n_tested)
{
  // isPointWithinMap, :260-270
  if (b.x < g.min_x || g.max_x < b.x || b.y < g.min_y || g.max_y < b.y || b.z < g.min_z || g.max_z < b.z)
    return ST_LONG;
  // setRay, :66-104
  F3 d;
  d.x = fsub(e.x, b.x);
  d.y = fsub(e.y, b.y);
  d.z = fsub(e.z, b.z);
  const float nrm = __fsqrt_rn(dot3(d, d));
  F3 dir;
  dir.x = fdiv(d.x, nrm);
  dir.y = fdiv(d.y, nrm);
  dir.z = fdiv(d.z, nrm);
  const float ex = fadd(e.x, fmul(dir.x, g.hit_tolerance));
  const float ey = fadd(e.y, fmul(dir.y, g.hit_tolerance));
  const float ez = fadd(e.z, fmul(dir.z, g.hit_tolerance));
  const int bx = dda_to_index(b.x, g.min_x, g.grid);
  const int by = dda_to_index(b.y, g.min_y, g.grid);
  const int bz = dda_to_index(b.z, g.min_z, g.grid);
  // (end indices are clamped to +-2^29 so that garbage end points — infinite or 1e30 coordinates, for which the
  // reference's own int conversion is undefined — cannot overflow the differences; the begin index is inside the grid)
  constexpr int kIdxLim = 1 << 29;
  const int dix = max(-kIdxLim, min(kIdxLim, dda_to_index(ex, g.min_x, g.grid))) - bx;
  const int diy = max(-kIdxLim, min(kIdxLim, dda_to_index(ey, g.min_y, g.grid))) - by;
  const int diz = max(-kIdxLim, min(kIdxLim, dda_to_index(ez, g.min_z, g.grid))) - bz;
  const int max_movement = abs(dix) + abs(diy) + abs(diz);
  const int sx = dix < 0 ? -1 : 1, sy = diy < 0 ? -1 : 1, sz = diz < 0 ? -1 : 1;
  const float inf = __int_as_float(0x7f800000);
  float e0x = inf, e0y = inf, e0z = inf, tdx = inf, tdy = inf, tdz = inf;
  if (dix != 0)
  {
    // nearest = index * grid + min_p in double; |(nearest - begin) / dir| and |grid / dir| stored as float (:94-99)
    const double nearest = dadd(dmul(static_cast<double>(dir.x < 0 ? bx : bx + 1), g.grid), static_cast<double>(g.min_x));
    e0x = __double2float_rn(fabs(ddiv(dsub(nearest, static_cast<double>(b.x)), static_cast<double>(dir.x))));
    tdx = __double2float_rn(fabs(ddiv(g.grid, static_cast<double>(dir.x))));
  }
  if (diy != 0)
  {
    const double nearest = dadd(dmul(static_cast<double>(dir.y < 0 ? by : by + 1), g.grid), static_cast<double>(g.min_y));
    e0y = __double2float_rn(fabs(ddiv(dsub(nearest, static_cast<double>(b.y)), static_cast<double>(dir.y))));
    tdy = __double2float_rn(fabs(ddiv(g.grid, static_cast<double>(dir.y))));
  }
  if (diz != 0)
  {
    const double nearest = dadd(dmul(static_cast<double>(dir.z < 0 ? bz : bz + 1), g.grid), static_cast<double>(g.min_z));
    e0z = __double2float_rn(fabs(ddiv(dsub(nearest, static_cast<double>(b.z)), static_cast<double>(dir.z))));
    tdz = __double2float_rn(fabs(ddiv(g.grid, static_cast<double>(dir.z))));
  }
  float tx = e0x, ty = e0y, tz = e0z;
  int cx = bx, cy = by, cz = bz;
  float kx = 0.0f, ky = 0.0f, kz = 0.0f;  // float(|current - begin|): small integers, exact in float
  const int nxy = g.nx * g.ny;
  int last_w = -1;
  uint32_t word = 0;
  // getNextCastResult, :106-159: at most max_movement-1 cells; begin and end cells are never tested
  for (int pos = 1; pos < max_movement; ++pos)
  {
    ++n_steps;
    // strict-'<' ladder (:114-147): x if tx<ty && tx<tz; y if !(tx<ty) && ty<tz; else z (on ties z beats
    // y beats x).  incrementIndex (:192-203) recomputes t_max from the start:
    // float(edge0) + float(t_delta) * float(|index - begin|).
    const bool lt_xy = tx < ty;
    const bool ax = lt_xy && (tx < tz);
    const bool ay = !lt_xy && (ty < tz);
    const bool az = !ax && !ay;
    cx += ax ? sx : 0;
    cy += ay ? sy : 0;
    cz += az ? sz : 0;
    kx = ax ? fadd(kx, 1.0f) : kx;
    ky = ay ? fadd(ky, 1.0f) : ky;
    kz = az ? fadd(kz, 1.0f) : kz;
    tx = ax ? fadd(e0x, fmul(tdx, kx)) : tx;
    ty = ay ? fadd(e0y, fmul(tdy, ky)) : ty;
    tz = az ? fadd(e0z, fmul(tdz, kz)) : tz;
    if (static_cast<unsigned>(cx) >= static_cast<unsigned>(g.nx) || static_cast<unsigned>(cy) >= static_cast<unsigned>(g.ny) ||
        static_cast<unsigned>(cz) >= static_cast<unsigned>(g.nz))
      return ST_LONG;  // left the grid (:197-201)
    const int cell = cx + cy * g.nx + cz * nxy;
    const int w = cell >> 5;
    if (w != last_w)
    {
      word = __ldg(g.occ + w);
      last_w = w;
    }
    if (!((word >> (cell & 31)) & 1u))
      continue;
    // hasIntersection, :237-258: first point of the cell, in map order, inside the cone
    const uint32_t s0 = __ldg(g.cell_start + cell);
    const uint32_t s1 = __ldg(g.cell_start + cell + 1);
    ++n_occ;
    for (uint32_t s = s0; s < s1; ++s)
    {
      ++n_tested;
      const float4 m = __ldg(g.pts + s);
      F3 rel;
      rel.x = fsub(m.x, b.x);
      rel.y = fsub(m.y, b.y);
      rel.z = fsub(m.z, b.z);
      const double foot = static_cast<double>(fabsf(dot3(rel, dir)));
      const double a0 = dmul(g.ray_angle_half, foot);
      const double thr = fmax(dmul(a0, a0), g.min_dist_thr_sq);
      const double dsq = dsub(static_cast<double>(dot3(rel, rel)), dmul(foot, foot));
      if (dsq < thr)
      {
        // getBeamStatus, beam.cpp:164-189
        if (__float_as_uint(m.w) > g.filter_label_max)
          break;  // this cell's collision is filtered; the walk continues with the next cell
        if (1.0f > g.sin_total_ref)
        {
          const double ddx = static_cast<double>(fsub(e.x, m.x));
          const double ddy = static_cast<double>(fsub(e.y, m.y));
          const double ddz = static_cast<double>(fsub(e.z, m.z));
          const float dist_sq = __double2float_rn(dadd(dadd(dmul(ddx, ddx), dmul(ddy, ddy)), dmul(ddz, ddz)));
          return dist_sq < g.hit_range_sq ? ST_HIT : ST_SHORT;
        }
        return ST_TOTAL_REFLECTION;
      }
    }
  }
  return ST_LONG;
}

// 1-NN within a radius with the index of the winner (ChunkedKdtree::radiusSearch(p, r, id, d2, 1)); ties go to
// the lowest original index, like the CPU checkers.  General window (any radius below a few cells).
__device__ __forceinline__ bool nn_search_arg(const NnGridDev& g, float qx, float qy, float qz, float rpad, float r_sq,
                                              float& best, uint32_t& best_orig, uint32_t& n_tested)
{
  int lx = __float2int_rd(fmul(fsub(fsub(qx, rpad), g.ox), g.inv_cell));
  int ly = __float2int_rd(fmul(fsub(fsub(qy, rpad), g.oy), g.inv_cell));
  int lz = __float2int_rd(fmul(fsub(fsub(qz, rpad), g.oz), g.inv_cell));
  int hx = __float2int_rd(fmul(fsub(fadd(qx, rpad), g.ox), g.inv_cell));
  int hy = __float2int_rd(fmul(fsub(fadd(qy, rpad), g.oy), g.inv_cell));
  int hz = __float2int_rd(fmul(fsub(fadd(qz, rpad), g.oz), g.inv_cell));
  lx = max(lx, 0);
  ly = max(ly, 0);
  lz = max(lz, 0);
  hx = min(hx, g.nx - 1);
  hy = min(hy, g.ny - 1);
  hz = min(hz, g.nz - 1);
  best = r_sq;
  best_orig = 0xffffffffu;
  for (int iz = lz; iz <= hz; ++iz)
    for (int iy = ly; iy <= hy; ++iy)
    {
      if (lx > hx)
        break;
      const int row = (iz * g.ny + iy) * g.nx;
      const uint32_t s0 = __ldg(g.cell_start + row + lx);
      const uint32_t s1 = __ldg(g.cell_start + row + hx + 1);
      for (uint32_t s = s0; s < s1; ++s)
      {
        ++n_tested;
        const float4 m = __ldg(g.pts + s);
        const float dx = fsub(qx, m.x);
        const float dy = fsub(qy, m.y);
        const float dz = fsub(qz, m.z);
        const float d = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));  // flann::L2_Simple
        const uint32_t orig = __float_as_uint(m.w);
        if (d < best || (d == best && orig < best_orig && best_orig != 0xffffffffu))
        {
          best = d;
          best_orig = orig;
        }
      }
    }
  return best_orig != 0xffffffffu;
}

// One ray with RaycastUsingKDTree (raycasts/raycast_using_kdtree.h:57-110) + getBeamStatus (beam.cpp:157-192).
__device__ __forceinline__ int cast_ray_kd(const KdRayDev& k, const NnGridDev& nn, const DdaGridDev& g, const F3& b,
                                           const F3& e, uint32_t& n_steps, uint32_t& n_occ, uint32_t& n_tested)
{
  // setRay, :57-64
  F3 d;
  d.x = fsub(e.x, b.x);
  d.y = fsub(e.y, b.y);
  d.z = fsub(e.z, b.z);
  const float nrm = __fsqrt_rn(dot3(d, d));
  // The march is capped at 65 536 steps (6.5 km at the default 0.1 m step; the node clips scans at <= 10 m): an end
  // point at 1e30 or +inf must not keep a warp busy for 2^31 steps.  (The reference's float -> int conversion of such a
  // length is undefined; on x86 it yields INT_MIN, i.e. no march at all.)
  const int length = min(__float2int_rz(floorf(fdiv(fadd(nrm, k.hit_tolerance), k.grid_min))), 65536);
  F3 inc;
  inc.x = fmul(fdiv(d.x, nrm), k.grid_min);
  inc.y = fmul(fdiv(d.y, nrm), k.grid_min);
  inc.z = fmul(fdiv(d.z, nrm), k.grid_min);
  F3 pos;
  pos.x = fadd(b.x, inc.x);
  pos.y = fadd(b.y, inc.y);
  pos.z = fadd(b.z, inc.z);
#if MCL3DL_KD_SKIP
  // Steps that a clear bit of the far field covers: the march moves |inc * w| per step in the rescaled metric (plus the
  // rounding of the sequential adds, bounded generously by 1e-5 |q| per step), and a map point within r1 of a later
  // position would be within far_margin + r1 of this one.
  const float step_w = __fsqrt_rn(fadd(fadd(fmul(fmul(inc.x, nn.wx), fmul(inc.x, nn.wx)), fmul(fmul(inc.y, nn.wy), fmul(inc.y, nn.wy))),
                                      fmul(fmul(inc.z, nn.wz), fmul(inc.z, nn.wz))));
#endif
  // getNextCastResult, :66-110
  for (int count = 1; count < length; ++count)
  {
#if MCL3DL_KD_SKIP
    if (k.far.bits)
    {
      const float fx = fmul(pos.x, nn.wx), fy = fmul(pos.y, nn.wy), fz = fmul(pos.z, nn.wz);
      if (!near_maybe(k.far, fx, fy, fz))
      {
        const float per_step = fadd(step_w, fmul(1e-5f, fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fmaxf(fabsf(fz), 1.0f))));
        // this position and the next `extra` ones cannot collide: advance over them with the same sequential adds
        const int extra = min(__float2int_rz(fdiv(k.far_margin, per_step)), 64);
        const int adv = min(extra + 1, length - count);
        for (int i = 0; i < adv; ++i)
        {
          ++n_steps;
          pos.x = fadd(pos.x, inc.x);
          pos.y = fadd(pos.y, inc.y);
          pos.z = fadd(pos.z, inc.z);
        }
        count += adv - 1;  // the loop header adds the last one
        continue;
      }
    }
#endif
    ++n_steps;
    float d2;
    uint32_t id;
    const float qx = fmul(pos.x, nn.wx), qy = fmul(pos.y, nn.wy), qz = fmul(pos.z, nn.wz);
#if MCL3DL_NEAR_BITS
    // free space: the marching search cannot find anything, skip it (most steps of most rays)
    if (near_maybe(k.near, qx, qy, qz) && nn_search_arg(nn, qx, qy, qz, k.r1_pad, k.r1_sq, d2, id, n_tested))
#else
    if (nn_search_arg(nn, qx, qy, qz, k.r1_pad, k.r1_sq, d2, id, n_tested))
#endif
    {
      ++n_occ;
      const float4 m = __ldg(k.raw_pts + id);
      if (!(__float_as_uint(m.w) > g.filter_label_max))  // beam.cpp:168
      {
        const float d0 = __fsqrt_rn(d2);
        // pos_prev = pos_ - inc_ * 2.0 (:91), second search radius grid_min*2 + sqrt(2)*grid_max/2 (:95)
        const float px = fsub(pos.x, fmul(inc.x, 2.0f)), py = fsub(pos.y, fmul(inc.y, 2.0f)), pz = fsub(pos.z, fmul(inc.z, 2.0f));
        float d2b;
        uint32_t idb;
        float sin_ang = 1.0f;
        if (nn_search_arg(nn, fmul(px, nn.wx), fmul(py, nn.wy), fmul(pz, nn.wz), k.r2_pad, k.r2_sq, d2b, idb, n_tested))
        {
          const float d1 = __fsqrt_rn(d2b);
          sin_ang = __double2float_rn(ddiv(fabs(static_cast<double>(fsub(d1, d0))), k.sin_den));  // :98
        }
        if (sin_ang > g.sin_total_ref)
        {
          const double ddx = static_cast<double>(fsub(e.x, m.x));
          const double ddy = static_cast<double>(fsub(e.y, m.y));
          const double ddz = static_cast<double>(fsub(e.z, m.z));
          const float dist_sq = __double2float_rn(dadd(dadd(dmul(ddx, ddx), dmul(ddy, ddy)), dmul(ddz, ddz)));
          return dist_sq < g.hit_range_sq ? ST_HIT : ST_SHORT;
        }
        return ST_TOTAL_REFLECTION;
      }
    }
    pos.x = fadd(pos.x, inc.x);
    pos.y = fadd(pos.y, inc.y);
    pos.z = fadd(pos.z, inc.z);
  }
  return ST_LONG;
}

// begin = s.pos_ + s.rot_ * origins[label] with the RAW rot_ (beam.cpp:145)
__device__ __forceinline__ F3 ray_origin(const F3& pos, const Q4& q_raw, const float* __restrict__ origins_xyz,
                                         uint32_t label)
{
  F3 o;
  o.x = __ldg(origins_xyz + 3 * label);
  o.y = __ldg(origins_xyz + 3 * label + 1);
  o.z = __ldg(origins_xyz + 3 * label + 2);
  const F3 ro = qrot(q_raw, o);
  F3 begin;
  begin.x = fadd(pos.x, ro.x);
  begin.y = fadd(pos.y, ro.y);
  begin.z = fadd(pos.z, ro.z);
  return begin;
}

}  // namespace mcl3dl
