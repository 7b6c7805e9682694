// cuda_shim.h — TEST INFRASTRUCTURE ONLY.  Host stand-ins for the handful of CUDA scalar intrinsics used by
// mcl_3dl_b200/csrc/device_math.cuh and device_funcs.cuh, so that those very files compile with g++ and every
// per-thread device function can be checked against the oracle without a GPU (tests/test_hostsim.py).
// Compiled with -O2 -ffp-contract=off on x86-64 (SSE2): float/double +,-,*,/ and sqrt are IEEE round-to-nearest,
// which is what __f*_rn / __d*_rn / __fsqrt_rn guarantee on the device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

#define __device__
#define __host__
#define __forceinline__ inline
#define MCL3DL_HOSTSIM 1

struct float4
{
  float x, y, z, w;
};
struct uint2
{
  unsigned int x, y;
};
struct uint4
{
  unsigned int x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __double2float_rn(double a) { return static_cast<float>(a); }

// float/double -> int conversions saturate and map NaN to 0 on the device
inline int __double2int_rz(double a)
{
  if (a != a) return 0;
  if (a >= 2147483647.0) return std::numeric_limits<int>::max();
  if (a <= -2147483648.0) return std::numeric_limits<int>::min();
  return static_cast<int>(a);
}
inline int __float2int_rz(float a) { return __double2int_rz(static_cast<double>(a)); }
inline int __float2int_rd(float a) { return __double2int_rz(std::floor(static_cast<double>(a))); }

inline unsigned int __float_as_uint(float f)
{
  unsigned int u;
  std::memcpy(&u, &f, 4);
  return u;
}
inline float __uint_as_float(unsigned int u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline float __int_as_float(int i)
{
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}
template <typename T>
inline T __ldg(const T* p)
{
  return *p;
}
inline unsigned int atomicOr(unsigned int* p, unsigned int v)  // single-threaded host build
{
  const unsigned int old = *p;
  *p |= v;
  return old;
}
inline unsigned int atomicMin(unsigned int* p, unsigned int v)  // single-threaded host build
{
  const unsigned int old = *p;
  if (v < old) *p = v;
  return old;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
