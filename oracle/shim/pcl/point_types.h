// TEST INFRASTRUCTURE ONLY (oracle/_ref build): PCL point macros/types named by the reference.
#ifndef ORACLE_SHIM_PCL_POINT_TYPES_H
#define ORACLE_SHIM_PCL_POINT_TYPES_H
#include <cstdint>
#include <Eigen/Core>
#define PCL_ADD_POINT4D \
  union EIGEN_ALIGN16   \
  {                     \
    float data[4];      \
    struct              \
    {                   \
      float x;          \
      float y;          \
      float z;          \
    };                  \
  };
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)
namespace pcl
{
struct EIGEN_ALIGN16 PointXYZ
{
  PCL_ADD_POINT4D;
  inline PointXYZ() { x = y = z = 0.0f; data[3] = 1.0f; }
  inline PointXYZ(float _x, float _y, float _z) { x = _x; y = _y; z = _z; data[3] = 1.0f; }
};
}  // namespace pcl
#endif
