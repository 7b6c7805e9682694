// TEST INFRASTRUCTURE ONLY: pcl::getMinMax3D restated (component-wise min/max of x,y,z; dense clouds).
#ifndef ORACLE_SHIM_PCL_COMMON_H
#define ORACLE_SHIM_PCL_COMMON_H
#include <cfloat>
#include <pcl/point_cloud.h>
namespace pcl
{
template <typename PointT>
inline void getMinMax3D(const PointCloud<PointT>& cloud, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt)
{
  min_pt = Eigen::Vector4f(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
  max_pt = Eigen::Vector4f(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (const auto& p : cloud.points)
  {
    const float c[4] = {p.x, p.y, p.z, p.data[3]};
    for (int i = 0; i < 4; ++i)
    {
      if (c[i] < min_pt[i]) min_pt[i] = c[i];
      if (c[i] > max_pt[i]) max_pt[i] = c[i];
    }
  }
}
}  // namespace pcl
#endif
