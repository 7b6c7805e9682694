// TEST INFRASTRUCTURE ONLY (oracle/_ref build): the subset of pcl::PointCloud the reference
// hot path uses (points/header/width/height, begin/end/erase/size/push_back/makeShared/+=).
#ifndef ORACLE_SHIM_PCL_POINT_CLOUD_H
#define ORACLE_SHIM_PCL_POINT_CLOUD_H
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <Eigen/Core>
namespace pcl
{
struct PCLHeader
{
  uint32_t seq;
  uint64_t stamp;
  std::string frame_id;
  PCLHeader() : seq(0), stamp(0) {}
};
template <typename PointT>
class PointCloud
{
public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  using iterator = typename std::vector<PointT>::iterator;
  using const_iterator = typename std::vector<PointT>::const_iterator;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width;
  uint32_t height;
  bool is_dense;
  PointCloud() : width(0), height(0), is_dense(true) {}
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const PointT& p)
  {
    points.push_back(p);
    width = static_cast<uint32_t>(points.size());
    height = 1;
  }
  iterator erase(iterator first, iterator last)
  {
    iterator it = points.erase(first, last);
    width = static_cast<uint32_t>(points.size());
    height = 1;
    return it;
  }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& rhs)
  {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = static_cast<uint32_t>(points.size());
    height = 1;
    return *this;
  }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};
}  // namespace pcl
#endif
