// TEST INFRASTRUCTURE ONLY: pcl::PointRepresentation semantics restated from PCL's published
// behaviour: vectorize() copies the first nr_dimensions_ floats and multiplies by alpha_ if set.
#ifndef ORACLE_SHIM_PCL_POINT_REPRESENTATION_H
#define ORACLE_SHIM_PCL_POINT_REPRESENTATION_H
#include <memory>
#include <vector>
namespace pcl
{
template <typename PointT>
class PointRepresentation
{
public:
  using Ptr = std::shared_ptr<PointRepresentation<PointT>>;
  using ConstPtr = std::shared_ptr<const PointRepresentation<PointT>>;
  PointRepresentation() : nr_dimensions_(3), trivial_(false) {}
  virtual ~PointRepresentation() {}
  virtual void copyToFloatArray(const PointT& p, float* out) const = 0;
  virtual bool isValid(const PointT&) const { return true; }
  int getNumberOfDimensions() const { return nr_dimensions_; }
  void setRescaleValues(const float* rescale_array)
  {
    alpha_.assign(rescale_array, rescale_array + nr_dimensions_);
  }
  template <typename OutT>
  void vectorize(const PointT& p, OutT& out) const
  {
    float tmp[8];
    copyToFloatArray(p, tmp);
    if (alpha_.empty())
      for (int i = 0; i < nr_dimensions_; ++i) out[i] = tmp[i];
    else
      for (int i = 0; i < nr_dimensions_; ++i) out[i] = tmp[i] * alpha_[i];
  }
protected:
  int nr_dimensions_;
  std::vector<float> alpha_;
  bool trivial_;
};
template <typename PointT>
class DefaultPointRepresentation : public PointRepresentation<PointT>
{
public:
  void copyToFloatArray(const PointT& p, float* out) const override
  {
    out[0] = p.x;
    out[1] = p.y;
    out[2] = p.z;
  }
};
}  // namespace pcl
#endif
