// TEST INFRASTRUCTURE ONLY (oracle/_ref build): stand-in for pcl::KdTreeFLANN<PointT>.
//
// PCL/FLANN are third-party, un-vendored and absent here, so the call surface used by the
// reference (chunked_kdtree.h:108,117,207-211,231) is restated from the published behaviour:
//   * points and queries are vectorised through the PointRepresentation (xyz * rescale alpha);
//   * distance is flann::L2_Simple<float>: sequential float sum of squared differences;
//   * radiusSearch(p, r, ids, d2, max_nn=1) returns the single nearest neighbour with
//     d2 < r*r (FLANN KNNRadiusResultSet::addPoint rejects dist >= worst_dist_), d2 in the
//     rescaled space.
// Deliberate difference: the search here is EXACT.  The real node sets eps = map_grid_min/16
// (mcl_3dl.cpp:1328) which lets FLANN prune branches with mindist*(1+eps) > worst; that
// approximation is documented, not emulated.  setEpsilon() is accepted and ignored.
// The tree is a balanced median-split kd-tree with small leaves (same family as FLANN's
// KDTreeSingleIndex) so that CPU-baseline timings are representative of a kd-tree descent.
#ifndef ORACLE_SHIM_PCL_KDTREE_FLANN_H
#define ORACLE_SHIM_PCL_KDTREE_FLANN_H
#include <algorithm>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <vector>
#include <pcl/kdtree/kdtree.h>
namespace pcl
{
template <typename PointT>
class KdTreeFLANN
{
public:
  using Ptr = std::shared_ptr<KdTreeFLANN<PointT>>;
  using ConstPtr = std::shared_ptr<const KdTreeFLANN<PointT>>;
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  using PointRepresentationConstPtr = typename PointRepresentation<PointT>::ConstPtr;

  KdTreeFLANN() : epsilon_(0.0f), point_rep_(new DefaultPointRepresentation<PointT>) {}
  void setEpsilon(float eps) { epsilon_ = eps; }
  void setPointRepresentation(const PointRepresentationConstPtr& rep) { point_rep_ = rep; }
  PointCloudConstPtr getInputCloud() const { return input_; }
  void setInputCloud(const PointCloudConstPtr& cloud)
  {
    input_ = cloud;
    const size_t n = cloud->points.size();
    data_.resize(n * 3);
    for (size_t i = 0; i < n; ++i)
    {
      float* d = &data_[i * 3];
      point_rep_->vectorize(cloud->points[i], d);
    }
    order_.resize(n);
    for (size_t i = 0; i < n; ++i) order_[i] = static_cast<int>(i);
    nodes_.clear();
    if (n) build(0, static_cast<int>(n));
  }
  int radiusSearch(const PointT& p, double radius, std::vector<int>& k_indices,
                   std::vector<float>& k_sqr_distances, unsigned int max_nn = 0) const
  {
    if (max_nn != 1)
      throw std::runtime_error("shim KdTreeFLANN: only max_nn == 1 is implemented");
    float q[3];
    point_rep_->vectorize(p, q);
    float best = static_cast<float>(radius * radius);
    int best_id = -1;
    if (!nodes_.empty())
    {
      float off[3] = {0.f, 0.f, 0.f};
      search(0, q, 0.f, off, best, best_id);
    }
    if (best_id < 0)
    {
      k_indices.clear();
      k_sqr_distances.clear();
      return 0;
    }
    k_indices.assign(1, best_id);
    k_sqr_distances.assign(1, best);
    return 1;
  }

private:
  struct Node
  {
    int left, right;  // children, or [begin,end) into order_ for a leaf
    int axis;         // -1 for leaf
    float split;
  };
  static constexpr int LEAF = 10;
  int build(int b, int e)
  {
    const int id = static_cast<int>(nodes_.size());
    nodes_.push_back(Node());
    if (e - b <= LEAF)
    {
      nodes_[id] = Node{b, e, -1, 0.f};
      return id;
    }
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) lo[k] = hi[k] = data_[order_[b] * 3 + k];
    for (int i = b + 1; i < e; ++i)
      for (int k = 0; k < 3; ++k)
      {
        const float v = data_[order_[i] * 3 + k];
        lo[k] = std::min(lo[k], v);
        hi[k] = std::max(hi[k], v);
      }
    int axis = 0;
    for (int k = 1; k < 3; ++k)
      if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
    const int m = (b + e) / 2;
    std::nth_element(order_.begin() + b, order_.begin() + m, order_.begin() + e,
                     [&](int a, int c) { return data_[a * 3 + axis] < data_[c * 3 + axis]; });
    const float split = data_[order_[m] * 3 + axis];
    const int l = build(b, m);
    const int r = build(m, e);
    nodes_[id] = Node{l, r, axis, split};
    return id;
  }
  // d2 uses the same operation order as flann::L2_Simple (sequential accumulate).
  static float dist2(const float* a, const float* b)
  {
    float r = 0.f;
    for (int k = 0; k < 3; ++k)
    {
      const float d = a[k] - b[k];
      r += d * d;
    }
    return r;
  }
  void search(int id, const float* q, float mind, float* off, float& best, int& best_id) const
  {
    const Node& nd = nodes_[id];
    if (nd.axis < 0)
    {
      for (int i = nd.left; i < nd.right; ++i)
      {
        const int pid = order_[i];
        const float d = dist2(q, &data_[pid * 3]);
        if (d < best || (d == best && best_id >= 0 && pid < best_id))
        {
          best = d;
          best_id = pid;
        }
      }
      return;
    }
    const float diff = q[nd.axis] - nd.split;
    const int near = diff < 0 ? nd.left : nd.right;
    const int far = diff < 0 ? nd.right : nd.left;
    search(near, q, mind, off, best, best_id);
    const float old = off[nd.axis];
    // exact lower bound on the far side, slightly relaxed so float rounding can never prune a true NN
    const float far_mind = mind - old * old + diff * diff;
    if (far_mind * 0.999f <= best)
    {
      off[nd.axis] = diff;
      search(far, q, far_mind, off, best, best_id);
      off[nd.axis] = old;
    }
  }
  float epsilon_;
  PointRepresentationConstPtr point_rep_;
  PointCloudConstPtr input_;
  std::vector<float> data_;
  std::vector<int> order_;
  std::vector<Node> nodes_;
};
}  // namespace pcl
#endif
