// Test-infrastructure shim (NOT product code): the minimum of <pcl/point_types.h>
// that the reference's ikd_Tree.{h,cpp} needs to compile in place from
// /root/reference.  PCL itself is absent from this image.
// Layout follows PCL's documented 48-byte PointXYZINormal:
//   {x,y,z,pad}{normal_x,normal_y,normal_z,pad}{intensity,curvature,pad,pad}
#pragma once
#include <cmath>
#include <memory>
#include <vector>
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad0; };
struct alignas(16) PointXYZI { float x, y, z, pad0; float intensity, pad1[3]; };
struct alignas(16) PointXYZINormal {
  float x, y, z, pad0;
  float normal_x, normal_y, normal_z, pad1;
  float intensity, curvature, pad2[2];
};
}  // namespace pcl
namespace Eigen {
template <class T> using aligned_allocator = std::allocator<T>;
}
