// PropagationDistanceFieldB200 — host-side C++ mirror of distance_field::PropagationDistanceField /
// distance_field::DistanceField over the b200df C ABI: the object behind CollisionEnvDistanceField::getDistanceField()
// and DistanceFieldCacheEntry::distance_field_, which CHOMP, SBPL and user code query directly.
//
// Same method names, argument meaning and return conventions as the reference:
//   distance_field/include/moveit/distance_field/distance_field.h:88-640         (DistanceField API)
//   distance_field/include/moveit/distance_field/propagation_distance_field.h     (constructors, stream I/O)
//   distance_field/src/distance_field.cpp:62-86, 196-330                          (gradient, shapes, octree)
//   distance_field/src/propagation_distance_field.cpp:47-94, 120-218, 507-525, 636-761
// The EDT behind it is the engine's exact one; getNearestCell is not offered (the engine keeps no
// closest-point field, nothing on the collision path reads it).
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <istream>
#include <iterator>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "collision_env_b200.hpp"

namespace distance_field_b200
{
using collision_detection_b200::Isometry3d;
using collision_detection_b200::Vector3d;
namespace shapes = collision_detection_b200::shapes;

class PropagationDistanceFieldB200
{
public:
  // propagation_distance_field.cpp:47-56 (origin = the grid corner)
  PropagationDistanceFieldB200(double size_x, double size_y, double size_z, double resolution, double origin_x,
                               double origin_y, double origin_z, double max_distance,
                               bool propagate_negative_distances = false)
    : max_distance_(max_distance), propagate_negative_(propagate_negative_distances)
  {
    create(size_x, size_y, size_z, resolution, origin_x, origin_y, origin_z);
  }

  // propagation_distance_field.cpp:64-74: a field read from a stream written by writeToStream
  PropagationDistanceFieldB200(std::istream& is, double max_distance, bool propagate_negative_distances = false)
    : max_distance_(max_distance), propagate_negative_(propagate_negative_distances)
  {
    readFromStream(is);
  }

  // a non-owning view of a field that lives elsewhere (CollisionEnvB200::getDistanceField(), a cache entry's self
  // field): queries and edits go to that field, the handle is not destroyed with the view
  explicit PropagationDistanceFieldB200(b200df_field* borrowed) : max_distance_(0.0), propagate_negative_(false), owns_(false)
  {
    b200df_field_desc d{};
    check(b200df_field_get_desc(borrowed, &d), "b200df_field_get_desc");
    field_ = borrowed;
    for (int i = 0; i < 3; ++i)
    {
      size_[i] = d.size[i];
      origin_[i] = d.origin[i];
    }
    resolution_ = d.resolution;
    max_distance_ = d.max_distance;
    propagate_negative_ = d.propagate_negative != 0;
    std::int32_t dims[4];
    check(b200df_field_dims(field_, dims), "b200df_field_dims");
    n_[0] = dims[0];
    n_[1] = dims[1];
    n_[2] = dims[2];
    max_distance_sq_ = dims[3];
  }
  // the view's cached host copy is refreshed on the next per-cell query (call after the owner changed the field)
  void invalidateHostCopy()
  {
    host_d2_valid_ = false;
  }

  PropagationDistanceFieldB200(const PropagationDistanceFieldB200&) = delete;
  PropagationDistanceFieldB200& operator=(const PropagationDistanceFieldB200&) = delete;

  ~PropagationDistanceFieldB200()
  {
    if (owns_)
      b200df_field_destroy(field_);
  }

  bool valid() const  // false after a stream constructor that could not read a field
  {
    return field_ != nullptr;
  }
  b200df_field* handle() const
  {
    return field_;
  }

  // ---- obstacles ---------------------------------------------------------------------------------------------
  void addPointsToField(const std::vector<Vector3d>& points)  // :176-195
  {
    const std::vector<double> p = flatten(points);
    check(b200df_field_add_points(field_, p.data(), (int64_t)points.size()), "b200df_field_add_points");
    host_d2_valid_ = false;
  }
  void removePointsFromField(const std::vector<Vector3d>& points)  // :197-218
  {
    const std::vector<double> p = flatten(points);
    check(b200df_field_remove_points(field_, p.data(), (int64_t)points.size()), "b200df_field_remove_points");
    host_d2_valid_ = false;
  }
  void updatePointsInField(const std::vector<Vector3d>& old_points, const std::vector<Vector3d>& new_points)  // :120-174
  {
    const std::vector<double> a = flatten(old_points), b = flatten(new_points);
    check(b200df_field_update_points(field_, a.data(), (int64_t)old_points.size(), b.data(),
                                     (int64_t)new_points.size()),
          "b200df_field_update_points");
    host_d2_valid_ = false;
  }
  void addShapeToField(const shapes::Shape* shape, const Isometry3d& pose)  // distance_field.cpp:221-226
  {
    check(b200df_field_add_shape(field_, shape->type, shape->dims, pose.m, 1.0, 0.0), "b200df_field_add_shape");
    host_d2_valid_ = false;
  }
  void removeShapeFromField(const shapes::Shape* shape, const Isometry3d& pose)  // distance_field.cpp:320-330
  {
    check(b200df_field_remove_shape(field_, shape->type, shape->dims, pose.m, 1.0, 0.0), "b200df_field_remove_shape");
    host_d2_valid_ = false;
  }
  // distance_field.cpp:290-318: the voxels of the old pose are exchanged for those of the new pose
  void moveShapeInField(const shapes::Shape* shape, const Isometry3d& old_pose, const Isometry3d& new_pose)
  {
    removeShapeFromField(shape, old_pose);
    addShapeToField(shape, new_pose);
  }
  // addOcTreeToField (distance_field.cpp:282-287) on the occupied leaves (centre xyz + edge length) of the tree
  void addOcTreeLeavesToField(const std::vector<double>& leaves4)
  {
    check(b200df_field_add_octree_leaves(field_, leaves4.data(), (int64_t)(leaves4.size() / 4)),
          "b200df_field_add_octree_leaves");
    host_d2_valid_ = false;
  }
  void reset()  // :507-525
  {
    check(b200df_field_reset(field_), "b200df_field_reset");
    host_d2_valid_ = false;
  }

  // ---- queries -----------------------------------------------------------------------------------------------
  // getDistance(double, double, double) (propagation_distance_field.h:411-431 + voxel_grid.h operator()): the
  // distance of the cell the point falls into, max_distance outside the grid
  double getDistance(double x, double y, double z) const
  {
    int ix, iy, iz;
    if (!worldToGrid(x, y, z, ix, iy, iz))
      return max_distance_;
    return getDistance(ix, iy, iz);
  }
  double getDistance(int x, int y, int z) const
  {
    if (!isCellValid(x, y, z))
      return max_distance_;
    syncHost();
    const std::size_t i = ((std::size_t)x * n_[1] + y) * n_[2] + z;
    double d = std::sqrt((double)host_d2_[i]) * resolution_;  // sqrt_table_[d2] (:88-91)
    if (propagate_negative_)
      d -= std::sqrt((double)host_neg_[i]) * resolution_;
    return d;
  }
  // DistanceField::getDistanceGradient (distance_field.cpp:62-86)
  double getDistanceGradient(double x, double y, double z, double& gradient_x, double& gradient_y, double& gradient_z,
                             bool& in_bounds) const
  {
    const double p[3] = { x, y, z };
    double d = 0.0, g[3] = { 0, 0, 0 };
    std::uint8_t inb = 0;
    check(b200df_field_distance_gradient(field_, p, 1, &d, g, &inb), "b200df_field_distance_gradient");
    gradient_x = g[0];
    gradient_y = g[1];
    gradient_z = g[2];
    in_bounds = inb != 0;
    return d;
  }
  // the batched form of the same query
  void getDistanceGradients(const std::vector<Vector3d>& points, std::vector<double>& dist, std::vector<Vector3d>& grad,
                            std::vector<std::uint8_t>& in_bounds) const
  {
    const std::vector<double> p = flatten(points);
    std::vector<double> g(points.size() * 3);
    dist.resize(points.size());
    in_bounds.resize(points.size());
    check(b200df_field_distance_gradient(field_, p.data(), (int64_t)points.size(), dist.data(), g.data(),
                                         in_bounds.data()),
          "b200df_field_distance_gradient");
    grad.resize(points.size());
    for (std::size_t i = 0; i < points.size(); ++i)
      grad[i] = Vector3d{ g[3 * i], g[3 * i + 1], g[3 * i + 2] };
  }

  // ---- geometry (voxel_grid.h:408-411, 515-538) ----------------------------------------------------------------
  bool isCellValid(int x, int y, int z) const
  {
    return x >= 0 && x < n_[0] && y >= 0 && y < n_[1] && z >= 0 && z < n_[2];
  }
  int getXNumCells() const
  {
    return n_[0];
  }
  int getYNumCells() const
  {
    return n_[1];
  }
  int getZNumCells() const
  {
    return n_[2];
  }
  bool gridToWorld(int x, int y, int z, double& world_x, double& world_y, double& world_z) const
  {
    world_x = origin_[0] + x * resolution_;
    world_y = origin_[1] + y * resolution_;
    world_z = origin_[2] + z * resolution_;
    return true;
  }
  bool worldToGrid(double world_x, double world_y, double world_z, int& x, int& y, int& z) const
  {
    const double oo = 1.0 / resolution_;
    x = (int)std::floor((world_x - (origin_[0] - 0.5 * resolution_)) * oo);
    y = (int)std::floor((world_y - (origin_[1] - 0.5 * resolution_)) * oo);
    z = (int)std::floor((world_z - (origin_[2] - 0.5 * resolution_)) * oo);
    return isCellValid(x, y, z);
  }
  double getUninitializedDistance() const
  {
    return max_distance_;
  }
  double getResolution() const
  {
    return resolution_;
  }
  double getSizeX() const
  {
    return size_[0];
  }
  double getSizeY() const
  {
    return size_[1];
  }
  double getSizeZ() const
  {
    return size_[2];
  }
  double getOriginX() const
  {
    return origin_[0];
  }
  double getOriginY() const
  {
    return origin_[1];
  }
  double getOriginZ() const
  {
    return origin_[2];
  }
  int getMaximumDistanceSquared() const
  {
    return max_distance_sq_;
  }
  // the whole field, x-major (what areDistanceFieldsDistancesEqual walks, test_distance_field.cpp:222-243)
  const std::vector<std::int32_t>& distanceSquares() const
  {
    syncHost();
    return host_d2_;
  }

  // ---- stream I/O (propagation_distance_field.cpp:636-761) -----------------------------------------------------
  // seven "key: value" lines in the default ostream formatting, then the occupancy bits (packed on the GPU) as one
  // zlib stream - what boost::iostreams::zlib_compressor with default parameters produces
  bool writeToStream(std::ostream& os) const
  {
    os << "resolution: " << resolution_ << std::endl;
    os << "size_x: " << size_[0] << std::endl;
    os << "size_y: " << size_[1] << std::endl;
    os << "size_z: " << size_[2] << std::endl;
    os << "origin_x: " << origin_[0] << std::endl;
    os << "origin_y: " << origin_[1] << std::endl;
    os << "origin_z: " << origin_[2] << std::endl;
    std::int64_t n = 0;
    if (b200df_field_pack_occupancy(field_, nullptr, 0, &n) != B200DF_OK)
      return false;
    std::vector<std::uint8_t> bits((std::size_t)n);
    if (b200df_field_pack_occupancy(field_, bits.data(), n, &n) != B200DF_OK)
      return false;
    uLongf zn = compressBound((uLong)bits.size());
    std::vector<std::uint8_t> z(zn);
    if (compress2(z.data(), &zn, bits.data(), (uLong)bits.size(), Z_DEFAULT_COMPRESSION) != Z_OK)
      return false;
    os.write(reinterpret_cast<const char*>(z.data()), (std::streamsize)zn);
    os.flush();
    return os.good();
  }

  // the field takes the geometry stored in the stream; max_distance and propagate_negative keep their values (:712)
  bool readFromStream(std::istream& is)
  {
    if (!is.good())
      return false;
    static const char* keys[7] = { "resolution:", "size_x:", "size_y:", "size_z:", "origin_x:", "origin_y:", "origin_z:" };
    double v[7];
    std::string temp;
    for (int i = 0; i < 7; ++i)
    {
      is >> temp;
      if (temp != keys[i])
        return false;
      is >> v[i];
      if (is.fail())
        return false;
    }
    create(v[1], v[2], v[3], v[0], v[4], v[5], v[6]);
    char nl;
    is.get(nl);  // "this should be newline"
    const std::vector<char> z((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    const std::size_t want = (std::size_t)n_[0] * n_[1] * ((n_[2] + 7) / 8);
    std::vector<std::uint8_t> bits(want);
    uLongf got = (uLongf)want;
    const int rc = uncompress(bits.data(), &got, reinterpret_cast<const Bytef*>(z.data()), (uLong)z.size());
    if ((rc != Z_OK && rc != Z_BUF_ERROR) || got < want)
      return false;  // the decompressor ran dry (:739-742)
    if (b200df_field_unpack_occupancy(field_, bits.data(), (std::int64_t)want) != B200DF_OK)
      return false;
    host_d2_valid_ = false;
    return true;
  }

private:
  static void check(int status, const char* what)
  {
    if (status != B200DF_OK)
      throw std::runtime_error(std::string(what) + ": " + b200df_last_error());
  }
  static std::vector<double> flatten(const std::vector<Vector3d>& pts)
  {
    std::vector<double> p(pts.size() * 3);
    for (std::size_t i = 0; i < pts.size(); ++i)
    {
      p[3 * i] = pts[i].x;
      p[3 * i + 1] = pts[i].y;
      p[3 * i + 2] = pts[i].z;
    }
    return p;
  }
  void create(double sx, double sy, double sz, double resolution, double ox, double oy, double oz)
  {
    if (field_ && owns_)
      b200df_field_destroy(field_);
    field_ = nullptr;
    owns_ = true;
    size_[0] = sx;
    size_[1] = sy;
    size_[2] = sz;
    origin_[0] = ox;
    origin_[1] = oy;
    origin_[2] = oz;
    resolution_ = resolution;
    b200df_field_desc d{};
    for (int i = 0; i < 3; ++i)
    {
      d.size[i] = size_[i];
      d.origin[i] = origin_[i];
    }
    d.resolution = resolution_;
    d.max_distance = max_distance_;
    d.propagate_negative = propagate_negative_ ? 1 : 0;
    check(b200df_field_create(&d, &field_), "b200df_field_create");
    std::int32_t dims[4];
    check(b200df_field_dims(field_, dims), "b200df_field_dims");
    n_[0] = dims[0];
    n_[1] = dims[1];
    n_[2] = dims[2];
    max_distance_sq_ = dims[3];
    host_d2_valid_ = false;
  }
  void syncHost() const
  {
    if (host_d2_valid_)
      return;
    const std::size_t total = (std::size_t)n_[0] * n_[1] * n_[2];
    host_d2_.resize(total);
    check(b200df_field_download_d2(field_, host_d2_.data(), 0), "b200df_field_download_d2");
    if (propagate_negative_)
    {
      host_neg_.resize(total);
      check(b200df_field_download_d2(field_, host_neg_.data(), 1), "b200df_field_download_d2");
    }
    host_d2_valid_ = true;
  }

  b200df_field* field_ = nullptr;
  double size_[3] = { 0, 0, 0 }, origin_[3] = { 0, 0, 0 }, resolution_ = 1.0, max_distance_;
  bool propagate_negative_;
  bool owns_ = true;
  int n_[3] = { 0, 0, 0 }, max_distance_sq_ = 0;
  mutable std::vector<std::int32_t> host_d2_, host_neg_;  // per-cell queries read a lazily refreshed host copy
  mutable bool host_d2_valid_ = false;
};
}  // namespace distance_field_b200
