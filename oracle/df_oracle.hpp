// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped engine.
//
// CPU restatement (double precision, no dependencies) of MoveIt's sphere-decomposed
// distance-field collision path: moveit_core/distance_field + moveit_core/collision_distance_field
// + the forward-kinematics slice of moveit_core/robot_state / robot_model.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call anything in this directory.  The product (moveit_b200/, include/) never does.
//
// Parity status: PINNED against every restatable known-answer test the reference holds for this
// path (tests/test_oracle_golden.py: test_voxel_grid.cpp:43-95, test_distance_field.cpp:329-429,
// :431-499, :601-649, :853-858, :882-928, test_collision_distance_field.cpp:336-383,
// robot_state_test.cpp:432-559).  The reference itself cannot be compiled here (Eigen, Boost,
// geometric_shapes, octomap, urdfdom, ROS absent), so there is no oracle/_ref.
//
// Third-party arithmetic that is NOT under /root/reference and is restated from its published
// behaviour (unpinned at the 1-ulp level):
//   * Eigen 3.3 (Ubuntu focal 3.3.7): fixed-size products.  Restated as left-to-right sums
//     ((a0*b0 + a1*b1) + a2*b2) + a3*b3, no FMA contraction (compile with -ffp-contract=off),
//     Isometry*Vector3 = t + ((r0*v0 + r1*v1) + r2*v2), squaredNorm = (x*x + y*y) + z*z.
//   * geometric_shapes (noetic-devel, >=0.5.2, moveit_core/package.xml:32): bodies::Box /
//     Cylinder / Sphere containsPoint, computeBoundingSphere, computeBoundingCylinder,
//     mergeBoundingSpheres.
//   * octomap: only "occupied leaf centre + size" reaches this path; leaves are inputs here.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/moveit_core unless stated otherwise).
#pragma once
#include <zlib.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace orc
{
struct V3
{
  double x, y, z;
};
struct I3
{
  int x, y, z;
};

static inline V3 operator-(const V3& a, const V3& b)
{
  return V3{ a.x - b.x, a.y - b.y, a.z - b.z };
}
static inline double sqnorm(const V3& v)
{
  return (v.x * v.x + v.y * v.y) + v.z * v.z;
}
static inline double norm(const V3& v)
{
  return std::sqrt(sqnorm(v));
}
static inline double dot(const V3& a, const V3& b)
{
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}

// Rigid transform, rows of the 3x4 affine part; the implied 4th row is (0,0,0,1).
struct Iso
{
  double m[3][4];
  static Iso identity()
  {
    Iso r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        r.m[i][j] = (i == j) ? 1.0 : 0.0;
    return r;
  }
};

// A.affine() * B.matrix()  (Eigen 3x4 * 4x4), robot_state.cpp:727-737
static inline Iso mul(const Iso& a, const Iso& b)
{
  static const double last_row[4] = { 0.0, 0.0, 0.0, 1.0 };
  Iso r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      r.m[i][j] = ((a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j]) + a.m[i][3] * last_row[j];
  return r;
}

// Isometry3d * Vector3d, collision_distance_field_types.cpp:393-396
static inline V3 apply(const Iso& t, const V3& v)
{
  V3 r;
  r.x = t.m[0][3] + ((t.m[0][0] * v.x + t.m[0][1] * v.y) + t.m[0][2] * v.z);
  r.y = t.m[1][3] + ((t.m[1][0] * v.x + t.m[1][1] * v.y) + t.m[1][2] * v.z);
  r.z = t.m[2][3] + ((t.m[2][0] * v.x + t.m[2][1] * v.y) + t.m[2][2] * v.z);
  return r;
}

// Isometry3d::inverse() (Eigen Isometry mode): R^T, -(R^T t)
static inline Iso inverse(const Iso& t)
{
  Iso r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r.m[i][j] = t.m[j][i];
  for (int i = 0; i < 3; ++i)
    r.m[i][3] = -((r.m[i][0] * t.m[0][3] + r.m[i][1] * t.m[1][3]) + r.m[i][2] * t.m[2][3]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// VoxelGrid geometry — distance_field/include/moveit/distance_field/voxel_grid.h
// ------------------------------------------------------------------------------------------------
struct GridGeom
{
  double size[3], origin[3], origin_minus[3], resolution, oo_resolution;
  int num_cells[3];
  int stride1, stride2;
  long num_cells_total;

  // voxel_grid.h:367-399 (resize)
  void init(double sx, double sy, double sz, double res, double ox, double oy, double oz)
  {
    size[0] = sx;
    size[1] = sy;
    size[2] = sz;
    origin[0] = ox;
    origin[1] = oy;
    origin[2] = oz;
    for (int i = 0; i < 3; ++i)
      origin_minus[i] = origin[i] - 0.5 * res;
    num_cells_total = 1;
    resolution = res;
    oo_resolution = 1.0 / resolution;
    for (int i = 0; i < 3; ++i)
    {
      num_cells[i] = size[i] * oo_resolution;  // double -> int truncation, voxel_grid.h:387
      num_cells_total *= num_cells[i];
    }
    stride1 = num_cells[1] * num_cells[2];
    stride2 = num_cells[2];
  }
  // voxel_grid.h:408-411
  bool isCellValid(int x, int y, int z) const
  {
    return (x >= 0 && x < num_cells[0] && y >= 0 && y < num_cells[1] && z >= 0 && z < num_cells[2]);
  }
  // voxel_grid.h:426-429 (z fastest)
  long ref(int x, int y, int z) const
  {
    return (long)x * stride1 + (long)y * stride2 + z;
  }
  // voxel_grid.h:515-532
  int getCellFromLocation(int dim, double loc) const
  {
    return int(std::floor((loc - origin_minus[dim]) * oo_resolution));
  }
  // voxel_grid.h:535-538
  double getLocationFromCell(int dim, int cell) const
  {
    return origin[dim] + resolution * (double(cell));
  }
  // voxel_grid.h:563-569
  bool worldToGrid(double wx, double wy, double wz, int& x, int& y, int& z) const
  {
    x = getCellFromLocation(0, wx);
    y = getCellFromLocation(1, wy);
    z = getCellFromLocation(2, wz);
    return isCellValid(x, y, z);
  }
};

// ------------------------------------------------------------------------------------------------
// PropagationDistanceField — distance_field/src/propagation_distance_field.cpp (faithful, F4/Q1:
// this is NOT an exact EDT) and DistanceField::getDistanceGradient (distance_field.cpp:62-86).
// ------------------------------------------------------------------------------------------------
struct Voxel  // propagation_distance_field.h:75-109
{
  int distance_square;
  int negative_distance_square;
  I3 closest_point;
  I3 closest_negative_point;
  int update_direction;
  int negative_update_direction;
};

class PropagationDistanceField
{
public:
  static constexpr int UNINITIALIZED = -1;

  PropagationDistanceField(double sx, double sy, double sz, double res, double ox, double oy, double oz,
                           double max_distance, bool propagate_negative)
    : propagate_negative_(propagate_negative), max_distance_(max_distance)
  {
    size_[0] = sx;
    size_[1] = sy;
    size_[2] = sz;
    origin_[0] = ox;
    origin_[1] = oy;
    origin_[2] = oz;
    resolution_ = res;
    inv_twice_resolution_ = 1.0 / (2.0 * resolution_);  // int! distance_field.h:622, distance_field.cpp:56 (Q3)
    initialize();
  }

  // propagation_distance_field.cpp:76-94
  void initialize()
  {
    max_distance_sq_ = std::ceil(max_distance_ / resolution_) * std::ceil(max_distance_ / resolution_);
    g_.init(size_[0], size_[1], size_[2], resolution_, origin_[0], origin_[1], origin_[2]);
    cells_.assign(g_.num_cells_total > 0 ? g_.num_cells_total : 0, defaultVoxel());
    initNeighborhoods();
    bucket_queue_.assign(max_distance_sq_ + 1, {});
    negative_bucket_queue_.assign(max_distance_sq_ + 1, {});
    sqrt_table_.resize(max_distance_sq_ + 1);
    for (int i = 0; i <= max_distance_sq_; ++i)
      sqrt_table_[i] = std::sqrt(double(i)) * resolution_;
    reset();
  }

  // propagation_distance_field.cpp:507-525
  void reset()
  {
    std::fill(cells_.begin(), cells_.end(), defaultVoxel());
    for (int x = 0; x < g_.num_cells[0]; x++)
      for (int y = 0; y < g_.num_cells[1]; y++)
        for (int z = 0; z < g_.num_cells[2]; z++)
        {
          Voxel& v = cell(x, y, z);
          v.closest_negative_point = I3{ x, y, z };
          v.negative_distance_square = 0;
        }
  }

  // propagation_distance_field.cpp:176-195
  void addPointsToField(const std::vector<V3>& points)
  {
    std::vector<I3> voxel_points;
    for (const V3& p : points)
    {
      I3 loc;
      bool valid = g_.worldToGrid(p.x, p.y, p.z, loc.x, loc.y, loc.z);
      if (valid)
        if (cell(loc.x, loc.y, loc.z).distance_square > 0)
          voxel_points.push_back(loc);
    }
    addNewObstacleVoxels(voxel_points);
  }

  // propagation_distance_field.cpp:197-218
  void removePointsFromField(const std::vector<V3>& points)
  {
    std::vector<I3> voxel_points;
    for (const V3& p : points)
    {
      I3 loc;
      bool valid = g_.worldToGrid(p.x, p.y, p.z, loc.x, loc.y, loc.z);
      if (valid)
        voxel_points.push_back(loc);
    }
    removeObstacleVoxels(voxel_points);
  }

  // propagation_distance_field.cpp:120-174; ordering functor propagation_distance_field.h:57-69 (z, then y, then x)
  void updatePointsInField(const std::vector<V3>& old_points, const std::vector<V3>& new_points)
  {
    auto cmp = [](const I3& a, const I3& b) {
      if (a.z != b.z)
        return a.z < b.z;
      else if (a.y != b.y)
        return a.y < b.y;
      else if (a.x != b.x)
        return a.x < b.x;
      return false;
    };
    std::set<I3, decltype(cmp)> old_set(cmp), new_set(cmp);
    for (const V3& p : old_points)
    {
      I3 loc;
      if (g_.worldToGrid(p.x, p.y, p.z, loc.x, loc.y, loc.z))
        old_set.insert(loc);
    }
    for (const V3& p : new_points)
    {
      I3 loc;
      if (g_.worldToGrid(p.x, p.y, p.z, loc.x, loc.y, loc.z))
        new_set.insert(loc);
    }
    std::vector<I3> old_not_new, new_not_old;
    std::set_difference(old_set.begin(), old_set.end(), new_set.begin(), new_set.end(),
                        std::inserter(old_not_new, old_not_new.end()), cmp);
    std::set_difference(new_set.begin(), new_set.end(), old_set.begin(), old_set.end(),
                        std::inserter(new_not_old, new_not_old.end()), cmp);
    std::vector<I3> new_not_in_current;
    for (const I3& loc : new_not_old)
      if (cell(loc.x, loc.y, loc.z).distance_square != 0)
        new_not_in_current.push_back(loc);
    removeObstacleVoxels(old_not_new);
    addNewObstacleVoxels(new_not_in_current);
  }

  // propagation_distance_field.cpp:220-300
  void addNewObstacleVoxels(const std::vector<I3>& voxel_points)
  {
    int initial_update_direction = getDirectionNumber(0, 0, 0);
    std::vector<I3> negative_stack;
    for (const I3& loc : voxel_points)
    {
      Voxel& voxel = cell(loc.x, loc.y, loc.z);
      voxel.distance_square = 0;
      voxel.closest_point = loc;
      voxel.update_direction = initial_update_direction;
      bucket_queue_[0].push_back(loc);
      if (propagate_negative_)
      {
        voxel.negative_distance_square = max_distance_sq_;
        voxel.closest_negative_point = I3{ UNINITIALIZED, UNINITIALIZED, UNINITIALIZED };
        negative_stack.push_back(loc);
      }
    }
    propagatePositive();

    if (propagate_negative_)
    {
      while (!negative_stack.empty())
      {
        I3 loc = negative_stack.back();
        negative_stack.pop_back();
        for (int neighbor = 0; neighbor < 27; neighbor++)
        {
          I3 diff = direction_number_to_direction_[neighbor];
          I3 nloc{ loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
          if (g_.isCellValid(nloc.x, nloc.y, nloc.z))
          {
            Voxel& nvoxel = cell(nloc.x, nloc.y, nloc.z);
            I3& close_point = nvoxel.closest_negative_point;
            if (!g_.isCellValid(close_point.x, close_point.y, close_point.z))
              close_point = nloc;
            Voxel& closest_point_voxel = cell(close_point.x, close_point.y, close_point.z);
            if (closest_point_voxel.negative_distance_square != 0)
            {
              if (nvoxel.negative_distance_square != max_distance_sq_)
              {
                nvoxel.negative_distance_square = max_distance_sq_;
                nvoxel.closest_negative_point = I3{ UNINITIALIZED, UNINITIALIZED, UNINITIALIZED };
                negative_stack.push_back(nloc);
              }
            }
            else
            {
              nvoxel.negative_update_direction = initial_update_direction;
              negative_bucket_queue_[0].push_back(nloc);
            }
          }
        }
      }
      propagateNegative();
    }
  }

  // propagation_distance_field.cpp:302-387
  void removeObstacleVoxels(const std::vector<I3>& voxel_points)
  {
    std::vector<I3> stack;
    int initial_update_direction = getDirectionNumber(0, 0, 0);
    for (const I3& p : voxel_points)
    {
      Voxel& voxel = cell(p.x, p.y, p.z);
      voxel.distance_square = max_distance_sq_;
      voxel.closest_point = p;
      voxel.update_direction = initial_update_direction;
      stack.push_back(p);
      if (propagate_negative_)
      {
        voxel.negative_distance_square = 0;
        voxel.closest_negative_point = p;
        voxel.negative_update_direction = initial_update_direction;
        negative_bucket_queue_[0].push_back(p);
      }
    }
    while (!stack.empty())
    {
      I3 loc = stack.back();
      stack.pop_back();
      for (int neighbor = 0; neighbor < 27; neighbor++)
      {
        I3 diff = direction_number_to_direction_[neighbor];
        I3 nloc{ loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
        if (g_.isCellValid(nloc.x, nloc.y, nloc.z))
        {
          Voxel& nvoxel = cell(nloc.x, nloc.y, nloc.z);
          I3& close_point = nvoxel.closest_point;
          if (!g_.isCellValid(close_point.x, close_point.y, close_point.z))
            close_point = nloc;
          Voxel& closest_point_voxel = cell(close_point.x, close_point.y, close_point.z);
          if (closest_point_voxel.distance_square != 0)
          {
            if (nvoxel.distance_square != max_distance_sq_)
            {
              nvoxel.distance_square = max_distance_sq_;
              nvoxel.closest_point = nloc;
              nvoxel.update_direction = initial_update_direction;
              stack.push_back(nloc);
            }
          }
          else
          {
            nvoxel.update_direction = initial_update_direction;
            bucket_queue_[0].push_back(nloc);
          }
        }
      }
    }
    propagatePositive();
    if (propagate_negative_)
      propagateNegative();
  }

  // propagation_distance_field.cpp:389-444.  The reference iterates bucket i with the end iterator
  // captured up front and clears the bucket afterwards; entries appended to the bucket being walked
  // are therefore dropped, and entries appended to an already-walked bucket wait for the next call.
  void propagatePositive()
  {
    for (unsigned int i = 0; i < bucket_queue_.size(); ++i)
    {
      const size_t n_at_start = bucket_queue_[i].size();
      for (size_t it = 0; it < n_at_start; ++it)
      {
        const I3 loc = bucket_queue_[i][it];
        Voxel* vptr = &cell(loc.x, loc.y, loc.z);
        int d = i;
        if (d > 1)
          d = 1;
        if (vptr->update_direction < 0 || vptr->update_direction > 26)
          continue;
        const std::vector<I3>& neighborhood = neighborhoods_[d][vptr->update_direction];
        for (const I3& diff : neighborhood)
        {
          I3 nloc{ loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
          if (!g_.isCellValid(nloc.x, nloc.y, nloc.z))
            continue;
          Voxel* neighbor = &cell(nloc.x, nloc.y, nloc.z);
          int dx = vptr->closest_point.x - nloc.x, dy = vptr->closest_point.y - nloc.y,
              dz = vptr->closest_point.z - nloc.z;
          int new_distance_sq = dx * dx + dy * dy + dz * dz;
          if (new_distance_sq > max_distance_sq_)
            continue;
          if (new_distance_sq < neighbor->distance_square)
          {
            neighbor->distance_square = new_distance_sq;
            neighbor->closest_point = vptr->closest_point;
            neighbor->update_direction = getDirectionNumber(diff.x, diff.y, diff.z);
            bucket_queue_[new_distance_sq].push_back(nloc);
          }
        }
      }
      bucket_queue_[i].clear();
    }
  }

  // propagation_distance_field.cpp:446-505
  void propagateNegative()
  {
    for (unsigned int i = 0; i < negative_bucket_queue_.size(); ++i)
    {
      const size_t n_at_start = negative_bucket_queue_[i].size();
      for (size_t it = 0; it < n_at_start; ++it)
      {
        const I3 loc = negative_bucket_queue_[i][it];
        Voxel* vptr = &cell(loc.x, loc.y, loc.z);
        int d = i;
        if (d > 1)
          d = 1;
        if (vptr->negative_update_direction < 0 || vptr->negative_update_direction > 26)
          continue;
        const std::vector<I3>& neighborhood = neighborhoods_[d][vptr->negative_update_direction];
        for (const I3& diff : neighborhood)
        {
          I3 nloc{ loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
          if (!g_.isCellValid(nloc.x, nloc.y, nloc.z))
            continue;
          Voxel* neighbor = &cell(nloc.x, nloc.y, nloc.z);
          int dx = vptr->closest_negative_point.x - nloc.x, dy = vptr->closest_negative_point.y - nloc.y,
              dz = vptr->closest_negative_point.z - nloc.z;
          int new_distance_sq = dx * dx + dy * dy + dz * dz;
          if (new_distance_sq > max_distance_sq_)
            continue;
          if (new_distance_sq < neighbor->negative_distance_square)
          {
            neighbor->negative_distance_square = new_distance_sq;
            neighbor->closest_negative_point = vptr->closest_negative_point;
            neighbor->negative_update_direction = getDirectionNumber(diff.x, diff.y, diff.z);
            negative_bucket_queue_[new_distance_sq].push_back(nloc);
          }
        }
      }
      negative_bucket_queue_[i].clear();
    }
  }

  // propagation_distance_field.cpp:527-583
  void initNeighborhoods()
  {
    direction_number_to_direction_.resize(27);
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz)
          direction_number_to_direction_[getDirectionNumber(dx, dy, dz)] = I3{ dx, dy, dz };
    neighborhoods_.assign(2, std::vector<std::vector<I3>>(27));
    for (int n = 0; n < 2; n++)
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dz = -1; dz <= 1; ++dz)
          {
            int direction_number = getDirectionNumber(dx, dy, dz);
            for (int tdx = -1; tdx <= 1; ++tdx)
              for (int tdy = -1; tdy <= 1; ++tdy)
                for (int tdz = -1; tdz <= 1; ++tdz)
                {
                  if (tdx == 0 && tdy == 0 && tdz == 0)
                    continue;
                  if (n >= 1)
                  {
                    if ((std::abs(tdx) + std::abs(tdy) + std::abs(tdz)) != 1)
                      continue;
                    if (dx * tdx < 0 || dy * tdy < 0 || dz * tdz < 0)
                      continue;
                  }
                  neighborhoods_[n][direction_number].push_back(I3{ tdx, tdy, tdz });
                }
          }
  }

  // propagation_distance_field.cpp:585-588
  static int getDirectionNumber(int dx, int dy, int dz)
  {
    return (dx + 1) * 9 + (dy + 1) * 3 + dz + 1;
  }

  // propagation_distance_field.h:597-600
  double getDistance(const Voxel& v) const
  {
    return sqrt_table_[v.distance_square] - sqrt_table_[v.negative_distance_square];
  }
  // propagation_distance_field.cpp:600-603
  double getDistance(int x, int y, int z) const
  {
    return getDistance(cell(x, y, z));
  }
  // propagation_distance_field.cpp:595-598 with voxel_grid.h:462-470 (out of grid -> default object)
  double getDistance(double x, double y, double z) const
  {
    int cx, cy, cz;
    if (!g_.worldToGrid(x, y, z, cx, cy, cz))
      return getDistance(defaultVoxel());
    return getDistance(cell(cx, cy, cz));
  }
  // propagation_distance_field.h:372-375
  double getUninitializedDistance() const
  {
    return max_distance_;
  }

  // distance_field.cpp:62-86
  double getDistanceGradient(double x, double y, double z, double& gradient_x, double& gradient_y, double& gradient_z,
                             bool& in_bounds) const
  {
    int gx, gy, gz;
    g_.worldToGrid(x, y, z, gx, gy, gz);
    if (gx < 1 || gy < 1 || gz < 1 || gx >= g_.num_cells[0] - 1 || gy >= g_.num_cells[1] - 1 ||
        gz >= g_.num_cells[2] - 1)
    {
      gradient_x = 0.0;
      gradient_y = 0.0;
      gradient_z = 0.0;
      in_bounds = false;
      return getUninitializedDistance();
    }
    gradient_x = (getDistance(gx + 1, gy, gz) - getDistance(gx - 1, gy, gz)) * inv_twice_resolution_;
    gradient_y = (getDistance(gx, gy + 1, gz) - getDistance(gx, gy - 1, gz)) * inv_twice_resolution_;
    gradient_z = (getDistance(gx, gy, gz + 1) - getDistance(gx, gy, gz - 1)) * inv_twice_resolution_;
    in_bounds = true;
    return getDistance(gx, gy, gz);
  }

  // propagation_distance_field.h:411-431
  // returns 0: nullptr, 1: non-null; dist / pos filled as the reference does
  int getNearestCell(int x, int y, int z, double& dist, I3& pos) const
  {
    const Voxel* c = &cell(x, y, z);
    if (c->distance_square > 0)
    {
      dist = sqrt_table_[c->distance_square];
      pos = c->closest_point;
      const Voxel* ncell = &cell(pos.x, pos.y, pos.z);
      return ncell == c ? 0 : 1;
    }
    if (c->negative_distance_square > 0)
    {
      dist = -sqrt_table_[c->negative_distance_square];
      pos = c->closest_negative_point;
      const Voxel* ncell = &cell(pos.x, pos.y, pos.z);
      return ncell == c ? 0 : 1;
    }
    dist = 0.0;
    pos = I3{ x, y, z };
    return 0;
  }

  // Occupancy bit stream of writeToStream (propagation_distance_field.cpp:652-671), before zlib:
  // x-major, y, then z in groups of 8, bit zi set iff d^2 == 0.
  std::vector<uint8_t> packOccupancy() const
  {
    std::vector<uint8_t> out;
    for (int x = 0; x < g_.num_cells[0]; x++)
      for (int y = 0; y < g_.num_cells[1]; y++)
        for (int z = 0; z < g_.num_cells[2]; z += 8)
        {
          uint8_t bs = 0;
          int zv = std::min(8, g_.num_cells[2] - z);
          for (int zi = 0; zi < zv; zi++)
            if (cell(x, y, z + zi).distance_square == 0)
              bs |= (uint8_t)(1u << zi);
          out.push_back(bs);
        }
    return out;
  }
  // readFromStream voxel part (propagation_distance_field.cpp:733-759)
  void unpackOccupancy(const uint8_t* bytes, size_t n)
  {
    std::vector<I3> obs_points;
    size_t k = 0;
    for (int x = 0; x < g_.num_cells[0]; x++)
      for (int y = 0; y < g_.num_cells[1]; y++)
        for (int z = 0; z < g_.num_cells[2]; z += 8)
        {
          if (k >= n)
            return;
          uint8_t b = bytes[k++];
          int zv = std::min(8, g_.num_cells[2] - z);
          for (int zi = 0; zi < zv; zi++)
            if ((b >> zi) & 1)
              obs_points.push_back(I3{ x, y, z + zi });
        }
    addNewObstacleVoxels(obs_points);
  }

  // writeToStream (propagation_distance_field.cpp:636-673): seven "key: value" lines with the default ostream
  // formatting of a double (operator<<(double) == printf("%g"), six significant digits), then the occupancy bits
  // through boost::iostreams::zlib_compressor (default parameters: a plain zlib stream, restated with zlib's own
  // compress2 at the default level).  Written into a byte vector: this library carries a static libstdc++, whose
  // iostreams must not be mixed with the host process's.
  bool writeToStream(std::vector<uint8_t>& os) const
  {
    const char* keys[7] = { "resolution", "size_x", "size_y", "size_z", "origin_x", "origin_y", "origin_z" };
    const double vals[7] = { resolution_, size_[0], size_[1], size_[2], origin_[0], origin_[1], origin_[2] };
    for (int i = 0; i < 7; ++i)
    {
      char line[96];
      const int n = std::snprintf(line, sizeof(line), "%s: %g\n", keys[i], vals[i]);
      os.insert(os.end(), line, line + n);
    }
    const std::vector<uint8_t> bits = packOccupancy();
    uLongf n = compressBound((uLong)bits.size());
    std::vector<uint8_t> z(n);
    if (compress2(z.data(), &n, bits.data(), (uLong)bits.size(), Z_DEFAULT_COMPRESSION) != Z_OK)
      return false;
    os.insert(os.end(), z.begin(), z.begin() + n);
    return true;
  }

  // readFromStream (propagation_distance_field.cpp:675-761); max_distance_ and propagate_negative_ keep the values
  // the field was constructed with (:712).  operator>>(std::string) / operator>>(double) read whitespace-separated
  // tokens.
  bool readFromStream(const uint8_t* data, size_t size)
  {
    if (size == 0)
      return false;  // !is.good()
    size_t pos = 0;
    auto token = [&](std::string& out) {
      while (pos < size && std::isspace(data[pos]))
        ++pos;
      const size_t b = pos;
      while (pos < size && !std::isspace(data[pos]))
        ++pos;
      out.assign(reinterpret_cast<const char*>(data) + b, pos - b);
      return !out.empty();
    };
    const char* keys[7] = { "resolution:", "size_x:", "size_y:", "size_z:", "origin_x:", "origin_y:", "origin_z:" };
    double* dst[7] = { &resolution_, &size_[0], &size_[1], &size_[2], &origin_[0], &origin_[1], &origin_[2] };
    for (int i = 0; i < 7; ++i)
    {
      std::string temp;
      if (!token(temp) || temp != keys[i])
        return false;
      if (!token(temp))
        return false;
      char* end = nullptr;
      *dst[i] = std::strtod(temp.c_str(), &end);
      if (end == temp.c_str())
        return false;
    }
    inv_twice_resolution_ = 1.0 / (2.0 * resolution_);
    initialize();
    if (pos < size)
      ++pos;  // "this should be newline" (:717-719)
    const size_t want = (size_t)g_.num_cells[0] * g_.num_cells[1] * ((g_.num_cells[2] + 7) / 8);
    std::vector<uint8_t> bits(want);
    uLongf got = (uLongf)want;
    const int rc = uncompress(bits.data(), &got, data + pos, (uLong)(size - pos));
    if ((rc != Z_OK && rc != Z_BUF_ERROR) || got < want)
      return false;  // the reference stops with false when the decompressor runs dry (:739-742)
    unpackOccupancy(bits.data(), want);
    return true;
  }

  Voxel& cell(int x, int y, int z)
  {
    return cells_[g_.ref(x, y, z)];
  }
  const Voxel& cell(int x, int y, int z) const
  {
    return cells_[g_.ref(x, y, z)];
  }
  Voxel defaultVoxel() const  // PropDistanceFieldVoxel(max_distance_sq_, 0), propagation_distance_field.h:575-585
  {
    Voxel v;
    v.distance_square = max_distance_sq_;
    v.negative_distance_square = 0;
    v.closest_point = I3{ UNINITIALIZED, UNINITIALIZED, UNINITIALIZED };
    v.closest_negative_point = I3{ UNINITIALIZED, UNINITIALIZED, UNINITIALIZED };
    v.update_direction = UNINITIALIZED;            // left uninitialised in the reference
    v.negative_update_direction = UNINITIALIZED;   // left uninitialised in the reference
    return v;
  }

  GridGeom g_;
  std::vector<Voxel> cells_;
  bool propagate_negative_;
  double max_distance_;
  int max_distance_sq_;
  double size_[3], origin_[3], resolution_;
  int inv_twice_resolution_;
  std::vector<double> sqrt_table_;
  std::vector<std::vector<I3>> bucket_queue_, negative_bucket_queue_;
  std::vector<std::vector<std::vector<I3>>> neighborhoods_;
  std::vector<I3> direction_number_to_direction_;
};

// ------------------------------------------------------------------------------------------------
// Exact clamped squared EDT (what north_star asks the GPU to produce; F4).  d2[v] = min over occupied
// q of |v-q|^2, clamped to max_d2.  Separable windowed min-plus passes; `brute` = O(V*occ) check.
// occ is z-fastest like VoxelGrid::ref.
// ------------------------------------------------------------------------------------------------
static inline void edt_exact(const uint8_t* occ, int nx, int ny, int nz, int max_d2, int32_t* out)
{
  const int INF = 1 << 29;
  int R = 0;
  while (R * R < max_d2)
    ++R;
  std::vector<int> a((size_t)nx * ny * nz), b((size_t)nx * ny * nz);
  auto ref = [&](int x, int y, int z) { return ((size_t)x * ny + y) * nz + z; };
  // z pass: 1-D distance by two sweeps
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
    {
      int last = -INF;
      for (int z = 0; z < nz; ++z)
      {
        if (occ[ref(x, y, z)])
          last = z;
        long d = (long)z - last;
        a[ref(x, y, z)] = (d > R) ? INF : (int)(d * d);
      }
      last = INF;
      for (int z = nz - 1; z >= 0; --z)
      {
        if (occ[ref(x, y, z)])
          last = z;
        long d = (long)last - z;
        int v = (d > R) ? INF : (int)(d * d);
        if (v < a[ref(x, y, z)])
          a[ref(x, y, z)] = v;
      }
    }
  // y pass
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z)
      {
        int best = INF;
        for (int yy = std::max(0, y - R); yy <= std::min(ny - 1, y + R); ++yy)
        {
          int f = a[ref(x, yy, z)];
          if (f >= INF)
            continue;
          int v = f + (y - yy) * (y - yy);
          if (v < best)
            best = v;
        }
        b[ref(x, y, z)] = best;
      }
  // x pass + clamp
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z)
      {
        int best = INF;
        for (int xx = std::max(0, x - R); xx <= std::min(nx - 1, x + R); ++xx)
        {
          int f = b[ref(xx, y, z)];
          if (f >= INF)
            continue;
          int v = f + (x - xx) * (x - xx);
          if (v < best)
            best = v;
        }
        out[ref(x, y, z)] = best > max_d2 ? max_d2 : best;
      }
}

static inline void edt_brute(const uint8_t* occ, int nx, int ny, int nz, int max_d2, int32_t* out)
{
  std::vector<I3> pts;
  auto ref = [&](int x, int y, int z) { return ((size_t)x * ny + y) * nz + z; };
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z)
        if (occ[ref(x, y, z)])
          pts.push_back(I3{ x, y, z });
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z)
      {
        int best = max_d2;
        for (const I3& p : pts)
        {
          int d = (p.x - x) * (p.x - x) + (p.y - y) * (p.y - y) + (p.z - z) * (p.z - z);
          if (d < best)
            best = d;
        }
        out[ref(x, y, z)] = best;
      }
}

// ------------------------------------------------------------------------------------------------
// geometric_shapes bodies (third party, restated; see header) and findInternalPointsConvex
// ------------------------------------------------------------------------------------------------
enum ShapeType
{
  SHAPE_SPHERE = 1,
  SHAPE_CYLINDER = 2,
  SHAPE_BOX = 4,
};

struct BoundingSphere
{
  V3 center;
  double radius;
};
struct BoundingCylinder
{
  Iso pose;
  double radius, length;
};

struct Body
{
  int type;
  double dims[3];  // sphere: r ; cylinder: r, length ; box: x, y, z
  Iso pose;
  double scale = 1.0, padding = 0.0;
  // derived (updateInternalData)
  V3 center;
  double radiusU, radius2, length2, width2, height2, radiusB;
  V3 n0, n1, n2;  // pose columns

  void update()
  {
    center = V3{ pose.m[0][3], pose.m[1][3], pose.m[2][3] };
    n0 = V3{ pose.m[0][0], pose.m[1][0], pose.m[2][0] };
    n1 = V3{ pose.m[0][1], pose.m[1][1], pose.m[2][1] };
    n2 = V3{ pose.m[0][2], pose.m[1][2], pose.m[2][2] };
    if (type == SHAPE_SPHERE)
    {
      radiusU = dims[0] * scale + padding;
      radius2 = radiusU * radiusU;
    }
    else if (type == SHAPE_CYLINDER)
    {
      radiusU = dims[0] * scale + padding;
      radius2 = radiusU * radiusU;
      length2 = scale * dims[1] / 2.0 + padding;
      radiusB = std::sqrt(length2 * length2 + radius2);
    }
    else
    {
      double s2 = scale / 2.0;
      length2 = dims[0] * s2 + padding;
      width2 = dims[1] * s2 + padding;
      height2 = dims[2] * s2 + padding;
      radius2 = length2 * length2 + width2 * width2 + height2 * height2;
      radiusB = std::sqrt(radius2);
    }
  }

  bool containsPoint(const V3& p) const
  {
    if (type == SHAPE_SPHERE)
      return sqnorm(center - p) <= radius2;
    V3 v = p - center;
    if (type == SHAPE_CYLINDER)
    {
      double pH = dot(v, n2);
      if (std::fabs(pH) > length2)
        return false;
      double pB1 = dot(v, n0);
      double remaining = radius2 - pB1 * pB1;
      if (remaining < 0.0)
        return false;
      double pB2 = dot(v, n1);
      return pB2 * pB2 <= remaining;
    }
    double pL = dot(v, n0);
    if (std::fabs(pL) > length2)
      return false;
    double pW = dot(v, n1);
    if (std::fabs(pW) > width2)
      return false;
    double pH = dot(v, n2);
    if (std::fabs(pH) > height2)
      return false;
    return true;
  }

  void computeBoundingSphere(BoundingSphere& s) const
  {
    s.center = center;
    s.radius = (type == SHAPE_SPHERE) ? radiusU : radiusB;
  }

  void computeBoundingCylinder(BoundingCylinder& c) const
  {
    if (type == SHAPE_SPHERE)
    {
      c.pose = pose;
      c.radius = radiusU;
      c.length = 2.0 * radiusU;
    }
    else if (type == SHAPE_CYLINDER)
    {
      c.pose = pose;
      c.radius = radiusU;
      c.length = scale * dims[1] + 2.0 * padding;
    }
    else
    {
      // cylinder axis = longest box dimension; radius = diagonal of the other two half extents
      double a, b;
      const double ang = 90.0f * (M_PI / 180.0f);
      if (length2 > width2 && length2 > height2)
      {
        c.length = length2 * 2.0;
        a = width2;
        b = height2;
        Iso rot = Iso::identity();  // AngleAxis(90deg, UnitY)
        double cs = std::cos(ang), sn = std::sin(ang);
        rot.m[0][0] = cs;
        rot.m[0][2] = sn;
        rot.m[2][0] = -sn;
        rot.m[2][2] = cs;
        c.pose = mul(pose, rot);
      }
      else if (width2 > height2)
      {
        c.length = width2 * 2.0;
        a = height2;
        b = length2;
        Iso rot = Iso::identity();  // AngleAxis(90deg, UnitX)
        double cs = std::cos(ang), sn = std::sin(ang);
        rot.m[1][1] = cs;
        rot.m[1][2] = -sn;
        rot.m[2][1] = sn;
        rot.m[2][2] = cs;
        c.pose = mul(pose, rot);
      }
      else
      {
        c.length = height2 * 2.0;
        a = width2;
        b = length2;
        c.pose = pose;
      }
      c.radius = std::sqrt(a * a + b * b);
    }
  }
};

// bodies::mergeBoundingSpheres (geometric_shapes, third party)
static inline void mergeBoundingSpheres(const std::vector<BoundingSphere>& spheres, BoundingSphere& merged)
{
  if (spheres.empty())
  {
    merged.center = V3{ 0.0, 0.0, 0.0 };
    merged.radius = 0.0;
    return;
  }
  merged = spheres[0];
  for (size_t i = 1; i < spheres.size(); ++i)
  {
    if (spheres[i].radius <= 0.0)
      continue;
    V3 diff = spheres[i].center - merged.center;
    double d = norm(diff);
    if (d + merged.radius <= spheres[i].radius)
    {
      merged = spheres[i];
    }
    else if (d + spheres[i].radius > merged.radius)
    {
      V3 delta = merged.center - spheres[i].center;
      double new_radius = (norm(delta) + spheres[i].radius + merged.radius) / 2.0;
      double n = norm(delta);
      double k = new_radius - spheres[i].radius;
      merged.center = V3{ delta.x / n * k + spheres[i].center.x, delta.y / n * k + spheres[i].center.y,
                          delta.z / n * k + spheres[i].center.z };
      merged.radius = new_radius;
    }
  }
}

// distance_field/src/find_internal_points.cpp:39-64 — lattice coordinates are accumulated with += (Q19)
static inline void findInternalPointsConvex(const Body& body, double resolution, std::vector<V3>& points)
{
  BoundingSphere sphere;
  body.computeBoundingSphere(sphere);
  double xval_s = std::floor((sphere.center.x - sphere.radius - resolution) / resolution) * resolution;
  double yval_s = std::floor((sphere.center.y - sphere.radius - resolution) / resolution) * resolution;
  double zval_s = std::floor((sphere.center.z - sphere.radius - resolution) / resolution) * resolution;
  double xval_e = sphere.center.x + sphere.radius + resolution;
  double yval_e = sphere.center.y + sphere.radius + resolution;
  double zval_e = sphere.center.z + sphere.radius + resolution;
  V3 pt;
  for (pt.x = xval_s; pt.x <= xval_e; pt.x += resolution)
    for (pt.y = yval_s; pt.y <= yval_e; pt.y += resolution)
      for (pt.z = zval_s; pt.z <= zval_e; pt.z += resolution)
        if (body.containsPoint(pt))
          points.push_back(pt);
}

// distance_field.cpp:236-280 (getOcTreePoints) with the octomap iteration replaced by an explicit list
// of occupied leaves (centre xyz + edge size) that lie in the field's bounding box.
static inline void octreeLeavesToPoints(double resolution, const double* leaves /*[n][4]*/, int n,
                                        std::vector<V3>& points)
{
  for (int i = 0; i < n; ++i)
  {
    const double cx = leaves[4 * i + 0], cy = leaves[4 * i + 1], cz = leaves[4 * i + 2], size = leaves[4 * i + 3];
    if (size <= resolution)
    {
      points.push_back(V3{ cx, cy, cz });
    }
    else
    {
      double ceil_val = std::ceil(size / resolution) * resolution / 2.0;
      for (double x = cx - ceil_val; x <= cx + ceil_val; x += resolution)
        for (double y = cy - ceil_val; y <= cy + ceil_val; y += resolution)
          for (double z = cz - ceil_val; z <= cz + ceil_val; z += resolution)
            points.push_back(V3{ x, y, z });
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Collision spheres / body decomposition — collision_distance_field_types.cpp
// ------------------------------------------------------------------------------------------------
struct CollisionSphere
{
  V3 relative_vec;
  double radius;
};

// collision_distance_field_types.cpp:46-79
static inline std::vector<CollisionSphere> determineCollisionSpheres(const Body& body)
{
  std::vector<CollisionSphere> css;
  if (body.type == SHAPE_SPHERE)
  {
    BoundingSphere bsphere;
    body.computeBoundingSphere(bsphere);
    if (bsphere.radius > 0.0)
      css.push_back(CollisionSphere{ bsphere.center, bsphere.radius });
  }
  else
  {
    BoundingCylinder cyl;
    body.computeBoundingCylinder(cyl);
    if (cyl.radius > 0.0 && cyl.length > 0.0)
    {
      unsigned int num_points = std::ceil(cyl.length / (cyl.radius / 2.0));
      double spacing = cyl.length / ((num_points * 1.0) - 1.0);
      for (unsigned int i = 1; i < num_points - 1; i++)  // Q11: both end spheres dropped; n<=2 gives none
      {
        CollisionSphere cs{ apply(cyl.pose, V3{ 0, 0, (-cyl.length / 2.0) + i * spacing }), cyl.radius };
        css.push_back(cs);
      }
    }
  }
  return css;
}

// collision_distance_field_types.cpp:290-334 (BodyDecomposition::init)
struct BodyDecomposition
{
  std::vector<Body> bodies;
  std::vector<CollisionSphere> collision_spheres;
  std::vector<V3> relative_collision_points;
  BoundingSphere relative_bounding_sphere;

  void init(const std::vector<Body>& shapes_posed, double resolution, double padding)
  {
    bodies = shapes_posed;
    for (Body& b : bodies)
    {
      b.padding = padding;  // BodyVector::addBody(shape, pose, padding)
      b.update();
    }
    collision_spheres.clear();
    relative_collision_points.clear();
    for (const Body& b : bodies)
    {
      std::vector<CollisionSphere> body_spheres = determineCollisionSpheres(b);
      collision_spheres.insert(collision_spheres.end(), body_spheres.begin(), body_spheres.end());
      std::vector<V3> pts;
      findInternalPointsConvex(b, resolution, pts);
      relative_collision_points.insert(relative_collision_points.end(), pts.begin(), pts.end());
    }
    std::vector<BoundingSphere> bs(bodies.size());
    for (size_t i = 0; i < bodies.size(); ++i)
      bodies[i].computeBoundingSphere(bs[i]);
    mergeBoundingSpheres(bs, relative_bounding_sphere);
  }
};

// ------------------------------------------------------------------------------------------------
// Robot model slice + FK — robot_state.cpp:718-755 and the *JointModel::computeTransform functions
// ------------------------------------------------------------------------------------------------
enum JointType
{
  JOINT_FIXED = 0,
  JOINT_REVOLUTE = 1,
  JOINT_PRISMATIC = 2,
  JOINT_PLANAR = 3,
  JOINT_FLOATING = 4,
};

struct LinkDesc
{
  int parent;  // link index of the parent, -1 for the root (links are in DFS pre-order: parent < child,
               // robot_model.cpp:783-838)
  int joint_type;
  double axis[3];
  Iso origin;  // joint origin transform (LinkModel::getJointOriginTransform)
  int origin_is_identity;
  int var_index;   // first variable of the parent joint, -1 if none
  int mimic_var;   // variable this joint mimics, -1 if none (robot_state.h:1854-1863)
  double mimic_mult, mimic_offset;
};

struct LinkGeom
{
  int has_geometry = 0;
  std::vector<CollisionSphere> spheres;
  std::vector<V3> points;  // relative collision points (BodyDecomposition)
  BoundingSphere bsphere{ V3{ 0, 0, 0 }, 0.0 };
};

struct RobotModel
{
  std::vector<LinkDesc> links;
  std::vector<LinkGeom> geom;
  int nvar = 0;
  // revolute axis precompute, revolute_joint_model.cpp:66-75 (axis normalised first)
  struct AxisPre
  {
    double ax, ay, az, x2, y2, z2, xy, xz, yz;
  };
  std::vector<AxisPre> pre;

  void finalize()
  {
    pre.resize(links.size());
    for (size_t i = 0; i < links.size(); ++i)
    {
      const double* a = links[i].axis;
      AxisPre p{ a[0], a[1], a[2], 0, 0, 0, 0, 0, 0 };
      if (links[i].joint_type == JOINT_REVOLUTE)
      {
        double n = std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);  // Eigen normalized(): v / norm
        p.ax = a[0] / n;
        p.ay = a[1] / n;
        p.az = a[2] / n;
      }
      else if (links[i].joint_type == JOINT_PRISMATIC)
      {
        double n = std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);  // prismatic_joint_model.cpp setAxis
        p.ax = a[0] / n;
        p.ay = a[1] / n;
        p.az = a[2] / n;
      }
      p.x2 = p.ax * p.ax;
      p.y2 = p.ay * p.ay;
      p.z2 = p.az * p.az;
      p.xy = p.ax * p.ay;
      p.xz = p.ax * p.az;
      p.yz = p.ay * p.az;
      pre[i] = p;
    }
  }

  double jointValue(const LinkDesc& l, const double* q, int k) const
  {
    if (l.mimic_var >= 0)  // robot_state.h:1856-1861: position = factor * v + offset
      return l.mimic_mult * q[l.mimic_var + k] + l.mimic_offset;
    return q[l.var_index + k];
  }

  Iso jointTransform(int i, const double* q) const
  {
    const LinkDesc& l = links[i];
    const AxisPre& p = pre[i];
    Iso t = Iso::identity();
    switch (l.joint_type)
    {
      case JOINT_REVOLUTE:  // revolute_joint_model.cpp:222-259
      {
        const double v = jointValue(l, q, 0);
        const double c = std::cos(v);
        const double s = std::sin(v);
        const double tt = 1.0 - c;
        const double txy = tt * p.xy;
        const double txz = tt * p.xz;
        const double tyz = tt * p.yz;
        const double zs = p.az * s;
        const double ys = p.ay * s;
        const double xs = p.ax * s;
        t.m[0][0] = tt * p.x2 + c;
        t.m[1][0] = txy + zs;
        t.m[2][0] = txz - ys;
        t.m[0][1] = txy - zs;
        t.m[1][1] = tt * p.y2 + c;
        t.m[2][1] = tyz + xs;
        t.m[0][2] = txz + ys;
        t.m[1][2] = tyz - xs;
        t.m[2][2] = tt * p.z2 + c;
        break;
      }
      case JOINT_PRISMATIC:  // prismatic_joint_model.cpp:119-144
      {
        const double v = jointValue(l, q, 0);
        t.m[0][3] = p.ax * v;
        t.m[1][3] = p.ay * v;
        t.m[2][3] = p.az * v;
        break;
      }
      case JOINT_PLANAR:  // planar_joint_model.cpp:327-331: Translation(x,y,0) * AngleAxis(theta, UnitZ)
      {
        const double x = jointValue(l, q, 0), y = jointValue(l, q, 1), th = jointValue(l, q, 2);
        const double s = std::sin(th), c = std::cos(th);
        const double c1 = 1.0 - c;  // Eigen AngleAxis::toRotationMatrix with axis (0,0,1)
        t.m[0][0] = c1 * 0.0 * 0.0 + c;
        t.m[1][1] = c1 * 0.0 * 0.0 + c;
        t.m[2][2] = c1 * 1.0 * 1.0 + c;
        t.m[0][1] = c1 * 0.0 * 0.0 - s * 1.0;
        t.m[1][0] = c1 * 0.0 * 0.0 + s * 1.0;
        t.m[0][3] = x;
        t.m[1][3] = y;
        t.m[2][3] = 0.0;
        break;
      }
      case JOINT_FLOATING:  // floating_joint_model.cpp:224-229: Translation * Quaternion(w=q6,x=q3,y=q4,z=q5).normalized()
      {
        double qx = jointValue(l, q, 3), qy = jointValue(l, q, 4), qz = jointValue(l, q, 5), qw = jointValue(l, q, 6);
        double n = std::sqrt(((qx * qx + qy * qy) + qz * qz) + qw * qw);
        qx /= n;
        qy /= n;
        qz /= n;
        qw /= n;
        const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
        const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
        const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        t.m[0][0] = 1.0 - (tyy + tzz);
        t.m[0][1] = txy - twz;
        t.m[0][2] = txz + twy;
        t.m[1][0] = txy + twz;
        t.m[1][1] = 1.0 - (txx + tzz);
        t.m[1][2] = tyz - twx;
        t.m[2][0] = txz - twy;
        t.m[2][1] = tyz + twx;
        t.m[2][2] = 1.0 - (txx + tyy);
        t.m[0][3] = jointValue(l, q, 0);
        t.m[1][3] = jointValue(l, q, 1);
        t.m[2][3] = jointValue(l, q, 2);
        break;
      }
      default:  // fixed_joint_model.cpp:94-97
        break;
    }
    return t;
  }

  // robot_state.cpp:718-755 (updateLinkTransformsInternal)
  void fk(const double* q, Iso* out) const
  {
    for (size_t i = 0; i < links.size(); ++i)
    {
      const LinkDesc& l = links[i];
      if (l.parent >= 0)
      {
        if (l.joint_type == JOINT_FIXED)
          out[i] = mul(out[l.parent], l.origin);
        else if (l.origin_is_identity)
          out[i] = mul(out[l.parent], jointTransform(i, q));
        else
          out[i] = mul(mul(out[l.parent], l.origin), jointTransform(i, q));
      }
      else
      {
        if (l.origin_is_identity)
          out[i] = jointTransform(i, q);
        else
          out[i] = mul(l.origin, jointTransform(i, q));
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Sphere-vs-field tests — collision_distance_field_types.cpp:143-231, PosedDistanceField types.h:114-192
// ------------------------------------------------------------------------------------------------
enum CollisionType
{
  NONE = 0,
  SELF = 1,
  INTRA = 2,
  ENVIRONMENT = 3,
};

struct GradientInfo  // collision_distance_field_types.h:78-105
{
  double closest_distance = DBL_MAX;
  bool collision = false;
  std::vector<V3> sphere_locations;
  std::vector<double> distances;
  std::vector<V3> gradients;
  std::vector<int> types;
};

// collision_distance_field_types.cpp:206-231
static inline bool getCollisionSphereCollision(const PropagationDistanceField* df,
                                               const std::vector<CollisionSphere>& sphere_list,
                                               const std::vector<V3>& sphere_centers, double maximum_value,
                                               double tolerance)
{
  for (unsigned int i = 0; i < sphere_list.size(); i++)
  {
    V3 p = sphere_centers[i];
    V3 grad;
    bool in_bounds = true;
    double dist = df->getDistanceGradient(p.x, p.y, p.z, grad.x, grad.y, grad.z, in_bounds);
    if (!in_bounds && norm(grad) > 0)  // dead branch (Q4)
      return true;
    if ((maximum_value > dist) && (sphere_list[i].radius - dist > tolerance))
      return true;
  }
  return false;
}

// collision_distance_field_types.cpp:233-267
static inline bool getCollisionSphereCollision(const PropagationDistanceField* df,
                                               const std::vector<CollisionSphere>& sphere_list,
                                               const std::vector<V3>& sphere_centers, double maximum_value,
                                               double tolerance, unsigned int num_coll, std::vector<unsigned int>& colls)
{
  colls.clear();
  for (unsigned int i = 0; i < sphere_list.size(); i++)
  {
    V3 p = sphere_centers[i];
    V3 grad;
    bool in_bounds = true;
    double dist = df->getDistanceGradient(p.x, p.y, p.z, grad.x, grad.y, grad.z, in_bounds);
    if (!in_bounds && (norm(grad) > 0))
      return true;
    if (maximum_value > dist && (sphere_list[i].radius - dist > tolerance))
    {
      if (num_coll == 0)
        return true;
      colls.push_back(i);
      if (colls.size() >= num_coll)
        return true;
    }
  }
  return !colls.empty();
}

// collision_distance_field_types.cpp:143-204
static inline bool getCollisionSphereGradients(const PropagationDistanceField* df,
                                               const std::vector<CollisionSphere>& sphere_list,
                                               const std::vector<V3>& sphere_centers, GradientInfo& gradient, int type,
                                               double tolerance, bool subtract_radii, double maximum_value,
                                               bool stop_at_first_collision)
{
  const double EPSILON = 0.0001;  // collision_distance_field_types.cpp:44
  bool in_collision = false;
  for (unsigned int i = 0; i < sphere_list.size(); i++)
  {
    V3 p = sphere_centers[i];
    V3 grad;
    bool in_bounds;
    double dist = df->getDistanceGradient(p.x, p.y, p.z, grad.x, grad.y, grad.z, in_bounds);
    if (!in_bounds && norm(grad) > EPSILON)
      return true;
    if (dist < maximum_value)
    {
      if (subtract_radii)
      {
        dist -= sphere_list[i].radius;
        if ((dist < 0) && (-dist >= tolerance))
          in_collision = true;
      }
      else
      {
        if (sphere_list[i].radius - dist > tolerance)
          in_collision = true;
      }
      if (dist < gradient.closest_distance)
        gradient.closest_distance = dist;
      if (dist < gradient.distances[i])
      {
        gradient.types[i] = type;
        gradient.distances[i] = dist;
        gradient.gradients[i] = grad;
      }
    }
    if (stop_at_first_collision && in_collision)
      return true;
  }
  return in_collision;
}

// collision_distance_field_types.h:114-192
class PosedDistanceField : public PropagationDistanceField
{
public:
  PosedDistanceField(const V3& size, const V3& origin, double resolution, double max_distance, bool neg)
    : PropagationDistanceField(size.x, size.y, size.z, resolution, origin.x, origin.y, origin.z, max_distance, neg)
    , pose_(Iso::identity())
  {
  }
  void updatePose(const Iso& t)
  {
    pose_ = t;
  }
  // types.h:156-168 — the gradient is mapped with the FULL transform (rotation + translation, Q6)
  double getDistanceGradientPosed(double x, double y, double z, double& gradient_x, double& gradient_y,
                                  double& gradient_z, bool& in_bounds) const
  {
    V3 rel_pos = apply(inverse(pose_), V3{ x, y, z });
    double gx, gy, gz;
    double res = PropagationDistanceField::getDistanceGradient(rel_pos.x, rel_pos.y, rel_pos.z, gx, gy, gz, in_bounds);
    V3 grad = apply(pose_, V3{ gx, gy, gz });
    gradient_x = grad.x;
    gradient_y = grad.y;
    gradient_z = grad.z;
    return res;
  }
  // collision_distance_field_types.cpp:81-141 (Q7: returns at the first out-of-bounds sphere)
  bool getCollisionSphereGradients(const std::vector<CollisionSphere>& sphere_list,
                                   const std::vector<V3>& sphere_centers, GradientInfo& gradient, int type,
                                   double tolerance, bool subtract_radii, double maximum_value,
                                   bool stop_at_first_collision) const
  {
    bool in_collision = false;
    for (unsigned int i = 0; i < sphere_list.size(); i++)
    {
      V3 p = sphere_centers[i];
      V3 grad{ 0, 0, 0 };
      bool in_bounds;
      double dist = getDistanceGradientPosed(p.x, p.y, p.z, grad.x, grad.y, grad.z, in_bounds);
      if (!in_bounds && norm(grad) > 0)
        return true;
      if (dist < maximum_value)
      {
        if (subtract_radii)
        {
          dist -= sphere_list[i].radius;
          if ((dist < 0) && (-dist >= tolerance))
            in_collision = true;
          dist = std::abs(dist);
        }
        else
        {
          if (sphere_list[i].radius - dist > tolerance)
            in_collision = true;
        }
        if (dist < gradient.closest_distance)
          gradient.closest_distance = dist;
        if (dist < gradient.distances[i])
        {
          gradient.types[i] = type;
          gradient.distances[i] = dist;
          gradient.gradients[i] = grad;
        }
      }
      if (stop_at_first_collision && in_collision)
        return true;
    }
    return in_collision;
  }
  Iso pose_;
};

// ------------------------------------------------------------------------------------------------
// AllowedCollisionMatrix slice — collision_detection/src/collision_matrix.cpp:46-64,113-124
// ------------------------------------------------------------------------------------------------
enum AllowedType
{
  ACM_NEVER = 0,
  ACM_ALWAYS = 1,
  ACM_CONDITIONAL = 2,
};
struct ACM
{
  std::map<std::string, std::map<std::string, int>> entries;
  void setEntry(const std::string& a, const std::string& b, bool allowed)  // collision_matrix.cpp:140-145
  {
    int v = allowed ? ACM_ALWAYS : ACM_NEVER;
    entries[a][b] = entries[b][a] = v;
  }
  bool getEntry(const std::string& a, const std::string& b, int& t) const  // collision_matrix.cpp:113-124
  {
    auto it1 = entries.find(a);
    if (it1 == entries.end())
      return false;
    auto it2 = it1->second.find(b);
    if (it2 == it1->second.end())
      return false;
    t = it2->second;
    return true;
  }
  size_t getSize() const
  {
    return entries.size();
  }
};

// ------------------------------------------------------------------------------------------------
// CollisionEnvDistanceField slice — collision_env_distance_field.cpp
// ------------------------------------------------------------------------------------------------
struct DistanceFieldCacheEntry  // collision_common_distance_field.h:56-167 (subset; no attached bodies)
{
  std::vector<int> link_indices;  // group links, in robot-model order (getUpdatedLinkModelNames / all links for "")
  std::vector<bool> link_has_geometry;
  std::vector<bool> self_collision_enabled;
  std::vector<std::vector<bool>> intra_group_collision_enabled;
  std::shared_ptr<PropagationDistanceField> distance_field;  // the "self" field from non-group links
  // acm_ view used by getSelfProximityGradients (:388-404): entry present and != NEVER -> skip
  std::vector<std::vector<int>> acm_entry;  // -1 = no entry, else AllowedType; indexed by group position
  bool acm_nonempty = false;
};

struct GroupStateRepresentation
{
  std::vector<std::vector<V3>> sphere_centers;  // per group link
  std::vector<V3> bsphere_center;               // per group link
  std::vector<Iso> link_pose;                   // per group link (PosedDistanceField::pose_)
  std::vector<GradientInfo> gradients;
};

struct EnvParams  // collision_env_distance_field.h:48-54 defaults
{
  double size[3] = { 3.0, 3.0, 4.0 };
  double origin[3] = { 0.0, 0.0, 0.0 };  // CENTRE of the field (Q14)
  bool use_signed = false;
  double resolution = 0.02;
  double collision_tolerance = 0.0;
  double max_propagation_distance = 0.25;
};

class CollisionEnvDistanceField
{
public:
  CollisionEnvDistanceField(std::shared_ptr<const RobotModel> model, const EnvParams& p) : model_(model), p_(p)
  {
    // generateDistanceFieldCacheEntryWorld, collision_env_distance_field.cpp:1800-1816
    world_field_ = makeField();
  }

  std::shared_ptr<PropagationDistanceField> makeField() const  // :946-948, :1804-1806
  {
    return std::make_shared<PropagationDistanceField>(
        p_.size[0], p_.size[1], p_.size[2], p_.resolution, p_.origin[0] - 0.5 * p_.size[0],
        p_.origin[1] - 0.5 * p_.size[1], p_.origin[2] - 0.5 * p_.size[2], p_.max_propagation_distance, p_.use_signed);
  }

  // generateDistanceFieldCacheEntry :728-975 with the name-keyed ACM resolved to per-link tables.
  //   group_links : link indices of the group in model order
  //   acm_entry   : [L*L] over ALL model links: -1 no entry, else AllowedType (getEntry semantics, Q12);
  //                 nullptr = "acm is null" branch (:835-839)
  //   self_points_state : if non-null, joint values used to pose the non-group links whose points build the
  //                 self field (:913-973); nullptr = generate_distance_field false
  void setGroup(const std::vector<int>& group_links, const int* acm_entry, const double* self_points_state)
  {
    const int L = (int)model_->links.size();
    auto dfce = std::make_shared<DistanceFieldCacheEntry>();
    dfce->link_indices = group_links;
    const size_t n = group_links.size();
    dfce->self_collision_enabled.assign(n, true);
    dfce->intra_group_collision_enabled.resize(n);
    std::vector<bool> all_true(n, true), all_false(n, false);
    dfce->acm_nonempty = false;
    dfce->acm_entry.assign(n, std::vector<int>(n, -1));
    if (acm_entry)
    {
      for (int k = 0; k < L * L; ++k)
        if (acm_entry[k] >= 0)
          dfce->acm_nonempty = true;
      for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j)
          dfce->acm_entry[i][j] = acm_entry[group_links[i] * L + group_links[j]];
    }
    for (size_t i = 0; i < n; ++i)
    {
      const int li = group_links[i];
      if (model_->geom[li].has_geometry)
      {
        dfce->link_has_geometry.push_back(true);
        if (acm_entry)
        {
          int t = acm_entry[li * L + li];
          if (t == ACM_ALWAYS)
            dfce->self_collision_enabled[i] = false;
          dfce->intra_group_collision_enabled[i] = all_true;
          for (size_t j = i + 1; j < n; ++j)
          {
            int tj = acm_entry[li * L + group_links[j]];
            if (tj == ACM_ALWAYS)
              dfce->intra_group_collision_enabled[i][j] = false;
          }
        }
        else
        {
          dfce->self_collision_enabled[i] = true;
          dfce->intra_group_collision_enabled[i] = all_true;
        }
      }
      else
      {
        dfce->link_has_geometry.push_back(false);
        dfce->self_collision_enabled[i] = false;
        dfce->intra_group_collision_enabled[i] = all_false;
      }
    }
    if (self_points_state)
    {
      // :922-970 — every link with geometry that is NOT in the group's updated-link set
      std::vector<bool> in_group(L, false);
      for (int li : group_links)
        in_group[li] = true;
      std::vector<Iso> T(L);
      model_->fk(self_points_state, T.data());
      std::vector<V3> all_points;
      for (int li = 0; li < L; ++li)
      {
        if (!model_->geom[li].has_geometry || in_group[li])
          continue;
        for (const V3& rp : model_->geom[li].points)  // PosedBodyPointDecomposition::updatePose types.cpp:368-379
          all_points.push_back(apply(T[li], rp));
      }
      dfce->distance_field = makeField();
      dfce->distance_field->addPointsToField(all_points);
    }
    dfce_ = dfce;
    link_fields_.clear();
  }

  // getGroupStateRepresentation :1172-1268 / updateGroupStateRepresentationState :1118-1170
  void poseGroup(const double* q, GroupStateRepresentation& gsr) const
  {
    const size_t n = dfce_->link_indices.size();
    std::vector<Iso> T(model_->links.size());
    model_->fk(q, T.data());
    gsr.sphere_centers.resize(n);
    gsr.bsphere_center.resize(n);
    gsr.link_pose.resize(n);
    gsr.gradients.resize(n);
    for (size_t i = 0; i < n; ++i)
    {
      if (!dfce_->link_has_geometry[i])
        continue;
      const int li = dfce_->link_indices[i];
      const LinkGeom& g = model_->geom[li];
      // PosedBodySphereDecomposition::updatePose types.cpp:390-397
      gsr.bsphere_center[i] = apply(T[li], g.bsphere.center);
      gsr.sphere_centers[i].resize(g.spheres.size());
      for (size_t k = 0; k < g.spheres.size(); ++k)
        gsr.sphere_centers[i][k] = apply(T[li], g.spheres[k].relative_vec);
      gsr.link_pose[i] = T[li];
      GradientInfo& gi = gsr.gradients[i];  // :1128-1134
      gi.closest_distance = DBL_MAX;
      gi.collision = false;
      gi.types.assign(g.spheres.size(), NONE);
      gi.distances.assign(g.spheres.size(), DBL_MAX);
      gi.gradients.assign(g.spheres.size(), V3{ 0.0, 0.0, 0.0 });
      gi.sphere_locations = gsr.sphere_centers[i];
    }
  }

  // getEnvironmentCollisions :1580-1662, req.contacts == false
  bool getEnvironmentCollisions(const GroupStateRepresentation& gsr) const
  {
    for (size_t i = 0; i < dfce_->link_indices.size(); i++)
    {
      if (!dfce_->link_has_geometry[i])
        continue;
      const LinkGeom& g = model_->geom[dfce_->link_indices[i]];
      if (getCollisionSphereCollision(world_field_.get(), g.spheres, gsr.sphere_centers[i],
                                      p_.max_propagation_distance, p_.collision_tolerance))
        return true;
    }
    return false;
  }

  // getSelfCollisions :277-356, req.contacts == false
  bool getSelfCollisions(const GroupStateRepresentation& gsr) const
  {
    if (!dfce_->distance_field)
      return false;
    for (size_t i = 0; i < dfce_->link_indices.size(); i++)
    {
      if (!dfce_->link_has_geometry[i] || !dfce_->self_collision_enabled[i])
        continue;
      const LinkGeom& g = model_->geom[dfce_->link_indices[i]];
      if (getCollisionSphereCollision(dfce_->distance_field.get(), g.spheres, gsr.sphere_centers[i],
                                      p_.max_propagation_distance, p_.collision_tolerance))
        return true;
    }
    return false;
  }

  // doBoundingSpheresIntersect types.cpp:410-420 (Q2: squared distance vs un-squared radius sum)
  bool doBoundingSpheresIntersect(const GroupStateRepresentation& gsr, size_t i, size_t j) const
  {
    double p1_radius = model_->geom[dfce_->link_indices[i]].bsphere.radius;
    double p2_radius = model_->geom[dfce_->link_indices[j]].bsphere.radius;
    double dist = sqnorm(gsr.bsphere_center[i] - gsr.bsphere_center[j]);
    return dist < (p1_radius + p2_radius);
  }

  // getIntraGroupCollisions :433-660, req.contacts == false, no attached bodies
  bool getIntraGroupCollisions(const GroupStateRepresentation& gsr) const
  {
    const size_t num_links = dfce_->link_indices.size();
    for (size_t i = 0; i < num_links; i++)
      for (size_t j = i + 1; j < num_links; j++)
      {
        if (!dfce_->link_has_geometry[i] || !dfce_->link_has_geometry[j])
          continue;
        if (!dfce_->intra_group_collision_enabled[i][j])
          continue;
        if (!doBoundingSpheresIntersect(gsr, i, j))
          continue;
        const auto& s1 = model_->geom[dfce_->link_indices[i]].spheres;
        const auto& s2 = model_->geom[dfce_->link_indices[j]].spheres;
        for (size_t k = 0; k < s1.size(); k++)
          for (size_t l = 0; l < s2.size(); l++)
          {
            V3 gradient = gsr.sphere_centers[i][k] - gsr.sphere_centers[j][l];
            double dist = norm(gradient);
            if (dist < s1[k].radius + s2[l].radius)
              return true;
          }
      }
    return false;
  }

  // checkRobotCollision :1500-1519
  bool checkRobotCollision(const double* q) const
  {
    GroupStateRepresentation gsr;
    poseGroup(q, gsr);
    return getEnvironmentCollisions(gsr);
  }
  // checkSelfCollisionHelper :182-204
  bool checkSelfCollision(const double* q) const
  {
    GroupStateRepresentation gsr;
    poseGroup(q, gsr);
    bool done = getSelfCollisions(gsr);
    if (!done)
      done = getIntraGroupCollisions(gsr);
    return done;
  }
  // checkCollision :1441-1464 (self field, intra group, then environment)
  bool checkCollision(const double* q) const
  {
    GroupStateRepresentation gsr;
    poseGroup(q, gsr);
    bool done = getSelfCollisions(gsr);
    if (!done)
      done = getIntraGroupCollisions(gsr);
    if (!done)
      done = getEnvironmentCollisions(gsr);
    return done;
  }
  // PlanningScene::checkCollision order (planning_scene.cpp:492-503): robot-vs-world, then self
  bool checkSceneOrder(const double* q, bool robot, bool self) const
  {
    GroupStateRepresentation gsr;
    poseGroup(q, gsr);
    if (robot && getEnvironmentCollisions(gsr))
      return true;
    if (self)
    {
      if (getSelfCollisions(gsr))
        return true;
      if (getIntraGroupCollisions(gsr))
        return true;
    }
    return false;
  }

  // ---- req.contacts == true ------------------------------------------------------------------------------
  // One contact: the group-link index of body 1, body 2 (a group-link index, kSelf or kEnvironment), the sphere of
  // body 1 whose centre is Contact::pos, and that centre.  res.contacts is a std::map keyed by the name pair; the
  // caller re-keys by name, the insertion order inside a key is the order below.
  enum
  {
    kSelf = -1,
    kEnvironment = -2,
  };
  struct ContactRec
  {
    int link1, body2, sphere;
    V3 pos;
  };
  struct ContactResult
  {
    bool collision = false;
    size_t contact_count = 0;
    std::vector<ContactRec> contacts;
    size_t pairCount(int a, int b) const
    {
      size_t n = 0;
      for (const ContactRec& c : contacts)
        n += (c.link1 == a && c.body2 == b);
      return n;
    }
  };
  // getSelfCollisions :277-356 / getEnvironmentCollisions :1580-1662 with req.contacts
  bool fieldContacts(const PropagationDistanceField* df, const GroupStateRepresentation& gsr, bool self, int body2,
                     size_t max_contacts, size_t max_contacts_per_pair, ContactResult& res) const
  {
    for (size_t i = 0; i < dfce_->link_indices.size(); i++)
    {
      if (!dfce_->link_has_geometry[i] || (self && !dfce_->self_collision_enabled[i]))
        continue;
      const LinkGeom& g = model_->geom[dfce_->link_indices[i]];
      std::vector<unsigned int> colls;
      bool coll = getCollisionSphereCollision(df, g.spheres, gsr.sphere_centers[i], p_.max_propagation_distance,
                                              p_.collision_tolerance,
                                              (unsigned int)std::min(max_contacts_per_pair, max_contacts - res.contact_count),
                                              colls);
      if (coll)
      {
        res.collision = true;
        for (unsigned int col : colls)
        {
          res.contacts.push_back(ContactRec{ (int)i, body2, (int)col, gsr.sphere_centers[i][col] });
          res.contact_count++;
        }
        if (res.contact_count >= max_contacts)
          return true;
      }
    }
    return res.contact_count >= max_contacts;
  }
  // getIntraGroupCollisions :433-660 with req.contacts, no attached bodies
  bool intraContacts(const GroupStateRepresentation& gsr, size_t max_contacts, size_t max_contacts_per_pair,
                     ContactResult& res) const
  {
    const size_t num_links = dfce_->link_indices.size();
    for (size_t i = 0; i < num_links; i++)
      for (size_t j = i + 1; j < num_links; j++)
      {
        if (!dfce_->link_has_geometry[i] || !dfce_->link_has_geometry[j])
          continue;
        if (!dfce_->intra_group_collision_enabled[i][j])
          continue;
        if (!doBoundingSpheresIntersect(gsr, i, j))
          continue;
        int num_pair = (int)res.pairCount((int)i, (int)j);
        const auto& s1 = model_->geom[dfce_->link_indices[i]].spheres;
        const auto& s2 = model_->geom[dfce_->link_indices[j]].spheres;
        for (size_t k = 0; k < s1.size() && num_pair < (int)max_contacts_per_pair; k++)
          for (size_t l = 0; l < s2.size() && num_pair < (int)max_contacts_per_pair; l++)
          {
            V3 gradient = gsr.sphere_centers[i][k] - gsr.sphere_centers[j][l];
            double dist = norm(gradient);
            if (dist < s1[k].radius + s2[l].radius)
            {
              res.collision = true;
              res.contacts.push_back(ContactRec{ (int)i, (int)j, (int)k, gsr.sphere_centers[i][k] });
              res.contact_count++;
              num_pair++;
              if (res.contact_count >= max_contacts)
                return true;
            }
          }
      }
    return false;
  }
  // checkCollision :1441-1464 with req.contacts == true
  ContactResult checkCollisionContacts(const double* q, size_t max_contacts, size_t max_contacts_per_pair) const
  {
    GroupStateRepresentation gsr;
    poseGroup(q, gsr);
    ContactResult res;
    bool done = false;
    if (dfce_->distance_field)
      done = fieldContacts(dfce_->distance_field.get(), gsr, true, kSelf, max_contacts, max_contacts_per_pair, res);
    if (!done)
      done = intraContacts(gsr, max_contacts, max_contacts_per_pair, res);
    if (!done)
      done = fieldContacts(world_field_.get(), gsr, false, kEnvironment, max_contacts, max_contacts_per_pair, res);
    return res;
  }

  // per-link PosedDistanceField, getGroupStateRepresentation :1196-1213
  void buildLinkFields()
  {
    link_fields_.clear();
    for (size_t i = 0; i < dfce_->link_indices.size(); ++i)
    {
      if (!dfce_->link_has_geometry[i])
      {
        link_fields_.push_back(nullptr);
        continue;
      }
      const LinkGeom& g = model_->geom[dfce_->link_indices[i]];
      double diameter = 2 * g.bsphere.radius;
      V3 link_size{ diameter, diameter, diameter };
      V3 link_origin{ g.bsphere.center.x - 0.5 * link_size.x, g.bsphere.center.y - 0.5 * link_size.y,
                      g.bsphere.center.z - 0.5 * link_size.z };
      auto f = std::make_shared<PosedDistanceField>(link_size, link_origin, p_.resolution,
                                                    p_.max_propagation_distance, p_.use_signed);
      // link_bd->getCollisionPoints() at construction = identity-posed relative points (types.cpp:381-388)
      f->addPointsToField(g.points);
      link_fields_.push_back(f);
    }
  }

  // getSelfProximityGradients :358-431
  void getSelfProximityGradients(GroupStateRepresentation& gsr)
  {
    const size_t n = dfce_->link_indices.size();
    for (size_t i = 0; i < n; i++)
    {
      if (!dfce_->link_has_geometry[i] || !dfce_->self_collision_enabled[i])
        continue;
      const LinkGeom& g = model_->geom[dfce_->link_indices[i]];
      if (dfce_->acm_nonempty)
      {
        for (size_t j = 0; j < n; j++)
        {
          if (i == j)
            continue;
          int t = dfce_->acm_entry[i][j];
          if (t >= 0 && t != ACM_NEVER)
            continue;
          if (link_fields_.size() == n && link_fields_[j])
          {
            link_fields_[j]->updatePose(gsr.link_pose[j]);
            link_fields_[j]->getCollisionSphereGradients(g.spheres, gsr.sphere_centers[i], gsr.gradients[i], SELF,
                                                         p_.collision_tolerance, false, p_.max_propagation_distance,
                                                         false);
          }
        }
      }
      if (dfce_->distance_field)
        getCollisionSphereGradients(dfce_->distance_field.get(), g.spheres, gsr.sphere_centers[i], gsr.gradients[i],
                                    SELF, p_.collision_tolerance, false, p_.max_propagation_distance, false);
    }
  }

  // getIntraGroupProximityGradients :662-727
  void getIntraGroupProximityGradients(GroupStateRepresentation& gsr) const
  {
    const size_t num_links = dfce_->link_indices.size();
    for (size_t i = 0; i < num_links; i++)
      for (size_t j = i + 1; j < num_links; j++)
      {
        if (!dfce_->link_has_geometry[i] || !dfce_->link_has_geometry[j])
          continue;
        if (!dfce_->intra_group_collision_enabled[i][j])
          continue;
        const size_t n1 = gsr.sphere_centers[i].size(), n2 = gsr.sphere_centers[j].size();
        for (size_t k = 0; k < n1; k++)
          for (size_t l = 0; l < n2; l++)
          {
            V3 gradient = gsr.sphere_centers[i][k] - gsr.sphere_centers[j][l];
            double dist = norm(gradient);
            if (dist < gsr.gradients[i].distances[k])
            {
              gsr.gradients[i].distances[k] = dist;
              gsr.gradients[i].gradients[k] = gradient;
              gsr.gradients[i].types[k] = INTRA;
            }
            if (dist < gsr.gradients[j].distances[l])
            {
              gsr.gradients[j].distances[l] = dist;
              gsr.gradients[j].gradients[l] = V3{ -gradient.x, -gradient.y, -gradient.z };
              gsr.gradients[j].types[l] = INTRA;
            }
          }
      }
  }

  // getEnvironmentProximityGradients :1664-1700
  void getEnvironmentProximityGradients(GroupStateRepresentation& gsr) const
  {
    for (size_t i = 0; i < dfce_->link_indices.size(); i++)
    {
      if (!dfce_->link_has_geometry[i])
        continue;
      const LinkGeom& g = model_->geom[dfce_->link_indices[i]];
      getCollisionSphereGradients(world_field_.get(), g.spheres, gsr.sphere_centers[i], gsr.gradients[i], ENVIRONMENT,
                                  p_.collision_tolerance, false, p_.max_propagation_distance, false);
    }
  }

  // getCollisionGradients :1536-1557
  void getCollisionGradients(const double* q, GroupStateRepresentation& gsr, bool self, bool intra, bool env)
  {
    poseGroup(q, gsr);
    if (self)
      getSelfProximityGradients(gsr);
    if (intra)
      getIntraGroupProximityGradients(gsr);
    if (env)
      getEnvironmentProximityGradients(gsr);
  }

  std::shared_ptr<const RobotModel> model_;
  EnvParams p_;
  std::shared_ptr<PropagationDistanceField> world_field_;
  std::shared_ptr<DistanceFieldCacheEntry> dfce_;
  std::vector<std::shared_ptr<PosedDistanceField>> link_fields_;
};

}  // namespace orc
