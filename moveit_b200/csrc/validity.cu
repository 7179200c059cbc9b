// Batched state validity (kernels 1-4): forward kinematics, collision-sphere posing, sphere-vs-field lookups,
// ACM-pruned intra-group sphere pairs.  EXACT path: fp64 arithmetic in the reference's operation order.
//
// Reference being replaced, per state:
//   RobotState::updateLinkTransformsInternal           robot_state/src/robot_state.cpp:718-755
//   *JointModel::computeTransform                       robot_model/src/{revolute,prismatic,fixed,planar,floating}_joint_model.cpp
//   PosedBodySphereDecomposition::updatePose            collision_distance_field/src/collision_distance_field_types.cpp:390-408
//   getCollisionSphereCollision                         collision_distance_field_types.cpp:206-231
//   DistanceField::getDistanceGradient                  distance_field/src/distance_field.cpp:62-86
//   CollisionEnvDistanceField::getEnvironmentCollisions collision_env_distance_field.cpp:1580-1662
//   CollisionEnvDistanceField::getSelfCollisions        :277-356
//   CollisionEnvDistanceField::getIntraGroupCollisions  :433-660, doBoundingSpheresIntersect types.cpp:410-420
//
// Layout: one thread per state.  Model tables are read through uniform (warp-broadcast) loads; every per-state
// intermediate that has to survive the FK sweep (branch-parent transforms, posed sphere centres, posed bounding
// sphere centres) lives in shared memory as [item][component][thread], so a thread only ever touches its own
// column (conflict-free, no block synchronisation after the joint values are staged).
//
// Compiled with --fmad=false so that mul/add round like the reference's non-FMA x86-64 build.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

#include "b200df_internal.cuh"

using namespace b200df;

namespace b200df
{
struct DevLink
{
  int parent;
  int joint_type;
  int var;        // first variable read for this joint (already redirected to the mimicked joint)
  int is_mimic;
  double mimic_mult, mimic_off;
  double ax, ay, az, x2, y2, z2, xy, xz, yz;  // RevoluteJointModel::setAxis precompute
  double origin[12];
  int origin_is_identity;
  int parent_is_prev;  // parent == i-1: parent transform is still in registers
  int store_slot;      // >=0: save this link's transform for a later non-adjacent child
  int parent_slot;     // slot holding the parent's transform when !parent_is_prev
};

struct DevGroupLink  // per group link with geometry, in group order
{
  int link;  // model link index
  int sphere_begin, sphere_end;
  int self_enabled;
  double bs[4];  // relative bounding sphere centre + radius
};

struct DevPair  // enabled link pair (i < j), indices into the group-link table
{
  int gi, gj;
  double radius_sum;  // p1_radius + p2_radius of doBoundingSpheresIntersect
};

struct GroupDev
{
  const DevLink* links;
  int n_links;
  int n_vars;
  int n_slots;
  const int* link_gslot;  // [n_links] index into glinks or -1
  const DevGroupLink* glinks;
  int n_glinks;
  const double* sph;      // [S][4] rel xyz + radius
  const int* sph_thr;     // [S] unsigned-field verdict threshold: collide iff d2 < thr
  const int* sph_cls;     // [S] radius class
  int n_spheres;
  const DevPair* pairs;
  int n_pairs;
  const double* pair_thr;  // [U][U] : ||c1-c2|| < r1+r2  <=>  squaredNorm < pair_thr[cls1][cls2]
  int n_cls;
  double tol, max_prop;
};
}  // namespace b200df

struct b200df_model
{
  int device = 0;
  int n_links = 0, n_vars = 0;
  std::vector<DevLink> links;
  std::vector<int> has_geometry;
  std::vector<int> sphere_offset;
  std::vector<double> spheres;
  std::vector<double> bsphere;
  std::vector<int64_t> point_offset;
  std::vector<double> points;
  int n_slots = 0;
  DevLink* d_links = nullptr;
};

struct b200df_group
{
  const b200df_model* model = nullptr;
  int device = 0;
  b200df_group_desc desc{};
  std::vector<DevGroupLink> glinks;
  std::vector<int> link_gslot;
  std::vector<double> sph;
  std::vector<DevPair> pairs;
  int n_cls = 0;
  GroupDev dev{};
  std::vector<void*> allocations;
};

namespace
{
// ------------------------------------------------------------------------------------------------------------
// device maths (operation order = Eigen 3.3 fixed-size products as the reference evaluates them: left-to-right sums)
// ------------------------------------------------------------------------------------------------------------
// r = a.affine() * b.matrix() with implied last row (0,0,0,1).  Zero terms are dropped (x + a*0 == x for the
// finite values that occur here), the "+ a3*1" term is kept as "+ a3".
__device__ __forceinline__ void iso_mul(const double* a, const double* b, double* r)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const double a0 = a[4 * i], a1 = a[4 * i + 1], a2 = a[4 * i + 2], a3 = a[4 * i + 3];
    r[4 * i + 0] = (a0 * b[0] + a1 * b[4]) + a2 * b[8];
    r[4 * i + 1] = (a0 * b[1] + a1 * b[5]) + a2 * b[9];
    r[4 * i + 2] = (a0 * b[2] + a1 * b[6]) + a2 * b[10];
    r[4 * i + 3] = ((a0 * b[3] + a1 * b[7]) + a2 * b[11]) + a3;
  }
}

__device__ __forceinline__ void iso_apply(const double* t, double x, double y, double z, double& ox, double& oy,
                                          double& oz)
{
  ox = t[3] + ((t[0] * x + t[1] * y) + t[2] * z);
  oy = t[7] + ((t[4] * x + t[5] * y) + t[6] * z);
  oz = t[11] + ((t[8] * x + t[9] * y) + t[10] * z);
}

template <typename QT>
__device__ __forceinline__ double joint_value(const DevLink& l, const double* qs, int k, int stride)
{
  const double v = qs[(l.var + k) * stride];
  return l.is_mimic ? l.mimic_mult * v + l.mimic_off : v;
}

// *JointModel::computeTransform
template <typename QT>
__device__ __forceinline__ void joint_transform(const DevLink& l, const double* qs, int stride, double* t)
{
  t[0] = 1.0; t[1] = 0.0; t[2] = 0.0; t[3] = 0.0;
  t[4] = 0.0; t[5] = 1.0; t[6] = 0.0; t[7] = 0.0;
  t[8] = 0.0; t[9] = 0.0; t[10] = 1.0; t[11] = 0.0;
  switch (l.joint_type)
  {
    case B200DF_JOINT_REVOLUTE:
    {
      const double v = joint_value<QT>(l, qs, 0, stride);
      double s, c;
      sincos(v, &s, &c);
      const double tt = 1.0 - c;
      const double txy = tt * l.xy, txz = tt * l.xz, tyz = tt * l.yz;
      const double zs = l.az * s, ys = l.ay * s, xs = l.ax * s;
      t[0] = tt * l.x2 + c;
      t[4] = txy + zs;
      t[8] = txz - ys;
      t[1] = txy - zs;
      t[5] = tt * l.y2 + c;
      t[9] = tyz + xs;
      t[2] = txz + ys;
      t[6] = tyz - xs;
      t[10] = tt * l.z2 + c;
      break;
    }
    case B200DF_JOINT_PRISMATIC:
    {
      const double v = joint_value<QT>(l, qs, 0, stride);
      t[3] = l.ax * v;
      t[7] = l.ay * v;
      t[11] = l.az * v;
      break;
    }
    case B200DF_JOINT_PLANAR:
    {
      const double x = joint_value<QT>(l, qs, 0, stride), y = joint_value<QT>(l, qs, 1, stride),
                   th = joint_value<QT>(l, qs, 2, stride);
      double s, c;
      sincos(th, &s, &c);
      const double c1 = 1.0 - c;
      t[0] = c1 * 0.0 * 0.0 + c;
      t[5] = c1 * 0.0 * 0.0 + c;
      t[10] = c1 * 1.0 * 1.0 + c;
      t[1] = c1 * 0.0 * 0.0 - s * 1.0;
      t[4] = c1 * 0.0 * 0.0 + s * 1.0;
      t[3] = x;
      t[7] = y;
      t[11] = 0.0;
      break;
    }
    case B200DF_JOINT_FLOATING:
    {
      double qx = joint_value<QT>(l, qs, 3, stride), qy = joint_value<QT>(l, qs, 4, stride),
             qz = joint_value<QT>(l, qs, 5, stride), qw = joint_value<QT>(l, qs, 6, stride);
      const double n = sqrt(((qx * qx + qy * qy) + qz * qz) + qw * qw);
      qx /= n;
      qy /= n;
      qz /= n;
      qw /= n;
      const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
      const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
      const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
      const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
      t[0] = 1.0 - (tyy + tzz);
      t[1] = txy - twz;
      t[2] = txz + twy;
      t[4] = txy + twz;
      t[5] = 1.0 - (txx + tzz);
      t[6] = tyz - twx;
      t[8] = txz - twy;
      t[9] = tyz + twx;
      t[10] = 1.0 - (txx + tyy);
      t[3] = joint_value<QT>(l, qs, 0, stride);
      t[7] = joint_value<QT>(l, qs, 1, stride);
      t[11] = joint_value<QT>(l, qs, 2, stride);
      break;
    }
    default:
      break;
  }
}

// RobotState::updateLinkTransformsInternal for link i; cur holds link i-1 on entry and link i on exit
template <typename QT>
__device__ __forceinline__ void fk_step(const DevLink& l, const double* qs, double* slots, int stride, double* cur)
{
  double par[12];
  if (l.parent >= 0)
  {
    if (l.parent_is_prev)
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        par[k] = cur[k];
    }
    else
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        par[k] = slots[(l.parent_slot * 12 + k) * stride];
    }
  }
  if (l.joint_type == B200DF_JOINT_FIXED)
  {
    if (l.parent >= 0)
      iso_mul(par, l.origin, cur);
    else
    {
      // root with a fixed joint: origin identity -> J (= identity); else origin.affine() * identity
      double j[12];
      joint_transform<QT>(l, qs, stride, j);
      if (l.origin_is_identity)
      {
#pragma unroll
        for (int k = 0; k < 12; ++k)
          cur[k] = j[k];
      }
      else
        iso_mul(l.origin, j, cur);
    }
  }
  else
  {
    double j[12];
    joint_transform<QT>(l, qs, stride, j);
    if (l.parent >= 0)
    {
      if (l.origin_is_identity)
        iso_mul(par, j, cur);
      else
      {
        double tmp[12];
        iso_mul(par, l.origin, tmp);
        iso_mul(tmp, j, cur);
      }
    }
    else
    {
      if (l.origin_is_identity)
      {
#pragma unroll
        for (int k = 0; k < 12; ++k)
          cur[k] = j[k];
      }
      else
        iso_mul(l.origin, j, cur);
    }
  }
  if (l.store_slot >= 0)
  {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      slots[(l.store_slot * 12 + k) * stride] = cur[k];
  }
}

// getDistanceGradient's cell lookup without the gradient: returns false when the centre is outside the
// 1-cell-inset box (the reference then reports max_distance => never a collision, Q4)
__device__ __forceinline__ bool inset_cell(const GridDev& g, double x, double y, double z, long& idx)
{
  const double fx = floor((x - g.om[0]) * g.oo), fy = floor((y - g.om[1]) * g.oo), fz = floor((z - g.om[2]) * g.oo);
  if (!(fx >= 1.0 && fy >= 1.0 && fz >= 1.0 && fx < (double)(g.n[0] - 1) && fy < (double)(g.n[1] - 1) &&
        fz < (double)(g.n[2] - 1)))
    return false;
  idx = (long)fx * g.stride1 + (long)fy * g.stride2 + (long)fz;
  return true;
}

template <typename FT>
__device__ __forceinline__ double cell_dist(const FieldDev& f, long idx)
{
  double d = f.sqrt_table[static_cast<const FT*>(f.d2)[idx]];
  if (f.neg)
    d = d - f.sqrt_table[static_cast<const FT*>(f.neg)[idx]];
  else
    d = d - f.sqrt_table[0];
  return d;
}

// getCollisionSphereCollision's per-sphere predicate (collision_distance_field_types.cpp:213-227)
template <typename FT>
__device__ __forceinline__ bool sphere_collides(const FieldDev& f, double x, double y, double z, double radius,
                                                int thr, double tol, double max_prop)
{
  long idx;
  if (!inset_cell(f.g, x, y, z, idx))
    return false;
  if (!f.neg)
    return (int)static_cast<const FT*>(f.d2)[idx] < thr;  // host-resolved, exactly equivalent threshold
  const double dist = cell_dist<FT>(f, idx);
  return (max_prop > dist) && (radius - dist > tol);
}

enum OutMode
{
  OUT_VERDICT = 0,
  OUT_CENTERS = 1,
  OUT_TRANSFORMS = 2,
  OUT_CLEARANCE = 3,
};

// ------------------------------------------------------------------------------------------------------------
// The EXACT validity kernel
// ------------------------------------------------------------------------------------------------------------
template <typename QT, typename FT, int MODE>
__global__ void validity_exact_kernel(GroupDev g, FieldDev wf, FieldDev sf, const QT* __restrict__ q, long n,
                                      int flags, uint8_t* __restrict__ verdict, double* __restrict__ out_f64)
{
  extern __shared__ double smem[];
  const int B = blockDim.x;
  const int tid = threadIdx.x;
  const long base = (long)blockIdx.x * B;
  double* qs = smem;                               // [n_vars][B]
  double* slots = qs + (size_t)g.n_vars * B;       // [n_slots][12][B]
  double* cen = slots + (size_t)g.n_slots * 12 * B;  // [S][3][B]
  double* bsc = cen + (size_t)g.n_spheres * 3 * B;   // [n_glinks][3][B]

  // stage the block's joint values: coalesced read, transposed into [var][thread]
  {
    const long total = (long)B * g.n_vars;
    const long limit = (n - base) * g.n_vars;
    for (long e = tid; e < total; e += B)
    {
      const int s = (int)(e / g.n_vars), v = (int)(e % g.n_vars);
      qs[v * B + s] = (e < limit) ? (double)q[base * g.n_vars + e] : 0.0;
    }
  }
  __syncthreads();
  const long state = base + tid;
  const bool active = state < n;
  const double* my_q = qs + tid;
  double* my_slots = slots + tid;
  double* my_cen = cen + tid;
  double* my_bsc = bsc + tid;

  bool hit = false;
  double clearance = DBL_MAX;
  double cur[12];
  const bool want_pairs = (flags & B200DF_CHECK_INTRA) != 0 && g.n_pairs > 0;

  for (int i = 0; i < g.n_links; ++i)
  {
    const DevLink& l = g.links[i];
    fk_step<QT>(l, my_q, my_slots, B, cur);
    if (MODE == OUT_TRANSFORMS)
    {
      if (active)
      {
#pragma unroll
        for (int k = 0; k < 12; ++k)
          out_f64[(state * g.n_links + i) * 12 + k] = cur[k];
      }
      continue;
    }
    const int gs = g.link_gslot[i];
    if (gs < 0)
      continue;
    const DevGroupLink& gl = g.glinks[gs];
    // PosedBodySphereDecomposition::updatePose
    if (want_pairs)
    {
      double bx, by, bz;
      iso_apply(cur, gl.bs[0], gl.bs[1], gl.bs[2], bx, by, bz);
      my_bsc[(gs * 3 + 0) * B] = bx;
      my_bsc[(gs * 3 + 1) * B] = by;
      my_bsc[(gs * 3 + 2) * B] = bz;
    }
    for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
    {
      const double* rel = g.sph + 4 * s;
      double cx, cy, cz;
      iso_apply(cur, rel[0], rel[1], rel[2], cx, cy, cz);
      if (want_pairs)
      {
        my_cen[(s * 3 + 0) * B] = cx;
        my_cen[(s * 3 + 1) * B] = cy;
        my_cen[(s * 3 + 2) * B] = cz;
      }
      if (MODE == OUT_CENTERS)
      {
        if (active)
        {
          out_f64[(state * g.n_spheres + s) * 3 + 0] = cx;
          out_f64[(state * g.n_spheres + s) * 3 + 1] = cy;
          out_f64[(state * g.n_spheres + s) * 3 + 2] = cz;
        }
        continue;
      }
      if (MODE == OUT_CLEARANCE)
      {
        long idx;
        double d = wf.max_distance;
        if (inset_cell(wf.g, cx, cy, cz, idx))
          d = cell_dist<FT>(wf, idx);
        d = d - rel[3];
        if (d < clearance)
          clearance = d;
        continue;
      }
      if (flags & B200DF_CHECK_ROBOT)
        hit |= sphere_collides<FT>(wf, cx, cy, cz, rel[3], g.sph_thr[s], g.tol, g.max_prop);
      if ((flags & B200DF_CHECK_SELF) && gl.self_enabled)
        hit |= sphere_collides<FT>(sf, cx, cy, cz, rel[3], g.sph_thr[s], g.tol, g.max_prop);
    }
  }
  if (MODE == OUT_TRANSFORMS || MODE == OUT_CENTERS)
    return;
  if (MODE == OUT_CLEARANCE)
  {
    if (active)
      out_f64[state] = clearance;
    return;
  }

  // intra-group sphere pairs (kernel 4).  A warp stops as soon as every one of its states has a verdict.
  if (want_pairs)
  {
    for (int p = 0; p < g.n_pairs; ++p)
    {
      if (__all_sync(0xffffffffu, hit || !active))
        break;
      const DevPair pr = g.pairs[p];
      const double dx = my_bsc[(pr.gi * 3 + 0) * B] - my_bsc[(pr.gj * 3 + 0) * B];
      const double dy = my_bsc[(pr.gi * 3 + 1) * B] - my_bsc[(pr.gj * 3 + 1) * B];
      const double dz = my_bsc[(pr.gi * 3 + 2) * B] - my_bsc[(pr.gj * 3 + 2) * B];
      const double d2 = (dx * dx + dy * dy) + dz * dz;
      // doBoundingSpheresIntersect: squared distance against the UN-squared radius sum (Q2)
      if (hit || !(d2 < pr.radius_sum))
        continue;
      const DevGroupLink& gi = g.glinks[pr.gi];
      const DevGroupLink& gj = g.glinks[pr.gj];
      for (int k = gi.sphere_begin; k < gi.sphere_end; ++k)
      {
        const double ax = my_cen[(k * 3 + 0) * B], ay = my_cen[(k * 3 + 1) * B], az = my_cen[(k * 3 + 2) * B];
        const double* thr_row = g.pair_thr + (size_t)g.sph_cls[k] * g.n_cls;
        for (int m = gj.sphere_begin; m < gj.sphere_end; ++m)
        {
          const double ex = ax - my_cen[(m * 3 + 0) * B];
          const double ey = ay - my_cen[(m * 3 + 1) * B];
          const double ez = az - my_cen[(m * 3 + 2) * B];
          const double e2 = (ex * ex + ey * ey) + ez * ez;
          hit |= e2 < thr_row[g.sph_cls[m]];
        }
      }
    }
  }
  if (active)
    verdict[state] = hit ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------
// Gradient kernel (getCollisionGradients :1536-1557): self field, intra group, environment — in that order,
// each keeping the per-sphere minimum with strict '<' like the reference.
// ------------------------------------------------------------------------------------------------------------
template <typename FT>
__device__ __forceinline__ void field_gradient_update(const FieldDev& f, double x, double y, double z, double max_value,
                                                      int type, double& best, double& gx, double& gy, double& gz,
                                                      int& btype)
{
  // getCollisionSphereGradients (collision_distance_field_types.cpp:143-204), subtract_radii == false
  long idx;
  double dist, dgx = 0.0, dgy = 0.0, dgz = 0.0;
  if (!inset_cell(f.g, x, y, z, idx))
    dist = f.max_distance;  // in_bounds = false, gradient 0: the "out of bounds" return is dead (Q4)
  else
  {
    const double itr = (double)f.inv_twice_res;
    dgx = (cell_dist<FT>(f, idx + f.g.stride1) - cell_dist<FT>(f, idx - f.g.stride1)) * itr;
    dgy = (cell_dist<FT>(f, idx + f.g.stride2) - cell_dist<FT>(f, idx - f.g.stride2)) * itr;
    dgz = (cell_dist<FT>(f, idx + 1) - cell_dist<FT>(f, idx - 1)) * itr;
    dist = cell_dist<FT>(f, idx);
  }
  if (dist < max_value && dist < best)
  {
    best = dist;
    gx = dgx;
    gy = dgy;
    gz = dgz;
    btype = type;
  }
}

template <typename QT, typename FT>
__global__ void gradients_kernel(GroupDev g, FieldDev wf, FieldDev sf, const QT* __restrict__ q, long n, int flags,
                                 double* __restrict__ out_dist, double* __restrict__ out_grad,
                                 uint8_t* __restrict__ out_type, double* __restrict__ out_cen)
{
  extern __shared__ double smem[];
  const int B = blockDim.x;
  const int tid = threadIdx.x;
  const long base = (long)blockIdx.x * B;
  double* qs = smem;
  double* slots = qs + (size_t)g.n_vars * B;
  double* cen = slots + (size_t)g.n_slots * 12 * B;
  {
    const long total = (long)B * g.n_vars;
    const long limit = (n - base) * g.n_vars;
    for (long e = tid; e < total; e += B)
    {
      const int s = (int)(e / g.n_vars), v = (int)(e % g.n_vars);
      qs[v * B + s] = (e < limit) ? (double)q[base * g.n_vars + e] : 0.0;
    }
  }
  __syncthreads();
  const long state = base + tid;
  if (state >= n)
    return;
  const double* my_q = qs + tid;
  double* my_slots = slots + tid;
  double* my_cen = cen + tid;
  const int S = g.n_spheres;
  double* dist = out_dist + state * S;
  double* grad = out_grad + state * S * 3;
  uint8_t* type = out_type + state * S;

  double cur[12];
  for (int i = 0; i < g.n_links; ++i)
  {
    const DevLink& l = g.links[i];
    fk_step<QT>(l, my_q, my_slots, B, cur);
    const int gs = g.link_gslot[i];
    if (gs < 0)
      continue;
    const DevGroupLink& gl = g.glinks[gs];
    for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
    {
      const double* rel = g.sph + 4 * s;
      double cx, cy, cz;
      iso_apply(cur, rel[0], rel[1], rel[2], cx, cy, cz);
      my_cen[(s * 3 + 0) * B] = cx;
      my_cen[(s * 3 + 1) * B] = cy;
      my_cen[(s * 3 + 2) * B] = cz;
      if (out_cen)
      {
        out_cen[(state * S + s) * 3 + 0] = cx;
        out_cen[(state * S + s) * 3 + 1] = cy;
        out_cen[(state * S + s) * 3 + 2] = cz;
      }
      // updateGroupStateRepresentationState :1128-1133
      double best = DBL_MAX, gx = 0.0, gy = 0.0, gz = 0.0;
      int bt = B200DF_TYPE_NONE;
      // getSelfProximityGradients: self field term (:420-422)
      if ((flags & B200DF_CHECK_SELF) && gl.self_enabled)
        field_gradient_update<FT>(sf, cx, cy, cz, g.max_prop, B200DF_TYPE_SELF, best, gx, gy, gz, bt);
      dist[s] = best;
      grad[3 * s] = gx;
      grad[3 * s + 1] = gy;
      grad[3 * s + 2] = gz;
      type[s] = (uint8_t)bt;
    }
  }
  // getIntraGroupProximityGradients :662-727 — sequential over pairs, both spheres updated with strict '<'
  if (flags & B200DF_CHECK_INTRA)
  {
    for (int p = 0; p < g.n_pairs; ++p)
    {
      const DevPair pr = g.pairs[p];
      const DevGroupLink& gi = g.glinks[pr.gi];
      const DevGroupLink& gj = g.glinks[pr.gj];
      for (int k = gi.sphere_begin; k < gi.sphere_end; ++k)
      {
        const double ax = my_cen[(k * 3 + 0) * B], ay = my_cen[(k * 3 + 1) * B], az = my_cen[(k * 3 + 2) * B];
        double dk = dist[k];
        for (int m = gj.sphere_begin; m < gj.sphere_end; ++m)
        {
          const double ex = ax - my_cen[(m * 3 + 0) * B];
          const double ey = ay - my_cen[(m * 3 + 1) * B];
          const double ez = az - my_cen[(m * 3 + 2) * B];
          const double d = sqrt((ex * ex + ey * ey) + ez * ez);
          if (d < dk)
          {
            dk = d;
            dist[k] = d;
            grad[3 * k] = ex;
            grad[3 * k + 1] = ey;
            grad[3 * k + 2] = ez;
            type[k] = B200DF_TYPE_INTRA;
          }
          if (d < dist[m])
          {
            dist[m] = d;
            grad[3 * m] = -ex;
            grad[3 * m + 1] = -ey;
            grad[3 * m + 2] = -ez;
            type[m] = B200DF_TYPE_INTRA;
          }
        }
      }
    }
  }
  // getEnvironmentProximityGradients :1664-1700
  if (flags & B200DF_CHECK_ROBOT)
  {
    for (int s = 0; s < S; ++s)
    {
      double best = dist[s], gx = grad[3 * s], gy = grad[3 * s + 1], gz = grad[3 * s + 2];
      int bt = type[s];
      const double before = best;
      field_gradient_update<FT>(wf, my_cen[(s * 3 + 0) * B], my_cen[(s * 3 + 1) * B], my_cen[(s * 3 + 2) * B],
                                g.max_prop, B200DF_TYPE_ENVIRONMENT, best, gx, gy, gz, bt);
      if (best < before)
      {
        dist[s] = best;
        grad[3 * s] = gx;
        grad[3 * s + 1] = gy;
        grad[3 * s + 2] = gz;
        type[s] = (uint8_t)bt;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------------------
// smallest double T with: sqrt(x) < R  <=>  x < T  (sqrt is monotone and correctly rounded on both sides)
double sqrt_less_threshold(double R)
{
  if (!(R > 0.0))
    return 0.0;  // sqrt(x) < R never holds for x >= 0
  double x = R * R;
  while (x > 0.0 && std::sqrt(std::nextafter(x, 0.0)) >= R)
    x = std::nextafter(x, 0.0);
  while (std::sqrt(x) < R)
    x = std::nextafter(x, INFINITY);
  return x;
}

template <typename T>
int upload(b200df_group* grp, const std::vector<T>& host, const T** dev)
{
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(1, host.size()) * sizeof(T);
  B200DF_CUDA(cudaMalloc(&p, bytes));
  grp->allocations.push_back(p);
  if (!host.empty())
    B200DF_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev = static_cast<const T*>(p);
  return B200DF_OK;
}

struct ScratchBuffer
{
  void* p = nullptr;
  cudaStream_t stream = nullptr;
  ~ScratchBuffer()
  {
    if (p)
      cudaFreeAsync(p, stream);
  }
};

size_t smem_bytes(const GroupDev& g, int block, bool with_centers, bool with_bsc)
{
  size_t d = (size_t)g.n_vars + (size_t)g.n_slots * 12;
  if (with_centers)
    d += (size_t)g.n_spheres * 3;
  if (with_bsc)
    d += (size_t)g.n_glinks * 3;
  return d * block * sizeof(double);
}

int pick_block(const GroupDev& g, bool with_centers, bool with_bsc, int* block, size_t* bytes)
{
  // largest block (multiple of 32) whose shared-memory footprint still lets two CTAs share an SM, else one
  const size_t budget2 = 110 * 1024, budget1 = 220 * 1024;
  for (int b : { 128, 96, 64, 32 })
  {
    const size_t s = smem_bytes(g, b, with_centers, with_bsc);
    if (s <= budget2 || (b == 32 && s <= budget1))
    {
      *block = b;
      *bytes = s;
      return B200DF_OK;
    }
  }
  set_error("robot too large for the per-state shared-memory layout (%zu bytes for 32 states)",
            smem_bytes(g, 32, with_centers, with_bsc));
  return B200DF_ERR_UNSUPPORTED;
}

template <typename K>
int set_smem_limit(K kernel, size_t bytes)
{
  if (bytes > 48 * 1024)
    B200DF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B200DF_OK;
}

template <typename QT, typename FT, int MODE>
int launch_exact(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, const void* q, long n, int flags,
                 uint8_t* verdict, double* out_f64, cudaStream_t stream)
{
  const bool pairs = MODE == OUT_VERDICT && (flags & B200DF_CHECK_INTRA) && g.n_pairs > 0;
  int block;
  size_t bytes;
  int rc = pick_block(g, pairs, pairs, &block, &bytes);
  if (rc)
    return rc;
  rc = set_smem_limit(validity_exact_kernel<QT, FT, MODE>, bytes);
  if (rc)
    return rc;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  validity_exact_kernel<QT, FT, MODE>
      <<<grid, block, bytes, stream>>>(g, wf, sf, static_cast<const QT*>(q), n, flags, verdict, out_f64);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}

template <int MODE>
int dispatch_exact(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                   int q_dtype, long n, int flags, uint8_t* verdict, double* out_f64, cudaStream_t stream)
{
  if (q_dtype == B200DF_Q_F32)
  {
    if (field_elem == 1)
      return launch_exact<float, uint8_t, MODE>(g, wf, sf, q, n, flags, verdict, out_f64, stream);
    return launch_exact<float, uint16_t, MODE>(g, wf, sf, q, n, flags, verdict, out_f64, stream);
  }
  if (field_elem == 1)
    return launch_exact<double, uint8_t, MODE>(g, wf, sf, q, n, flags, verdict, out_f64, stream);
  return launch_exact<double, uint16_t, MODE>(g, wf, sf, q, n, flags, verdict, out_f64, stream);
}

int prepare_fields(const b200df_group* group, b200df_field* world, b200df_field* self, int flags, FieldDev* wf,
                   FieldDev* sf, int* elem)
{
  std::memset(wf, 0, sizeof(*wf));
  std::memset(sf, 0, sizeof(*sf));
  *elem = 1;
  bool have = false;
  auto check = [&](b200df_field* f, FieldDev* v) -> int {
    B200DF_REQUIRE(f->device == group->device, "field and group live on different devices");
    B200DF_REQUIRE(f->desc.resolution == group->desc.resolution, "field resolution differs from the group's");
    int rc = field_prepare(f, v, nullptr);
    if (rc)
      return rc;
    if (have)
      B200DF_REQUIRE(v->elem_size == *elem, "world and self field must share max_distance (voxel type)");
    *elem = v->elem_size;
    have = true;
    return B200DF_OK;
  };
  if (flags & B200DF_CHECK_ROBOT)
  {
    B200DF_REQUIRE(world, "B200DF_CHECK_ROBOT needs a world field");
    int rc = check(world, wf);
    if (rc)
      return rc;
  }
  if (flags & B200DF_CHECK_SELF)
  {
    B200DF_REQUIRE(self, "B200DF_CHECK_SELF needs a self field");
    int rc = check(self, sf);
    if (rc)
      return rc;
  }
  return B200DF_OK;
}

size_t q_elem_size(int q_dtype)
{
  return q_dtype == B200DF_Q_F32 ? sizeof(float) : sizeof(double);
}
}  // namespace

extern "C" {

int b200df_model_create(const b200df_model_desc* d, b200df_model** out)
{
  B200DF_REQUIRE(d && out, "null argument");
  B200DF_REQUIRE(d->n_links > 0 && d->n_vars >= 0, "empty model");
  B200DF_REQUIRE(d->parent && d->joint_type && d->axis && d->origin && d->origin_is_identity && d->var_index &&
                     d->mimic_var && d->mimic_mult && d->mimic_offset && d->has_geometry && d->sphere_offset &&
                     d->bsphere,
                 "null model array");
  int ndev = 0;
  B200DF_CUDA(cudaGetDeviceCount(&ndev));
  B200DF_REQUIRE(ndev > 0, "no CUDA device");
  std::unique_ptr<b200df_model> m(new b200df_model);
  m->device = current_device();
  const int L = d->n_links;
  m->n_links = L;
  m->n_vars = d->n_vars;
  m->links.resize(L);
  std::vector<int> needs_store(L, 0);
  for (int i = 0; i < L; ++i)
  {
    B200DF_REQUIRE(d->parent[i] < i && d->parent[i] >= -1, "links must be in DFS pre-order (parent < child)");
    if (d->parent[i] >= 0 && d->parent[i] != i - 1)
      needs_store[d->parent[i]] = 1;
  }
  std::vector<int> slot_of(L, -1);
  int n_slots = 0;
  for (int i = 0; i < L; ++i)
    if (needs_store[i])
      slot_of[i] = n_slots++;
  static const int nvars_of[5] = { 0, 1, 1, 3, 7 };
  for (int i = 0; i < L; ++i)
  {
    DevLink& l = m->links[i];
    std::memset(&l, 0, sizeof(l));
    l.parent = d->parent[i];
    l.joint_type = d->joint_type[i];
    B200DF_REQUIRE(l.joint_type >= 0 && l.joint_type <= 4, "unknown joint type");
    l.is_mimic = d->mimic_var[i] >= 0;
    l.var = l.is_mimic ? d->mimic_var[i] : d->var_index[i];
    if (nvars_of[l.joint_type] > 0)
      B200DF_REQUIRE(l.var >= 0 && l.var + nvars_of[l.joint_type] <= d->n_vars, "joint variable index out of range");
    l.mimic_mult = d->mimic_mult[i];
    l.mimic_off = d->mimic_offset[i];
    const double* a = d->axis + 3 * i;
    l.ax = a[0];
    l.ay = a[1];
    l.az = a[2];
    if (l.joint_type == B200DF_JOINT_REVOLUTE || l.joint_type == B200DF_JOINT_PRISMATIC)
    {
      const double nrm = std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);  // setAxis: axis.normalized()
      B200DF_REQUIRE(nrm > 0.0, "zero joint axis");
      l.ax = a[0] / nrm;
      l.ay = a[1] / nrm;
      l.az = a[2] / nrm;
    }
    l.x2 = l.ax * l.ax;
    l.y2 = l.ay * l.ay;
    l.z2 = l.az * l.az;
    l.xy = l.ax * l.ay;
    l.xz = l.ax * l.az;
    l.yz = l.ay * l.az;
    std::memcpy(l.origin, d->origin + 12 * i, sizeof(l.origin));
    l.origin_is_identity = d->origin_is_identity[i];
    l.parent_is_prev = (l.parent >= 0 && l.parent == i - 1);
    l.store_slot = slot_of[i];
    l.parent_slot = (l.parent >= 0 && !l.parent_is_prev) ? slot_of[l.parent] : -1;
  }
  m->n_slots = n_slots;
  m->has_geometry.assign(d->has_geometry, d->has_geometry + L);
  m->sphere_offset.assign(d->sphere_offset, d->sphere_offset + L + 1);
  const int S = m->sphere_offset[L];
  B200DF_REQUIRE(S == 0 || d->spheres, "null sphere array");
  m->spheres.assign(d->spheres, d->spheres + (size_t)S * 4);
  m->bsphere.assign(d->bsphere, d->bsphere + (size_t)L * 4);
  if (d->point_offset)
  {
    m->point_offset.assign(d->point_offset, d->point_offset + L + 1);
    const int64_t P = m->point_offset[L];
    B200DF_REQUIRE(P == 0 || d->points, "null point array");
    m->points.assign(d->points, d->points + (size_t)P * 3);
  }
  else
    m->point_offset.assign(L + 1, 0);
  DeviceGuard dg(m->device);
  B200DF_CUDA(cudaMalloc(&m->d_links, sizeof(DevLink) * L));
  B200DF_CUDA(cudaMemcpy(m->d_links, m->links.data(), sizeof(DevLink) * L, cudaMemcpyHostToDevice));
  *out = m.release();
  return B200DF_OK;
}

int b200df_model_destroy(b200df_model* m)
{
  if (!m)
    return B200DF_OK;
  DeviceGuard dg(m->device);
  cudaFree(m->d_links);
  delete m;
  return B200DF_OK;
}

int b200df_group_create(const b200df_model* model, const b200df_group_desc* d, b200df_group** out)
{
  B200DF_REQUIRE(model && d && out, "null argument");
  B200DF_REQUIRE(d->n_group_links >= 0 && (d->n_group_links == 0 || (d->group_links && d->self_enabled)),
                 "bad group arrays");
  B200DF_REQUIRE(d->resolution > 0.0, "resolution must be > 0");
  std::unique_ptr<b200df_group> g(new b200df_group);
  g->model = model;
  g->device = model->device;
  g->desc = *d;
  g->desc.group_links = nullptr;
  g->desc.self_enabled = nullptr;
  g->desc.intra_enabled = nullptr;
  const int n = d->n_group_links;
  const int L = model->n_links;
  g->link_gslot.assign(L, -1);
  std::vector<int> gslot_of_pos(n, -1);
  std::vector<int> sph_cls;
  std::vector<int> sph_thr;
  std::vector<double> radii;  // unique radii
  int prev_link = -1;
  for (int i = 0; i < n; ++i)
  {
    const int li = d->group_links[i];
    B200DF_REQUIRE(li >= 0 && li < L && li > prev_link, "group links must be ascending model link indices");
    prev_link = li;
    if (!model->has_geometry[li])
      continue;  // link_has_geometry_ false: skipped by every loop of the reference
    DevGroupLink gl;
    gl.link = li;
    gl.sphere_begin = (int)(g->sph.size() / 4);
    for (int s = model->sphere_offset[li]; s < model->sphere_offset[li + 1]; ++s)
    {
      const double* sp = &model->spheres[(size_t)s * 4];
      g->sph.insert(g->sph.end(), sp, sp + 4);
      const double r = sp[3];
      int cls = -1;
      for (size_t u = 0; u < radii.size(); ++u)
        if (radii[u] == r)
          cls = (int)u;
      if (cls < 0)
      {
        cls = (int)radii.size();
        radii.push_back(r);
      }
      sph_cls.push_back(cls);
      // verdict threshold for unsigned fields: number of leading d2 values for which
      // "max_prop > sqrt_table[d2] && r - sqrt_table[d2] > tol" holds (the predicate is monotone in d2)
      int thr = 0;
      while (thr < 65536)
      {
        const double dist = std::sqrt(double(thr)) * d->resolution - std::sqrt(0.0) * d->resolution;
        if ((d->max_propagation_distance > dist) && (r - dist > d->collision_tolerance))
          ++thr;
        else
          break;
      }
      sph_thr.push_back(thr);
    }
    gl.sphere_end = (int)(g->sph.size() / 4);
    gl.self_enabled = d->self_enabled[i] ? 1 : 0;
    std::memcpy(gl.bs, &model->bsphere[(size_t)li * 4], sizeof(gl.bs));
    gslot_of_pos[i] = (int)g->glinks.size();
    g->link_gslot[li] = (int)g->glinks.size();
    g->glinks.push_back(gl);
  }
  // getIntraGroupCollisions loop order: i ascending, j > i ascending
  if (d->intra_enabled)
  {
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j)
      {
        if (gslot_of_pos[i] < 0 || gslot_of_pos[j] < 0)
          continue;
        if (!d->intra_enabled[(size_t)i * n + j])
          continue;
        DevPair p;
        p.gi = gslot_of_pos[i];
        p.gj = gslot_of_pos[j];
        p.radius_sum = g->glinks[p.gi].bs[3] + g->glinks[p.gj].bs[3];
        g->pairs.push_back(p);
      }
  }
  const int U = (int)radii.size();
  g->n_cls = U;
  std::vector<double> pair_thr((size_t)U * U);
  for (int a = 0; a < U; ++a)
    for (int b = 0; b < U; ++b)
      pair_thr[(size_t)a * U + b] = sqrt_less_threshold(radii[a] + radii[b]);

  DeviceGuard dg(g->device);
  GroupDev& dev = g->dev;
  dev.links = model->d_links;
  dev.n_links = L;
  dev.n_vars = model->n_vars;
  dev.n_slots = model->n_slots;
  dev.n_glinks = (int)g->glinks.size();
  dev.n_spheres = (int)(g->sph.size() / 4);
  dev.n_pairs = (int)g->pairs.size();
  dev.n_cls = U;
  dev.tol = d->collision_tolerance;
  dev.max_prop = d->max_propagation_distance;
  int rc;
  if ((rc = upload(g.get(), g->link_gslot, &dev.link_gslot)))
    return rc;
  if ((rc = upload(g.get(), g->glinks, &dev.glinks)))
    return rc;
  if ((rc = upload(g.get(), g->sph, &dev.sph)))
    return rc;
  if ((rc = upload(g.get(), sph_thr, &dev.sph_thr)))
    return rc;
  if ((rc = upload(g.get(), sph_cls, &dev.sph_cls)))
    return rc;
  if ((rc = upload(g.get(), g->pairs, &dev.pairs)))
    return rc;
  if ((rc = upload(g.get(), pair_thr, &dev.pair_thr)))
    return rc;
  *out = g.release();
  return B200DF_OK;
}

int b200df_group_destroy(b200df_group* g)
{
  if (!g)
    return B200DF_OK;
  DeviceGuard dg(g->device);
  for (void* p : g->allocations)
    cudaFree(p);
  delete g;
  return B200DF_OK;
}

int b200df_group_sphere_count(const b200df_group* g, int32_t* n)
{
  B200DF_REQUIRE(g && n, "null argument");
  *n = g->dev.n_spheres;
  return B200DF_OK;
}

int b200df_check_batch_device(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q_dev,
                              int q_dtype, int64_t n, int flags, int precision, uint8_t* verdict_dev, void* stream)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q_dev && verdict_dev)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  B200DF_REQUIRE(precision == B200DF_PRECISION_EXACT || precision == B200DF_PRECISION_FAST, "unknown precision");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  int elem;
  int rc = prepare_fields(group, world, self, flags, &wf, &sf, &elem);
  if (rc)
    return rc;
  DeviceGuard dg(group->device);
  return dispatch_exact<OUT_VERDICT>(group->dev, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, nullptr,
                                     static_cast<cudaStream_t>(stream));
}

int b200df_check_batch(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q, int q_dtype,
                       int64_t n, int flags, int precision, uint8_t* verdict)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q && verdict)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  if (n == 0)
    return B200DF_OK;
  DeviceGuard dg(group->device);
  cudaStream_t stream;
  B200DF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t qbytes = (size_t)n * group->dev.n_vars * q_elem_size(q_dtype);
  int rc = B200DF_OK;
  {
    ScratchBuffer dq, dv;
    dq.stream = dv.stream = stream;
    cudaError_t e = cudaMallocAsync(&dq.p, std::max<size_t>(qbytes, 1), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dv.p, (size_t)n, stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dq.p, q, qbytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess)
    {
      set_error("staging the batch failed: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
    if (!rc)
      rc = b200df_check_batch_device(group, world, self, dq.p, q_dtype, n, flags, precision,
                                     static_cast<uint8_t*>(dv.p), stream);
    if (!rc)
    {
      e = cudaMemcpyAsync(verdict, dv.p, (size_t)n, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess)
      {
        set_error("reading the verdicts back failed: %s", cudaGetErrorString(e));
        rc = B200DF_ERR_CUDA;
      }
    }
  }
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
  return rc;
}

// shared host wrapper for the fp64-output modes
static int run_f64_output(const b200df_group* group, b200df_field* world, const void* q, int q_dtype, int64_t n,
                          int mode, size_t out_per_state, double* out)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q && out)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  std::memset(&wf, 0, sizeof(wf));
  std::memset(&sf, 0, sizeof(sf));
  int elem = 1;
  if (mode == OUT_CLEARANCE)
  {
    int rc = prepare_fields(group, world, nullptr, B200DF_CHECK_ROBOT, &wf, &sf, &elem);
    if (rc)
      return rc;
  }
  DeviceGuard dg(group->device);
  cudaStream_t stream;
  B200DF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t qbytes = (size_t)n * group->dev.n_vars * q_elem_size(q_dtype);
  const size_t obytes = (size_t)n * out_per_state * sizeof(double);
  int rc = B200DF_OK;
  {
    ScratchBuffer dq, dout;
    dq.stream = dout.stream = stream;
    cudaError_t e = cudaMallocAsync(&dq.p, std::max<size_t>(qbytes, 1), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dout.p, std::max<size_t>(obytes, 1), stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dq.p, q, qbytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess)
    {
      set_error("staging the batch failed: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
    if (!rc)
    {
      double* o = static_cast<double*>(dout.p);
      if (mode == OUT_CENTERS)
        rc = dispatch_exact<OUT_CENTERS>(group->dev, wf, sf, elem, dq.p, q_dtype, n, 0, nullptr, o, stream);
      else if (mode == OUT_TRANSFORMS)
        rc = dispatch_exact<OUT_TRANSFORMS>(group->dev, wf, sf, elem, dq.p, q_dtype, n, 0, nullptr, o, stream);
      else
        rc = dispatch_exact<OUT_CLEARANCE>(group->dev, wf, sf, elem, dq.p, q_dtype, n, 0, nullptr, o, stream);
    }
    if (!rc)
    {
      e = cudaMemcpyAsync(out, dout.p, obytes, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess)
      {
        set_error("reading results back failed: %s", cudaGetErrorString(e));
        rc = B200DF_ERR_CUDA;
      }
    }
  }
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
  return rc;
}

int b200df_sphere_centers(const b200df_group* group, const void* q, int q_dtype, int64_t n, double* centers)
{
  B200DF_REQUIRE(group, "null group");
  return run_f64_output(group, nullptr, q, q_dtype, n, OUT_CENTERS, (size_t)group->dev.n_spheres * 3, centers);
}

int b200df_distance_batch(const b200df_group* group, b200df_field* world, const void* q, int q_dtype, int64_t n,
                          double* min_clearance)
{
  B200DF_REQUIRE(group && world, "null argument");
  return run_f64_output(group, world, q, q_dtype, n, OUT_CLEARANCE, 1, min_clearance);
}

int b200df_model_fk(const b200df_model* model, const void* q, int q_dtype, int64_t n, double* link_transforms)
{
  B200DF_REQUIRE(model, "null model");
  // an empty group over the model gives the kernel its link table
  b200df_group_desc gd{};
  gd.n_group_links = 0;
  gd.resolution = 1.0;
  b200df_group* g = nullptr;
  int rc = b200df_group_create(model, &gd, &g);
  if (rc)
    return rc;
  rc = run_f64_output(g, nullptr, q, q_dtype, n, OUT_TRANSFORMS, (size_t)model->n_links * 12, link_transforms);
  b200df_group_destroy(g);
  return rc;
}

int b200df_gradients_batch(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                           int q_dtype, int64_t n, int flags, double* dist, double* grad, uint8_t* type,
                           double* centers)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q && dist && grad && type)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  int elem;
  int rc = prepare_fields(group, world, self, flags, &wf, &sf, &elem);
  if (rc)
    return rc;
  const GroupDev& g = group->dev;
  const size_t S = (size_t)g.n_spheres;
  DeviceGuard dg(group->device);
  cudaStream_t stream;
  B200DF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const size_t qbytes = (size_t)n * g.n_vars * q_elem_size(q_dtype);
  {
    ScratchBuffer dq, dd, dgr, dt, dc;
    dq.stream = dd.stream = dgr.stream = dt.stream = dc.stream = stream;
    cudaError_t e = cudaMallocAsync(&dq.p, std::max<size_t>(qbytes, 1), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dd.p, std::max<size_t>(n * S * sizeof(double), 1), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dgr.p, std::max<size_t>(n * S * 3 * sizeof(double), 1), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dt.p, std::max<size_t>(n * S, 1), stream);
    if (e == cudaSuccess && centers)
      e = cudaMallocAsync(&dc.p, std::max<size_t>(n * S * 3 * sizeof(double), 1), stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dq.p, q, qbytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess)
    {
      set_error("staging the batch failed: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
    if (!rc)
    {
      int block;
      size_t bytes;
      rc = pick_block(g, true, false, &block, &bytes);
      const unsigned grid = (unsigned)((n + block - 1) / block);
#define B200DF_LAUNCH_GRAD(QT, FT)                                                                                 \
  do                                                                                                               \
  {                                                                                                                \
    if (!rc)                                                                                                       \
      rc = set_smem_limit(gradients_kernel<QT, FT>, bytes);                                                        \
    if (!rc)                                                                                                       \
      gradients_kernel<QT, FT><<<grid, block, bytes, stream>>>(                                                    \
          g, wf, sf, static_cast<const QT*>(dq.p), n, flags, static_cast<double*>(dd.p),                           \
          static_cast<double*>(dgr.p), static_cast<uint8_t*>(dt.p), static_cast<double*>(dc.p));                   \
  } while (0)
      if (q_dtype == B200DF_Q_F32 && elem == 1)
        B200DF_LAUNCH_GRAD(float, uint8_t);
      else if (q_dtype == B200DF_Q_F32)
        B200DF_LAUNCH_GRAD(float, uint16_t);
      else if (elem == 1)
        B200DF_LAUNCH_GRAD(double, uint8_t);
      else
        B200DF_LAUNCH_GRAD(double, uint16_t);
#undef B200DF_LAUNCH_GRAD
      if (!rc)
      {
        B200DF_LAUNCHED();
        e = cudaGetLastError();
        if (e != cudaSuccess)
        {
          set_error("gradient kernel launch failed: %s", cudaGetErrorString(e));
          rc = B200DF_ERR_CUDA;
        }
      }
    }
    if (!rc)
    {
      e = cudaMemcpyAsync(dist, dd.p, n * S * sizeof(double), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(grad, dgr.p, n * S * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(type, dt.p, n * S, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess && centers)
        e = cudaMemcpyAsync(centers, dc.p, n * S * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess)
      {
        set_error("reading gradients back failed: %s", cudaGetErrorString(e));
        rc = B200DF_ERR_CUDA;
      }
    }
  }
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
  return rc;
}

}  // extern "C"
