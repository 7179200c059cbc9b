// Device-side model tables and the FK / sphere-vs-field primitives shared by the validity and gradient kernels.
//
// Reference being replaced (paths relative to /root/reference/moveit_core):
//   RobotState::updateLinkTransformsInternal   robot_state/src/robot_state.cpp:718-755
//   *JointModel::computeTransform               robot_model/src/{revolute,prismatic,fixed,planar,floating}_joint_model.cpp
//   DistanceField::getDistanceGradient          distance_field/src/distance_field.cpp:62-86
//   getCollisionSphereCollision                 collision_distance_field/src/collision_distance_field_types.cpp:206-231
//
// Everything here is fp64 in the reference's operation order; the including translation units are compiled with
// --fmad=false so mul/add round like the reference's non-FMA x86-64 build.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <map>
#include <atomic>
#include <thread>
#include <utility>
#include <vector>

#include "b200df_internal.cuh"

namespace b200df
{
struct DevLink
{
  int parent;
  int joint_type;
  int var;  // first variable read for this joint (already redirected to the mimicked joint)
  int is_mimic;
  double mimic_mult, mimic_off;
  double ax, ay, az, x2, y2, z2, xy, xz, yz;  // RevoluteJointModel::setAxis precompute
  double origin[12];
  int origin_is_identity;
  int parent_is_prev;  // parent == i-1: the parent transform is still in registers
  int store_slot;      // >=0: save this link's transform for a later non-adjacent child
  int parent_slot;     // slot holding the parent's transform when !parent_is_prev
  int jkind;           // revolute joints: 1/2/3 = axis exactly +-x/+-y/+-z (sparse joint matrix), 0 = general
};

// One group link with geometry, in group order.  Links that are the lower-index side of an enabled pair are
// "stored": their posed sphere centres stay in shared memory until the higher-index partner is swept.
struct DevGroupLink
{
  int link;  // model link index
  int sphere_begin, sphere_end;
  int self_enabled;
  double bs[4];  // relative bounding sphere centre + radius
  int store_link;    // index among stored links, -1 if not stored
  int store_sphere;  // first stored-sphere slot, -1 if not stored
  int partner_begin, partner_end;  // lower-index partners of this link (range in GroupDev::partners)
  int pair_const_begin;            // offset of this link's first sphere in the pair-constant arrays
  int pair_const_stride;           // pair constants per sphere of this link (sum of the partners' sphere counts)
};

struct DevPartner
{
  int store_link;    // stored-link index of the partner (bounding-sphere centre slot)
  int store_sphere;  // first stored-sphere slot of the partner
  int count;         // its sphere count
  int const_offset;  // offset of this partner's thresholds inside one sphere's stride of pair_T
  double radius_sum;  // p1_radius + p2_radius of doBoundingSpheresIntersect (compared with the SQUARED distance, Q2)
  double tight2;      // (R'_i + R'_j)^2, R' = radius around the same centre that really encloses the link's spheres:
                      // a sphere pair can only touch when the squared centre distance is below it (exact cull)
};

struct DevPair  // enabled link pair (gi < gj), indices into the group-link table (gradient kernel order)
{
  int gi, gj;
  double radius_sum;
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
  const double* sph;   // [S][4] rel xyz + radius
  const int* sph_thr;  // [S] unsigned-field verdict threshold: collide iff d2 < thr
  int n_spheres;
  int n_store_links, n_store_spheres;
  const DevPartner* partners;
  int n_partners;
  const double* pair_T;  // ||c1-c2|| < r1+r2  <=>  squaredNorm < pair_T; [own sphere][partner][partner sphere]
  long n_pair_consts;
  // tables of the generic sweep (gradient kernel, FK / centre / clearance outputs)
  const int* sph_cls;      // [S] radius class
  const double* pair_thr;  // [U][U] thresholds by radius class
  int n_cls;
  const DevPair* pairs;  // link pairs in the reference's (i, j>i) loop order
  int n_pairs;
  // per-sphere gather form of the intra-group pairs (gradient kernel): for group link gs, the enabled partner links
  // in the order the reference's pair loop touches gs (lower-index partners ascending, then higher-index ascending)
  const int* sph_glink;  // [S] group-link slot of each sphere
  const int* nbr_off;    // [n_glinks + 1]
  int n_nbr;             // nbr_off[n_glinks]
  const int4* nbr;       // {partner group-link slot, 1 if the partner is the lower-index (first) link of the pair,
                         //  the partner's sphere range [begin, end)}
  // per-link PosedDistanceField term of getSelfProximityGradients (:395-418), set by b200df_group_set_link_fields
  const FieldDev* link_fields;   // [n_glinks] views (d2 == nullptr: no field); all in their link's own frame
  const unsigned* lf_enabled;    // [n_glinks] bit j: the ACM lets link j's field act on this link's spheres
  int n_link_fields;             // 0: term off
  double tol, max_prop;
};

// ------------------------------------------------------------------------------------------------------------
// maths: Eigen 3.3 fixed-size products as the reference evaluates them (left-to-right sums, no FMA)
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

// Isometry3d * Vector3d: t + ((r0*x + r1*y) + r2*z)   (collision_distance_field_types.cpp:393-396)
__device__ __forceinline__ void iso_apply(const double* t, double x, double y, double z, double& ox, double& oy,
                                          double& oz)
{
  ox = t[3] + ((t[0] * x + t[1] * y) + t[2] * z);
  oy = t[7] + ((t[4] * x + t[5] * y) + t[6] * z);
  oz = t[11] + ((t[8] * x + t[9] * y) + t[10] * z);
}

template <typename QT>
__device__ __forceinline__ double joint_value(const DevLink& l, const QT* qs, int k, int stride)
{
  const double v = (double)qs[(l.var + k) * stride];
  return l.is_mimic ? l.mimic_mult * v + l.mimic_off : v;  // robot_state.h:1856-1861
}

// *JointModel::computeTransform
template <typename QT>
__device__ __forceinline__ void joint_transform(const DevLink& l, const QT* qs, int stride, double* t)
{
  t[0] = 1.0; t[1] = 0.0; t[2] = 0.0; t[3] = 0.0;
  t[4] = 0.0; t[5] = 1.0; t[6] = 0.0; t[7] = 0.0;
  t[8] = 0.0; t[9] = 0.0; t[10] = 1.0; t[11] = 0.0;
  switch (l.joint_type)
  {
    case B200DF_JOINT_REVOLUTE:  // revolute_joint_model.cpp:222-259
    {
      const double v = joint_value(l, qs, 0, stride);
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
    case B200DF_JOINT_PRISMATIC:  // prismatic_joint_model.cpp:119-144
    {
      const double v = joint_value(l, qs, 0, stride);
      t[3] = l.ax * v;
      t[7] = l.ay * v;
      t[11] = l.az * v;
      break;
    }
    case B200DF_JOINT_PLANAR:  // planar_joint_model.cpp:327-331 (Translation * AngleAxis(theta, UnitZ))
    {
      const double x = joint_value(l, qs, 0, stride), y = joint_value(l, qs, 1, stride),
                   th = joint_value(l, qs, 2, stride);
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
    case B200DF_JOINT_FLOATING:  // floating_joint_model.cpp:224-229
    {
      double qx = joint_value(l, qs, 3, stride), qy = joint_value(l, qs, 4, stride),
             qz = joint_value(l, qs, 5, stride), qw = joint_value(l, qs, 6, stride);
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
      t[3] = joint_value(l, qs, 0, stride);
      t[7] = joint_value(l, qs, 1, stride);
      t[11] = joint_value(l, qs, 2, stride);
      break;
    }
    default:  // fixed_joint_model.cpp:94-97
      break;
  }
}

// RobotState::updateLinkTransformsInternal for link i; cur holds link i-1 on entry and link i on exit.
// slots: per-thread column of saved branch-parent transforms, [slot*12 + k] * stride.
template <typename QT>
__device__ __forceinline__ void fk_step(const DevLink& l, const QT* qs, double* slots, int qstride, int stride,
                                        double* cur)
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
  if (l.joint_type == B200DF_JOINT_FIXED && l.parent >= 0)
  {
    iso_mul(par, l.origin, cur);  // robot_state.cpp:727-729
  }
  else
  {
    double j[12];
    joint_transform(l, qs, qstride, j);
    if (l.parent >= 0)
    {
      if (l.origin_is_identity)
        iso_mul(par, j, cur);  // :732-734
      else
      {
        double tmp[12];
        iso_mul(par, l.origin, tmp);  // :735-738, evaluated left to right
        iso_mul(tmp, j, cur);
      }
    }
    else
    {
      if (l.origin_is_identity)  // :743-744
      {
#pragma unroll
        for (int k = 0; k < 12; ++k)
          cur[k] = j[k];
      }
      else
        iso_mul(l.origin, j, cur);  // :745-747
    }
  }
  if (l.store_slot >= 0)
  {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      slots[(l.store_slot * 12 + k) * stride] = cur[k];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Sparse variant of fk_step for the verdict kernel.  A revolute joint about a coordinate axis has a joint matrix
// with structural zeros; the dropped terms are exact zeros, so every entry has the same VALUE as the full product
// (only the sign of a zero result can differ, which no later mul/add/compare/floor can observe).
// ------------------------------------------------------------------------------------------------------------
// r = a * J for J = rotation about +-x / +-y / +-z (kind 1/2/3); j = the 9 rotation entries of J, row-major
__device__ __forceinline__ void iso_mul_axis(const double* a, const double* j, int kind, double* r)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const double a0 = a[4 * i], a1 = a[4 * i + 1], a2 = a[4 * i + 2];
    if (kind == 3)
    {
      r[4 * i + 0] = a0 * j[0] + a1 * j[3];
      r[4 * i + 1] = a0 * j[1] + a1 * j[4];
      r[4 * i + 2] = a2 * j[8];
    }
    else if (kind == 2)
    {
      r[4 * i + 0] = a0 * j[0] + a2 * j[6];
      r[4 * i + 1] = a1 * j[4];
      r[4 * i + 2] = a0 * j[2] + a2 * j[8];
    }
    else
    {
      r[4 * i + 0] = a0 * j[0];
      r[4 * i + 1] = a1 * j[4] + a2 * j[7];
      r[4 * i + 2] = a1 * j[5] + a2 * j[8];
    }
    r[4 * i + 3] = a[4 * i + 3];
  }
}

template <typename QT>
__device__ __forceinline__ void fk_step_sparse(const DevLink& l, const QT* qs, double* slots, int qstride, int stride,
                                               double* cur)
{
  if (l.parent < 0 || l.joint_type == B200DF_JOINT_FIXED ||
      !((l.joint_type == B200DF_JOINT_REVOLUTE && l.jkind != 0) || l.joint_type == B200DF_JOINT_PRISMATIC))
  {
    fk_step(l, qs, slots, qstride, stride, cur);
    return;
  }
  double par[12];
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
  double base[12];
  if (l.origin_is_identity)
  {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      base[k] = par[k];
  }
  else
    iso_mul(par, l.origin, base);
  const double v = joint_value(l, qs, 0, qstride);
  if (l.joint_type == B200DF_JOINT_PRISMATIC)
  {
    // J = identity rotation + translation axis*v: the rotation block passes through unchanged
    const double tx = l.ax * v, ty = l.ay * v, tz = l.az * v;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      cur[4 * i + 0] = base[4 * i + 0];
      cur[4 * i + 1] = base[4 * i + 1];
      cur[4 * i + 2] = base[4 * i + 2];
      cur[4 * i + 3] = ((base[4 * i] * tx + base[4 * i + 1] * ty) + base[4 * i + 2] * tz) + base[4 * i + 3];
    }
  }
  else
  {
    double s, c;
    sincos(v, &s, &c);
    const double tt = 1.0 - c;
    const double txy = tt * l.xy, txz = tt * l.xz, tyz = tt * l.yz;
    const double zs = l.az * s, ys = l.ay * s, xs = l.ax * s;
    double j[9];
    j[0] = tt * l.x2 + c;
    j[3] = txy + zs;
    j[6] = txz - ys;
    j[1] = txy - zs;
    j[4] = tt * l.y2 + c;
    j[7] = tyz + xs;
    j[2] = txz + ys;
    j[5] = tyz - xs;
    j[8] = tt * l.z2 + c;
    iso_mul_axis(base, j, l.jkind, cur);
  }
  if (l.store_slot >= 0)
  {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      slots[(l.store_slot * 12 + k) * stride] = cur[k];
  }
}

template <typename QT>
__device__ __forceinline__ void fk_step(const DevLink& l, const QT* qs, double* slots, int stride, double* cur)
{
  fk_step(l, qs, slots, stride, stride, cur);
}

// ------------------------------------------------------------------------------------------------------------
// Cooperative FK: ONE state walked by ONE warp (the latency-bound small-batch kernels).  Lane i first evaluates the
// joint transform of link i (all sincos in parallel), then the chain T_link = (T_parent * T_origin) * T_joint runs in
// link order with one matrix entry per lane.  Every entry is the same left-to-right sum of products iso_mul / fk_step
// evaluate on a single thread, so the transforms are bit-identical to the batched kernels'.
// Shared-memory staging owned by the calling warp: J, O [n_links][12], par_flag [2 * n_links], TMP [16].
// Output: T [n_links][12] (global link transforms).
// ------------------------------------------------------------------------------------------------------------
constexpr int kCoopFixed = 1, kCoopOriginIdentity = 2;

// entry (r, c) of a.affine() * b.matrix(), the sums in iso_mul's order
__device__ __forceinline__ double iso_mul_entry(const double* a, const double* b, int r, int c)
{
  const double a0 = a[4 * r], a1 = a[4 * r + 1], a2 = a[4 * r + 2];
  const double v = (a0 * b[c] + a1 * b[4 + c]) + a2 * b[8 + c];
  return c == 3 ? v + a[4 * r + 3] : v;
}

__host__ __device__ inline size_t warp_fk_staging_doubles(int n_links)
{
  return (size_t)n_links * 24 + 16 + (size_t)(n_links + 1) / 2 * 2;  // J, O, TMP, parent + flags as ints
}

template <typename QT>
__device__ __forceinline__ void warp_fk(const DevLink* __restrict__ links, int L, const QT* my_q, int lane,
                                        double* staging, double* T)
{
  double* J = staging;
  double* O = J + (size_t)L * 12;
  double* TMP = O + (size_t)L * 12;
  int* parent = reinterpret_cast<int*>(TMP + 16);
  int* lflags = parent + L;
  for (int i = lane; i < L; i += 32)
  {
    const DevLink& l = links[i];
    double j[12];
    joint_transform(l, my_q, 1, j);
#pragma unroll
    for (int k = 0; k < 12; ++k)
    {
      J[i * 12 + k] = j[k];
      O[i * 12 + k] = l.origin[k];
    }
    parent[i] = l.parent;
    lflags[i] = (l.joint_type == B200DF_JOINT_FIXED ? kCoopFixed : 0) | (l.origin_is_identity ? kCoopOriginIdentity : 0);
  }
  __syncwarp();
  const int e = lane < 12 ? lane : 0;  // lanes 12..31 shadow lane 0 and do not store
  const int r = e >> 2, c = e & 3;
  for (int i = 0; i < L; ++i)
  {
    const int p = parent[i], f = lflags[i];
    const double* par = T + p * 12;
    const double* Ji = J + i * 12;
    const double* Oi = O + i * 12;
    double out;
    if (p >= 0 && (f & kCoopFixed))
      out = iso_mul_entry(par, Oi, r, c);  // robot_state.cpp:727-729
    else if (p >= 0)
    {
      if (f & kCoopOriginIdentity)
        out = iso_mul_entry(par, Ji, r, c);  // :732-734
      else
      {
        const double tmp = iso_mul_entry(par, Oi, r, c);  // :735-738, left to right
        if (lane < 12)
          TMP[e] = tmp;
        __syncwarp();
        out = iso_mul_entry(TMP, Ji, r, c);
      }
    }
    else
      out = (f & kCoopOriginIdentity) ? Ji[e] : iso_mul_entry(Oi, Ji, r, c);  // :743-747
    if (lane < 12)
      T[i * 12 + e] = out;
    __syncwarp();
  }
}

// getDistanceGradient's cell lookup without the gradient: false when the centre is outside the 1-cell-inset
// box (the reference then reports max_distance => never a collision, Q4).  NaN fails every comparison.
__device__ __forceinline__ bool inset_cell(const GridDev& g, double x, double y, double z, long& idx)
{
  const double fx = floor((x - g.om[0]) * g.oo), fy = floor((y - g.om[1]) * g.oo), fz = floor((z - g.om[2]) * g.oo);
  if (!(fx >= 1.0 && fy >= 1.0 && fz >= 1.0 && fx < (double)(g.n[0] - 1) && fy < (double)(g.n[1] - 1) &&
        fz < (double)(g.n[2] - 1)))
    return false;
  idx = (long)fx * g.stride1 + (long)fy * g.stride2 + (long)fz;
  return true;
}

// PropagationDistanceField::getDistance (propagation_distance_field.h:597-600)
template <typename FT>
__device__ __forceinline__ double cell_dist(const FieldDev& f, long idx)
{
  B200DF_CHECK_INDEX(idx, (long long)f.g.n[0] * f.g.stride1);
  B200DF_CHECK_INDEX(static_cast<const FT*>(f.d2)[idx], f.max_d2 + 1);
  double d = f.sqrt_table[static_cast<const FT*>(f.d2)[idx]];
  if (f.neg)
    d = d - f.sqrt_table[static_cast<const FT*>(f.neg)[idx]];
  // unsigned field: the reference subtracts sqrt_table_[0] = sqrt(0) * resolution = +0.0, and d - 0.0 == d bit for bit
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
  {
    B200DF_CHECK_INDEX(idx, (long long)f.g.n[0] * f.g.stride1);
    return (int)static_cast<const FT*>(f.d2)[idx] < thr;  // host-resolved, exactly equivalent threshold
  }
  const double dist = cell_dist<FT>(f, idx);
  return (max_prop > dist) && (radius - dist > tol);
}

// what the verdict kernel reads of a field: the fp64 world->grid mapping plus integer bounds
struct LookupDev
{
  double om[3];      // origin_minus_
  double oo;         // oo_resolution_
  unsigned lim[3];   // n-2 (0 when n < 2): cell c is inside the 1-cell inset box iff (unsigned)(c-1) < lim
  int stride1, stride2;
  const void* d2;
  const void* neg;
  const double* sqrt_table;
};

inline LookupDev make_lookup(const FieldDev& f)
{
  LookupDev l{};
  for (int i = 0; i < 3; ++i)
  {
    l.om[i] = f.g.om[i];
    l.lim[i] = f.g.n[i] >= 2 ? (unsigned)(f.g.n[i] - 2) : 0u;
  }
  l.oo = f.g.oo;
  l.stride1 = (int)f.g.stride1;
  l.stride2 = (int)f.g.stride2;
  l.d2 = f.d2;
  l.neg = f.neg;
  l.sqrt_table = f.sqrt_table;
  return l;
}

// getCollisionSphereCollision's per-sphere predicate (collision_distance_field_types.cpp:213-227) on top of
// getDistanceGradient's cell lookup (distance_field.cpp:62-86, voxel_grid.h:515-532).
// floor((p - origin_minus) * oo) is evaluated in fp64 exactly like the reference; cvt.rmi.s32.f64 IS that floor for
// every value inside the grid and saturates outside (NaN -> 0), where the unsigned range test fails anyway.
// A centre outside the 1-cell inset box reads as "max distance" in the reference => never a collision (Q4).
// world->grid cell of a sphere centre: flat voxel index (0 when outside) + "inside the 1-cell inset box"
__device__ __forceinline__ unsigned cell_index(const LookupDev& f, double x, double y, double z, bool& in)
{
  const int ix = __double2int_rd((x - f.om[0]) * f.oo);
  const int iy = __double2int_rd((y - f.om[1]) * f.oo);
  const int iz = __double2int_rd((z - f.om[2]) * f.oo);
  in = ((unsigned)(ix - 1) < f.lim[0]) & ((unsigned)(iy - 1) < f.lim[1]) & ((unsigned)(iz - 1) < f.lim[2]);
  return in ? (unsigned)(ix * f.stride1 + iy * f.stride2 + iz) : 0u;
}

template <typename FT>
__device__ __forceinline__ bool voxel_hits(const LookupDev& f, unsigned idx, bool in, int d2, double radius, int thr,
                                           double tol, double max_prop)
{
  if (!f.neg)
    return in & (d2 < thr);  // host-resolved threshold, exactly equivalent to the fp64 predicate
  const double dist = f.sqrt_table[d2] - f.sqrt_table[static_cast<const FT*>(f.neg)[idx]];
  return in && (max_prop > dist) && (radius - dist > tol);
}

// stage one tile of joint values into shared memory as [var][thread] (coalesced global read)
template <typename QT>
__device__ __forceinline__ void stage_q(const QT* __restrict__ q, long base, long n, int n_vars, int B, int tid,
                                        QT* qs)
{
  const long total = (long)B * n_vars;
  const long limit = (n - base) * n_vars;
  for (long e = tid; e < total; e += B)
  {
    const int s = (int)(e / n_vars), v = (int)(e % n_vars);
    qs[v * B + s] = (e < limit) ? q[base * n_vars + e] : QT(0);
  }
}
}  // namespace b200df

// host-side handles (shared by validity.cu and gradients.cu)
struct b200df_model
{
  int device = 0;
  int n_links = 0, n_vars = 0;
  std::vector<b200df::DevLink> links;
  std::vector<int> has_geometry;
  std::vector<int> sphere_offset;
  std::vector<double> spheres;
  std::vector<double> bsphere;
  std::vector<int64_t> point_offset;
  std::vector<double> points;
  int n_slots = 0;
  b200df::DevLink* d_links = nullptr;
  // b200df_model_fk's link table: an empty group over the model, built on first use and kept
  mutable std::mutex fk_mu;
  mutable b200df_group* fk_group = nullptr;
};

struct b200df_group
{
  const b200df_model* model = nullptr;
  int device = 0;
  b200df_group_desc desc{};
  std::vector<b200df::DevGroupLink> glinks;
  std::vector<int> link_gslot;
  std::vector<double> sph;
  std::vector<b200df::DevPair> pairs;
  b200df::GroupDev dev{};
  std::vector<void*> allocations;
  std::vector<b200df::DevPartner> partners;  // host copies the FAST tables are derived from
  std::vector<int> sph_thr;
  bool neg_irrelevant = false;  // verdicts of this group never depend on a field's negative distances (validity.cu)
  struct StagePool;  // host<->device staging contexts of the pipelined host-buffer entry points (validity.cu)
  StagePool* stage = nullptr;
  struct FastPath;   // fp32 screening tables + scratch (validity_fast.cu); null when the model does not fit them
  FastPath* fast = nullptr;
  struct SmallPath;  // one-CTA-per-state kernel for latency-bound small batches (validity_small.cu)
  SmallPath* small = nullptr;
  void* d_link_fields = nullptr;  // device copies behind dev.link_fields / dev.lf_enabled
  void* d_lf_enabled = nullptr;
  int link_field_elem = 0;        // voxel type of the link fields (1 or 2)
};

// ------------------------------------------------------------------------------------------------------------
// Staging contexts of the host-buffer entry point: a batch is cut into 2^20-state chunks that travel H2D -> kernel ->
// D2H on three streams, so the PCIe copies of one chunk overlap the kernel of another.  Contexts are cached per group
// (checks may run concurrently from several threads: each call takes its own context).
// ------------------------------------------------------------------------------------------------------------
constexpr int kSlots = 3;  // chunks in flight: H2D of one overlaps the kernel of another and the D2H of a third

struct b200df_group::StagePool
{
  struct Ctx
  {
    cudaStream_t stream[kSlots] = {};
    void* q[kSlots] = {};
    uint8_t* v[kSlots] = {};
    int* amb[kSlots] = {};  // [0] = count, [1..] = indices (FAST path scratch)
    size_t q_bytes = 0, v_bytes = 0;
    // small-call staging shared by every entry point (grow-only): one mapped pinned buffer the kernels read their
    // inputs from and write their results to across PCIe, plus device scratch
    void* pin = nullptr;
    size_t pin_bytes = 0;
    static constexpr int kScratch = 12;
    void* scratch[kScratch] = {};
    size_t scratch_bytes[kScratch] = {};
    // page-locked staging of large transfers from / to pageable caller memory, one per pipeline slot (grow-only)
    void* hstage[kSlots] = {};
    size_t hstage_bytes[kSlots] = {};
    // lanes of the narrowing path (pageable fp64 rows -> fp32 screening): one independent pipeline per host thread
    static constexpr int kLanes = 16;
    struct Lane
    {
      cudaStream_t st = nullptr;
      void* hq = nullptr;    // pinned: fp32 rows of one chunk, then the chunk's verdict bytes
      void* dq = nullptr;    // device: the same
      int* amb = nullptr;    // device scratch of the screening pass
      size_t states = 0;     // chunk capacity the buffers were sized for
      size_t row = 0;
    };
    Lane lane[kLanes];
  };
  std::mutex mu;
  std::vector<Ctx*> free_list;
  ~StagePool()
  {
    for (Ctx* c : free_list)
      release(c);
  }
  static void release(Ctx* c)
  {
    for (int i = 0; i < kSlots; ++i)
    {
      cudaFree(c->q[i]);
      cudaFree(c->v[i]);
      cudaFree(c->amb[i]);
      if (c->stream[i])
        cudaStreamDestroy(c->stream[i]);
    }
    if (c->pin)
      cudaFreeHost(c->pin);
    for (void* p : c->scratch)
      cudaFree(p);
    for (void* p : c->hstage)
      if (p)
        cudaFreeHost(p);
    for (Ctx::Lane& l : c->lane)
    {
      if (l.hq)
        cudaFreeHost(l.hq);
      cudaFree(l.dq);
      cudaFree(l.amb);
      if (l.st)
        cudaStreamDestroy(l.st);
    }
    delete c;
  }
  Ctx* acquire()
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!free_list.empty())
    {
      Ctx* c = free_list.back();
      free_list.pop_back();
      return c;
    }
    return new Ctx;
  }
  void give_back(Ctx* c)
  {
    std::lock_guard<std::mutex> lk(mu);
    free_list.push_back(c);
  }
  static cudaError_t ensure_stream(Ctx* c, int i)
  {
    return c->stream[i] ? cudaSuccess : cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking);
  }
  // mapped pinned buffer of at least `bytes`; *dev = the address kernels use for it
  static cudaError_t ensure_pin(Ctx* c, size_t bytes, void** dev)
  {
    cudaError_t e = cudaSuccess;
    if (bytes > c->pin_bytes)
    {
      if (c->pin)
        cudaFreeHost(c->pin);
      c->pin = nullptr;
      c->pin_bytes = 0;
      const size_t want = std::max<size_t>(bytes, 64 * 1024);
      e = cudaHostAlloc(&c->pin, want, cudaHostAllocMapped);
      if (e != cudaSuccess)
        return e;
      c->pin_bytes = want;
    }
    return cudaHostGetDevicePointer(dev, c->pin, 0);
  }
  static cudaError_t ensure_hstage(Ctx* c, int i, size_t bytes)
  {
    if (bytes <= c->hstage_bytes[i])
      return cudaSuccess;
    if (c->hstage[i])
      cudaFreeHost(c->hstage[i]);
    c->hstage[i] = nullptr;
    c->hstage_bytes[i] = 0;
    const cudaError_t e = cudaHostAlloc(&c->hstage[i], bytes, cudaHostAllocDefault);
    if (e == cudaSuccess)
      c->hstage_bytes[i] = bytes;
    return e;
  }
  static cudaError_t ensure_scratch(Ctx* c, int i, size_t bytes)
  {
    if (bytes <= c->scratch_bytes[i])
      return cudaSuccess;
    cudaFree(c->scratch[i]);
    c->scratch[i] = nullptr;
    c->scratch_bytes[i] = 0;
    const cudaError_t e = cudaMalloc(&c->scratch[i], bytes);
    if (e == cudaSuccess)
      c->scratch_bytes[i] = bytes;
    return e;
  }
};



namespace b200df
{
// FAST path (validity_fast.cu)
bool fast_tables_ok(const b200df_group* group);  // the robot fits FAST's constant-bank tables
int fast_create(b200df_group* group);   // builds the fp32 tables; leaves group->fast null when unsupported
void fast_destroy(b200df_group* group);
// true when the FAST kernel can serve this call (unsigned fields, model fits the tables)
bool fast_supported(const b200df_group* group, const FieldDev& wf, const FieldDev& sf, int q_dtype, int flags);
// fp32 screening pass: verdict[i] in {0, 1} or 2 = "rounding could matter"; the indices of the 2s are appended to
// amb_idx (capacity n) and counted in *amb_count, both device memory, *amb_count zeroed by this call
int fast_launch(const b200df_group* group, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                int q_dtype, long n, int flags, uint8_t* verdict, int* amb_idx, int* amb_count, cudaStream_t stream);

// small-batch path (validity_small.cu): one CTA per state, same verdicts as the batched exact kernel
int small_create(b200df_group* group);  // leaves group->small null when the robot does not fit shared memory
void small_destroy(b200df_group* group);
bool small_supported(const b200df_group* group);
// redo_idx / redo_count (device): evaluate only the listed states (the FAST path's undecided set)
int small_launch(const b200df_group* group, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                 int q_dtype, long n, int flags, uint8_t* verdict, cudaStream_t stream, const int* redo_idx = nullptr,
                 const int* redo_count = nullptr);

// true when p is ordinary pageable host memory (a cudaMemcpyAsync on it goes through the driver's own small staging
// buffer at a few GB/s and holds the calling thread)
inline bool is_pageable_host(const void* p)
{
  const char* off = getenv("B200DF_STAGE_PAGEABLE");  // "0": leave pageable memory to the driver (for A/B measurements)
  if (off && off[0] == '0')
    return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// memcpy split over a few host threads: one core moves ~10 GB/s, PCIe Gen5 five times that
inline void parallel_memcpy(void* dst, const void* src, size_t bytes)
{
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t max_threads = std::min<size_t>(8, std::max<unsigned>(1, hw / 2));
  const size_t n = std::min(max_threads, bytes / (size_t(4) << 20));  // at least 4 MB per thread (a spawn costs ~20 us)
  if (n <= 1)
  {
    std::memcpy(dst, src, bytes);
    return;
  }
  const size_t piece = ((bytes / n) + 4095) & ~size_t(4095);
  std::vector<std::thread> workers;
  for (size_t i = 1; i < n; ++i)
  {
    const size_t off = i * piece;
    if (off >= bytes)
      break;
    const size_t len = std::min(piece, bytes - off);
    workers.emplace_back([=] { std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
  }
  std::memcpy(dst, src, std::min(piece, bytes));
  for (std::thread& w : workers)
    w.join();
}

// several host copies at once: the pieces (>= 1 MB) of all jobs are dealt out over up to 16 threads
struct CopyJob
{
  void* dst;
  const void* src;
  size_t bytes;
};
inline void parallel_copy_many(const CopyJob* jobs, int n_jobs)
{
  struct Piece
  {
    char* dst;
    const char* src;
    size_t len;
  };
  std::vector<Piece> pieces;
  const size_t kPiece = size_t(1) << 20;
  for (int j = 0; j < n_jobs; ++j)
    for (size_t off = 0; off < jobs[j].bytes; off += kPiece)
      pieces.push_back({ static_cast<char*>(jobs[j].dst) + off, static_cast<const char*>(jobs[j].src) + off,
                         std::min(kPiece, jobs[j].bytes - off) });
  const unsigned hw = std::thread::hardware_concurrency();
  const size_t n_threads = std::min<size_t>(std::min<size_t>(16, std::max<unsigned>(1, hw)), pieces.size() / 2);
  if (n_threads <= 1)
  {
    for (const Piece& p : pieces)
      std::memcpy(p.dst, p.src, p.len);
    return;
  }
  std::atomic<size_t> next{ 0 };
  auto work = [&] {
    for (size_t i = next++; i < pieces.size(); i = next++)
      std::memcpy(pieces[i].dst, pieces[i].src, pieces[i].len);
  };
  std::vector<std::thread> workers;
  for (size_t t = 1; t < n_threads; ++t)
    workers.emplace_back(work);
  work();
  for (std::thread& w : workers)
    w.join();
}

// B200DF_SMALL_BATCH=0 switches the mapped-pinned small-call staging of the host-buffer entry points off
inline bool small_batch_enabled()
{
  const char* e = getenv("B200DF_SMALL_BATCH");
  return !(e && e[0] == '0');
}

// fills wf / sf for the requested flags (rebuilding dirty fields) and checks they fit the group
int prepare_fields(const b200df_group* group, b200df_field* world, b200df_field* self, int flags, FieldDev* wf,
                   FieldDev* sf, int* elem);

inline size_t q_elem_size(int q_dtype)
{
  return q_dtype == B200DF_Q_F32 ? sizeof(float) : sizeof(double);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device, size growth): the driver
// call is too slow to repeat on every chunk of a pipelined batch
inline cudaError_t allow_dynamic_smem_impl(const void* kernel, size_t bytes)
{
  if (bytes <= 48 * 1024)
    return cudaSuccess;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;  // (kernel, device) -> bytes already allowed
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  std::lock_guard<std::mutex> lk(mu);
  size_t& have = granted[std::make_pair(kernel, dev)];
  if (bytes <= have)
    return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess)
    have = bytes;
  return e;
}

template <typename K>
cudaError_t allow_dynamic_smem(K kernel, size_t bytes)
{
  return allow_dynamic_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

struct ScratchBuffer  // stream-ordered scratch
{
  void* p = nullptr;
  cudaStream_t stream = nullptr;
  ~ScratchBuffer()
  {
    if (p)
      cudaFreeAsync(p, stream);
  }
};
}  // namespace b200df
