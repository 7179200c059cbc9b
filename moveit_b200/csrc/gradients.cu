// Generic fp64 sweep (one thread per state, all sphere centres kept in shared memory): FK / sphere-centre /
// clearance outputs, the gradient kernel of getCollisionGradients, and the round-1 verdict kernel that is kept as an
// independent cross-check of the fused verdict kernel in validity.cu (precision value 2).
//
// Reference being replaced, per state:
//   RobotState::updateLinkTransformsInternal            robot_state/src/robot_state.cpp:718-755
//   PosedBodySphereDecomposition::updatePose            collision_distance_field/src/collision_distance_field_types.cpp:390-408
//   getCollisionSphereGradients                         collision_distance_field_types.cpp:143-204
//   CollisionEnvDistanceField::getCollisionGradients    collision_env_distance_field.cpp:1536-1557
//     getSelfProximityGradients :358-431, getIntraGroupProximityGradients :662-727,
//     getEnvironmentProximityGradients :1664-1700
//
// Compiled with --fmad=false so that mul/add round like the reference's non-FMA x86-64 build.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "validity_common.cuh"

using namespace b200df;

namespace
{
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
    fk_step(l, my_q, my_slots, B, cur);
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
// getCollisionGradients :1536-1557: self field, intra group, environment — in that order, each keeping the
// per-sphere minimum with strict '<' like the reference.
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

// The same query split in two, so that a caller can issue the seven voxel loads early and consume them after other
// work (the gather kernel puts its partner-sphere loop between the two halves: the L2 round trip hides behind it).
struct FieldTaps
{
  bool in;
  int v[7];  // d2 at the centre cell, +x, -x, +y, -y, +z, -z
  int n[7];  // negative_distance_square at the same cells (signed fields)
};

template <typename FT>
__device__ __forceinline__ void field_taps_load(const FieldDev& f, double x, double y, double z, FieldTaps& t)
{
  long idx;
  t.in = inset_cell(f.g, x, y, z, idx);
  if (!t.in)
    return;
  const long off[7] = { 0, f.g.stride1, -f.g.stride1, f.g.stride2, -f.g.stride2, 1, -1 };
#pragma unroll
  for (int k = 0; k < 7; ++k)
  {
    B200DF_CHECK_INDEX(idx + off[k], (long long)f.g.n[0] * f.g.stride1);
    t.v[k] = (int)static_cast<const FT*>(f.d2)[idx + off[k]];
  }
  if (f.neg)
  {
#pragma unroll
    for (int k = 0; k < 7; ++k)
      t.n[k] = (int)static_cast<const FT*>(f.neg)[idx + off[k]];
  }
}

__device__ __forceinline__ void field_taps_finish(const FieldDev& f, const double* sqrt_table, const FieldTaps& t,
                                                  double max_value, int type, double& best, double& gx, double& gy,
                                                  double& gz, int& btype)
{
  double dist, dgx = 0.0, dgy = 0.0, dgz = 0.0;
  if (!t.in)
    dist = f.max_distance;
  else
  {
    double d[7];
#pragma unroll
    for (int k = 0; k < 7; ++k)
    {
      d[k] = sqrt_table[t.v[k]];  // cell_dist: sqrt_table_[d2] - sqrt_table_[neg d2]
      if (f.neg)
        d[k] = d[k] - sqrt_table[t.n[k]];
    }
    const double itr = (double)f.inv_twice_res;
    dgx = (d[1] - d[2]) * itr;
    dgy = (d[3] - d[4]) * itr;
    dgz = (d[5] - d[6]) * itr;
    dist = d[0];
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

// ------------------------------------------------------------------------------------------------------------
// Gradient kernel, gather form.  Phase 1: one thread per state runs FK and poses every sphere into shared memory.
// Phase 2: one thread per (state, sphere) replays, for ITS sphere only, the candidate sequence the reference's loops
// produce for it — self field, then every sphere of every enabled partner link (lower-index partners ascending, then
// higher-index partners ascending: exactly the order in which getIntraGroupProximityGradients :662-727 touches this
// sphere), then the environment field — keeping the minimum with strict '<' like the reference.  No read-modify-write
// of global memory.  In the 32-states-per-CTA variant a warp holds ONE sphere index for 32 states (lane = state), so
// the partner loops run the same trip count on every lane; the small-call variant (4 states per CTA) keeps
// consecutive threads on consecutive spheres of a state.
// ------------------------------------------------------------------------------------------------------------
constexpr int kGradStates = 32;    // states per CTA
#ifndef B200DF_GRAD_THREADS
#define B200DF_GRAD_THREADS 128
#endif
#ifndef B200DF_GRAD_MINB
#define B200DF_GRAD_MINB 1
#endif
constexpr int kGradThreads = B200DF_GRAD_THREADS;
constexpr int kMaxStagedNbr = 1024;  // neighbour-table entries staged in shared memory (16 KB)

// PosedDistanceField::getDistanceGradient's point mapping (collision_distance_field_types.h:156-160):
// rel = pose.inverse() * p with Eigen's isometry inverse (R^T, -(R^T t))
__device__ __forceinline__ void posed_rel(const double* pose, double x, double y, double z, double& rx, double& ry,
                                          double& rz)
{
  double inv[12];
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    inv[4 * i + 0] = pose[i];
    inv[4 * i + 1] = pose[4 + i];
    inv[4 * i + 2] = pose[8 + i];
    inv[4 * i + 3] = -((inv[4 * i] * pose[3] + inv[4 * i + 1] * pose[7]) + inv[4 * i + 2] * pose[11]);
  }
  iso_apply(inv, x, y, z, rx, ry, rz);
}

// COOP (small batches): GS = 4 states per CTA and one warp walks one state (warp_fk), so a 100-waypoint trajectory
// spreads over 25 SMs and phase 1 costs one FK chain; otherwise 32 states per CTA, one thread per state's FK.
// OT: element type of the distance / gradient / centre outputs.  float = the fp64 results rounded once on the way out
// (north_star asks for 1e-5 m there; rounding a distance below 1 m or a gradient component below 8 costs < 5e-7), "no
// candidate" (DBL_MAX) becomes FLT_MAX: half the bytes to write and to carry over PCIe.
template <typename OT>
__device__ __forceinline__ OT to_out(double v)
{
  return (OT)v;
}
template <>
__device__ __forceinline__ float to_out<float>(double v)
{
  return v >= (double)FLT_MAX ? FLT_MAX : (float)v;
}

template <typename QT, typename FT, typename OT, bool LF, int GS, int NT, bool COOP>
__global__ void __launch_bounds__(NT, COOP ? 1 : B200DF_GRAD_MINB) gradients_gather_kernel(GroupDev g, FieldDev wf, FieldDev sf,
                                                                        const QT* __restrict__ q, long n, int flags,
                                                                        OT* __restrict__ out_dist,
                                                                        OT* __restrict__ out_grad,
                                                                        uint8_t* __restrict__ out_type,
                                                                        OT* __restrict__ out_cen)
{
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const long base = (long)blockIdx.x * GS;
  const int S = g.n_spheres;
  const int G = g.n_glinks;
  const int L = g.n_links;
  // XPOSE (the 32-states-per-CTA variant): lane = state, a warp works on ONE sphere index for 32 states, so the
  // partner loops have the same trip count on every lane (the per-sphere mapping left 12 of 32 lanes active on
  // average).  The per-state strides are odd so that 32 lanes reading the same element of 32 states hit different banks.
  constexpr bool XPOSE = !COOP && GS == 32;
  const int CS = (S * 3) | 1;  // doubles per state in cen
  const int US = S | 1;        // words per state in can / abt
  double* cen = smem;                        // [state][CS]: posed sphere centres, [k][3] inside a state
  double* slots = cen + (size_t)GS * CS;     // !COOP: [n_slots][12][state]   COOP: T [state][L][12]
  double* poses = slots + (COOP ? (size_t)GS * L * 12 : (size_t)g.n_slots * 12 * GS);  // !COOP && LF: [state][G][12]
  double* staging = poses + ((LF && !COOP) ? (size_t)GS * G * 12 : 0);                 // COOP: warp_fk staging per warp
  unsigned* can = reinterpret_cast<unsigned*>(staging + (COOP ? (NT / 32) * warp_fk_staging_doubles(L) : 0));
  unsigned* abt = can + (LF ? (size_t)GS * US : 0);                       // LF: [state][US] (can likewise)
  QT* qs = reinterpret_cast<QT*>(abt + (LF ? (size_t)GS * US : 0));  // [state][n_vars]
  // the pair tables of the gather phase, staged once per CTA (a global load per partner link in the inner loop was
  // the top stall): neighbour list (16-byte aligned), its offsets, the link slot of every sphere
  const bool stage_tables = g.n_nbr <= kMaxStagedNbr;
  // (offsets are taken relative to the shared-memory base so that the compiler keeps these in the shared address space)
  char* const sm_bytes = reinterpret_cast<char*>(smem);
  const size_t nbr_at = ((size_t)(reinterpret_cast<char*>(qs + (size_t)GS * g.n_vars) - sm_bytes) + 15) & ~(size_t)15;
  int4* s_nbr = reinterpret_cast<int4*>(sm_bytes + nbr_at);
  int* s_nbr_off = reinterpret_cast<int*>(s_nbr + (stage_tables ? g.n_nbr : 0));
  int* s_sph_glink = s_nbr_off + (G + 1);
  if (stage_tables)
  {
    for (int e = tid; e < g.n_nbr; e += NT)
      s_nbr[e] = g.nbr[e];
    for (int e = tid; e <= G; e += NT)
      s_nbr_off[e] = g.nbr_off[e];
    for (int e = tid; e < S; e += NT)
      s_sph_glink[e] = g.sph_glink[e];
  }
  // the fields' sqrt tables (<= 626 doubles each): 7 dependent lookups per sphere and field otherwise go to L1 / L2
  const size_t sqrt_at = ((size_t)(reinterpret_cast<char*>(s_sph_glink + S) - sm_bytes) + 7) & ~(size_t)7;
  double* s_sqrt_w = reinterpret_cast<double*>(sm_bytes + sqrt_at);
  double* s_sqrt_s = s_sqrt_w + (wf.d2 ? wf.max_d2 + 1 : 0);
  if (wf.d2)
  {
    for (int e = tid; e <= wf.max_d2; e += NT)
      s_sqrt_w[e] = wf.sqrt_table[e];
    wf.sqrt_table = s_sqrt_w;
  }
  if (sf.d2)
  {
    for (int e = tid; e <= sf.max_d2; e += NT)
      s_sqrt_s[e] = sf.sqrt_table[e];
    sf.sqrt_table = s_sqrt_s;
  }
  const int* sph_glink = stage_tables ? s_sph_glink : g.sph_glink;
  {
    const int total = GS * g.n_vars;
    const long limit = (n - base) * g.n_vars;
    const QT* src = q + base * g.n_vars;
    for (int e = tid; e < total; e += NT)
      qs[e] = (e < limit) ? src[e] : QT(0);
  }
  __syncthreads();
  if (COOP)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int st = warp; st < GS; st += NT / 32)
    {
      double* T = slots + (size_t)st * L * 12;
      warp_fk(g.links, L, qs + (size_t)st * g.n_vars, lane, staging + warp * warp_fk_staging_doubles(L), T);
      double* my_cen = cen + (size_t)st * CS;
      for (int s = lane; s < S; s += 32)
      {
        const double* rel = g.sph + 4 * s;
        const double* Tl = T + (size_t)g.glinks[g.sph_glink[s]].link * 12;
        iso_apply(Tl, rel[0], rel[1], rel[2], my_cen[3 * s], my_cen[3 * s + 1], my_cen[3 * s + 2]);
      }
    }
  }
  else if (tid < GS)
  {
    double cur[12];
    double* my_cen = cen + (size_t)tid * CS;
    for (int i = 0; i < g.n_links; ++i)
    {
      fk_step(g.links[i], qs + (size_t)tid * g.n_vars, slots + tid, 1, GS, cur);
      const int gs = g.link_gslot[i];
      if (gs < 0)
        continue;
      const DevGroupLink& gl = g.glinks[gs];
      for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
      {
        const double* rel = g.sph + 4 * s;
        iso_apply(cur, rel[0], rel[1], rel[2], my_cen[3 * s], my_cen[3 * s + 1], my_cen[3 * s + 2]);
      }
      if (LF)
      {
#pragma unroll
        for (int k = 0; k < 12; ++k)
          poses[((size_t)tid * G + gs) * 12 + k] = cur[k];  // PosedDistanceField::updatePose
      }
    }
  }
  __syncthreads();
  // pose of group link j in state st (PosedDistanceField::updatePose)
  auto link_pose = [&](int st, int j) -> const double* {
    return COOP ? slots + ((size_t)st * L + g.glinks[j].link) * 12 : poses + ((size_t)st * G + j) * 12;
  };
  const int items = GS * S;
  if (LF)
  {
    // Which link fields may act on which sphere.  PosedDistanceField::getCollisionSphereGradients
    // (collision_distance_field_types.cpp:81-141) walks a link's spheres in order and RETURNS at the first sphere
    // that is outside the other link's field (Q7; the "gradient norm > 0" guard is the norm of pose * 0 = the
    // pose's translation, Q6), so sphere k sees field j only if it is inside AND no earlier sphere of its link aborted.
    for (int it = tid; it < items; it += NT)
    {
      const int st = XPOSE ? it % GS : it / S, k = XPOSE ? it / GS : it - st * S;
      const double* sc = cen + (size_t)st * CS;
      const int gs = sph_glink[k];
      unsigned allowed = g.glinks[gs].self_enabled ? (g.lf_enabled[gs] & ~(1u << gs)) : 0u;
      unsigned inb = 0u, ab = 0u;
      while (allowed)
      {
        const int j = __ffs(allowed) - 1;
        allowed &= allowed - 1;
        const FieldDev& f = g.link_fields[j];
        if (!f.d2)
          continue;
        const double* pose = link_pose(st, j);
        double rx, ry, rz;
        posed_rel(pose, sc[3 * k], sc[3 * k + 1], sc[3 * k + 2], rx, ry, rz);
        long idx;
        if (inset_cell(f.g, rx, ry, rz, idx))
          inb |= 1u << j;
        else if (sqrt((pose[3] * pose[3] + pose[7] * pose[7]) + pose[11] * pose[11]) > 0.0)
          ab |= 1u << j;
      }
      can[st * US + k] = inb;
      abt[st * US + k] = ab;
    }
    __syncthreads();
    for (int it = tid; it < GS * G; it += NT)
    {
      const int st = it / G, gs = it - st * G;
      unsigned alive = 0xffffffffu;
      for (int k = g.glinks[gs].sphere_begin; k < g.glinks[gs].sphere_end; ++k)
      {
        const unsigned a = abt[st * US + k];
        can[st * US + k] &= alive;
        alive &= ~a;
      }
    }
    __syncthreads();
  }
  for (int it = tid; it < items; it += NT)
  {
    const int st = XPOSE ? it % GS : it / S, k = XPOSE ? it / GS : it - st * S;
    const long state = base + st;
    if (state >= n)
    {
      if (XPOSE)
        continue;
      break;
    }
    const double* sc = cen + (size_t)st * CS;
    const double cx = sc[3 * k], cy = sc[3 * k + 1], cz = sc[3 * k + 2];
    const int gs = sph_glink[k];
    // updateGroupStateRepresentationState :1128-1133
    double best = DBL_MAX, gx = 0.0, gy = 0.0, gz = 0.0;
    int bt = B200DF_TYPE_NONE;
    if (LF)
    {
      // getSelfProximityGradients: the other links' PosedDistanceFields, in link order (:388-418)
      unsigned m = can[st * US + k];
      while (m)
      {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const double* pose = link_pose(st, j);
        double rx, ry, rz;
        posed_rel(pose, cx, cy, cz, rx, ry, rz);
        double d = DBL_MAX, lx = 0.0, ly = 0.0, lz = 0.0;
        int lt = B200DF_TYPE_NONE;
        field_gradient_update<FT>(g.link_fields[j], rx, ry, rz, g.max_prop, B200DF_TYPE_SELF, d, lx, ly, lz, lt);
        if (lt == B200DF_TYPE_SELF && d < best)  // dist < maximum_value && dist < distances[i]
        {
          best = d;
          iso_apply(pose, lx, ly, lz, gx, gy, gz);  // pose * gradient: rotation AND translation (Q6)
          bt = B200DF_TYPE_SELF;
        }
      }
    }
    // getSelfProximityGradients: self field term (:420-422)
    if ((flags & B200DF_CHECK_SELF) && g.glinks[gs].self_enabled)
      field_gradient_update<FT>(sf, cx, cy, cz, g.max_prop, B200DF_TYPE_SELF, best, gx, gy, gz, bt);
    FieldTaps wt;
    if (flags & B200DF_CHECK_ROBOT)
      field_taps_load<FT>(wf, cx, cy, cz, wt);  // consumed after the partner loop
    if (flags & B200DF_CHECK_INTRA)
    {
      // sqrt is only taken for candidates that can still win: e2 > bound  =>  sqrt(e2) > best.
      // The reference's gradient is centre(first link) - centre(second link), negated for the second link's sphere;
      // both cases come out as -(other - own) bit for bit, and the squared norm does not see the sign.
      double bound = best * best * 1.000000000000001;
      auto consider = [&](double dx, double dy, double dz, double e2) {
        if (e2 > bound)
          return;
        const double d = sqrt(e2);
        if (d < best)
        {
          best = d;
          bound = d * d * 1.000000000000001;
          gx = -dx;
          gy = -dy;
          gz = -dz;
          bt = B200DF_TYPE_INTRA;
        }
      };
      // (a lambda instantiated once per table location, so that the staged copy is read with shared-memory loads)
      auto partner_loop = [&](const int4* nbr, int nb0, int nb1) {
      for (int nb = nb0; nb < nb1; ++nb)
      {
        const int4 o = nbr[nb];  // partner link: its sphere range is [z, w)
        int m = o.z;
        // four candidates at a time: independent loads and distance chains, then the reference's sequential
        // comparison (a candidate above the current bound stays above every later, smaller bound)
        for (; m + 4 <= o.w; m += 4)
        {
          double dx[4], dy[4], dz[4], e2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
          {
            dx[j] = sc[3 * (m + j)] - cx;
            dy[j] = sc[3 * (m + j) + 1] - cy;
            dz[j] = sc[3 * (m + j) + 2] - cz;
            e2[j] = (dx[j] * dx[j] + dy[j] * dy[j]) + dz[j] * dz[j];
          }
          if ((e2[0] > bound) & (e2[1] > bound) & (e2[2] > bound) & (e2[3] > bound))
            continue;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            consider(dx[j], dy[j], dz[j], e2[j]);
        }
        for (; m < o.w; ++m)
        {
          const double dx = sc[3 * m] - cx, dy = sc[3 * m + 1] - cy, dz = sc[3 * m + 2] - cz;
          consider(dx, dy, dz, (dx * dx + dy * dy) + dz * dz);
        }
      }
      };
      if (stage_tables)
        partner_loop(s_nbr, s_nbr_off[gs], s_nbr_off[gs + 1]);
      else
        partner_loop(g.nbr, g.nbr_off[gs], g.nbr_off[gs + 1]);
    }
    // getEnvironmentProximityGradients :1664-1700
    if (flags & B200DF_CHECK_ROBOT)
      field_taps_finish(wf, s_sqrt_w, wt, g.max_prop, B200DF_TYPE_ENVIRONMENT, best, gx, gy, gz, bt);
    const long o = state * S + k;
    out_dist[o] = to_out<OT>(best);
    out_grad[3 * o] = to_out<OT>(gx);
    out_grad[3 * o + 1] = to_out<OT>(gy);
    out_grad[3 * o + 2] = to_out<OT>(gz);
    out_type[o] = (uint8_t)bt;
    if (out_cen)
    {
      out_cen[3 * o] = to_out<OT>(cx);
      out_cen[3 * o + 1] = to_out<OT>(cy);
      out_cen[3 * o + 2] = to_out<OT>(cz);
    }
  }
}

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
}  // namespace

namespace b200df
{
// round-1 verdict kernel (all centres in shared memory, pair phase after the sweep); cross-check only
int launch_verdict_v1(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                      int q_dtype, long n, int flags, uint8_t* verdict, cudaStream_t stream)
{
  return dispatch_exact<OUT_VERDICT>(g, wf, sf, field_elem, q, q_dtype, n, flags, verdict, nullptr, stream);
}
}  // namespace b200df

namespace
{
constexpr int kCoopStates = 4, kCoopThreads = 128;  // small-batch variant
constexpr size_t kSmallGradBytes = 1 << 20;         // host-buffer calls up to this size use the mapped pinned staging
constexpr int64_t kCoopMaxStates = 4096;            // above this the 32-states-per-CTA variant has the GPU filled

size_t gather_smem_bytes(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, size_t q_elem, bool lf, int gs,
                         int nt, bool coop)
{
  size_t d = (size_t)gs * ((g.n_spheres * 3) | 1);
  if (coop)
    d += (size_t)gs * g.n_links * 12 + (size_t)(nt / 32) * warp_fk_staging_doubles(g.n_links);
  else
    d += (size_t)g.n_slots * 12 * gs + (lf ? (size_t)gs * g.n_glinks * 12 : 0);
  size_t b = d * sizeof(double) + (size_t)gs * g.n_vars * q_elem;
  if (lf)
    b += 2 * (size_t)gs * (g.n_spheres | 1) * sizeof(unsigned);
  b += 16 + (g.n_nbr <= kMaxStagedNbr ? (size_t)g.n_nbr * sizeof(int4) : 0) +
       (size_t)(g.n_glinks + 1 + g.n_spheres) * sizeof(int);
  b += 8 + (size_t)((wf.d2 ? wf.max_d2 + 1 : 0) + (sf.d2 ? sf.max_d2 + 1 : 0)) * sizeof(double);
  return b;
}

template <typename QT, typename FT, typename OT, bool LF, int GS, int NT, bool COOP>
int launch_gather_variant(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, const void* q, int64_t n,
                          int flags, OT* dist, OT* grad, uint8_t* type, OT* centers, cudaStream_t stream, size_t bytes)
{
  B200DF_CUDA(allow_dynamic_smem(gradients_gather_kernel<QT, FT, OT, LF, GS, NT, COOP>, bytes));
  const unsigned grid = (unsigned)((n + GS - 1) / GS);
  gradients_gather_kernel<QT, FT, OT, LF, GS, NT, COOP><<<grid, NT, bytes, stream>>>(
      g, wf, sf, static_cast<const QT*>(q), n, flags, dist, grad, type, centers);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}

template <typename QT, typename FT, typename OT>
int launch_gather(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, const void* q, int64_t n, int flags,
                  OT* dist, OT* grad, uint8_t* type, OT* centers, cudaStream_t stream)
{
  const bool lf = (flags & B200DF_CHECK_LINK_FIELDS) != 0;
  const size_t coop_bytes = gather_smem_bytes(g, wf, sf, sizeof(QT), lf, kCoopStates, kCoopThreads, true);
  const size_t big_bytes = gather_smem_bytes(g, wf, sf, sizeof(QT), lf, kGradStates, kGradThreads, false);
  // small calls, and robots whose 32-state tile does not fit an SM's shared memory at any size
  if ((n <= kCoopMaxStates || big_bytes > 220 * 1024) && coop_bytes <= 200 * 1024)
  {
    if (lf)
      return launch_gather_variant<QT, FT, OT, true, kCoopStates, kCoopThreads, true>(
          g, wf, sf, q, n, flags, dist, grad, type, centers, stream, coop_bytes);
    return launch_gather_variant<QT, FT, OT, false, kCoopStates, kCoopThreads, true>(
        g, wf, sf, q, n, flags, dist, grad, type, centers, stream, coop_bytes);
  }
  const size_t bytes = big_bytes;
  if (bytes > 220 * 1024)
  {
    set_error("robot too large for the gradient kernel's shared-memory layout (%zu bytes)", bytes);
    return B200DF_ERR_UNSUPPORTED;
  }
  if (lf)
    return launch_gather_variant<QT, FT, OT, true, kGradStates, kGradThreads, false>(
        g, wf, sf, q, n, flags, dist, grad, type, centers, stream, bytes);
  return launch_gather_variant<QT, FT, OT, false, kGradStates, kGradThreads, false>(
      g, wf, sf, q, n, flags, dist, grad, type, centers, stream, bytes);
}

template <typename OT>
int dispatch_gather(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, int elem, const void* q,
                    int q_dtype, int64_t n, int flags, OT* dist, OT* grad, uint8_t* type, OT* centers,
                    cudaStream_t stream)
{
  if (q_dtype == B200DF_Q_F32 && elem == 1)
    return launch_gather<float, uint8_t, OT>(g, wf, sf, q, n, flags, dist, grad, type, centers, stream);
  if (q_dtype == B200DF_Q_F32)
    return launch_gather<float, uint16_t, OT>(g, wf, sf, q, n, flags, dist, grad, type, centers, stream);
  if (elem == 1)
    return launch_gather<double, uint8_t, OT>(g, wf, sf, q, n, flags, dist, grad, type, centers, stream);
  return launch_gather<double, uint16_t, OT>(g, wf, sf, q, n, flags, dist, grad, type, centers, stream);
}

}  // namespace

extern "C" {

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
  const size_t qbytes = (size_t)n * group->dev.n_vars * q_elem_size(q_dtype);
  const size_t obytes = (size_t)n * out_per_state * sizeof(double);
  typedef b200df_group::StagePool Pool;
  Pool::Ctx* ctx = group->stage->acquire();
  cudaError_t e = Pool::ensure_stream(ctx, 0);
  int rc = B200DF_OK;
  // small calls: joint values in and results out through the context's mapped pinned buffer (one launch, one sync)
  const size_t o_out = (qbytes + 255) & ~(size_t)255;
  const bool mapped = o_out + obytes <= kSmallGradBytes && small_batch_enabled();
  void *dq = nullptr, *dout = nullptr;
  if (e == cudaSuccess && mapped)
  {
    void* dpin = nullptr;
    e = Pool::ensure_pin(ctx, o_out + obytes, &dpin);
    if (e == cudaSuccess)
    {
      std::memcpy(ctx->pin, q, qbytes);
      dq = dpin;
      dout = static_cast<char*>(dpin) + o_out;
    }
  }
  else if (e == cudaSuccess)
  {
    e = Pool::ensure_scratch(ctx, 0, std::max<size_t>(qbytes, 1));
    if (e == cudaSuccess)
      e = Pool::ensure_scratch(ctx, 1, std::max<size_t>(obytes, 1));
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ctx->scratch[0], q, qbytes, cudaMemcpyHostToDevice, ctx->stream[0]);
    dq = ctx->scratch[0];
    dout = ctx->scratch[1];
  }
  if (e == cudaSuccess)
  {
    double* o = static_cast<double*>(dout);
    cudaStream_t stream = ctx->stream[0];
    if (mode == OUT_CENTERS)
      rc = dispatch_exact<OUT_CENTERS>(group->dev, wf, sf, elem, dq, q_dtype, n, 0, nullptr, o, stream);
    else if (mode == OUT_TRANSFORMS)
      rc = dispatch_exact<OUT_TRANSFORMS>(group->dev, wf, sf, elem, dq, q_dtype, n, 0, nullptr, o, stream);
    else
      rc = dispatch_exact<OUT_CLEARANCE>(group->dev, wf, sf, elem, dq, q_dtype, n, 0, nullptr, o, stream);
    if (!rc && !mapped)
      e = cudaMemcpyAsync(out, dout, obytes, cudaMemcpyDeviceToHost, stream);
    const cudaError_t es = cudaStreamSynchronize(stream);
    if (e == cudaSuccess)
      e = es;
    if (!rc && e == cudaSuccess && mapped)
      std::memcpy(out, static_cast<char*>(ctx->pin) + o_out, obytes);
  }
  group->stage->give_back(ctx);
  if (!rc && e != cudaSuccess)
  {
    set_error("staging / running / reading back the batch failed: %s", cudaGetErrorString(e));
    rc = B200DF_ERR_CUDA;
  }
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
  // an empty group over the model gives the kernel its link table (built once per model)
  b200df_group* g = nullptr;
  {
    std::lock_guard<std::mutex> lk(model->fk_mu);
    if (!model->fk_group)
    {
      b200df_group_desc gd{};
      gd.n_group_links = 0;
      gd.resolution = 1.0;
      const int rc = b200df_group_create(model, &gd, &model->fk_group);
      if (rc)
        return rc;
    }
    g = model->fk_group;
  }
  return run_f64_output(g, nullptr, q, q_dtype, n, OUT_TRANSFORMS, (size_t)model->n_links * 12, link_transforms);
}

int b200df_group_link_count(const b200df_group* g, int32_t* n)
{
  B200DF_REQUIRE(g && n, "null argument");
  *n = g->dev.n_glinks;
  return B200DF_OK;
}

int b200df_group_set_link_fields(b200df_group* group, b200df_field* const* fields, const uint8_t* enabled)
{
  B200DF_REQUIRE(group, "null group");
  DeviceGuard dg(group->device);
  cudaFree(group->d_link_fields);
  cudaFree(group->d_lf_enabled);
  group->d_link_fields = group->d_lf_enabled = nullptr;
  group->dev.link_fields = nullptr;
  group->dev.lf_enabled = nullptr;
  group->dev.n_link_fields = 0;
  group->link_field_elem = 0;
  if (!fields)
    return B200DF_OK;
  B200DF_REQUIRE(enabled, "null enabled matrix");
  const int G = group->dev.n_glinks;
  B200DF_REQUIRE(G <= 32, "the link-field term supports at most 32 links with geometry per group");
  std::vector<FieldDev> views(std::max(G, 1));
  std::vector<unsigned> en(std::max(G, 1), 0u);
  std::memset(views.data(), 0, views.size() * sizeof(FieldDev));
  for (int j = 0; j < G; ++j)
  {
    if (!fields[j])
      continue;
    B200DF_REQUIRE(fields[j]->device == group->device, "link field and group live on different devices");
    B200DF_REQUIRE(fields[j]->desc.resolution == group->desc.resolution, "link field resolution differs from the group's");
    int rc = field_prepare(fields[j], &views[j], nullptr);
    if (rc)
      return rc;
    if (!views[j].d2)
      continue;  // empty grid
    B200DF_REQUIRE(group->link_field_elem == 0 || group->link_field_elem == views[j].elem_size,
                   "link fields must share max_distance (voxel type)");
    group->link_field_elem = views[j].elem_size;
  }
  for (int i = 0; i < G; ++i)
    for (int j = 0; j < G; ++j)
      if (i != j && enabled[(size_t)i * G + j])
        en[i] |= 1u << j;
  B200DF_CUDA(cudaMalloc(&group->d_link_fields, views.size() * sizeof(FieldDev)));
  B200DF_CUDA(cudaMalloc(&group->d_lf_enabled, en.size() * sizeof(unsigned)));
  B200DF_CUDA(cudaMemcpy(group->d_link_fields, views.data(), views.size() * sizeof(FieldDev), cudaMemcpyHostToDevice));
  B200DF_CUDA(cudaMemcpy(group->d_lf_enabled, en.data(), en.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
  group->dev.link_fields = static_cast<const FieldDev*>(group->d_link_fields);
  group->dev.lf_enabled = static_cast<const unsigned*>(group->d_lf_enabled);
  group->dev.n_link_fields = G;
  return B200DF_OK;
}

// voxel type the gradient kernel is instantiated for: world / self fields and the link fields must agree
static int resolve_elem(const b200df_group* group, int flags, int* elem)
{
  if (!(flags & B200DF_CHECK_LINK_FIELDS) || group->link_field_elem == 0)
    return B200DF_OK;
  if (flags & (B200DF_CHECK_ROBOT | B200DF_CHECK_SELF))
    B200DF_REQUIRE(*elem == group->link_field_elem, "link fields and world/self field must share max_distance");
  *elem = group->link_field_elem;
  return B200DF_OK;
}

}  // extern "C"

template <typename OT>
static int gradients_device_impl(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q_dev,
                                 int q_dtype, int64_t n, int flags, OT* dist_dev, OT* grad_dev, uint8_t* type_dev,
                                 OT* centers_dev, void* stream)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q_dev && dist_dev && grad_dev && type_dev)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  int elem;
  int rc = prepare_fields(group, world, self, flags, &wf, &sf, &elem);
  if (!rc)
    rc = resolve_elem(group, flags, &elem);
  if (rc)
    return rc;
  DeviceGuard dg(group->device);
  return dispatch_gather(group->dev, wf, sf, elem, q_dev, q_dtype, n, flags, dist_dev, grad_dev, type_dev, centers_dev,
                         static_cast<cudaStream_t>(stream));
}

// Host-buffer entry: chunks of states travel H2D -> kernel -> D2H on two streams, so the (large) result copies of
// one chunk overlap the kernel of the next.
template <typename OT>
static int gradients_host_impl(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                               int q_dtype, int64_t n, int flags, OT* dist, OT* grad, uint8_t* type, OT* centers)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q && dist && grad && type)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  int elem;
  int rc = prepare_fields(group, world, self, flags, &wf, &sf, &elem);
  if (!rc)
    rc = resolve_elem(group, flags, &elem);
  if (rc)
    return rc;
  const GroupDev& g = group->dev;
  const size_t S = (size_t)g.n_spheres;
  DeviceGuard dg(group->device);
  const size_t row = (size_t)g.n_vars * q_elem_size(q_dtype);
  typedef b200df_group::StagePool Pool;
  Pool::Ctx* ctx = group->stage->acquire();
  cudaError_t e = cudaSuccess;
  auto finish = [&](int status) {
    group->stage->give_back(ctx);
    if (!status && e != cudaSuccess)
    {
      set_error("gradient batch failed while staging / running / reading back: %s", cudaGetErrorString(e));
      status = B200DF_ERR_CUDA;
    }
    return status;
  };
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };

  // Small calls (CHOMP evaluates one trajectory, ~100 waypoints, per iteration): one mapped pinned buffer carries the
  // joint values in and the results out, so the call is one launch and one synchronise instead of five staged copies.
  const size_t o_q = 0, o_d = up((size_t)n * row), o_g = o_d + up(n * S * sizeof(OT)),
               o_t = o_g + up(n * S * 3 * sizeof(OT)), o_c = o_t + up(n * S),
               total = o_c + (centers ? up(n * S * 3 * sizeof(OT)) : 0);
  if (total <= kSmallGradBytes && small_batch_enabled())
  {
    void* dpin = nullptr;
    e = Pool::ensure_stream(ctx, 0);
    if (e == cudaSuccess)
      e = Pool::ensure_pin(ctx, total, &dpin);
    if (e != cudaSuccess)
      return finish(B200DF_OK);
    char* hp = static_cast<char*>(ctx->pin);
    char* dp = static_cast<char*>(dpin);
    std::memcpy(hp + o_q, q, (size_t)n * row);
    // result arrays the caller already keeps in pinned memory (b200df_host_alloc) are written in place
    auto pinned_alias = [](void* host) -> void* {
      cudaPointerAttributes a;
      if (host && cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost)
        return a.devicePointer;
      cudaGetLastError();
      return nullptr;
    };
    void* a_d = pinned_alias(dist);
    void* a_g = pinned_alias(grad);
    void* a_t = pinned_alias(type);
    void* a_c = pinned_alias(centers);
    rc = dispatch_gather(g, wf, sf, elem, dp + o_q, q_dtype, n, flags, static_cast<OT*>(a_d ? a_d : dp + o_d),
                         static_cast<OT*>(a_g ? a_g : dp + o_g), static_cast<uint8_t*>(a_t ? a_t : dp + o_t),
                         centers ? static_cast<OT*>(a_c ? a_c : dp + o_c) : nullptr, ctx->stream[0]);
    e = cudaStreamSynchronize(ctx->stream[0]);
    if (!rc && e == cudaSuccess)
    {
      if (!a_d)
        std::memcpy(dist, hp + o_d, n * S * sizeof(OT));
      if (!a_g)
        std::memcpy(grad, hp + o_g, n * S * 3 * sizeof(OT));
      if (!a_t)
        std::memcpy(type, hp + o_t, n * S);
      if (centers && !a_c)
        std::memcpy(centers, hp + o_c, n * S * 3 * sizeof(OT));
    }
    return finish(rc);
  }

  // Large calls: chunks of states travel H2D -> kernel -> D2H on two streams of the cached context, so the (large)
  // result copies of one chunk overlap the kernel of the next.  Result arrays in ordinary pageable memory (what a
  // planner's own vectors are) are not handed to the driver, whose staged copy runs at a third of the PCIe rate: each
  // chunk lands in a page-locked slot and a few host threads move it on while the next chunk computes and copies.
  const int64_t chunk = std::min<int64_t>(n, 16384);
  const size_t want[5] = { std::max<size_t>(chunk * row, 1), chunk * S * sizeof(OT), chunk * S * 3 * sizeof(OT),
                           chunk * S, centers ? chunk * S * 3 * sizeof(OT) : 0 };
  for (int k = 0; k < 2 && e == cudaSuccess; ++k)
  {
    e = Pool::ensure_stream(ctx, k);
    for (int b = 0; b < 5 && e == cudaSuccess; ++b)
      if (want[b])
        e = Pool::ensure_scratch(ctx, 5 * k + b, want[b]);
  }
  auto is_pinned = [](const void* host) {
    cudaPointerAttributes a;
    const bool yes = host && cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return yes;
  };
  static const bool stage_env = !(getenv("B200DF_GRAD_STAGE_OUT") && atoi(getenv("B200DF_GRAD_STAGE_OUT")) == 0);
  const bool stage_out = stage_env && !(is_pinned(dist) && is_pinned(grad) && is_pinned(type) && (!centers || is_pinned(centers)));
  const size_t s_d = 0, s_g = s_d + up(want[1]), s_t = s_g + up(want[2]), s_c = s_t + up(want[3]),
               slot_bytes = s_c + up(want[4]);
  char* pin = nullptr;
  if (stage_out && e == cudaSuccess)
  {
    void* dpin = nullptr;
    e = Pool::ensure_pin(ctx, 2 * slot_bytes, &dpin);
    pin = static_cast<char*>(ctx->pin);
  }
  auto drain = [&](int slot, int64_t s, int64_t m) {  // chunk [s, s + m) has landed in `slot`: on to the caller's arrays
    const cudaError_t es = cudaStreamSynchronize(ctx->stream[slot]);
    if (es != cudaSuccess)
    {
      if (e == cudaSuccess)
        e = es;
      return;
    }
    const char* src = pin + (size_t)slot * slot_bytes;
    const CopyJob jobs[4] = { { dist + s * S, src + s_d, (size_t)m * S * sizeof(OT) },
                              { grad + s * S * 3, src + s_g, (size_t)m * S * 3 * sizeof(OT) },
                              { type + s * S, src + s_t, (size_t)m * S },
                              { centers ? centers + s * S * 3 : nullptr, src + s_c,
                                centers ? (size_t)m * S * 3 * sizeof(OT) : 0 } };
    parallel_copy_many(jobs, 4);
  };
  int k = 0;
  int64_t prev_s = -1, prev_m = 0;
  for (int64_t s = 0; s < n && !rc && e == cudaSuccess; s += chunk, k ^= 1)
  {
    const int64_t m = std::min<int64_t>(chunk, n - s);
    void** buf = ctx->scratch + 5 * k;
    cudaStream_t st = ctx->stream[k];
    e = cudaMemcpyAsync(buf[0], static_cast<const char*>(q) + (size_t)s * row, (size_t)m * row, cudaMemcpyHostToDevice,
                        st);
    if (e != cudaSuccess)
      break;
    rc = dispatch_gather(g, wf, sf, elem, buf[0], q_dtype, m, flags, static_cast<OT*>(buf[1]),
                         static_cast<OT*>(buf[2]), static_cast<uint8_t*>(buf[3]),
                         centers ? static_cast<OT*>(buf[4]) : nullptr, st);
    if (rc)
      break;
    char* slot = stage_out ? pin + (size_t)k * slot_bytes : nullptr;
    e = cudaMemcpyAsync(stage_out ? static_cast<void*>(slot + s_d) : static_cast<void*>(dist + s * S), buf[1],
                        m * S * sizeof(OT), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(stage_out ? static_cast<void*>(slot + s_g) : static_cast<void*>(grad + s * S * 3), buf[2],
                          m * S * 3 * sizeof(OT), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(stage_out ? static_cast<void*>(slot + s_t) : static_cast<void*>(type + s * S), buf[3], m * S,
                          cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && centers)
      e = cudaMemcpyAsync(stage_out ? static_cast<void*>(slot + s_c) : static_cast<void*>(centers + s * S * 3), buf[4],
                          m * S * 3 * sizeof(OT), cudaMemcpyDeviceToHost, st);
    if (stage_out)
    {
      // the chunk before this one sits in the other slot: move it on while this one computes and copies
      if (prev_s >= 0 && e == cudaSuccess)
        drain(k ^ 1, prev_s, prev_m);
      prev_s = s;
      prev_m = m;
    }
  }
  if (stage_out && prev_s >= 0 && !rc && e == cudaSuccess)
    drain(k ^ 1, prev_s, prev_m);
  for (int i = 0; i < 2; ++i)
  {
    if (!ctx->stream[i])
      continue;
    const cudaError_t es = cudaStreamSynchronize(ctx->stream[i]);
    if (e == cudaSuccess)
      e = es;
  }
  return finish(rc);
}

extern "C" {
int b200df_gradients_batch_device(const b200df_group* group, b200df_field* world, b200df_field* self,
                                  const void* q_dev, int q_dtype, int64_t n, int flags, double* dist_dev,
                                  double* grad_dev, uint8_t* type_dev, double* centers_dev, void* stream)
{
  return gradients_device_impl<double>(group, world, self, q_dev, q_dtype, n, flags, dist_dev, grad_dev, type_dev,
                                       centers_dev, stream);
}

int b200df_gradients_batch(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                           int q_dtype, int64_t n, int flags, double* dist, double* grad, uint8_t* type,
                           double* centers)
{
  return gradients_host_impl<double>(group, world, self, q, q_dtype, n, flags, dist, grad, type, centers);
}

int b200df_gradients_batch_device_f32(const b200df_group* group, b200df_field* world, b200df_field* self,
                                      const void* q_dev, int q_dtype, int64_t n, int flags, float* dist_dev,
                                      float* grad_dev, uint8_t* type_dev, float* centers_dev, void* stream)
{
  return gradients_device_impl<float>(group, world, self, q_dev, q_dtype, n, flags, dist_dev, grad_dev, type_dev,
                                      centers_dev, stream);
}

int b200df_gradients_batch_f32(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                               int q_dtype, int64_t n, int flags, float* dist, float* grad, uint8_t* type,
                               float* centers)
{
  return gradients_host_impl<float>(group, world, self, q, q_dtype, n, flags, dist, grad, type, centers);
}

}  // extern "C"