// FAST precision of the batched validity check: the same fused sweep as validity.cu (FK, sphere posing, field
// lookups, ACM-pruned intra-group pairs, one thread per state) in fp32 with FMA, plus a proof obligation: every
// comparison the verdict depends on carries an error margin derived on the host from a forward error analysis of
// this robot's chain.  A state gets
//     1  when some test collides by more than its margin  (the exact evaluation collides too),
//     0  when every test is clear by more than its margin  (the exact evaluation is clear too),
//     2  otherwise; those states are appended to a list and re-run by the exact fp64 kernel of validity.cu.
// So the verdicts that leave b200df_check_batch are those of the exact path, bit for bit, at fp32 cost for all but
// the ~1 % of states that sit within micrometres of a voxel face or a sphere-pair threshold.
//
// Reference being replaced: the same functions as validity.cu (robot_state.cpp:718-755, *_joint_model.cpp,
// collision_distance_field_types.cpp:206-231,390-420, distance_field.cpp:62-86,
// collision_env_distance_field.cpp:277-356,433-660,1580-1662).
//
// This translation unit is compiled WITH fused multiply-add (no --fmad=false): nothing here has to round like the
// reference, it only has to stay inside the margins.
#include <algorithm>
#include <type_traits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "sincos_small.cuh"
#include "validity_common.cuh"

using namespace b200df;

namespace
{
constexpr int kMaxLinks = 40;
constexpr int kMaxGLinks = 40;
constexpr int kMaxSpheres = 128;
constexpr int kMaxPartners = 512;
constexpr int kBatch = 4;
constexpr double kU = 5.9604644775390625e-08;  // 2^-24, unit round-off of fp32

struct FLink
{
  int joint_type, jkind, var, is_mimic;
  int origin_is_identity, parent, parent_is_prev, store_slot, parent_slot, gslot;
  double mimic_mult, mimic_off;  // evaluated in fp64 like the exact kernel, then rounded once
  float ax, ay, az, x2, y2, z2, xy, xz, yz;
  float origin[12];
  float vmax;  // joint values beyond this leave the analysed range -> the state is sent to the exact kernel
};

struct FGLink
{
  int sphere_begin, sphere_end, self_enabled, store_link, store_sphere, partner_begin, partner_end;
  int pair_const_begin, pair_const_stride;
  float bs[3];
};

struct FSphere
{
  float x, y, z;
  float eps;  // bound on the distance between the fp32 centre and the exact (fp64) centre, metres
  int thr;    // collide iff d2 < thr (same host-resolved threshold as the exact kernel)
};

struct FPartner
{
  int store_link, store_sphere, count, const_offset;
  // negated bit patterns of the squared-distance thresholds below which the prune certainly / possibly passes:
  // min(exact-cull threshold + margin, doBoundingSpheresIntersect threshold (Q2) - / + its margin)
  int n_sure, n_maybe;
  int div_magic;  // floor(idx / count) == (idx * div_magic) >> 16 for every sphere-pair index of this link pair
  int pad;
};

struct FastTables  // passed by value: lives in the constant bank, read with warp-uniform indices
{
  int n_links, n_vars, n_slots, n_glinks, n_spheres, n_store_links, n_store_spheres, n_partners;
  FLink links[kMaxLinks];
  FGLink glinks[kMaxGLinks];
  FSphere sph[kMaxSpheres];
  FPartner partners[kMaxPartners];
};

struct FastLookup
{
  float oo;       // 1/resolution
  float k[3];     // -origin_minus/resolution: grid coordinate g = p*oo + k
  float gerr;     // rounding of that evaluation, in cells
  unsigned lim[3];
  int stride1, stride2;
  const void* d2;
};

// ------------------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mulf(const float* a, const float* b, float* r)  // 3x4 * 3x4 (implied last row)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const float a0 = a[4 * i], a1 = a[4 * i + 1], a2 = a[4 * i + 2];
    r[4 * i + 0] = fmaf(a2, b[8], fmaf(a1, b[4], a0 * b[0]));
    r[4 * i + 1] = fmaf(a2, b[9], fmaf(a1, b[5], a0 * b[1]));
    r[4 * i + 2] = fmaf(a2, b[10], fmaf(a1, b[6], a0 * b[2]));
    r[4 * i + 3] = fmaf(a2, b[11], fmaf(a1, b[7], fmaf(a0, b[3], a[4 * i + 3])));
  }
}

__device__ __forceinline__ void applyf(const float* t, float x, float y, float z, float& ox, float& oy, float& oz)
{
  ox = fmaf(t[2], z, fmaf(t[1], y, fmaf(t[0], x, t[3])));
  oy = fmaf(t[6], z, fmaf(t[5], y, fmaf(t[4], x, t[7])));
  oz = fmaf(t[10], z, fmaf(t[9], y, fmaf(t[8], x, t[11])));
}

// joint value in fp32; mimic joints evaluate mult*v+off in fp64 like the exact kernel and round once
template <typename QP>
__device__ __forceinline__ float jointf(const FLink& l, const QP* q, int k, bool& out_of_range)
{
  float v = (float)q[l.var + k];  // fp64 joint values are rounded once (covered by the margins)
  if (l.is_mimic)
    v = (float)(l.mimic_mult * (double)v + l.mimic_off);
  out_of_range |= !(fabsf(v) <= l.vmax);
  return v;
}

// FULL: floating joints, planar joints and revolute joints about an oblique axis are compiled only into the kernel
// variant for robots that have one (the hot kernel is sensitive to code size: the quaternion branch alone cost 6 % on a
// Panda, whose joints are all fixed, prismatic or revolute about a coordinate axis)
// saved: when non-null the robot has at most one branch point and its transform is parked in these 12 registers
// instead of a shared-memory slot (the pair pass of the two-pass form: 48 bytes per state less)
template <bool FULL, typename QP>
__device__ __forceinline__ void fk_step_f(const FLink& l, const QP* q, float* slots, int B, float* cur,
                                          bool& out_of_range, float* saved = nullptr)
{
  float base[12];
  if (l.parent < 0)
  {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      base[k] = l.origin[k];  // identity when origin_is_identity
  }
  else
  {
    float par[12];
    if (l.parent_is_prev)
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        par[k] = cur[k];
    }
    else if (saved)
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        par[k] = saved[k];
    }
    else
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        par[k] = slots[(l.parent_slot * 12 + k) * B];
    }
    if (l.origin_is_identity)
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        base[k] = par[k];
    }
    else
      mulf(par, l.origin, base);
  }
  if (l.joint_type == B200DF_JOINT_FIXED)
  {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      cur[k] = base[k];
  }
  else if (l.joint_type == B200DF_JOINT_PRISMATIC)
  {
    const float v = jointf(l, q, 0, out_of_range);
    const float tx = l.ax * v, ty = l.ay * v, tz = l.az * v;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
      cur[4 * i + 0] = base[4 * i + 0];
      cur[4 * i + 1] = base[4 * i + 1];
      cur[4 * i + 2] = base[4 * i + 2];
      cur[4 * i + 3] = fmaf(base[4 * i + 2], tz, fmaf(base[4 * i + 1], ty, fmaf(base[4 * i], tx, base[4 * i + 3])));
    }
  }
  else if (FULL && l.joint_type == B200DF_JOINT_FLOATING)  // translation + normalised quaternion (floating_joint_model.cpp:224-229)
  {
    float j[12];
    j[3] = jointf(l, q, 0, out_of_range);
    j[7] = jointf(l, q, 1, out_of_range);
    j[11] = jointf(l, q, 2, out_of_range);
    float qx = jointf(l, q, 3, out_of_range), qy = jointf(l, q, 4, out_of_range), qz = jointf(l, q, 5, out_of_range),
          qw = jointf(l, q, 6, out_of_range);
    const float n2 = fmaf(qw, qw, fmaf(qz, qz, fmaf(qy, qy, qx * qx)));
    out_of_range |= !(n2 > 1e-6f && n2 < 1e6f);  // the error analysis assumes a quaternion that can be normalised in fp32
    const float inv = 1.f / sqrtf(n2);
    qx *= inv;
    qy *= inv;
    qz *= inv;
    qw *= inv;
    const float tx = 2.f * qx, ty = 2.f * qy, tz = 2.f * qz;
    const float twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const float txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    j[0] = 1.f - (tyy + tzz);
    j[1] = txy - twz;
    j[2] = txz + twy;
    j[4] = txy + twz;
    j[5] = 1.f - (txx + tzz);
    j[6] = tyz - twx;
    j[8] = txz - twy;
    j[9] = tyz + twx;
    j[10] = 1.f - (txx + tyy);
    mulf(base, j, cur);
  }
  else  // revolute, or planar = translation (x, y) followed by a rotation about z
  {
    const bool planar = FULL && l.joint_type == B200DF_JOINT_PLANAR;
    float px = 0.f, py = 0.f;
    if (planar)
    {
      px = jointf(l, q, 0, out_of_range);
      py = jointf(l, q, 1, out_of_range);
    }
    const float v = jointf(l, q, planar ? 2 : 0, out_of_range);
    float s, c;
    if (FULL)
      sincosf(v, &s, &c);  // planar theta is analysed up to 16 rad
    else
      sincos_small(v, &s, &c);  // |v| <= kAngleRange or the state is flagged; no large-argument path to carry around
    const int kind = planar ? 3 : l.jkind;
    if (FULL && kind == 0)
    {
      const float tt = 1.f - c;
      float j[12];
      j[0] = fmaf(tt, l.x2, c);
      j[1] = fmaf(tt, l.xy, -l.az * s);
      j[2] = fmaf(tt, l.xz, l.ay * s);
      j[4] = fmaf(tt, l.xy, l.az * s);
      j[5] = fmaf(tt, l.y2, c);
      j[6] = fmaf(tt, l.yz, -l.ax * s);
      j[8] = fmaf(tt, l.xz, -l.ay * s);
      j[9] = fmaf(tt, l.yz, l.ax * s);
      j[10] = fmaf(tt, l.z2, c);
      j[3] = j[7] = j[11] = 0.f;
      mulf(base, j, cur);
    }
    else
    {
      // rotation by v about +-x / +-y / +-z: the axis sign only flips the sine
      const float sg = planar ? s : (kind == 1 ? l.ax : (kind == 2 ? l.ay : l.az)) * s;
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        const float a0 = base[4 * i], a1 = base[4 * i + 1], a2 = base[4 * i + 2];
        float r0, r1, r2;
        if (kind == 3)
        {
          r0 = fmaf(a1, sg, a0 * c);
          r1 = fmaf(a1, c, -a0 * sg);
          r2 = a2;
        }
        else if (kind == 2)
        {
          r0 = fmaf(a2, -sg, a0 * c);
          r1 = a1;
          r2 = fmaf(a2, c, a0 * sg);
        }
        else
        {
          r0 = a0;
          r1 = fmaf(a2, sg, a1 * c);
          r2 = fmaf(a2, c, -a1 * sg);
        }
        float t3 = base[4 * i + 3];
        if (planar)
          t3 = fmaf(a1, py, fmaf(a0, px, t3));
        cur[4 * i + 0] = r0;
        cur[4 * i + 1] = r1;
        cur[4 * i + 2] = r2;
        cur[4 * i + 3] = t3;
      }
    }
  }
  if (l.store_slot >= 0)
  {
    if (saved)
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        saved[k] = cur[k];
    }
    else
    {
#pragma unroll
      for (int k = 0; k < 12; ++k)
        slots[(l.store_slot * 12 + k) * B] = cur[k];
    }
  }
}

// one axis of the world->grid mapping: the cell and the position inside it (0 .. 1)
__device__ __forceinline__ int axis_cell(float p, float oo, float k, float& frac)
{
  const float g = fmaf(p, oo, k);
  const float fl = floorf(g);
  frac = g - fl;
  return (int)fl;  // |g| far outside the grid saturates and fails the range test like the exact kernel
}
// could the exact cell be the lower (-1) / upper (+1) neighbour?
__device__ __forceinline__ int axis_alt(float frac, float delta)
{
  return frac < delta ? -1 : (frac > 1.f - delta ? 1 : 0);
}

// Flags are carried as signed margins and combined with integer min / max (one instruction each) instead of bools
// (compare + select + and/or chains): a margin < 0 means "collides".
// margin of one cell: d2 - thr inside the inset box, a large positive number outside it (never a collision, Q4)
template <typename FT>
__device__ __forceinline__ int cell_margin(const FastLookup& f, int ix, int iy, int iz, int thr)
{
  const bool in = ((unsigned)(ix - 1) < f.lim[0]) & ((unsigned)(iy - 1) < f.lim[1]) & ((unsigned)(iz - 1) < f.lim[2]);
  const unsigned idx = in ? (unsigned)(ix * f.stride1 + iy * f.stride2 + iz) : 0u;
  B200DF_CHECK_INDEX(idx, (long long)(f.lim[0] + 2) * f.stride1);
  const int v = (int)static_cast<const FT*>(f.d2)[idx] - thr;
  return in ? v : 0x3fffffff;
}

// Sphere vs field.  sure (< 0): the exact evaluation certainly reports this sphere colliding; maybe (< 0): it could.
template <typename FT>
__device__ __forceinline__ void sphere_field(const FastLookup& f, float x, float y, float z, float eps, int thr,
                                             int& sure, int& maybe)
{
  const float delta = fmaf(eps, f.oo, f.gerr);
  float fx, fy, fz;
  const int ix = axis_cell(x, f.oo, f.k[0], fx);
  const int iy = axis_cell(y, f.oo, f.k[1], fy);
  const int iz = axis_cell(z, f.oo, f.k[2], fz);
  const int s0 = cell_margin<FT>(f, ix, iy, iz, thr);
  // the common case: the centre is clear of every voxel face, |frac - 1/2| <= 1/2 - delta on all three axes (one
  // subtraction and one compare per axis; the per-axis directions are only worked out on the rare path).  The two
  // roundings of this form (frac - 1/2 and 1/2 - delta, <= 1.5e-8 each) are covered by 1.2e-7 of extra margin, which
  // sends 4e-5 more of the states to the exact kernel.
  const float clear = 0.5f - delta - 1.2e-7f;
  if ((fabsf(fx - 0.5f) <= clear) & (fabsf(fy - 0.5f) <= clear) & (fabsf(fz - 0.5f) <= clear))
  {
    sure = min(sure, s0);
    maybe = min(maybe, s0);
    return;
  }
  const int ax = axis_alt(fx, delta), ay = axis_alt(fy, delta), az = axis_alt(fz, delta);
  // near a face: the exact cell is one of up to 8 candidates; the verdict is settled only if they all agree
  int worst = s0, best = s0;  // all collide <=> the largest margin is negative; any collides <=> the smallest is
  for (int m = 1; m < 8; ++m)
  {
    if (((m & 1) && !ax) || ((m & 2) && !ay) || ((m & 4) && !az))
      continue;
    const int sm = cell_margin<FT>(f, ix + ((m & 1) ? ax : 0), iy + ((m & 2) ? ay : 0), iz + ((m & 4) ? az : 0), thr);
    worst = max(worst, sm);
    best = min(best, sm);
  }
  sure = min(sure, worst);
  maybe = min(maybe, best);
}

// SELF: the self-field lookups are compiled in only when the call asks for them (code size again)
// PAIRS: likewise the whole intra-group pair machinery (a checkRobotCollision batch does not need it)
// MODE 0: one sweep; 1: first pass of two (appends the states left at 0 to surv_idx); 2: second pass (walks `list`)
template <typename QT, typename FT, int B, bool PAIRS, bool PAIR_SMEM, bool FULL, bool SELF, int MODE>
__global__ void __launch_bounds__(B) validity_fast_kernel(const __grid_constant__ FastTables T, FastLookup wf,
                                                          FastLookup sf, const int2* __restrict__ pair_T_global,
                                                          const QT* __restrict__ q, long n, int flags,
                                                          uint8_t* __restrict__ verdict, int* __restrict__ amb_idx,
                                                          int* __restrict__ amb_count, int n_pair_smem,
                                                          const int* __restrict__ list,
                                                          const int* __restrict__ list_count,
                                                          int* __restrict__ surv_idx, int* __restrict__ surv_count)
{
  // Two-pass form (fast_launch): pass A (field lookups only, no pair code, three times the resident warps) appends the
  // states it leaves at verdict 0 to surv_idx; pass B (pairs only) walks that list - `list` non-null: slot -> state
  // through the list, rows gathered, only verdicts 1 / 2 written.  A state that already collides with the world costs
  // pass B nothing, where the fused sweep drags it along as a masked lane until the whole warp collides.
  extern __shared__ float fsm[];
  constexpr bool LIST = MODE == 2;
  const int tid = threadIdx.x;
  const long base = (long)blockIdx.x * B;
  const long total = LIST ? (long)*list_count : n;
  if (LIST && base >= total)
    return;
  // (kept as a run-time value in the PAIRS variant: the constant-folded form of that variant measured 1.5 % slower)
  const bool do_pairs = PAIRS && (flags & B200DF_CHECK_INTRA) != 0 && T.n_partners > 0;
  // the pair pass keeps a lone branch-point transform in registers and reads its joint values straight from the
  // gathered rows (L1 / L2 hits: the field pass has just read them): 84 bytes of shared memory per state less, one
  // more CTA per SM
  const bool slot_in_regs = LIST && T.n_slots <= 1;
  float* slots = fsm;                                                    // [n_slots][12][B]
  float* stc = slots + (slot_in_regs ? 0 : T.n_slots * 12 * B);          // [n_store_spheres][3][B]
  float* stb = stc + (do_pairs ? T.n_store_spheres * 3 * B : 0);         // [n_store_links][3][B]
  float* qs = stb + (do_pairs ? T.n_store_links * 3 * B : 0);            // [B][n_vars], not in the pair pass
  const bool active = base + tid < total;
  const long state = LIST ? (active ? (long)list[base + tid] : 0L) : base + tid;
  B200DF_CHECK_INDEX(active ? state : 0L, n);  // padding lanes of the last CTA touch nothing
  if (!LIST)
  {
    const int count = B * T.n_vars;
    const long limit = (n - base) * T.n_vars;
    const QT* src = q + base * T.n_vars;
    for (int e = tid; e < count; e += B)
      qs[e] = (e < limit) ? (float)src[e] : 0.f;  // fp64 joint values are rounded once (covered by the margins)
  }
  // the pair thresholds are read once per sphere-pair test by every warp: keep them in shared memory when they fit
  // (L1 is nearly all shared-memory carve-out here and what is left is thrashed by the voxel gathers)
  int2* pair_T_shared = reinterpret_cast<int2*>(qs + (LIST ? 0 : B * T.n_vars));
  if (PAIR_SMEM)
  {
    for (int e = tid; e < n_pair_smem; e += B)
      pair_T_shared[e] = pair_T_global[e];
  }
  __syncthreads();
  const float* my_q = qs + tid * T.n_vars;
  const QT* my_row = q + state * T.n_vars;  // padding lanes: row 0, values unused
  float saved[12];
  float* my_slots = slots + tid;
  float* my_stc = stc + tid;
  float* my_stb = stb + tid;
  const bool do_robot = (flags & B200DF_CHECK_ROBOT) != 0;
  constexpr bool do_self = SELF;

  // running minima of signed margins: negative = a certain collision (padding lanes start there: no work) / some test
  // fell inside its error margin
  int hit_m = active ? 0x3fffffff : -1;
  int amb_m = 0x3fffffff;
  bool oor = false;    // a joint value outside the analysed range: nothing computed here can be trusted
  float cur[12];
  for (int i = 0; i < T.n_links; ++i)
  {
    const FLink& l = T.links[i];
    if (LIST)
      fk_step_f<FULL>(l, my_row, my_slots, B, cur, oor, slot_in_regs ? saved : nullptr);
    else
      fk_step_f<FULL>(l, my_q, my_slots, B, cur, oor);
    const int gs = l.gslot;
    if (gs < 0)
      continue;
    const FGLink& gl = T.glinks[gs];
    const int np = do_pairs ? gl.partner_end - gl.partner_begin : 0;
    const bool stored = do_pairs && gl.store_link >= 0;
    const bool self_here = do_self && gl.self_enabled;
    float bx = 0.f, by = 0.f, bz = 0.f;
    if (np > 0 || stored)
    {
      applyf(cur, gl.bs[0], gl.bs[1], gl.bs[2], bx, by, bz);
      if (stored)
      {
        my_stb[(gl.store_link * 3 + 0) * B] = bx;
        my_stb[(gl.store_link * 3 + 1) * B] = by;
        my_stb[(gl.store_link * 3 + 2) * B] = bz;
      }
    }
    const int n_chunks = np > 0 ? (np + 31) / 32 : 1;
    for (int c = 0; c < n_chunks; ++c)
    {
      const int p0 = gl.partner_begin + 32 * c;
      unsigned sure_mask = 0u, maybe_mask = 0u;  // prune certainly passed / could pass
      if (np > 0)
      {
        const int pc = min(32, gl.partner_end - p0);
        // highest partner first, so that shifting the masks left by one per partner leaves partner b in bit b;
        // "d2 < threshold" is the sign of bits(d2) + (negated threshold bits), see the pair loop
#pragma unroll 2
        for (int b = pc - 1; b >= 0; --b)
        {
          const FPartner& pr = T.partners[p0 + b];
          const float dx = my_stb[(pr.store_link * 3 + 0) * B] - bx;
          const float dy = my_stb[(pr.store_link * 3 + 1) * B] - by;
          const float dz = my_stb[(pr.store_link * 3 + 2) * B] - bz;
          const int d2 = __float_as_int(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
          sure_mask = (sure_mask << 1) + ((unsigned)(d2 + pr.n_sure) >> 31);
          maybe_mask = (maybe_mask << 1) + ((unsigned)(d2 + pr.n_maybe) >> 31);
        }
        if (hit_m < 0)  // a state that already collides opens no more pair work
          sure_mask = maybe_mask = 0u;
      }
      const unsigned warp_mask = __reduce_or_sync(0xffffffffu, maybe_mask);
      // nothing to look up, nothing to keep for later links and no partner within reach of any lane: the link's
      // spheres need not be posed at all (the pair pass of the two-pass form, or a call that asks for pairs only)
      if (PAIRS && !do_robot && !self_here && !stored && warp_mask == 0u)
        continue;
      for (int s0 = gl.sphere_begin; s0 < gl.sphere_end; s0 += kBatch)
      {
        const int cnt = min(kBatch, gl.sphere_end - s0);
        float cx[kBatch], cy[kBatch], cz[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
        {
          if (k < cnt)
          {
            const FSphere& sp = T.sph[s0 + k];
            applyf(cur, sp.x, sp.y, sp.z, cx[k], cy[k], cz[k]);
            if (c == 0 && stored)
            {
              const int slot = gl.store_sphere + (s0 + k - gl.sphere_begin);
              my_stc[(slot * 3 + 0) * B] = cx[k];
              my_stc[(slot * 3 + 1) * B] = cy[k];
              my_stc[(slot * 3 + 2) * B] = cz[k];
            }
          }
        }
        if (c == 0)
        {
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
          {
            if (k < cnt)
            {
              const FSphere& sp = T.sph[s0 + k];
              // (maybe && !sure) is what matters; a certain hit overrides it at the end
              if (do_robot)
                sphere_field<FT>(wf, cx[k], cy[k], cz[k], sp.eps, sp.thr, hit_m, amb_m);
              if (self_here)
                sphere_field<FT>(sf, cx[k], cy[k], cz[k], sp.eps, sp.thr, hit_m, amb_m);
            }
          }
        }
        unsigned mm = warp_mask;
        const int2* t_row = (PAIR_SMEM ? pair_T_shared : pair_T_global) + gl.pair_const_begin +
                              (long)(s0 - gl.sphere_begin) * gl.pair_const_stride;
        while (mm)
        {
          const int b = __ffs(mm) - 1;
          mm &= mm - 1;
          const FPartner& pr = T.partners[p0 + b];
          const int2* t = t_row + pr.const_offset;
          const float* pc = my_stc + pr.store_sphere * 3 * B;
          const int stride = gl.pair_const_stride;
          // A squared distance is a non-negative float, so its bit pattern orders like the value.  The thresholds are
          // stored as the negated patterns {-bits(lo2), -(bits(hi2) + 1)}: "e2 < lo2" and "e2 <= hi2" become "the sum
          // is negative", and a running minimum of the sums (one VIADDMNMX each) replaces compare + select + or.
          // NaN and +inf have patterns above every finite threshold and never count, like the float comparison.
          int a_sure = 0x7fffffff, a_maybe = 0x7fffffff;
          // (unroll factors forced by hand measured within 2 % of the compiler's own choice: 1: -4 %, 2 / 4: +-1 %)
          for (int m = 0; m < pr.count; ++m)
          {
            const float px = pc[(m * 3 + 0) * B], py = pc[(m * 3 + 1) * B], pz = pc[(m * 3 + 2) * B];
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
            {
              if (k < cnt)
              {
                const float ex = cx[k] - px, ey = cy[k] - py, ez = cz[k] - pz;
                const int e2 = __float_as_int(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
                const int2 th = t[k * stride + m];  // {certainly touching below, certainly apart above}, encoded
                a_sure = __viaddmin_s32(e2, th.x, a_sure);
                a_maybe = __viaddmin_s32(e2, th.y, a_maybe);
              }
            }
          }
          // the sphere-pair result counts for the lanes whose own bounding-sphere prune passed
          hit_m = min(hit_m, ((sure_mask >> b) & 1u) ? a_sure : 0x3fffffff);
          amb_m = min(amb_m, ((maybe_mask >> b) & 1u) ? a_maybe : 0x3fffffff);
        }
      }
    }
    if (__all_sync(0xffffffffu, hit_m < 0))
      break;
  }
  const int v = oor ? 2 : (hit_m < 0 ? 1 : (amb_m < 0 ? 2 : 0));
  if (active)
  {
    if (!LIST || v != 0)  // a listed state already holds the 0 of the first pass
      verdict[state] = (uint8_t)v;
    if (v == 2)
    {
      const int at = atomicAdd(amb_count, 1);
      B200DF_CHECK_INDEX(at, n);
      amb_idx[at] = (int)state;
    }
  }
  if (MODE == 1)
  {
    const unsigned m = __ballot_sync(0xffffffffu, active && v == 0);
    if (m)
    {
      const int lane = tid & 31;
      int at = 0;
      if (lane == __ffs(m) - 1)
        at = atomicAdd(surv_count, __popc(m));
      at = __shfl_sync(0xffffffffu, at, __ffs(m) - 1);
      if ((m >> lane) & 1u)
      {
        B200DF_CHECK_INDEX(at + __popc(m & ((1u << lane) - 1u)), n);
        surv_idx[at + __popc(m & ((1u << lane) - 1u))] = (int)state;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Pair tests, ONE WARP PER STATE (robots with many spheres and link pairs).
//
// The one-thread-per-state kernels keep a state's stored centres in a shared-memory column; a PR2-sized tree (33 links,
// 124 spheres, ~480 enabled link pairs) needs 1.9 KB per state there, which leaves the pair pass THREE warps per SM.
// Here a warp owns one state: lane i builds joint matrix i (all sincos at once), twelve lanes walk the chain with one
// matrix entry each, the lanes then pose the spheres and the bounding centres four at a time, and the enabled link
// pairs are dealt out one per lane (bounding-sphere prune with its two thresholds, then the sphere pairs of the pairs
// that pass).  5 KB of shared memory per warp, 44 warps per SM.  Same tables, same thresholds and the same error
// analysis as the one-thread-per-state kernel: every transform entry is a three-term FMA chain of fp32 factors, which
// is what the per-link bounds (9.1 u per product) assume; a pair is "surely touching" / "possibly touching" by the same
// integer comparisons.  Verdicts: 1 = certain, 2 = the exact kernel decides, 0 = certainly no pair touches.
// Floating joints are not built here (robots that have one stay on the other path).
// ------------------------------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ float entryf(const float* a, const float* b, int r, int c)  // (a * b)[r][c], like mulf
{
  const float a0 = a[4 * r], a1 = a[4 * r + 1], a2 = a[4 * r + 2];
  const float acc = c == 3 ? fmaf(a0, b[3], a[4 * r + 3]) : a0 * b[c];
  return fmaf(a2, b[8 + c], fmaf(a1, b[4 + c], acc));
}

// the joint's own transform (fixed: identity; the same formulas as fk_step_f, written out as a matrix)
template <bool FULL, typename QP>
__device__ __forceinline__ void joint_matrix_f(const FLink& l, const QP* q, float* j, bool& out_of_range)
{
#pragma unroll
  for (int k = 0; k < 12; ++k)
    j[k] = (k == 0 || k == 5 || k == 10) ? 1.f : 0.f;
  if (l.joint_type == B200DF_JOINT_PRISMATIC)
  {
    const float v = jointf(l, q, 0, out_of_range);
    j[3] = l.ax * v;
    j[7] = l.ay * v;
    j[11] = l.az * v;
  }
  else if (l.joint_type == B200DF_JOINT_REVOLUTE || (FULL && l.joint_type == B200DF_JOINT_PLANAR))
  {
    const bool planar = FULL && l.joint_type == B200DF_JOINT_PLANAR;
    if (planar)
    {
      j[3] = jointf(l, q, 0, out_of_range);
      j[7] = jointf(l, q, 1, out_of_range);
    }
    const float v = jointf(l, q, planar ? 2 : 0, out_of_range);
    float s, c;
    if (FULL)
      sincosf(v, &s, &c);
    else
      sincos_small(v, &s, &c);
    const int kind = planar ? 3 : l.jkind;
    if (FULL && kind == 0)
    {
      const float tt = 1.f - c;
      j[0] = fmaf(tt, l.x2, c);
      j[1] = fmaf(tt, l.xy, -l.az * s);
      j[2] = fmaf(tt, l.xz, l.ay * s);
      j[4] = fmaf(tt, l.xy, l.az * s);
      j[5] = fmaf(tt, l.y2, c);
      j[6] = fmaf(tt, l.yz, -l.ax * s);
      j[8] = fmaf(tt, l.xz, -l.ay * s);
      j[9] = fmaf(tt, l.yz, l.ax * s);
      j[10] = fmaf(tt, l.z2, c);
    }
    else
    {
      const float sg = planar ? s : (kind == 1 ? l.ax : (kind == 2 ? l.ay : l.az)) * s;
      if (kind == 3)
      {
        j[0] = c;
        j[1] = -sg;
        j[4] = sg;
        j[5] = c;
      }
      else if (kind == 2)
      {
        j[0] = c;
        j[2] = sg;
        j[8] = -sg;
        j[10] = c;
      }
      else
      {
        j[5] = c;
        j[6] = -sg;
        j[9] = sg;
        j[10] = c;
      }
    }
  }
}

template <typename QT, bool FULL>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
    validity_fast_warp_pairs_kernel(const __grid_constant__ FastTables T, const FastTables* __restrict__ Tg,
                                    const int* __restrict__ maps, const int2* __restrict__ pair_T,
                                    const QT* __restrict__ q, long n, uint8_t* __restrict__ verdict,
                                    int* __restrict__ amb_idx, int* __restrict__ amb_count,
                                    const int* __restrict__ list, const int* __restrict__ list_count)
{
  extern __shared__ float wsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long slot = (long)blockIdx.x * kWarpsPerCta + warp;
  const long total = list ? (long)*list_count : n;
  if (slot >= total)
    return;  // whole warps leave; nothing below synchronises across warps
  const long state = list ? (long)list[slot] : slot;
  B200DF_CHECK_INDEX(state, n);
  const int L = T.n_links, G = T.n_glinks, S = T.n_spheres, P = T.n_partners;
  const int* sph_link = maps;
  const int* glink_link = sph_link + S;
  const int* owner = glink_link + G;
  const int* partner_glink = owner + P;
  float* Tm = wsm + (size_t)warp * ((size_t)L * 24 + 16 + (size_t)S * 3 + (size_t)G * 3 + (size_t)P);  // [L][12] transforms
  float* J = Tm + (size_t)L * 12;                                                          // [L][12] joint matrices
  float* TMP = J + (size_t)L * 12;                                                         // [16]
  float* C = TMP + 16;                                                                     // [S][3] sphere centres
  float* BC = C + (size_t)S * 3;                                                           // [G][3] bounding centres
  const QT* row = q + state * T.n_vars;
  bool oor = false;
  for (int i = lane; i < L; i += 32)
  {
    float j[12];
    joint_matrix_f<FULL>(Tg->links[i], row, j, oor);
#pragma unroll
    for (int k = 0; k < 12; ++k)
      J[i * 12 + k] = j[k];
  }
  __syncwarp();
  {
    const int e = lane < 12 ? lane : 0;  // lanes 12..31 shadow lane 0 and do not store
    const int r = e >> 2, c = e & 3;
    for (int i = 0; i < L; ++i)
    {
      const FLink& l = T.links[i];  // warp-uniform: constant bank
      float out;
      if (l.parent < 0)
        out = l.origin_is_identity ? J[i * 12 + e] : entryf(l.origin, J + i * 12, r, c);
      else
      {
        const float* par = Tm + l.parent * 12;
        if (l.joint_type == B200DF_JOINT_FIXED)
          out = l.origin_is_identity ? par[e] : entryf(par, l.origin, r, c);
        else if (l.origin_is_identity)
          out = entryf(par, J + i * 12, r, c);
        else
        {
          const float base = entryf(par, l.origin, r, c);
          if (lane < 12)
            TMP[e] = base;
          __syncwarp();
          out = entryf(TMP, J + i * 12, r, c);
        }
      }
      if (lane < 12)
        Tm[i * 12 + e] = out;
      __syncwarp();
    }
  }
  for (int sidx = lane; sidx < S; sidx += 32)
  {
    const FSphere sp = Tg->sph[sidx];
    applyf(Tm + sph_link[sidx] * 12, sp.x, sp.y, sp.z, C[3 * sidx], C[3 * sidx + 1], C[3 * sidx + 2]);
  }
  for (int gidx = lane; gidx < G; gidx += 32)
  {
    const FGLink& gl = Tg->glinks[gidx];
    applyf(Tm + glink_link[gidx] * 12, gl.bs[0], gl.bs[1], gl.bs[2], BC[3 * gidx], BC[3 * gidx + 1], BC[3 * gidx + 2]);
  }
  __syncwarp();
  // phase 1: the bounding-sphere prune of every enabled link pair, one pair per lane; the pairs that may touch are
  // queued (bit 31: the prune CERTAINLY passes, so a certain sphere contact is a certain collision)
  unsigned* queue = reinterpret_cast<unsigned*>(BC + (size_t)G * 3);  // [P]
  int n_queued = 0;
  for (int p0 = 0; p0 < P; p0 += 32)
  {
    const int p = p0 + lane;
    bool sure = false, maybe = false;
    if (p < P)
    {
      const int own = owner[p], pg = partner_glink[p];
      const int n_sure = Tg->partners[p].n_sure, n_maybe = Tg->partners[p].n_maybe;
      const float dx = BC[3 * pg] - BC[3 * own], dy = BC[3 * pg + 1] - BC[3 * own + 1], dz = BC[3 * pg + 2] - BC[3 * own + 2];
      const int d2 = __float_as_int(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
      sure = d2 + n_sure < 0;
      maybe = d2 + n_maybe < 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, maybe);
    if (maybe)
      queue[n_queued + __popc(m & ((1u << lane) - 1u))] = (unsigned)p | (sure ? 0x80000000u : 0u);
    n_queued += __popc(m);
  }
  __syncwarp();
  // phase 2: the sphere pairs of one queued link pair at a time, dealt out over the lanes
  bool hit = false, amb = false;
  for (int k = 0; k < n_queued && !hit; ++k)
  {
    const unsigned entry = queue[k];
    const int p = (int)(entry & 0x7fffffffu);
    const FPartner& pr = T.partners[p];  // warp-uniform from here on: constant bank
    const FGLink& gl = T.glinks[owner[p]];
    const int ps0 = T.glinks[partner_glink[p]].sphere_begin;
    const int cnt = pr.count, n_pairs = (gl.sphere_end - gl.sphere_begin) * cnt;
    const int2* t0 = pair_T + gl.pair_const_begin + pr.const_offset;
    int a_sure = 0x7fffffff, a_maybe = 0x7fffffff;
    for (int idx = lane; idx < n_pairs; idx += 32)
    {
      const int io = pr.div_magic ? (idx * pr.div_magic) >> 16 : idx / cnt, mm = idx - io * cnt;
      const int so = gl.sphere_begin + io, sq = ps0 + mm;
      const float ex = C[3 * so] - C[3 * sq], ey = C[3 * so + 1] - C[3 * sq + 1], ez = C[3 * so + 2] - C[3 * sq + 2];
      const int e2 = __float_as_int(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
      const int2 th = t0[(long)io * gl.pair_const_stride + mm];
      a_sure = __viaddmin_s32(e2, th.x, a_sure);
      a_maybe = __viaddmin_s32(e2, th.y, a_maybe);
    }
    hit = (entry >> 31) != 0u && __any_sync(0xffffffffu, a_sure < 0);
    amb = amb || __any_sync(0xffffffffu, a_maybe < 0);
  }
  oor = __any_sync(0xffffffffu, oor);
  if (lane == 0)
  {
    const int v = oor ? 2 : (hit ? 1 : (amb ? 2 : 0));
    if (!list || v != 0)  // a listed state already holds the 0 of the field pass
      verdict[state] = (uint8_t)v;
    if (v == 2)
    {
      const int at = atomicAdd(amb_count, 1);
      B200DF_CHECK_INDEX(at, n);
      amb_idx[at] = (int)state;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host: tables + forward error analysis
// ------------------------------------------------------------------------------------------------------------
double norm3(const double* v)
{
  return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

// Travel assumed for translational joint variables and the angle range assumed for rotational ones; values
// outside are detected at run time and sent to the exact kernel.
constexpr double kPrismaticTravel = 2.0, kPlanarTravel = 16.0, kAngleRange = 8.0, kFloatingTravel = 4.0;

struct LinkErr
{
  double eR = 0.0;     // || fp32 rotation - exact rotation ||_F
  double et = 0.0;     // || fp32 translation - exact translation ||_2
  double reach = 0.0;  // bound on || exact translation ||_2
};
}  // namespace

struct b200df_group::FastPath
{
  FastTables tables;
  int2* pair_T = nullptr;  // device: {-bits(lo2), -(bits(hi2) + 1)} per sphere pair, see the pair loop
  // the warp-per-state pair kernel (robots whose per-state shared-memory column leaves the one-thread-per-state pair
  // pass a handful of warps per SM): the tables again in global memory, where 32 lanes may read 32 different entries,
  // plus the flat maps it needs
  FastTables* d_tables = nullptr;
  int* d_warp_maps = nullptr;  // [sph_link[S] | glink_link[G] | owner_glink[P] | partner_glink[P]]
  bool warp_ok = false;        // no floating joint (the one joint type that kernel does not build)
  bool ok = false;
  bool needs_full = false;  // a floating / planar / oblique-axis revolute joint: the kernel variant that knows them
};

namespace b200df
{
int fast_create(b200df_group* g)
{
  const b200df_model* m = g->model;
  const int L = m->n_links, G = (int)g->glinks.size(), S = (int)(g->sph.size() / 4);
  const int P = (int)g->partners.size();
  if (L > kMaxLinks || G > kMaxGLinks || S > kMaxSpheres || P > kMaxPartners)
    return B200DF_OK;  // FAST silently degrades to EXACT for robots that do not fit the constant-bank tables
  std::unique_ptr<b200df_group::FastPath> fp(new b200df_group::FastPath);
  FastTables& T = fp->tables;
  std::memset(&T, 0, sizeof(T));
  T.n_links = L;
  T.n_vars = m->n_vars;
  T.n_slots = m->n_slots;
  T.n_glinks = G;
  T.n_spheres = S;
  T.n_store_links = g->dev.n_store_links;
  T.n_store_spheres = g->dev.n_store_spheres;
  T.n_partners = P;

  // ---- forward error analysis, link by link (u = 2^-24; rotations have unit 2-norm, so errors add up instead
  // of multiplying; every product of fp32 3x3 blocks adds at most 9u in Frobenius norm: 3 terms per entry,
  // |row| |col| <= 1 by Cauchy-Schwarz)
  const double u = kU;
  // sincosf: <= 2 ulp, ulp(1) = 2^-23 = 2u; sincos_small (the SIMPLE variant, |x| <= 8): 0.78 ulp over every fp32 value in
  // range (tools/sincos_check.cu)
  const double sc = 2.0 * 2.0 * u;
  std::vector<LinkErr> E(L);
  for (int i = 0; i < L; ++i)
  {
    const DevLink& l = m->links[i];
    FLink& f = T.links[i];
    f.joint_type = l.joint_type;
    f.jkind = l.jkind;
    f.var = l.var;
    f.is_mimic = l.is_mimic;
    f.origin_is_identity = l.origin_is_identity;
    f.parent = l.parent;
    f.parent_is_prev = l.parent_is_prev;
    f.store_slot = l.store_slot;
    f.parent_slot = l.parent_slot;
    f.gslot = g->link_gslot[i];
    f.mimic_mult = l.mimic_mult;
    f.mimic_off = l.mimic_off;
    f.ax = (float)l.ax;
    f.ay = (float)l.ay;
    f.az = (float)l.az;
    f.x2 = (float)l.x2;
    f.y2 = (float)l.y2;
    f.z2 = (float)l.z2;
    f.xy = (float)l.xy;
    f.xz = (float)l.xz;
    f.yz = (float)l.yz;
    for (int k = 0; k < 12; ++k)
      f.origin[k] = (float)l.origin[k];
    const double o[3] = { l.origin[3], l.origin[7], l.origin[11] };
    const double on = norm3(o);
    LinkErr par;  // root: identity, exact
    if (l.parent >= 0)
      par = E[l.parent];
    LinkErr b = par;  // base = parent * origin
    if (!l.origin_is_identity || l.parent < 0)
    {
      const double eO = u * std::sqrt(3.0);  // entries rounded to fp32: <= u/2 relative, ||R||_F = sqrt(3)
      if (l.parent < 0)
      {
        b.eR = eO;
        b.et = u * on;
      }
      else
      {
        b.eR = par.eR * (1.0 + eO) + eO + 9.1 * u;
        b.et = par.et + par.eR * on + (1.0 + par.eR) * u * on + 4.0 * u * std::sqrt(3.0) * (std::sqrt(3.0) * on + par.reach);
      }
      b.reach = par.reach + on;
    }
    LinkErr e = b;
    f.vmax = 0.f;
    if (l.joint_type == B200DF_JOINT_REVOLUTE || l.joint_type == B200DF_JOINT_PLANAR)
    {
      const bool planar = l.joint_type == B200DF_JOINT_PLANAR;
      if (planar || l.jkind == 0)
        fp->needs_full = true;
      f.vmax = (float)(planar ? kPlanarTravel : kAngleRange);  // planar: x, y and theta share the bound
      // the joint value reaches the fp32 FK rounded once (fp64 inputs, mimic joints evaluated in fp64 from the
      // rounded source value): angle error <= u*|v| (1 + |mimic factor|)
      const double dv = u * (planar ? kPlanarTravel : kAngleRange) * (1.0 + (l.is_mimic ? std::fabs(l.mimic_mult) : 0.0));
      const double dsc = sc + dv;
      // axis-aligned: entries are c, +-s (error dsc) and one ~1; general Rodrigues: every entry within 4*dsc + 4u
      const double eJ = (planar || l.jkind != 0) ? std::sqrt(4.0 * dsc * dsc + 4.0 * u * u) : 3.0 * (4.0 * dsc + 4.0 * u);
      e.eR = b.eR * (1.0 + eJ) + eJ + 9.1 * u;
      if (planar)
      {
        const double tr = kPlanarTravel * std::sqrt(2.0);
        e.et = b.et + b.eR * tr + 2.0 * u * tr + 4.0 * u * std::sqrt(3.0) * (std::sqrt(3.0) * tr + b.reach);
        e.reach = b.reach + tr;
      }
    }
    else if (l.joint_type == B200DF_JOINT_FLOATING)
    {
      fp->needs_full = true;
      // The seven variables reach the kernel rounded to fp32 (relative u each).  Normalised quaternion: input
      // rounding moves it by <= 2u, the fp32 normalisation (squared norm, sqrt, reciprocal, product) by <= 5u more.
      // Every entry of R(q) is quadratic in q with gradient norm <= 4 on the unit sphere, plus ~6u of rounding while
      // forming it: <= 34u per entry, 102u in Frobenius norm.
      f.vmax = (float)kFloatingTravel;
      const double eJ = 102.0 * u;
      const double tr = kFloatingTravel * std::sqrt(3.0);
      e.eR = b.eR * (1.0 + eJ) + eJ + 9.1 * u;
      e.et = b.et + b.eR * tr + 2.0 * u * tr + 4.0 * u * std::sqrt(3.0) * (std::sqrt(3.0) * tr + b.reach);
      e.reach = b.reach + tr;
    }
    else if (l.joint_type == B200DF_JOINT_PRISMATIC)
    {
      f.vmax = (float)kPrismaticTravel;
      const double tr = kPrismaticTravel;
      // axis rounded to fp32 (u*tr), axis*v rounded (u*tr), joint value rounded (u*tr (1 + |mimic factor|)), R*t + t0
      e.et = b.et + b.eR * tr + (3.0 + (l.is_mimic ? std::fabs(l.mimic_mult) : 0.0)) * u * tr +
             4.0 * u * std::sqrt(3.0) * (std::sqrt(3.0) * tr + b.reach);
      e.reach = b.reach + tr;
    }
    E[i] = e;
  }
  auto point_eps = [&](int link, const double* rel) {
    const LinkErr& e = E[link];
    const double rn = norm3(rel);
    // R~ * rel~ + t~ : rotation error on |rel|, translation error, rel rounded to fp32, rounding of the evaluation
    const double eps = e.eR * rn + e.et + u * rn * (1.0 + e.eR) + 4.0 * u * std::sqrt(3.0) * (std::sqrt(3.0) * rn + e.reach);
    return eps * 1.25 + 1e-9;  // slack for the second-order terms dropped above
  };
  std::vector<double> bs_eps(G), sph_eps(S);
  for (int gs = 0; gs < G; ++gs)
  {
    const DevGroupLink& gl = g->glinks[gs];
    FGLink& f = T.glinks[gs];
    f.sphere_begin = gl.sphere_begin;
    f.sphere_end = gl.sphere_end;
    f.self_enabled = gl.self_enabled;
    f.store_link = gl.store_link;
    f.store_sphere = gl.store_sphere;
    f.partner_begin = gl.partner_begin;
    f.partner_end = gl.partner_end;
    f.pair_const_begin = gl.pair_const_begin;
    f.pair_const_stride = gl.pair_const_stride;
    for (int k = 0; k < 3; ++k)
      f.bs[k] = (float)gl.bs[k];
    bs_eps[gs] = point_eps(gl.link, gl.bs);
    for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
    {
      const double* sp = &g->sph[(size_t)s * 4];
      sph_eps[s] = point_eps(gl.link, sp);
      T.sph[s].x = (float)sp[0];
      T.sph[s].y = (float)sp[1];
      T.sph[s].z = (float)sp[2];
      T.sph[s].eps = (float)(sph_eps[s] * (1.0 + 2.0 * u));
      T.sph[s].thr = g->sph_thr[s];
    }
  }
  // thresholds on fp32 squared distances: the fp32 distance is within E = eps_a + eps_b of the exact one, and the
  // fp32 evaluation of the squared norm is within a relative 8u of its own inputs' squared distance
  auto lo2 = [&](double R, double Esum) {
    const double d = R - Esum;
    return d > 0.0 ? (float)std::max(0.0, d * d * (1.0 - 16.0 * u) - 1e-30) : 0.f;
  };
  auto hi2 = [&](double R, double Esum) {
    const double d = R + Esum;
    return std::nextafter((float)(d * d * (1.0 + 16.0 * u)), INFINITY);
  };
  std::vector<int2> pair_T((size_t)std::max<long>(1, g->dev.n_pair_consts), make_int2(0, 0));
  auto bits = [](float v) {
    int i;
    std::memcpy(&i, &v, sizeof(i));
    return i;
  };
  for (int gs = 0; gs < G; ++gs)
  {
    const DevGroupLink& gl = g->glinks[gs];
    for (int p = gl.partner_begin; p < gl.partner_end; ++p)
    {
      const DevPartner& pr = g->partners[p];
      // which group link is this partner?  the one that owns stored-sphere slot pr.store_sphere
      int lo = -1;
      for (int h = 0; h < G; ++h)
        if (g->glinks[h].store_link == pr.store_link)
          lo = h;
      FPartner& f = T.partners[p];
      f.store_link = pr.store_link;
      f.store_sphere = pr.store_sphere;
      f.count = pr.count;
      f.const_offset = pr.const_offset;
      const double Eb = bs_eps[gs] + bs_eps[lo];
      // the reference compares the SQUARED distance with radius_sum (Q2): distance threshold sqrt(radius_sum)
      const double rs = std::sqrt(std::max(0.0, pr.radius_sum));
      const float tight_hi = hi2(std::sqrt(pr.tight2), Eb);
      auto fbits = [](float v) {
        int i;
        std::memcpy(&i, &v, sizeof(i));
        return i;
      };
      f.n_sure = -fbits(std::min(tight_hi, lo2(rs, Eb)));
      f.n_maybe = -fbits(std::min(tight_hi, hi2(rs, Eb)));
      {
        const int cnt = std::max(1, pr.count), n_pairs = (gl.sphere_end - gl.sphere_begin) * cnt;
        f.div_magic = 65536 / cnt + 1;
        for (int idx = 0; idx < n_pairs; ++idx)  // 128 spheres at most: the products stay far below 2^31
          if ((int)(((long)idx * f.div_magic) >> 16) != idx / cnt)
            f.div_magic = 0;  // never seen; the kernel then divides
      }
      for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
        for (int mm = 0; mm < pr.count; ++mm)
        {
          const int ps = g->glinks[lo].sphere_begin + mm;
          const double R = g->sph[(size_t)s * 4 + 3] + g->sph[(size_t)ps * 4 + 3];
          const double Es = sph_eps[s] + sph_eps[ps];
          const size_t at = (size_t)gl.pair_const_begin + (size_t)(s - gl.sphere_begin) * gl.pair_const_stride +
                            pr.const_offset + mm;
          pair_T[at] = make_int2(-bits(lo2(R, Es)), -(bits(hi2(R, Es)) + 1));
        }
    }
  }
  DeviceGuard dg(g->device);
  B200DF_CUDA(cudaMalloc(&fp->pair_T, pair_T.size() * sizeof(int2)));
  B200DF_CUDA(cudaMemcpy(fp->pair_T, pair_T.data(), pair_T.size() * sizeof(int2), cudaMemcpyHostToDevice));
  {
    std::vector<int> maps((size_t)S + G + 2 * (size_t)P, 0);
    int* sph_link = maps.data();
    int* glink_link = sph_link + S;
    int* owner = glink_link + G;
    int* partner_glink = owner + P;
    bool has_floating = false;
    for (int i = 0; i < L; ++i)
      has_floating = has_floating || m->links[i].joint_type == B200DF_JOINT_FLOATING;
    for (int gs = 0; gs < G; ++gs)
    {
      const DevGroupLink& gl = g->glinks[gs];
      glink_link[gs] = gl.link;
      for (int sp = gl.sphere_begin; sp < gl.sphere_end; ++sp)
        sph_link[sp] = gl.link;
      for (int pp = gl.partner_begin; pp < gl.partner_end; ++pp)
      {
        owner[pp] = gs;
        for (int h = 0; h < G; ++h)
          if (g->glinks[h].store_link == g->partners[pp].store_link)
            partner_glink[pp] = h;
      }
    }
    B200DF_CUDA(cudaMalloc(&fp->d_tables, sizeof(FastTables)));
    B200DF_CUDA(cudaMemcpy(fp->d_tables, &T, sizeof(FastTables), cudaMemcpyHostToDevice));
    B200DF_CUDA(cudaMalloc(&fp->d_warp_maps, maps.size() * sizeof(int)));
    B200DF_CUDA(cudaMemcpy(fp->d_warp_maps, maps.data(), maps.size() * sizeof(int), cudaMemcpyHostToDevice));
    fp->warp_ok = !has_floating;
  }
  fp->ok = true;
  g->fast = fp.release();
  return B200DF_OK;
}

void fast_destroy(b200df_group* g)
{
  if (!g->fast)
    return;
  cudaFree(g->fast->pair_T);
  cudaFree(g->fast->d_tables);
  cudaFree(g->fast->d_warp_maps);
  delete g->fast;
  g->fast = nullptr;
}

bool fast_tables_ok(const b200df_group* g)
{
  return g->fast && g->fast->ok;
}

bool fast_supported(const b200df_group* g, const FieldDev& wf, const FieldDev& sf, int q_dtype, int flags)
{
  if (!g->fast || !g->fast->ok || (q_dtype != B200DF_Q_F32 && q_dtype != B200DF_Q_F64))
    return false;
  if ((flags & B200DF_CHECK_ROBOT) && wf.neg)
    return false;  // signed fields: the predicate needs the fp64 distances
  if ((flags & B200DF_CHECK_SELF) && sf.neg)
    return false;
  return true;
}

}  // namespace b200df

namespace
{
FastLookup make_fast_lookup(const FieldDev& f)
{
  FastLookup l{};
  double gmax = 4.0;
  for (int i = 0; i < 3; ++i)
  {
    const double k = -f.g.om[i] * f.g.oo;
    l.k[i] = (float)k;
    l.lim[i] = f.g.n[i] >= 2 ? (unsigned)(f.g.n[i] - 2) : 0u;
    gmax = std::max(gmax, (double)f.g.n[i] + 2.0 * std::fabs(k) + 4.0);
  }
  l.oo = (float)f.g.oo;
  // g = fma(p, oo~, k~): oo rounded (u/2 relative on |p*oo| <= n+|k|), k rounded (u/2 |k|), fma rounded (u/2 |g|);
  // doubled for slack.  Centres farther out than the grid are decided by the range test with cells to spare.
  l.gerr = (float)(2.0 * kU * gmax);
  l.stride1 = (int)f.g.stride1;
  l.stride2 = (int)f.g.stride2;
  l.d2 = f.d2;
  return l;
}

constexpr long kPairSmemMax = 1024;  // sphere pairs whose thresholds are staged in shared memory (8 KB)

struct TwoPass  // see the kernel's header comment
{
  const int* list = nullptr;
  const int* list_count = nullptr;
  int* surv_idx = nullptr;
  int* surv_count = nullptr;
};

template <typename QT, typename FT, int B>
int launch_fast(const b200df_group* g, const FieldDev& wf, const FieldDev& sf, const QT* q, long n, int flags,
                uint8_t* verdict, int* amb_idx, int* amb_count, size_t bytes, cudaStream_t stream,
                TwoPass tp = TwoPass())
{
  const bool pairs = (flags & B200DF_CHECK_INTRA) && g->fast->tables.n_partners > 0;
  const int n_pair_smem = (pairs && g->dev.n_pair_consts <= kPairSmemMax) ? (int)g->dev.n_pair_consts : 0;
  bytes += (size_t)n_pair_smem * sizeof(int2);
  {
    // measurement knob: unused shared memory per CTA, to see what fewer resident warps cost (tools/fast_loop.py)
    static const long pad = getenv("B200DF_FAST_PAD_SMEM") ? atol(getenv("B200DF_FAST_PAD_SMEM")) : 0;
    bytes += (size_t)std::max(0L, pad);
  }
  const unsigned grid = (unsigned)((n + B - 1) / B);
  auto go = [&](auto kernel) -> cudaError_t {
    const cudaError_t e = allow_dynamic_smem(kernel, bytes);
    if (e != cudaSuccess)
      return e;
    kernel<<<grid, B, bytes, stream>>>(g->fast->tables, make_fast_lookup(wf), make_fast_lookup(sf), g->fast->pair_T, q, n,
                                       flags, verdict, amb_idx, amb_count, n_pair_smem, tp.list, tp.list_count,
                                       tp.surv_idx, tp.surv_count);
    return cudaSuccess;
  };
  const bool fl = g->fast->needs_full;
  const bool self = (flags & B200DF_CHECK_SELF) != 0;
  auto pick = [&](auto with_pairs, auto pair_smem, auto full) -> cudaError_t {
    constexpr bool W = decltype(with_pairs)::value, P = decltype(pair_smem)::value, F = decltype(full)::value;
    if constexpr (!W)
    {
      if (tp.surv_idx)  // first pass of two: field lookups only
        return self ? go(validity_fast_kernel<QT, FT, B, W, P, F, true, 1>) :
                      go(validity_fast_kernel<QT, FT, B, W, P, F, false, 1>);
    }
    else
    {
      if (tp.list)  // second pass: pairs only
        return go(validity_fast_kernel<QT, FT, B, W, P, F, false, 2>);
    }
    return self ? go(validity_fast_kernel<QT, FT, B, W, P, F, true, 0>) :
                  go(validity_fast_kernel<QT, FT, B, W, P, F, false, 0>);
  };
  using T_ = std::true_type;
  using F_ = std::false_type;
  if (!pairs)
    B200DF_CUDA(fl ? pick(F_(), F_(), T_()) : pick(F_(), F_(), F_()));
  else if (n_pair_smem > 0)
    B200DF_CUDA(fl ? pick(T_(), T_(), T_()) : pick(T_(), T_(), F_()));
  else
    B200DF_CUDA(fl ? pick(T_(), F_(), T_()) : pick(T_(), F_(), F_()));
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}
}  // namespace

namespace
{
// smallest batch the two-pass form is used for: below it the second launch and the list cost more than they save
// (B200DF_FAST_TWO_PASS_MIN overrides: measurement knob; 65 536 states: fused 0.073 ms, two passes 0.090; 131 072: 0.124
// against 0.107; 200 000: 0.173 against 0.163)
constexpr long kTwoPassMinStatesDefault = 1L << 17;
static const long kTwoPassMinStates =
    getenv("B200DF_FAST_TWO_PASS_MIN") ? atol(getenv("B200DF_FAST_TWO_PASS_MIN")) : kTwoPassMinStatesDefault;

bool two_pass_enabled()
{
  static const bool on = !(getenv("B200DF_FAST_TWO_PASS") && atoi(getenv("B200DF_FAST_TWO_PASS")) == 0);
  return on;
}

template <typename QT>
int fast_launch_typed(const b200df_group* g, const FieldDev& wf, const FieldDev& sf, int field_elem, const QT* q, long n,
                      int flags, uint8_t* verdict, int* amb_idx, int* amb_count, cudaStream_t stream)
{
  const FastTables& T = g->fast->tables;
  const bool has_pairs = (flags & B200DF_CHECK_INTRA) && T.n_partners > 0;
  // (the pair pass of the two-pass form keeps neither the joint values nor a lone branch-point slot in shared memory)
  auto smem = [&](int block, bool pairs, bool pair_pass) {
    size_t fl = pair_pass ? (T.n_slots <= 1 ? 0 : (size_t)T.n_slots * 12) : (size_t)T.n_slots * 12 + T.n_vars;
    if (pairs)
      fl += (size_t)T.n_store_spheres * 3 + (size_t)T.n_store_links * 3;
    return fl * block * sizeof(float);
  };
  auto launch = [&](int fl, TwoPass tp) -> int {
    const bool pairs = (fl & B200DF_CHECK_INTRA) && T.n_partners > 0;
    const bool pair_pass = tp.list != nullptr;
    const size_t b128 = smem(128, pairs, pair_pass);
    if (b128 <= 100 * 1024)
    {
      if (field_elem == 1)
        return launch_fast<QT, uint8_t, 128>(g, wf, sf, q, n, fl, verdict, amb_idx, amb_count, b128, stream, tp);
      return launch_fast<QT, uint16_t, 128>(g, wf, sf, q, n, fl, verdict, amb_idx, amb_count, b128, stream, tp);
    }
    const size_t b32 = smem(32, pairs, pair_pass);
    B200DF_REQUIRE(b32 <= 220 * 1024, "robot too large for the per-state shared-memory layout");
    if (field_elem == 1)
      return launch_fast<QT, uint8_t, 32>(g, wf, sf, q, n, fl, verdict, amb_idx, amb_count, b32, stream, tp);
    return launch_fast<QT, uint16_t, 32>(g, wf, sf, q, n, fl, verdict, amb_idx, amb_count, b32, stream, tp);
  };
  B200DF_CUDA(cudaMemsetAsync(amb_count, 0, sizeof(int), stream));
  const int field_flags = flags & (B200DF_CHECK_ROBOT | B200DF_CHECK_SELF);
  // the pair tests one warp per state (list may be null: every state)
  auto launch_warp_pairs = [&](const int* list, const int* list_count) -> int {
    const size_t per_warp = ((size_t)T.n_links * 24 + 16 + (size_t)T.n_spheres * 3 + (size_t)T.n_glinks * 3 +
                             (size_t)T.n_partners) * sizeof(float);
    const size_t bytes = per_warp * kWarpsPerCta;
    const unsigned grid = (unsigned)((n + kWarpsPerCta - 1) / kWarpsPerCta);
    auto go = [&](auto kernel) -> cudaError_t {
      const cudaError_t e = allow_dynamic_smem(kernel, bytes);
      if (e != cudaSuccess)
        return e;
      kernel<<<grid, 32 * kWarpsPerCta, bytes, stream>>>(T, g->fast->d_tables, g->fast->d_warp_maps, g->fast->pair_T, q, n,
                                                         verdict, amb_idx, amb_count, list, list_count);
      return cudaSuccess;
    };
    B200DF_CUDA(g->fast->needs_full ? go(validity_fast_warp_pairs_kernel<QT, true>) :
                                      go(validity_fast_warp_pairs_kernel<QT, false>));
    B200DF_LAUNCHED();
    B200DF_CUDA(cudaGetLastError());
    return B200DF_OK;
  };
  // robots whose stored centres leave the one-thread-per-state pair pass a few warps per SM (its 128-state CTA does
  // not fit): the pair tests go to the warp-per-state kernel, behind the field pass when field lookups are asked for
  static const bool warp_pairs_on = !(getenv("B200DF_FAST_WARP_PAIRS") && atoi(getenv("B200DF_FAST_WARP_PAIRS")) == 0);
  const bool heavy = has_pairs && g->fast->warp_ok && warp_pairs_on && smem(128, true, false) > 100 * 1024 &&
                     ((size_t)T.n_links * 24 + 16 + (size_t)T.n_spheres * 3 + (size_t)T.n_glinks * 3 + (size_t)T.n_partners) *
                             sizeof(float) * kWarpsPerCta <= 200 * 1024;
  if (getenv("B200DF_DEBUG_FAST"))
    fprintf(stderr, "b200df fast: links %d glinks %d spheres %d partners %d slots %d store_spheres %d store_links %d "
                    "smem128 %zu heavy %d warp_ok %d\n", T.n_links, T.n_glinks, T.n_spheres, T.n_partners, T.n_slots,
            T.n_store_spheres, T.n_store_links, smem(128, true, false), (int)heavy, (int)g->fast->warp_ok);
  if (heavy && !field_flags)
    return launch_warp_pairs(nullptr, nullptr);
  if (!has_pairs || !field_flags || (!heavy && (n < kTwoPassMinStates || !two_pass_enabled())))
    return launch(flags, TwoPass());
  // pass A: the field lookups of every state; pass B: the sphere pairs of the states pass A left at 0
  int* surv = nullptr;
  keep_async_pool(g->device);
  B200DF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&surv), ((size_t)n + 1) * sizeof(int), stream));
  B200DF_CUDA(cudaMemsetAsync(surv, 0, sizeof(int), stream));
  TwoPass a, b;
  a.surv_count = surv;
  a.surv_idx = surv + 1;
  b.list_count = surv;
  b.list = surv + 1;
  int rc = launch(field_flags, a);
  if (!rc)
    rc = heavy ? launch_warp_pairs(b.list, b.list_count) : launch(flags & ~(B200DF_CHECK_ROBOT | B200DF_CHECK_SELF), b);
  cudaFreeAsync(surv, stream);
  return rc;
}
}  // namespace

namespace b200df
{
int fast_launch(const b200df_group* g, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q, int q_dtype,
                long n, int flags, uint8_t* verdict, int* amb_idx, int* amb_count, cudaStream_t stream)
{
  if (q_dtype == B200DF_Q_F32)
    return fast_launch_typed(g, wf, sf, field_elem, static_cast<const float*>(q), n, flags, verdict, amb_idx, amb_count,
                             stream);
  return fast_launch_typed(g, wf, sf, field_elem, static_cast<const double*>(q), n, flags, verdict, amb_idx, amb_count,
                           stream);
}
}  // namespace b200df
