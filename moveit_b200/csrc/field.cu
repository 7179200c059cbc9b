// Distance field: occupancy grid + exact clamped squared EDT, shape rasteriser, point scatter, getDistanceGradient.
//
// Replaces distance_field::PropagationDistanceField / DistanceField / findInternalPointsConvex
// (reference: moveit_core/distance_field/src/propagation_distance_field.cpp, distance_field.cpp,
// find_internal_points.cpp, include/moveit/distance_field/voxel_grid.h).
//
// This translation unit is compiled with --fmad=false: world->grid rounding and the containsPoint tests must
// round exactly like the reference's (non-FMA) double arithmetic, because lattice points sit exactly on shape
// and voxel boundaries by construction.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "b200df_internal.cuh"

using namespace b200df;

namespace
{
constexpr int kInf16 = 0xFFFF;

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------
// VoxelGrid::getCellFromLocation (voxel_grid.h:515-532), kept in double so that absurd coordinates cannot
// overflow an int; the caller range-checks before converting.
__device__ __forceinline__ double cell_coord(const GridDev& g, int dim, double loc)
{
  return floor((loc - g.om[dim]) * g.oo);
}

__device__ __forceinline__ bool world_to_grid(const GridDev& g, double x, double y, double z, int& ix, int& iy, int& iz)
{
  const double fx = cell_coord(g, 0, x), fy = cell_coord(g, 1, y), fz = cell_coord(g, 2, z);
  // isCellValid (voxel_grid.h:408-411); written so that NaN fails
  if (!(fx >= 0.0 && fx < (double)g.n[0] && fy >= 0.0 && fy < (double)g.n[1] && fz >= 0.0 && fz < (double)g.n[2]))
    return false;
  ix = (int)fx;
  iy = (int)fy;
  iz = (int)fz;
  return true;
}

// addPointsToField / removePointsFromField voxelisation (propagation_distance_field.cpp:176-218)
__global__ void scatter_points_kernel(GridDev g, const double* __restrict__ xyz, long n, uint8_t* __restrict__ vol,
                                      uint8_t value, bool or_value, uint32_t* __restrict__ bits = nullptr, int wpr = 0)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int ix, iy, iz;
  if (!world_to_grid(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], ix, iy, iz))
    return;
  const long r = ix * g.stride1 + iy * g.stride2 + iz;
  B200DF_CHECK_INDEX(r, (long long)g.n[0] * g.stride1);
  if (or_value)
    vol[r] |= value;  // only used with the same value from every thread of a launch: benign race
  else
    vol[r] = value;
  if (bits)  // keep the bit-packed mirror of the occupancy current (only when setting voxels)
    atomicOr(&bits[((long)ix * g.n[1] + iy) * wpr + (iz >> 5)], 1u << (iz & 31));
}

__global__ void scatter_voxels_kernel(GridDev g, const int32_t* __restrict__ ijk, long n, uint8_t* __restrict__ vol,
                                      uint8_t value)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int ix = ijk[3 * i], iy = ijk[3 * i + 1], iz = ijk[3 * i + 2];
  if (ix < 0 || ix >= g.n[0] || iy < 0 || iy >= g.n[1] || iz < 0 || iz >= g.n[2])
    return;
  B200DF_CHECK_INDEX(ix * g.stride1 + iy * g.stride2 + iz, (long long)g.n[0] * g.stride1);
  vol[ix * g.stride1 + iy * g.stride2 + iz] = value;
}

// updatePointsInField (propagation_distance_field.cpp:120-174): marks bit0 = in old set, bit1 = in new set;
// old-only voxels are cleared, new-only voxels are set.
__global__ void apply_update_marks_kernel(const uint8_t* __restrict__ marks, uint8_t* __restrict__ occ, long n)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint8_t m = marks[i];
  if (m == 1)
    occ[i] = 0;
  else if (m == 2)
    occ[i] = 1;
}

// bodies::Body derived data (geometric_shapes, third party; restated from its published behaviour)
struct BodyDev
{
  int type;
  double center[3];
  double n0[3], n1[3], n2[3];  // columns of the pose rotation
  double radius2, length2, width2, height2;
};

__device__ __forceinline__ double dot3(const double* a, double bx, double by, double bz)
{
  return (a[0] * bx + a[1] * by) + a[2] * bz;
}

__device__ __forceinline__ bool body_contains(const BodyDev& b, double px, double py, double pz)
{
  if (b.type == B200DF_SHAPE_SPHERE)
  {
    const double dx = b.center[0] - px, dy = b.center[1] - py, dz = b.center[2] - pz;
    return ((dx * dx + dy * dy) + dz * dz) <= b.radius2;
  }
  const double vx = px - b.center[0], vy = py - b.center[1], vz = pz - b.center[2];
  if (b.type == B200DF_SHAPE_CYLINDER)
  {
    const double pH = dot3(b.n2, vx, vy, vz);
    if (fabs(pH) > b.length2)
      return false;
    const double pB1 = dot3(b.n0, vx, vy, vz);
    const double remaining = b.radius2 - pB1 * pB1;
    if (remaining < 0.0)
      return false;
    const double pB2 = dot3(b.n1, vx, vy, vz);
    return pB2 * pB2 <= remaining;
  }
  const double pL = dot3(b.n0, vx, vy, vz);
  if (fabs(pL) > b.length2)
    return false;
  const double pW = dot3(b.n1, vx, vy, vz);
  if (fabs(pW) > b.width2)
    return false;
  const double pH = dot3(b.n2, vx, vy, vz);
  if (fabs(pH) > b.height2)
    return false;
  return true;
}

struct PoseDev
{
  double m[12];
};

// findInternalPointsConvex (find_internal_points.cpp:39-64) with the lattice coordinates accumulated on the
// host (Q19).  One thread per lattice point.  Optionally poses the point (PosedBodyPointDecomposition::updatePose,
// collision_distance_field_types.cpp:368-379) before voxelising it; optionally records a per-point flag.
__global__ void rasterize_kernel(BodyDev b, const double* __restrict__ xs, int nxs, const double* __restrict__ ys,
                                 int nys, const double* __restrict__ zs, int nzs, int apply_pose, PoseDev pose,
                                 GridDev g, uint8_t* __restrict__ vol, uint8_t value, uint8_t* __restrict__ flags)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)nxs * nys * nzs;
  if (i >= total)
    return;
  const int k = (int)(i % nzs);
  const int j = (int)((i / nzs) % nys);
  const int a = (int)(i / ((long)nzs * nys));
  double px = xs[a], py = ys[j], pz = zs[k];
  const bool inside = body_contains(b, px, py, pz);
  if (flags)
    flags[i] = inside ? 1 : 0;
  if (!inside || !vol)
    return;
  if (apply_pose)
  {
    const double* m = pose.m;  // Isometry3d * Vector3d: t + ((r0*x + r1*y) + r2*z)
    const double wx = m[3] + ((m[0] * px + m[1] * py) + m[2] * pz);
    const double wy = m[7] + ((m[4] * px + m[5] * py) + m[6] * pz);
    const double wz = m[11] + ((m[8] * px + m[9] * py) + m[10] * pz);
    px = wx;
    py = wy;
    pz = wz;
  }
  int ix, iy, iz;
  if (world_to_grid(g, px, py, pz, ix, iy, iz))
  {
    B200DF_CHECK_INDEX(ix * g.stride1 + iy * g.stride2 + iz, (long long)g.n[0] * g.stride1);
    vol[ix * g.stride1 + iy * g.stride2 + iz] = value;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Exact clamped squared EDT, separable (kernel 5).  d2(v) = min over obstacle voxels q of |v-q|^2, clamped to
// max_d2.  Any voxel whose true distance is <= max_d2 has every axis offset <= R (R*R >= max_d2), so windows of
// +-R per axis are exact after the final clamp.
// ------------------------------------------------------------------------------------------------------------
// z pass (contiguous axis): squared distance to the nearest obstacle voxel in the same z-row, kInf16 if none in +-R
template <bool INVERT>
__global__ void edt_pass_z_kernel(const uint8_t* __restrict__ occ, uint16_t* __restrict__ out, int nz, long n_vox,
                                  int R)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_vox)
    return;
  const int z = (int)(i % nz);
  const uint8_t* row = occ + (i - z);
  auto obstacle = [&](int zz) -> bool {
    const uint8_t o = row[zz];
    return INVERT ? (o == 0) : (o != 0);
  };
  int best = kInf16;
  if (obstacle(z))
    best = 0;
  else
  {
    for (int d = 1; d <= R; ++d)
    {
      if ((z - d >= 0 && obstacle(z - d)) || (z + d < nz && obstacle(z + d)))
      {
        best = d * d;
        break;
      }
    }
  }
  out[i] = (uint16_t)best;
}

// y / x pass: min-plus with the parabola d*d over a +-R window along an axis with the given stride
template <typename OutT, bool FINAL>
__global__ void edt_pass_axis_kernel(const uint16_t* __restrict__ in, OutT* __restrict__ out, int n_axis, long stride,
                                     long n_vox, int R, int max_d2)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_vox)
    return;
  const int c = (int)((i / stride) % n_axis);
  int best = in[i];
  for (int d = 1; d <= R && d * d < best; ++d)
  {
    const int dd = d * d;
    if (c - d >= 0)
    {
      const int v = in[i - d * stride];
      if (v != kInf16 && v + dd < best)
        best = v + dd;
    }
    if (c + d < n_axis)
    {
      const int v = in[i + d * stride];
      if (v != kInf16 && v + dd < best)
        best = v + dd;
    }
  }
  if (FINAL)
    out[i] = (OutT)(best > max_d2 ? max_d2 : best);
  else
    out[i] = (OutT)(best > max_d2 ? kInf16 : best);  // anything beyond the clamp can never win later
}

// getOcTreePoints (distance_field.cpp:254-279) on the device.  A leaf no bigger than a voxel contributes its centre;
// a bigger leaf a lattice `for (x = c - ceil_val; x <= c + ceil_val; x += resolution)` per axis, whose number of steps
// and coordinates depend on the accumulated rounding of the literal += loop (Q19) - so the loops are run literally:
// kernel 1 counts the steps per axis, kernel 2 (one thread per lattice point, leaf found by binary search in the
// prefix sums) re-accumulates its own coordinates and voxelises the point.  Nothing is expanded on the host and the
// point list never exists in memory.
__global__ void octree_count_kernel(const double* __restrict__ leaves4, long n, double res, int* __restrict__ counts3,
                                    unsigned long long* __restrict__ totals)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const double size = leaves4[4 * i + 3];
  int c[3] = { 1, 1, 1 };
  if (!(size <= res))
  {
    const double ceil_val = ceil(size / res) * res / 2.0;
    for (int a = 0; a < 3; ++a)
    {
      const double ctr = leaves4[4 * i + a];
      int k = 0;
      for (double v = ctr - ceil_val; v <= ctr + ceil_val; v += res)
        ++k;
      c[a] = k;
    }
  }
  counts3[3 * i] = c[0];
  counts3[3 * i + 1] = c[1];
  counts3[3 * i + 2] = c[2];
  totals[i] = (unsigned long long)c[0] * c[1] * c[2];
}

__device__ __forceinline__ double lattice_coord(double ctr, double ceil_val, double res, int steps)
{
  double v = ctr - ceil_val;
  for (int k = 0; k < steps; ++k)
    v += res;
  return v;
}

__global__ void octree_scatter_kernel(GridDev g, const double* __restrict__ leaves4, long n,
                                      const int* __restrict__ counts3, const unsigned long long* __restrict__ offsets,
                                      unsigned long long total, double res, uint8_t* __restrict__ vol, uint8_t value,
                                      uint32_t* __restrict__ bits, int wpr)
{
  const unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (p >= total)
    return;
  long lo = 0, hi = n - 1;  // last leaf whose offset is <= p
  while (lo < hi)
  {
    const long mid = (lo + hi + 1) >> 1;
    if (offsets[mid] <= p)
      lo = mid;
    else
      hi = mid - 1;
  }
  const long leaf = lo;
  const unsigned long long local = p - offsets[leaf];
  const int ny = counts3[3 * leaf + 1], nz = counts3[3 * leaf + 2];
  const int iz = (int)(local % nz), iy = (int)((local / nz) % ny), ix = (int)(local / ((unsigned long long)nz * ny));
  const double size = leaves4[4 * leaf + 3];
  double x = leaves4[4 * leaf], y = leaves4[4 * leaf + 1], z = leaves4[4 * leaf + 2];
  if (!(size <= res))
  {
    const double ceil_val = ceil(size / res) * res / 2.0;
    x = lattice_coord(x, ceil_val, res, ix);
    y = lattice_coord(y, ceil_val, res, iy);
    z = lattice_coord(z, ceil_val, res, iz);
  }
  int gx, gy, gz;
  if (!world_to_grid(g, x, y, z, gx, gy, gz))
    return;
  B200DF_CHECK_INDEX(gx * g.stride1 + gy * g.stride2 + gz, (long long)g.n[0] * g.stride1);
  vol[gx * g.stride1 + gy * g.stride2 + gz] = value;
  if (bits)
    atomicOr(&bits[((long)gx * g.n[1] + gy) * wpr + (gz >> 5)], 1u << (gz & 31));
}

// exclusive prefix sum of n counts, single block (octree leaf lists are thousands to a few million entries; this is a
// set-up step next to the rebuild it triggers)
__global__ void exclusive_scan_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
                                      long n, unsigned long long* __restrict__ total)
{
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const long per = (n + T - 1) / T;
  const long a = min(n, (long)t * per), b = min(n, a + per);
  unsigned long long sum = 0;
  for (long i = a; i < b; ++i)
    sum += in[i];
  part[t] = sum;
  __syncthreads();
  if (t == 0)
  {
    unsigned long long run = 0;
    for (int k = 0; k < T; ++k)
    {
      const unsigned long long v = part[k];
      part[k] = run;
      run += v;
    }
    *total = run;
  }
  __syncthreads();
  unsigned long long run = part[t];
  for (long i = a; i < b; ++i)
  {
    out[i] = run;
    run += in[i];
  }
}

// b200df_field_upload_d2: int32 d2 (and negative d2) -> the field's element type; occupancy = (d2 == 0).
// Out-of-range values raise *bad (the upload is then refused).
template <typename T>
__global__ void narrow_kernel(const int32_t* __restrict__ d2, const int32_t* __restrict__ neg, T* __restrict__ out,
                              T* __restrict__ out_neg, uint8_t* __restrict__ occ, long n, int max_d2,
                              int* __restrict__ bad)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int v = d2[i];
  bool ok = v >= 0 && v <= max_d2;
  out[i] = (T)v;
  occ[i] = v == 0 ? 1 : 0;
  if (neg)
  {
    const int w = neg[i];
    ok = ok && w >= 0 && w <= max_d2;
    out_neg[i] = (T)w;
  }
  if (!ok)
    atomicExch(bad, 1);
}

// PropagationDistanceField::getNearestCell (propagation_distance_field.h:411-431) without a stored closest point:
// the first voxel (x-major scan of the offsets) of the other kind at exactly the cell's squared distance.
template <typename T>
__global__ void nearest_cell_kernel(GridDev g, const T* __restrict__ d2, const T* __restrict__ neg,
                                    const uint8_t* __restrict__ occ, const double* __restrict__ sqrt_table, int max_d2,
                                    const int32_t* __restrict__ ijk, long n, double* __restrict__ dist,
                                    int32_t* __restrict__ pos, uint8_t* __restrict__ found)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int x = ijk[3 * i], y = ijk[3 * i + 1], z = ijk[3 * i + 2];
  pos[3 * i] = x;
  pos[3 * i + 1] = y;
  pos[3 * i + 2] = z;
  dist[i] = 0.0;
  found[i] = 0;
  if (x < 0 || x >= g.n[0] || y < 0 || y >= g.n[1] || z < 0 || z >= g.n[2])
    return;
  const long idx = x * g.stride1 + y * g.stride2 + z;
  int target = d2[idx];
  uint8_t want = 1;  // looking for an obstacle voxel
  double sign = 1.0;
  if (target == 0)
  {
    target = neg ? (int)neg[idx] : 0;
    want = 0;  // inside an obstacle: looking for a free voxel
    sign = -1.0;
    if (target == 0)
      return;
  }
  dist[i] = sign * sqrt_table[target];
  if (target >= max_d2)
    return;  // clamped: the nearest cell is unknown (the reference returns an uninitialised closest point here)
  int r = 0;
  while (r * r < target)
    ++r;
  for (int dx = -r; dx <= r; ++dx)
  {
    const int xx = x + dx;
    if (xx < 0 || xx >= g.n[0])
      continue;
    for (int dy = -r; dy <= r; ++dy)
    {
      const int yy = y + dy;
      const int rest = target - dx * dx - dy * dy;
      if (yy < 0 || yy >= g.n[1] || rest < 0)
        continue;
      int dz = (int)sqrtf((float)rest);
      while (dz * dz > rest)
        --dz;
      while ((dz + 1) * (dz + 1) <= rest)
        ++dz;
      if (dz * dz != rest)
        continue;
      for (int s = -1; s <= 1; s += 2)
      {
        const int zz = z + s * dz;
        if (zz < 0 || zz >= g.n[2])
          continue;
        if ((occ[xx * g.stride1 + yy * g.stride2 + zz] != 0) == (want != 0))
        {
          pos[3 * i] = xx;
          pos[3 * i + 1] = yy;
          pos[3 * i + 2] = zz;
          found[i] = 1;
          return;
        }
        if (dz == 0)
          break;
      }
    }
  }
}

template <typename T>
__global__ void widen_kernel(const T* __restrict__ in, int32_t* __restrict__ out, long n)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = (int32_t)in[i];
}

// DistanceField::getDistanceGradient (distance_field.cpp:62-86) + PropagationDistanceField::getDistance
// (propagation_distance_field.h:597-600)
template <typename T>
__device__ __forceinline__ double cell_distance(const FieldDev& f, long idx)
{
  const T* d2 = static_cast<const T*>(f.d2);
  double d = f.sqrt_table[d2[idx]];
  if (f.neg)
    d = d - f.sqrt_table[static_cast<const T*>(f.neg)[idx]];
  else
    d = d - f.sqrt_table[0];
  return d;
}

template <typename T>
__global__ void distance_gradient_kernel(FieldDev f, const double* __restrict__ xyz, long n, double* __restrict__ dist,
                                         double* __restrict__ grad, uint8_t* __restrict__ in_bounds)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const GridDev& g = f.g;
  const double fx = cell_coord(g, 0, xyz[3 * i]), fy = cell_coord(g, 1, xyz[3 * i + 1]),
               fz = cell_coord(g, 2, xyz[3 * i + 2]);
  const bool ok = fx >= 1.0 && fy >= 1.0 && fz >= 1.0 && fx < (double)(g.n[0] - 1) && fy < (double)(g.n[1] - 1) &&
                  fz < (double)(g.n[2] - 1);
  if (!ok)
  {
    dist[i] = f.max_distance;
    grad[3 * i] = grad[3 * i + 1] = grad[3 * i + 2] = 0.0;
    in_bounds[i] = 0;
    return;
  }
  const long idx = (long)fx * g.stride1 + (long)fy * g.stride2 + (long)fz;
  const double itr = (double)f.inv_twice_res;
  grad[3 * i + 0] = (cell_distance<T>(f, idx + g.stride1) - cell_distance<T>(f, idx - g.stride1)) * itr;
  grad[3 * i + 1] = (cell_distance<T>(f, idx + g.stride2) - cell_distance<T>(f, idx - g.stride2)) * itr;
  grad[3 * i + 2] = (cell_distance<T>(f, idx + 1) - cell_distance<T>(f, idx - 1)) * itr;
  dist[i] = cell_distance<T>(f, idx);
  in_bounds[i] = 1;
}

// writeToStream payload (propagation_distance_field.cpp:652-671): one byte per 8 z-cells, bit zi <-> z+zi
__global__ void pack_occupancy_kernel(const uint8_t* __restrict__ occ, uint8_t* __restrict__ out, int nz, int zbytes,
                                      long n_rows)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_rows * zbytes)
    return;
  const long row = i / zbytes;
  const int zb = (int)(i % zbytes);
  uint8_t b = 0;
  for (int zi = 0; zi < 8; ++zi)
  {
    const int z = zb * 8 + zi;
    if (z < nz && occ[row * nz + z])
      b |= (uint8_t)(1u << zi);
  }
  out[i] = b;
}

__global__ void unpack_occupancy_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ occ, int nz, int zbytes,
                                        long n_rows)
{
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_rows * nz)
    return;
  const long row = i / nz;
  const int z = (int)(i % nz);
  if ((in[row * zbytes + z / 8] >> (z % 8)) & 1)
    occ[i] = 1;  // readFromStream ADDS the stored obstacle voxels (addNewObstacleVoxels)
}

inline size_t bits_bytes(const b200df_field* f)
{
  return (size_t)f->grid.n[0] * f->grid.n[1] * bits_wpr(f->grid.n[2]) * sizeof(uint32_t);
}

inline unsigned blocks_for(long n, int threads)
{
  return (unsigned)((n + threads - 1) / threads);
}

// ------------------------------------------------------------------------------------------------------------
// host-side bodies::Body maths (geometric_shapes semantics)
// ------------------------------------------------------------------------------------------------------------
struct HostBody
{
  BodyDev dev;
  double radiusU = 0, radiusB = 0, bound_radius = 0;
  double scaled_length = 0;  // cylinder: scale*length + 2*padding
  double pose[12];
};

int make_body(int type, const double* dims, const double* pose12, double scale, double padding, HostBody* out)
{
  HostBody hb;
  std::memcpy(hb.pose, pose12, sizeof(hb.pose));
  BodyDev& b = hb.dev;
  b.type = type;
  b.center[0] = pose12[3];
  b.center[1] = pose12[7];
  b.center[2] = pose12[11];
  for (int r = 0; r < 3; ++r)
  {
    b.n0[r] = pose12[4 * r + 0];
    b.n1[r] = pose12[4 * r + 1];
    b.n2[r] = pose12[4 * r + 2];
  }
  b.radius2 = b.length2 = b.width2 = b.height2 = 0.0;
  if (type == B200DF_SHAPE_SPHERE)
  {
    hb.radiusU = dims[0] * scale + padding;
    b.radius2 = hb.radiusU * hb.radiusU;
    hb.bound_radius = hb.radiusU;
  }
  else if (type == B200DF_SHAPE_CYLINDER)
  {
    hb.radiusU = dims[0] * scale + padding;
    b.radius2 = hb.radiusU * hb.radiusU;
    b.length2 = scale * dims[1] / 2.0 + padding;
    hb.radiusB = std::sqrt(b.length2 * b.length2 + b.radius2);
    hb.bound_radius = hb.radiusB;
    hb.scaled_length = scale * dims[1] + 2.0 * padding;
  }
  else if (type == B200DF_SHAPE_BOX)
  {
    const double s2 = scale / 2.0;
    b.length2 = dims[0] * s2 + padding;
    b.width2 = dims[1] * s2 + padding;
    b.height2 = dims[2] * s2 + padding;
    b.radius2 = b.length2 * b.length2 + b.width2 * b.width2 + b.height2 * b.height2;
    hb.radiusB = std::sqrt(b.radius2);
    hb.bound_radius = hb.radiusB;
  }
  else
  {
    set_error("unsupported shape type %d (sphere=1, cylinder=2, box=4)", type);
    return B200DF_ERR_INVALID;
  }
  *out = hb;
  return B200DF_OK;
}

// lattice of findInternalPointsConvex along one axis, accumulated exactly like the reference loop
void lattice_axis(double center, double radius, double resolution, std::vector<double>& out)
{
  const double start = std::floor((center - radius - resolution) / resolution) * resolution;
  const double end = center + radius + resolution;
  out.clear();
  for (double v = start; v <= end; v += resolution)
    out.push_back(v);
}

struct DeviceLattice
{
  double *xs = nullptr, *ys = nullptr, *zs = nullptr;
  int nx = 0, ny = 0, nz = 0;
  ~DeviceLattice()
  {
    cudaFree(xs);
    cudaFree(ys);
    cudaFree(zs);
  }
};

int upload_lattice(const HostBody& hb, double resolution, cudaStream_t stream, DeviceLattice* lat)
{
  std::vector<double> xs, ys, zs;
  lattice_axis(hb.dev.center[0], hb.bound_radius, resolution, xs);
  lattice_axis(hb.dev.center[1], hb.bound_radius, resolution, ys);
  lattice_axis(hb.dev.center[2], hb.bound_radius, resolution, zs);
  lat->nx = (int)xs.size();
  lat->ny = (int)ys.size();
  lat->nz = (int)zs.size();
  B200DF_CUDA(cudaMalloc(&lat->xs, std::max<size_t>(1, xs.size()) * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&lat->ys, std::max<size_t>(1, ys.size()) * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&lat->zs, std::max<size_t>(1, zs.size()) * sizeof(double)));
  B200DF_CUDA(cudaMemcpyAsync(lat->xs, xs.data(), xs.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  B200DF_CUDA(cudaMemcpyAsync(lat->ys, ys.data(), ys.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  B200DF_CUDA(cudaMemcpyAsync(lat->zs, zs.data(), zs.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  B200DF_CUDA(cudaStreamSynchronize(stream));  // the host vectors die at scope exit
  return B200DF_OK;
}

int rasterize_into(b200df_field* f, int shape_type, const double* dims3, const double* body_pose12,
                   const double* apply_pose12, double scale, double padding, uint8_t value)
{
  HostBody hb;
  int rc = make_body(shape_type, dims3, body_pose12, scale, padding, &hb);
  if (rc)
    return rc;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  DeviceLattice lat;
  rc = upload_lattice(hb, f->desc.resolution, f->stream, &lat);
  if (rc)
    return rc;
  const long total = (long)lat.nx * lat.ny * lat.nz;
  if (total > 0)
  {
    PoseDev pd{};
    if (apply_pose12)
      std::memcpy(pd.m, apply_pose12, sizeof(pd.m));
    rasterize_kernel<<<blocks_for(total, 256), 256, 0, f->stream>>>(hb.dev, lat.xs, lat.nx, lat.ys, lat.ny, lat.zs,
                                                                    lat.nz, apply_pose12 ? 1 : 0, pd, f->grid, f->occ,
                                                                    value, nullptr);
    B200DF_LAUNCHED();
    B200DF_CUDA(cudaGetLastError());
    B200DF_CUDA(cudaStreamSynchronize(f->stream));
  }
  f->dirty = true;
  f->bits_valid = false;
  return B200DF_OK;
}

template <typename T>
int run_edt(b200df_field* f, bool invert, T* out)
{
  const long n = f->n_voxels;
  const int nz = f->grid.n[2], ny = f->grid.n[1], nx = f->grid.n[0];
  const int threads = 256;
  const unsigned blocks = blocks_for(n, threads);
  if (edt_fast_applies(f))
    return edt_fast(f, invert, out);  // edt.cu: bit rows -> plane kernel (z + y) -> axis kernel (x)
  if (invert)
    edt_pass_z_kernel<true><<<blocks, threads, 0, f->stream>>>(f->occ, f->tmp_a, nz, n, f->radius);
  else
    edt_pass_z_kernel<false><<<blocks, threads, 0, f->stream>>>(f->occ, f->tmp_a, nz, n, f->radius);
  B200DF_LAUNCHED();
  edt_pass_axis_kernel<uint16_t, false>
      <<<blocks, threads, 0, f->stream>>>(f->tmp_a, f->tmp_b, ny, (long)nz, n, f->radius, f->max_d2);
  B200DF_LAUNCHED();
  edt_pass_axis_kernel<T, true>
      <<<blocks, threads, 0, f->stream>>>(f->tmp_b, out, nx, (long)ny * nz, n, f->radius, f->max_d2);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}

int rebuild_locked(b200df_field* f)
{
  if (!f->dirty)
    return B200DF_OK;
  if (f->n_voxels == 0)
  {
    f->dirty = false;
    return B200DF_OK;
  }
  B200DF_CUDA(cudaEventRecord(f->ev0, f->stream));
  int rc;
  if (f->elem_size == 1)
    rc = run_edt<uint8_t>(f, false, static_cast<uint8_t*>(f->d2));
  else
    rc = run_edt<uint16_t>(f, false, static_cast<uint16_t*>(f->d2));
  if (rc)
    return rc;
  if (f->desc.propagate_negative)
  {
    if (f->elem_size == 1)
      rc = run_edt<uint8_t>(f, true, static_cast<uint8_t*>(f->neg));
    else
      rc = run_edt<uint16_t>(f, true, static_cast<uint16_t*>(f->neg));
    if (rc)
      return rc;
  }
  B200DF_CUDA(cudaEventRecord(f->ev1, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  B200DF_CUDA(cudaEventElapsedTime(&f->last_rebuild_ms, f->ev0, f->ev1));
  f->dirty = false;
  f->uploaded = false;  // the arrays hold the engine's own exact EDT again
  return B200DF_OK;
}

void fill_view(const b200df_field* f, FieldDev* v)
{
  v->g = f->grid;
  v->d2 = f->d2;
  v->neg = f->desc.propagate_negative ? f->neg : nullptr;
  v->sqrt_table = f->sqrt_table;
  v->elem_size = f->elem_size;
  v->max_d2 = f->max_d2;
  v->max_distance = f->desc.max_distance;
  v->inv_twice_res = f->inv_twice_res;
}

struct DeviceBuffer
{
  void* p = nullptr;
  ~DeviceBuffer()
  {
    cudaFree(p);
  }
};
}  // namespace

namespace b200df
{
int field_prepare(b200df_field* f, FieldDev* view, cudaStream_t /*wait_on*/)
{
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  int rc = rebuild_locked(f);
  if (rc)
    return rc;
  fill_view(f, view);
  return B200DF_OK;
}
}  // namespace b200df

extern "C" {

int b200df_field_create(const b200df_field_desc* desc, b200df_field** out)
{
  B200DF_REQUIRE(desc && out, "null argument");
  B200DF_REQUIRE(desc->resolution > 0.0 && desc->max_distance >= 0.0, "resolution must be > 0 and max_distance >= 0");
  int ndev = 0;
  B200DF_CUDA(cudaGetDeviceCount(&ndev));
  B200DF_REQUIRE(ndev > 0, "no CUDA device");
  struct Destroy  // an early return (e.g. out of device memory half way) releases what was allocated so far
  {
    void operator()(b200df_field* p) const
    {
      b200df_field_destroy(p);
    }
  };
  std::unique_ptr<b200df_field, Destroy> f(new b200df_field);
  f->device = current_device();
  f->desc = *desc;
  // VoxelGrid::resize (voxel_grid.h:367-399)
  const double res = desc->resolution;
  const double oo = 1.0 / res;
  f->grid.oo = oo;
  long total = 1;
  for (int i = 0; i < 3; ++i)
  {
    f->grid.om[i] = desc->origin[i] - 0.5 * res;
    const double cells = desc->size[i] * oo;
    B200DF_REQUIRE(cells >= 0.0 && cells < 65536.0, "field dimension out of range");
    f->grid.n[i] = (int)cells;  // truncation, voxel_grid.h:387
    total *= f->grid.n[i];
  }
  f->grid.stride1 = (long)f->grid.n[1] * f->grid.n[2];
  f->grid.stride2 = f->grid.n[2];
  f->n_voxels = total;
  // PropagationDistanceField::initialize (propagation_distance_field.cpp:76-94)
  const double c = std::ceil(desc->max_distance / res);
  B200DF_REQUIRE(c * c <= 21609.0, "max_distance/resolution too large (max_distance_sq_ must be <= 147^2)");
  f->max_d2 = (int)(c * c);
  int R = 0;
  while (R * R < f->max_d2)
    ++R;
  f->radius = R;
  f->elem_size = f->max_d2 <= 255 ? 1 : 2;
  f->inv_twice_res = (int)(1.0 / (2.0 * res));  // int member, distance_field.h:622 (Q3)
  f->sqrt_table_host.resize(f->max_d2 + 1);
  for (int i = 0; i <= f->max_d2; ++i)
    f->sqrt_table_host[i] = std::sqrt(double(i)) * res;

  DeviceGuard dg(f->device);
  B200DF_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  B200DF_CUDA(cudaEventCreate(&f->ev0));
  B200DF_CUDA(cudaEventCreate(&f->ev1));
  B200DF_CUDA(cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < b200df_field::kEdtStreams; ++i)
  {
    B200DF_CUDA(cudaStreamCreateWithFlags(&f->edt_streams[i], cudaStreamNonBlocking));
    B200DF_CUDA(cudaEventCreateWithFlags(&f->ev_join[i], cudaEventDisableTiming));
  }
  const size_t nv = (size_t)std::max<long>(total, 1);
  B200DF_CUDA(cudaMalloc(&f->occ, nv));
  B200DF_CUDA(cudaMalloc(&f->d2, nv * f->elem_size));
  if (desc->propagate_negative)
    B200DF_CUDA(cudaMalloc(&f->neg, nv * f->elem_size));
  B200DF_CUDA(cudaMalloc(&f->tmp_a, nv * sizeof(uint16_t)));
  B200DF_CUDA(cudaMalloc(&f->tmp_b, nv * sizeof(uint16_t)));
  if (total > 0 && f->grid.n[2] % 2 == 0 && f->radius <= 31 && !getenv("B200DF_EDT_WINDOWED"))
  {
    const size_t bit_bytes = (size_t)f->grid.n[0] * f->grid.n[1] * bits_wpr(f->grid.n[2]) * sizeof(uint32_t);
    B200DF_CUDA(cudaMalloc(&f->bits, bit_bytes));
    B200DF_CUDA(cudaMemset(f->bits, 0, bit_bytes));  // padding bits of a row stay 0 (= no obstacle) forever
  }
  B200DF_CUDA(cudaMalloc(&f->sqrt_table, f->sqrt_table_host.size() * sizeof(double)));
  B200DF_CUDA(cudaMemcpy(f->sqrt_table, f->sqrt_table_host.data(), f->sqrt_table_host.size() * sizeof(double),
                         cudaMemcpyHostToDevice));
  B200DF_CUDA(cudaMemsetAsync(f->occ, 0, nv, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  f->dirty = true;
  f->bits_valid = false;
  *out = f.release();
  return B200DF_OK;
}

int b200df_field_destroy(b200df_field* f)
{
  if (!f)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  cudaFree(f->occ);
  cudaFree(f->scratch8);
  cudaFree(f->pts_dev);
  if (f->pts_pin)
    cudaFreeHost(f->pts_pin);
  cudaFree(f->d2);
  cudaFree(f->neg);
  cudaFree(f->tmp_a);
  cudaFree(f->tmp_b);
  cudaFree(f->bits);
  cudaFree(f->sqrt_table);
  if (f->ev0)
    cudaEventDestroy(f->ev0);
  if (f->ev1)
    cudaEventDestroy(f->ev1);
  if (f->ev_fork)
    cudaEventDestroy(f->ev_fork);
  for (int i = 0; i < b200df_field::kEdtStreams; ++i)
  {
    if (f->ev_join[i])
      cudaEventDestroy(f->ev_join[i]);
    if (f->edt_streams[i])
      cudaStreamDestroy(f->edt_streams[i]);
  }
  if (f->stream)
    cudaStreamDestroy(f->stream);
  delete f;
  return B200DF_OK;
}

int b200df_field_dims(const b200df_field* f, int32_t* out4)
{
  B200DF_REQUIRE(f && out4, "null argument");
  out4[0] = f->grid.n[0];
  out4[1] = f->grid.n[1];
  out4[2] = f->grid.n[2];
  out4[3] = f->max_d2;
  return B200DF_OK;
}

int b200df_field_get_desc(const b200df_field* f, b200df_field_desc* out)
{
  B200DF_REQUIRE(f && out, "null argument");
  *out = f->desc;
  return B200DF_OK;
}

int b200df_field_reset(b200df_field* f)
{
  B200DF_REQUIRE(f, "null field");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  if (f->n_voxels > 0)
    B200DF_CUDA(cudaMemsetAsync(f->occ, 0, (size_t)f->n_voxels, f->stream));
  if (f->bits)
    B200DF_CUDA(cudaMemsetAsync(f->bits, 0, bits_bytes(f), f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  f->dirty = true;
  f->bits_valid = f->bits != nullptr;
  return B200DF_OK;
}

// Host point list -> voxels.  Small lists: one copy + one kernel on the field's stream.  Large lists (a config-3
// cloud is 24 MB): the list is cut into slices; a few host threads each copy their slice into the field's page-locked
// staging buffer and queue its H2D copy and its scatter kernel on a stream of their own, so the CPU-side staging of
// one slice overlaps the DMA and the scatter of the others (measured before: 2.09 ms for 1 M points through a
// cudaMalloc + pageable cudaMemcpy + kernel + cudaFree sequence against 0.03 ms for the scatter kernel alone).
static int scatter_host_points(b200df_field* f, const double* xyz, int64_t n, uint8_t* vol, uint8_t value, bool orv)
{
  if (n <= 0 || f->n_voxels == 0)
    return B200DF_OK;
  const size_t bytes = (size_t)n * 3 * sizeof(double);
  if (f->pts_dev_bytes < bytes)
  {
    cudaFree(f->pts_dev);
    f->pts_dev = nullptr;
    f->pts_dev_bytes = 0;
    B200DF_CUDA(cudaMalloc(&f->pts_dev, bytes));
    f->pts_dev_bytes = bytes;
  }
  const bool mirror = vol == f->occ && value == 1 && !orv && f->bits && f->bits_valid;
  const int wpr = bits_wpr(f->grid.n[2]);
  const int64_t kSlicePoints = (int64_t)(2 << 20) / 24;  // ~2 MB of points per slice
  const unsigned hw = std::thread::hardware_concurrency();
  // (B200DF_ADD_POINTS_WORKERS: measurement knob; workers beyond the field's four side streams share them.  1 M points
  // on a 16-core box: 4 workers 0.75 ms, 8 workers 0.82 ms; 0.89 ms with 4 MB slices)
  static const int max_workers = getenv("B200DF_ADD_POINTS_WORKERS") ? atoi(getenv("B200DF_ADD_POINTS_WORKERS")) : 4;
  const int n_workers = (int)std::min<int64_t>(std::min<unsigned>((unsigned)std::max(1, max_workers), std::max(1u, hw / 2)),
                                               n / kSlicePoints);
  if (n_workers <= 1)
  {
    B200DF_CUDA(cudaMemcpyAsync(f->pts_dev, xyz, bytes, cudaMemcpyHostToDevice, f->stream));
    scatter_points_kernel<<<blocks_for(n, 256), 256, 0, f->stream>>>(f->grid, f->pts_dev, n, vol, value, orv,
                                                                     mirror ? f->bits : nullptr, wpr);
    B200DF_LAUNCHED();
    B200DF_CUDA(cudaGetLastError());
    B200DF_CUDA(cudaStreamSynchronize(f->stream));
    return B200DF_OK;
  }
  if (f->pts_pin_bytes < bytes)
  {
    if (f->pts_pin)
      cudaFreeHost(f->pts_pin);
    f->pts_pin = nullptr;
    f->pts_pin_bytes = 0;
    B200DF_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&f->pts_pin), bytes, cudaHostAllocDefault));
    f->pts_pin_bytes = bytes;
  }
  B200DF_CUDA(cudaStreamSynchronize(f->stream));  // earlier edits of the occupancy are done before the workers write
  const int64_t per = (n + n_workers - 1) / n_workers;
  std::vector<cudaError_t> err((size_t)n_workers, cudaSuccess);
  auto work = [&](int w) {
    const int64_t p0 = (int64_t)w * per, p1 = std::min<int64_t>(n, p0 + per);
    if (p0 >= p1)
      return;
    cudaError_t e = cudaSetDevice(f->device);
    cudaStream_t st = f->edt_streams[w % b200df_field::kEdtStreams];
    // sub-slices keep the copy engine busy while this thread is still staging the rest of its share
    for (int64_t q0 = p0; q0 < p1 && e == cudaSuccess; q0 += kSlicePoints)
    {
      const int64_t q1 = std::min<int64_t>(p1, q0 + kSlicePoints);
      const size_t off = (size_t)q0 * 24, len = (size_t)(q1 - q0) * 24;
      stream_copy(f->pts_pin + off, reinterpret_cast<const char*>(xyz) + off, len);
      e = cudaMemcpyAsync(reinterpret_cast<char*>(f->pts_dev) + off, f->pts_pin + off, len, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess)
        break;
      scatter_points_kernel<<<blocks_for(q1 - q0, 256), 256, 0, st>>>(f->grid, f->pts_dev + 3 * q0, q1 - q0, vol, value,
                                                                      orv, mirror ? f->bits : nullptr, wpr);
      B200DF_LAUNCHED();
      e = cudaGetLastError();
    }
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(st);
    err[(size_t)w] = e;
  };
  std::vector<std::thread> workers;
  for (int w = 1; w < n_workers; ++w)
    workers.emplace_back(work, w);
  work(0);
  for (std::thread& t : workers)
    t.join();
  for (cudaError_t e : err)
    B200DF_CUDA(e);
  return B200DF_OK;
}

int b200df_field_add_points(b200df_field* f, const double* xyz, int64_t n)
{
  B200DF_REQUIRE(f && (xyz || n == 0) && n >= 0, "bad arguments");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  const bool keep = f->bits_valid;  // scatter_host_points mirrors the new voxels into valid bit rows
  int rc = scatter_host_points(f, xyz, n, f->occ, 1, false);
  f->dirty = true;
  f->bits_valid = keep && rc == B200DF_OK;
  return rc;
}

int b200df_field_add_points_device(b200df_field* f, const double* xyz_dev, int64_t n, void* stream)
{
  B200DF_REQUIRE(f && (xyz_dev || n == 0) && n >= 0, "bad arguments");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  if (n > 0 && f->n_voxels > 0)
  {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    scatter_points_kernel<<<blocks_for(n, 256), 256, 0, st>>>(f->grid, xyz_dev, n, f->occ, 1, false,
                                                              f->bits_valid ? f->bits : nullptr, bits_wpr(f->grid.n[2]));
    B200DF_LAUNCHED();
    B200DF_CUDA(cudaGetLastError());
    // the rebuild runs on the field's own stream: order it after the caller's scatter
    B200DF_CUDA(cudaEventRecord(f->ev1, st));
    B200DF_CUDA(cudaStreamWaitEvent(f->stream, f->ev1, 0));
  }
  f->dirty = true;  // bits_valid unchanged: the scatter mirrored the voxels when the bit rows were current
  return B200DF_OK;
}

int b200df_field_remove_points(b200df_field* f, const double* xyz, int64_t n)
{
  B200DF_REQUIRE(f && (xyz || n == 0) && n >= 0, "bad arguments");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  int rc = scatter_host_points(f, xyz, n, f->occ, 0, false);  // Q8: cleared even if another object shares the voxel
  f->dirty = true;
  f->bits_valid = false;
  return rc;
}

int b200df_field_update_points(b200df_field* f, const double* old_xyz, int64_t n_old, const double* new_xyz,
                               int64_t n_new)
{
  B200DF_REQUIRE(f && n_old >= 0 && n_new >= 0 && (old_xyz || n_old == 0) && (new_xyz || n_new == 0), "bad arguments");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  if (f->n_voxels == 0)
    return B200DF_OK;
  if (!f->scratch8)
    B200DF_CUDA(cudaMalloc(&f->scratch8, (size_t)f->n_voxels));
  B200DF_CUDA(cudaMemsetAsync(f->scratch8, 0, (size_t)f->n_voxels, f->stream));
  int rc = scatter_host_points(f, old_xyz, n_old, f->scratch8, 1, true);
  if (rc)
    return rc;
  rc = scatter_host_points(f, new_xyz, n_new, f->scratch8, 2, true);
  if (rc)
    return rc;
  apply_update_marks_kernel<<<blocks_for(f->n_voxels, 256), 256, 0, f->stream>>>(f->scratch8, f->occ, f->n_voxels);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  f->dirty = true;
  f->bits_valid = false;
  return B200DF_OK;
}

int b200df_field_add_voxels(b200df_field* f, const int32_t* ijk, int64_t n)
{
  B200DF_REQUIRE(f && (ijk || n == 0) && n >= 0, "bad arguments");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  if (n > 0 && f->n_voxels > 0)
  {
    DeviceBuffer buf;
    B200DF_CUDA(cudaMalloc(&buf.p, (size_t)n * 3 * sizeof(int32_t)));
    B200DF_CUDA(cudaMemcpyAsync(buf.p, ijk, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, f->stream));
    scatter_voxels_kernel<<<blocks_for(n, 256), 256, 0, f->stream>>>(f->grid, static_cast<const int32_t*>(buf.p), n,
                                                                     f->occ, 1);
    B200DF_LAUNCHED();
    B200DF_CUDA(cudaGetLastError());
    B200DF_CUDA(cudaStreamSynchronize(f->stream));
  }
  f->dirty = true;
  f->bits_valid = false;
  return B200DF_OK;
}

int b200df_field_add_shape(b200df_field* f, int32_t shape_type, const double* dims3, const double* pose12, double scale,
                           double padding)
{
  B200DF_REQUIRE(f && dims3 && pose12, "null argument");
  return rasterize_into(f, shape_type, dims3, pose12, nullptr, scale, padding, 1);
}

int b200df_field_remove_shape(b200df_field* f, int32_t shape_type, const double* dims3, const double* pose12,
                              double scale, double padding)
{
  B200DF_REQUIRE(f && dims3 && pose12, "null argument");
  return rasterize_into(f, shape_type, dims3, pose12, nullptr, scale, padding, 0);
}

int b200df_field_add_posed_body(b200df_field* f, int32_t shape_type, const double* dims3, const double* pose12,
                                double padding)
{
  B200DF_REQUIRE(f && dims3 && pose12, "null argument");
  static const double identity[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
  return rasterize_into(f, shape_type, dims3, identity, pose12, 1.0, padding, 1);
}

int b200df_field_remove_posed_body(b200df_field* f, int32_t shape_type, const double* dims3, const double* pose12,
                                   double padding)
{
  B200DF_REQUIRE(f && dims3 && pose12, "null argument");
  static const double identity[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
  return rasterize_into(f, shape_type, dims3, identity, pose12, 1.0, padding, 0);  // Q8: shared voxels are cleared too
}

static int octree_leaves(b200df_field* f, const double* leaves4, int64_t n, uint8_t value)
{
  B200DF_REQUIRE(f && (leaves4 || n == 0) && n >= 0, "bad arguments");
  if (n == 0 || f->n_voxels == 0)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  const double res = f->desc.resolution;
  DeviceBuffer dl, dc, dt, doff, dtot;
  B200DF_CUDA(cudaMalloc(&dl.p, (size_t)n * 4 * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&dc.p, (size_t)n * 3 * sizeof(int)));
  B200DF_CUDA(cudaMalloc(&dt.p, (size_t)n * sizeof(unsigned long long)));
  B200DF_CUDA(cudaMalloc(&doff.p, (size_t)n * sizeof(unsigned long long)));
  B200DF_CUDA(cudaMalloc(&dtot.p, sizeof(unsigned long long)));
  B200DF_CUDA(cudaMemcpyAsync(dl.p, leaves4, (size_t)n * 4 * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  octree_count_kernel<<<blocks_for(n, 256), 256, 0, f->stream>>>(static_cast<const double*>(dl.p), n, res,
                                                                 static_cast<int*>(dc.p),
                                                                 static_cast<unsigned long long*>(dt.p));
  B200DF_LAUNCHED();
  exclusive_scan_kernel<<<1, 1024, 0, f->stream>>>(static_cast<const unsigned long long*>(dt.p),
                                                   static_cast<unsigned long long*>(doff.p), n,
                                                   static_cast<unsigned long long*>(dtot.p));
  B200DF_LAUNCHED();
  unsigned long long total = 0;
  B200DF_CUDA(cudaMemcpyAsync(&total, dtot.p, sizeof(total), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  B200DF_REQUIRE(total < (1ull << 40), "octree leaves expand to an absurd number of lattice points");
  const bool mirror = value == 1 && f->bits && f->bits_valid;
  if (total > 0)
  {
    octree_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, f->stream>>>(
        f->grid, static_cast<const double*>(dl.p), n, static_cast<const int*>(dc.p),
        static_cast<const unsigned long long*>(doff.p), total, res, f->occ, value, mirror ? f->bits : nullptr,
        bits_wpr(f->grid.n[2]));
    B200DF_LAUNCHED();
    B200DF_CUDA(cudaGetLastError());
    B200DF_CUDA(cudaStreamSynchronize(f->stream));
  }
  f->dirty = true;
  f->bits_valid = mirror;  // set voxels were mirrored into current bit rows; anything else re-packs at the rebuild
  return B200DF_OK;
}

int b200df_field_add_octree_leaves(b200df_field* f, const double* leaves4, int64_t n)
{
  return octree_leaves(f, leaves4, n, 1);
}

int b200df_field_remove_octree_leaves(b200df_field* f, const double* leaves4, int64_t n)
{
  return octree_leaves(f, leaves4, n, 0);  // Q8: voxels shared with other objects are cleared too
}

int b200df_field_rebuild(b200df_field* f, float* ms)
{
  B200DF_REQUIRE(f, "null field");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  const bool was_dirty = f->dirty;
  int rc = rebuild_locked(f);
  if (ms)
    *ms = was_dirty ? f->last_rebuild_ms : 0.f;
  return rc;
}

int b200df_field_download_d2(b200df_field* f, int32_t* out, int negative)
{
  B200DF_REQUIRE(f && out, "null argument");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  int rc = rebuild_locked(f);
  if (rc)
    return rc;
  if (f->n_voxels == 0)
    return B200DF_OK;
  if (negative && !f->desc.propagate_negative)
  {
    std::memset(out, 0, (size_t)f->n_voxels * sizeof(int32_t));  // reset(): negative_distance_square_ = 0
    return B200DF_OK;
  }
  DeviceBuffer wide;
  B200DF_CUDA(cudaMalloc(&wide.p, (size_t)f->n_voxels * sizeof(int32_t)));
  const void* src = negative ? f->neg : f->d2;
  if (f->elem_size == 1)
    widen_kernel<uint8_t><<<blocks_for(f->n_voxels, 256), 256, 0, f->stream>>>(static_cast<const uint8_t*>(src),
                                                                                static_cast<int32_t*>(wide.p),
                                                                                f->n_voxels);
  else
    widen_kernel<uint16_t><<<blocks_for(f->n_voxels, 256), 256, 0, f->stream>>>(static_cast<const uint16_t*>(src),
                                                                                 static_cast<int32_t*>(wide.p),
                                                                                 f->n_voxels);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  B200DF_CUDA(cudaMemcpyAsync(out, wide.p, (size_t)f->n_voxels * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  return B200DF_OK;
}

int b200df_field_download_occupancy(b200df_field* f, uint8_t* out)
{
  B200DF_REQUIRE(f && out, "null argument");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  if (f->n_voxels == 0)
    return B200DF_OK;
  B200DF_CUDA(cudaMemcpyAsync(out, f->occ, (size_t)f->n_voxels, cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  return B200DF_OK;
}

int b200df_field_distance_gradient(b200df_field* f, const double* xyz, int64_t n, double* dist, double* grad,
                                   uint8_t* in_bounds)
{
  B200DF_REQUIRE(f && n >= 0 && (n == 0 || (xyz && dist && grad && in_bounds)), "bad arguments");
  if (n == 0)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  int rc = rebuild_locked(f);
  if (rc)
    return rc;
  FieldDev view;
  fill_view(f, &view);
  DeviceBuffer dx, dd, dgr, dib;
  B200DF_CUDA(cudaMalloc(&dx.p, (size_t)n * 3 * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&dd.p, (size_t)n * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&dgr.p, (size_t)n * 3 * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&dib.p, (size_t)n));
  B200DF_CUDA(cudaMemcpyAsync(dx.p, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, f->stream));
  if (f->elem_size == 1)
    distance_gradient_kernel<uint8_t><<<blocks_for(n, 256), 256, 0, f->stream>>>(
        view, static_cast<const double*>(dx.p), n, static_cast<double*>(dd.p), static_cast<double*>(dgr.p),
        static_cast<uint8_t*>(dib.p));
  else
    distance_gradient_kernel<uint16_t><<<blocks_for(n, 256), 256, 0, f->stream>>>(
        view, static_cast<const double*>(dx.p), n, static_cast<double*>(dd.p), static_cast<double*>(dgr.p),
        static_cast<uint8_t*>(dib.p));
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  B200DF_CUDA(cudaMemcpyAsync(dist, dd.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaMemcpyAsync(grad, dgr.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaMemcpyAsync(in_bounds, dib.p, (size_t)n, cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  return B200DF_OK;
}

int b200df_field_pack_occupancy(b200df_field* f, uint8_t* out, int64_t cap, int64_t* n_bytes)
{
  B200DF_REQUIRE(f && n_bytes, "null argument");
  const int nz = f->grid.n[2];
  const int zbytes = (nz + 7) / 8;
  const long rows = (long)f->grid.n[0] * f->grid.n[1];
  *n_bytes = rows * zbytes;
  if (!out || cap < *n_bytes || *n_bytes == 0)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  DeviceBuffer buf;
  B200DF_CUDA(cudaMalloc(&buf.p, (size_t)*n_bytes));
  pack_occupancy_kernel<<<blocks_for(*n_bytes, 256), 256, 0, f->stream>>>(f->occ, static_cast<uint8_t*>(buf.p), nz,
                                                                          zbytes, rows);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  B200DF_CUDA(cudaMemcpyAsync(out, buf.p, (size_t)*n_bytes, cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  return B200DF_OK;
}

int b200df_field_unpack_occupancy(b200df_field* f, const uint8_t* bytes, int64_t n_bytes)
{
  B200DF_REQUIRE(f && bytes, "null argument");
  const int nz = f->grid.n[2];
  const int zbytes = (nz + 7) / 8;
  const long rows = (long)f->grid.n[0] * f->grid.n[1];
  B200DF_REQUIRE(n_bytes == rows * zbytes, "payload size does not match the field dimensions");
  if (n_bytes == 0)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  DeviceBuffer buf;
  B200DF_CUDA(cudaMalloc(&buf.p, (size_t)n_bytes));
  B200DF_CUDA(cudaMemcpyAsync(buf.p, bytes, (size_t)n_bytes, cudaMemcpyHostToDevice, f->stream));
  unpack_occupancy_kernel<<<blocks_for(f->n_voxels, 256), 256, 0, f->stream>>>(static_cast<const uint8_t*>(buf.p),
                                                                              f->occ, nz, zbytes, rows);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  f->dirty = true;
  f->bits_valid = false;
  return B200DF_OK;
}

int b200df_field_device_arrays(b200df_field* f, void** d2, void** neg_d2, int64_t* n_bytes, int32_t* elem_size)
{
  B200DF_REQUIRE(f, "null field");
  if (d2)
    *d2 = f->d2;
  if (neg_d2)
    *neg_d2 = f->neg;
  if (n_bytes)
    *n_bytes = (int64_t)f->n_voxels * f->elem_size;
  if (elem_size)
    *elem_size = f->elem_size;
  return B200DF_OK;
}

int b200df_field_mark_built(b200df_field* f)
{
  B200DF_REQUIRE(f, "null field");
  std::lock_guard<std::mutex> lk(f->mu);
  f->dirty = false;
  return B200DF_OK;
}

int b200df_field_copy_from(b200df_field* dst, b200df_field* src)
{
  B200DF_REQUIRE(dst && src && dst != src, "bad arguments");
  B200DF_REQUIRE(dst->n_voxels == src->n_voxels && dst->elem_size == src->elem_size &&
                     dst->grid.n[0] == src->grid.n[0] && dst->grid.n[1] == src->grid.n[1] &&
                     dst->grid.n[2] == src->grid.n[2] &&
                     (dst->desc.propagate_negative != 0) == (src->desc.propagate_negative != 0),
                 "fields have different geometry");
  {
    DeviceGuard dg(src->device);
    std::lock_guard<std::mutex> lk(src->mu);
    int rc = rebuild_locked(src);
    if (rc)
      return rc;
  }
  DeviceGuard dg(dst->device);
  std::lock_guard<std::mutex> lk(dst->mu);
  const size_t nb = (size_t)src->n_voxels * src->elem_size;
  if (nb == 0)
    return B200DF_OK;
  // NVLink peer copy (or plain D2D on one device): one replication per rebuild, no other collective on this path
  B200DF_CUDA(cudaMemcpyPeerAsync(dst->d2, dst->device, src->d2, src->device, nb, dst->stream));
  B200DF_CUDA(cudaMemcpyPeerAsync(dst->occ, dst->device, src->occ, src->device, (size_t)src->n_voxels, dst->stream));
  if (src->desc.propagate_negative)
    B200DF_CUDA(cudaMemcpyPeerAsync(dst->neg, dst->device, src->neg, src->device, nb, dst->stream));
  B200DF_CUDA(cudaStreamSynchronize(dst->stream));
  dst->dirty = false;
  dst->bits_valid = false;
  return B200DF_OK;
}

int b200df_field_export_device(b200df_field* f, void* occ_dev, void* d2_dev, void* neg_d2_dev, void* stream)
{
  B200DF_REQUIRE(f, "null field");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  int rc = rebuild_locked(f);
  if (rc)
    return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nb = (size_t)f->n_voxels * f->elem_size;
  if (occ_dev && f->n_voxels)
    B200DF_CUDA(cudaMemcpyAsync(occ_dev, f->occ, (size_t)f->n_voxels, cudaMemcpyDeviceToDevice, st));
  if (d2_dev && nb)
    B200DF_CUDA(cudaMemcpyAsync(d2_dev, f->d2, nb, cudaMemcpyDeviceToDevice, st));
  if (neg_d2_dev && nb && f->neg)
    B200DF_CUDA(cudaMemcpyAsync(neg_d2_dev, f->neg, nb, cudaMemcpyDeviceToDevice, st));
  B200DF_CUDA(cudaStreamSynchronize(st));
  return B200DF_OK;
}

int b200df_field_import_device(b200df_field* f, const void* occ_dev, const void* d2_dev, const void* neg_d2_dev,
                               void* stream)
{
  B200DF_REQUIRE(f, "null field");
  // a signed field's two arrays belong together: d2 without its negative half would leave a stale negative array
  B200DF_REQUIRE(!(d2_dev && f->neg && !neg_d2_dev), "a signed field needs neg_d2_dev together with d2_dev");
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nb = (size_t)f->n_voxels * f->elem_size;
  if (occ_dev && f->n_voxels)
    B200DF_CUDA(cudaMemcpyAsync(f->occ, occ_dev, (size_t)f->n_voxels, cudaMemcpyDeviceToDevice, st));
  if (d2_dev && nb)
    B200DF_CUDA(cudaMemcpyAsync(f->d2, d2_dev, nb, cudaMemcpyDeviceToDevice, st));
  if (neg_d2_dev && nb && f->neg)
    B200DF_CUDA(cudaMemcpyAsync(f->neg, neg_d2_dev, nb, cudaMemcpyDeviceToDevice, st));
  B200DF_CUDA(cudaStreamSynchronize(st));
  if (d2_dev)
    f->dirty = false;  // the imported distances are served as they are (they belong to the imported occupancy)
  else if (occ_dev)
    f->dirty = true;   // occupancy alone: the distances are stale and are rebuilt before the next query
  if (occ_dev)
    f->bits_valid = false;
  return B200DF_OK;
}

int b200df_field_upload_d2(b200df_field* f, const int32_t* d2, const int32_t* neg_d2)
{
  B200DF_REQUIRE(f && d2, "null argument");
  B200DF_REQUIRE(!f->desc.propagate_negative || neg_d2, "a signed field needs the negative distances too");
  if (f->n_voxels == 0)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  const size_t nb = (size_t)f->n_voxels * sizeof(int32_t);
  DeviceBuffer a, b, bad;
  B200DF_CUDA(cudaMalloc(&a.p, nb));
  B200DF_CUDA(cudaMalloc(&bad.p, sizeof(int)));
  B200DF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), f->stream));
  B200DF_CUDA(cudaMemcpyAsync(a.p, d2, nb, cudaMemcpyHostToDevice, f->stream));
  const bool with_neg = f->desc.propagate_negative && neg_d2;
  if (with_neg)
  {
    B200DF_CUDA(cudaMalloc(&b.p, nb));
    B200DF_CUDA(cudaMemcpyAsync(b.p, neg_d2, nb, cudaMemcpyHostToDevice, f->stream));
  }
  // staged into the EDT scratch first, so that a refused upload leaves the field untouched
  const unsigned blocks = blocks_for(f->n_voxels, 256);
  if (!f->scratch8)
    B200DF_CUDA(cudaMalloc(&f->scratch8, (size_t)f->n_voxels));
  if (f->elem_size == 1)
    narrow_kernel<uint8_t><<<blocks, 256, 0, f->stream>>>(
        static_cast<const int32_t*>(a.p), static_cast<const int32_t*>(b.p), reinterpret_cast<uint8_t*>(f->tmp_a),
        reinterpret_cast<uint8_t*>(f->tmp_b), f->scratch8, f->n_voxels, f->max_d2, static_cast<int*>(bad.p));
  else
    narrow_kernel<uint16_t><<<blocks, 256, 0, f->stream>>>(static_cast<const int32_t*>(a.p),
                                                           static_cast<const int32_t*>(b.p), f->tmp_a, f->tmp_b,
                                                           f->scratch8, f->n_voxels, f->max_d2,
                                                           static_cast<int*>(bad.p));
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  int host_bad = 0;
  B200DF_CUDA(cudaMemcpyAsync(&host_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  B200DF_REQUIRE(!host_bad, "uploaded squared distances must lie in [0, max_distance_sq_]");
  const size_t ne = (size_t)f->n_voxels * f->elem_size;
  B200DF_CUDA(cudaMemcpyAsync(f->d2, f->tmp_a, ne, cudaMemcpyDeviceToDevice, f->stream));
  if (with_neg)
    B200DF_CUDA(cudaMemcpyAsync(f->neg, f->tmp_b, ne, cudaMemcpyDeviceToDevice, f->stream));
  B200DF_CUDA(cudaMemcpyAsync(f->occ, f->scratch8, (size_t)f->n_voxels, cudaMemcpyDeviceToDevice, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  f->dirty = false;
  f->bits_valid = false;
  f->uploaded = true;
  return B200DF_OK;
}

int b200df_field_values_uploaded(const b200df_field* f, int32_t* uploaded)
{
  B200DF_REQUIRE(f && uploaded, "null argument");
  *uploaded = f->uploaded && !f->dirty ? 1 : 0;
  return B200DF_OK;
}

int b200df_field_nearest_cells(b200df_field* f, const int32_t* ijk, int64_t n, double* dist, int32_t* pos,
                               uint8_t* found)
{
  B200DF_REQUIRE(f && n >= 0 && (n == 0 || (ijk && dist && pos && found)), "bad arguments");
  if (n == 0)
    return B200DF_OK;
  DeviceGuard dg(f->device);
  std::lock_guard<std::mutex> lk(f->mu);
  int rc = rebuild_locked(f);
  if (rc)
    return rc;
  DeviceBuffer dq, dd, dp, df;
  B200DF_CUDA(cudaMalloc(&dq.p, (size_t)n * 3 * sizeof(int32_t)));
  B200DF_CUDA(cudaMalloc(&dd.p, (size_t)n * sizeof(double)));
  B200DF_CUDA(cudaMalloc(&dp.p, (size_t)n * 3 * sizeof(int32_t)));
  B200DF_CUDA(cudaMalloc(&df.p, (size_t)n));
  B200DF_CUDA(cudaMemcpyAsync(dq.p, ijk, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, f->stream));
  const void* neg = f->desc.propagate_negative ? f->neg : nullptr;
  if (f->elem_size == 1)
    nearest_cell_kernel<uint8_t><<<blocks_for(n, 128), 128, 0, f->stream>>>(
        f->grid, static_cast<const uint8_t*>(f->d2), static_cast<const uint8_t*>(neg), f->occ, f->sqrt_table, f->max_d2,
        static_cast<const int32_t*>(dq.p), n, static_cast<double*>(dd.p), static_cast<int32_t*>(dp.p),
        static_cast<uint8_t*>(df.p));
  else
    nearest_cell_kernel<uint16_t><<<blocks_for(n, 128), 128, 0, f->stream>>>(
        f->grid, static_cast<const uint16_t*>(f->d2), static_cast<const uint16_t*>(neg), f->occ, f->sqrt_table,
        f->max_d2, static_cast<const int32_t*>(dq.p), n, static_cast<double*>(dd.p), static_cast<int32_t*>(dp.p),
        static_cast<uint8_t*>(df.p));
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  B200DF_CUDA(cudaMemcpyAsync(dist, dd.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaMemcpyAsync(pos, dp.p, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaMemcpyAsync(found, df.p, (size_t)n, cudaMemcpyDeviceToHost, f->stream));
  B200DF_CUDA(cudaStreamSynchronize(f->stream));
  return B200DF_OK;
}

// BodyDecomposition (collision_distance_field_types.cpp:290-334)
int b200df_body_decomposition(int32_t n_shapes, const int32_t* shape_types, const double* dims, const double* poses,
                              double resolution, double padding, double* spheres, int64_t sphere_cap,
                              int64_t* n_spheres, double* points, int64_t point_cap, int64_t* n_points, double* bsphere)
{
  B200DF_REQUIRE(n_shapes >= 0 && n_spheres && n_points && bsphere && resolution > 0.0, "bad arguments");
  B200DF_REQUIRE(n_shapes == 0 || (shape_types && dims && poses), "null shape arrays");
  std::vector<double> sph, pts;
  struct BS
  {
    double c[3], r;
  };
  std::vector<BS> bounds;
  int ndev = 0;
  B200DF_CUDA(cudaGetDeviceCount(&ndev));
  B200DF_REQUIRE(ndev > 0, "no CUDA device");
  DeviceGuard dg(current_device());
  for (int s = 0; s < n_shapes; ++s)
  {
    HostBody hb;
    int rc = make_body(shape_types[s], dims + 3 * s, poses + 12 * s, 1.0, padding, &hb);
    if (rc)
      return rc;
    const double* P = poses + 12 * s;
    // determineCollisionSpheres (collision_distance_field_types.cpp:46-79)
    if (hb.dev.type == B200DF_SHAPE_SPHERE)
    {
      if (hb.radiusU > 0.0)
      {
        sph.insert(sph.end(), { hb.dev.center[0], hb.dev.center[1], hb.dev.center[2], hb.radiusU });
      }
    }
    else
    {
      // bodies::*::computeBoundingCylinder
      double cyl_pose[12];
      std::memcpy(cyl_pose, P, sizeof(cyl_pose));
      double cyl_radius, cyl_length;
      if (hb.dev.type == B200DF_SHAPE_CYLINDER)
      {
        cyl_radius = hb.radiusU;
        cyl_length = hb.scaled_length;
      }
      else
      {
        const double l2 = hb.dev.length2, w2 = hb.dev.width2, h2 = hb.dev.height2;
        const double ang = 90.0f * (M_PI / 180.0f);
        const double cs = std::cos(ang), sn = std::sin(ang);
        double rot[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        double a, b;
        if (l2 > w2 && l2 > h2)
        {
          cyl_length = l2 * 2.0;
          a = w2;
          b = h2;
          const double ry[9] = { cs, 0, sn, 0, 1, 0, -sn, 0, cs };  // AngleAxis(90deg, UnitY)
          std::memcpy(rot, ry, sizeof(rot));
        }
        else if (w2 > h2)
        {
          cyl_length = w2 * 2.0;
          a = h2;
          b = l2;
          const double rx[9] = { 1, 0, 0, 0, cs, -sn, 0, sn, cs };  // AngleAxis(90deg, UnitX)
          std::memcpy(rot, rx, sizeof(rot));
        }
        else
        {
          cyl_length = h2 * 2.0;
          a = w2;
          b = l2;
        }
        cyl_radius = std::sqrt(a * a + b * b);
        // cylinder.pose = pose_ * rot   (3x4 * 4x4, translation column unchanged because rot has none)
        for (int r = 0; r < 3; ++r)
        {
          for (int c = 0; c < 3; ++c)
            cyl_pose[4 * r + c] = ((P[4 * r + 0] * rot[c] + P[4 * r + 1] * rot[3 + c]) + P[4 * r + 2] * rot[6 + c]) +
                                  P[4 * r + 3] * 0.0;
          cyl_pose[4 * r + 3] = ((P[4 * r + 0] * 0.0 + P[4 * r + 1] * 0.0) + P[4 * r + 2] * 0.0) + P[4 * r + 3] * 1.0;
        }
      }
      if (cyl_radius > 0.0 && cyl_length > 0.0)
      {
        const unsigned int num_points = (unsigned int)std::ceil(cyl_length / (cyl_radius / 2.0));
        const double spacing = cyl_length / ((num_points * 1.0) - 1.0);
        for (unsigned int i = 1; i + 1 < num_points; i++)  // Q11: both end spheres are dropped
        {
          const double z = (-cyl_length / 2.0) + i * spacing;
          const double cx = cyl_pose[3] + ((cyl_pose[0] * 0.0 + cyl_pose[1] * 0.0) + cyl_pose[2] * z);
          const double cy = cyl_pose[7] + ((cyl_pose[4] * 0.0 + cyl_pose[5] * 0.0) + cyl_pose[6] * z);
          const double cz = cyl_pose[11] + ((cyl_pose[8] * 0.0 + cyl_pose[9] * 0.0) + cyl_pose[10] * z);
          sph.insert(sph.end(), { cx, cy, cz, cyl_radius });
        }
      }
    }
    // findInternalPointsConvex on the GPU, point order = x-major lattice order like the reference loop
    {
      cudaStream_t stream = nullptr;
      DeviceLattice lat;
      rc = upload_lattice(hb, resolution, stream, &lat);
      if (rc)
        return rc;
      const long total = (long)lat.nx * lat.ny * lat.nz;
      if (total > 0)
      {
        DeviceBuffer flags;
        B200DF_CUDA(cudaMalloc(&flags.p, (size_t)total));
        GridDev none{};
        PoseDev pd{};
        rasterize_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(hb.dev, lat.xs, lat.nx, lat.ys, lat.ny, lat.zs,
                                                                     lat.nz, 0, pd, none, nullptr, 0,
                                                                     static_cast<uint8_t*>(flags.p));
        B200DF_LAUNCHED();
        B200DF_CUDA(cudaGetLastError());
        std::vector<uint8_t> hflags((size_t)total);
        B200DF_CUDA(cudaMemcpy(hflags.data(), flags.p, (size_t)total, cudaMemcpyDeviceToHost));
        std::vector<double> xs, ys, zs;
        lattice_axis(hb.dev.center[0], hb.bound_radius, resolution, xs);
        lattice_axis(hb.dev.center[1], hb.bound_radius, resolution, ys);
        lattice_axis(hb.dev.center[2], hb.bound_radius, resolution, zs);
        long k = 0;
        for (size_t a = 0; a < xs.size(); ++a)
          for (size_t b = 0; b < ys.size(); ++b)
            for (size_t c = 0; c < zs.size(); ++c, ++k)
              if (hflags[k])
                pts.insert(pts.end(), { xs[a], ys[b], zs[c] });
      }
    }
    bounds.push_back(BS{ { hb.dev.center[0], hb.dev.center[1], hb.dev.center[2] }, hb.bound_radius });
  }
  // bodies::mergeBoundingSpheres
  BS merged{ { 0, 0, 0 }, 0 };
  if (!bounds.empty())
  {
    merged = bounds[0];
    for (size_t i = 1; i < bounds.size(); ++i)
    {
      const BS& s = bounds[i];
      if (s.r <= 0.0)
        continue;
      const double dx = s.c[0] - merged.c[0], dy = s.c[1] - merged.c[1], dz = s.c[2] - merged.c[2];
      const double d = std::sqrt((dx * dx + dy * dy) + dz * dz);
      if (d + merged.r <= s.r)
        merged = s;
      else if (d + s.r > merged.r)
      {
        const double ex = merged.c[0] - s.c[0], ey = merged.c[1] - s.c[1], ez = merged.c[2] - s.c[2];
        const double en = std::sqrt((ex * ex + ey * ey) + ez * ez);
        const double new_r = (en + s.r + merged.r) / 2.0;
        const double k = new_r - s.r;
        merged.c[0] = ex / en * k + s.c[0];
        merged.c[1] = ey / en * k + s.c[1];
        merged.c[2] = ez / en * k + s.c[2];
        merged.r = new_r;
      }
    }
  }
  *n_spheres = (int64_t)(sph.size() / 4);
  *n_points = (int64_t)(pts.size() / 3);
  if (spheres && sphere_cap >= *n_spheres)
    std::memcpy(spheres, sph.data(), sph.size() * sizeof(double));
  if (points && point_cap >= *n_points)
    std::memcpy(points, pts.data(), pts.size() * sizeof(double));
  bsphere[0] = merged.c[0];
  bsphere[1] = merged.c[1];
  bsphere[2] = merged.c[2];
  bsphere[3] = merged.r;
  return B200DF_OK;
}

}  // extern "C"
