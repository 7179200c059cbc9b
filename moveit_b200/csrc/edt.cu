// Kernel 5, fast path: exact clamped squared EDT from bit-packed occupancy in TWO launches.
//
// Replaces PropagationDistanceField::propagatePositive / propagateNegative (reference:
// moveit_core/distance_field/src/propagation_distance_field.cpp:389-505) for grids with an even nz and a clamp radius
// <= 31 cells; layout from voxel_grid.h:426-429 (z fastest).  d2(v) = min over obstacle voxels q of |v-q|^2, clamped
// to max_d2.  Any voxel whose true distance is below max_d2 has every axis offset <= reach = radius-1, so windows of
// +-reach per axis are exact after the final clamp (a tap at the radius itself adds >= max_d2).
//
//   plane kernel (edt_plane_kernel): a thread owns the voxel pair (x, z0, z0+1) and marches along y.  The input of a
//       step is the pair's squared z distance to the nearest obstacle of row (x, y), computed on the fly from the
//       row's bit words (two bit scans) - the former z pass and its 2 bytes/voxel round trip through HBM are gone.
//       Output: the 2-D (y, z) squared distance as u16x2, values >= max_d2 normalised to INF.
//   axis kernel (edt_axis_kernel): the same march along x over the plane kernel's output, final clamp, u8 / u16 out.
//
// The march: input row t is added to the 2R+1 open outputs t-R..t+R, one VIADDMNMX.U16x2 (DPX) per (input, output)
// pair and voxel pair, no re-reads.  The open outputs live in registers as a window of W + U - 1 values that is
// addressed statically inside a block of U = 8 steps and shifted by U registers between blocks (6 moves per step).
// Round 1 unrolled a full revolution of the ring instead (2R+1 steps x 2R+1 taps = 50 KB of code; its top stall was
// instruction fetch); a block is 8 x 49 taps = 7 KB and stays in the instruction cache.  A block whose inputs are all
// INF while the window holds nothing but INF (most of a sparse scene) emits INF rows and touches neither the taps nor
// the shift.  Long lines are cut into segments (grid.y) that re-read a halo of R rows on either side: more and
// shorter work items, which is what balances the SM sub-partitions (one warp per 64 voxels of a line is all the
// parallelism a 400^3 pass has: 2500 warps for 592 sub-partitions).
#include <algorithm>
#include <cstdlib>

#include "b200df_internal.cuh"

using namespace b200df;

namespace
{
constexpr unsigned kInfPair = 0x40004000u;  // leaves head-room for two additions of reach^2 <= 900 in 16 bits
constexpr int kBlockSteps = 8;              // U: steps per statically addressed block
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned z_dist2(unsigned lo, unsigned hi, int reach)
{
  // lo = bits [z-32, z) (bit 31 is z-1), hi = bits [z, z+32) (bit 0 is z)
  const int up = hi ? __ffs(hi) - 1 : 64;
  const int dn = lo ? __clz(lo) + 1 : 64;
  const int d = min(up, dn);
  return d <= reach ? (unsigned)(d * d) : 0x4000u;
}

// values >= c (per half) -> 0x4000.  Halves are < 0x8000, so adding 0x8000 - c raises bit 15 exactly for those.
__device__ __forceinline__ unsigned normalise_pair(unsigned v, unsigned bias_pair)
{
  const unsigned hit = ((v + bias_pair) >> 15) & 0x00010001u;
  const unsigned mask = hit * 0xffffu;
  return (v & ~mask) | (kInfPair & mask);
}

// ---- sources: what a marching thread reads per step ------------------------------------------------------------
// All indices are 32-bit element indices from warp-uniform base pointers: base + t * stride compiles to one IMAD and
// one IMAD.WIDE on the FMA pipe, which the marching loop leaves idle (its taps and compares are ALU-pipe work).
struct BitsSource  // plane kernel: the pair's squared z distance, from the bit row of (x, t)
{
  const uint32_t* bits;
  unsigned row0;  // word index of (x, y = 0, word wi + 1)
  unsigned wpr;
  int o, reach;
  bool has0, has2;
  struct Raw
  {
    unsigned a, b, c;
  };
  __device__ __forceinline__ Raw load(int t) const
  {
    const uint32_t* br = bits + (row0 + (unsigned)t * wpr);
    Raw r;
    r.a = has0 ? __ldg(br - 1) : 0u;
    r.b = __ldg(br);
    r.c = has2 ? __ldg(br + 1) : 0u;
    return r;
  }
  static constexpr bool kExactTest = false;  // words in reach of the pair may still hold no obstacle within reach
  static __device__ __forceinline__ Raw none()
  {
    return Raw{ 0u, 0u, 0u };
  }
  __device__ __forceinline__ bool maybe_finite(const Raw& r) const
  {
    return (r.a | r.b | r.c) != 0u;
  }
  __device__ __forceinline__ unsigned convert(const Raw& r) const
  {
    // voxel z0: bits below in lo0 (bit 31 is z0-1), bits from z0 up in hi0 (bit 0 is z0); voxel z0+1 likewise.
    // Two scans instead of four: the distance up is found for z0+1 and carried down, the distance down for z0 and
    // carried up.
    const unsigned lo0 = __funnelshift_r(r.a, r.b, o), hi0 = __funnelshift_r(r.b, r.c, o);
    const unsigned above1 = hi0 >> 1;  // bits [z0+1, z0+32): enough for reach <= 30
    const int up1 = above1 ? __ffs(above1) - 1 : 64;
    const int dn0 = lo0 ? __clz(lo0) + 1 : 64;
    const bool occ0 = hi0 & 1u;
    const int up0 = occ0 ? 0 : up1 + 1;
    const int dn1 = occ0 ? 1 : dn0 + 1;
    const int d0 = min(up0, dn0), d1 = min(up1, dn1);
    const unsigned v0 = d0 <= reach ? (unsigned)(d0 * d0) : 0x4000u;
    const unsigned v1 = d1 <= reach ? (unsigned)(d1 * d1) : 0x4000u;
    return v0 | (v1 << 16);
  }
};

struct PairSource  // axis kernel: u16x2 pairs written by the plane kernel
{
  const uint32_t* in;
  unsigned base, stride;
  typedef unsigned Raw;
  static constexpr bool kExactTest = true;
  __device__ __forceinline__ Raw load(int t) const
  {
    return __ldg(in + (base + (unsigned)t * stride));
  }
  static __device__ __forceinline__ Raw none()
  {
    return kInfPair;
  }
  __device__ __forceinline__ bool maybe_finite(const Raw& r) const
  {
    return r != kInfPair;
  }
  __device__ __forceinline__ unsigned convert(const Raw& r) const
  {
    return r;
  }
};

// ---- sinks: where a finished output row goes -------------------------------------------------------------------
struct PairSink  // intermediate: u16x2, >= max_d2 -> INF (can never win below the clamp later)
{
  uint32_t* out;
  unsigned base, stride;
  unsigned bias_pair;
  bool live;
  __device__ __forceinline__ void emit(int o, unsigned v) const
  {
    if (live)
      out[base + (unsigned)o * stride] = normalise_pair(v, bias_pair);
  }
  __device__ __forceinline__ void emit_inf(int o) const
  {
    if (live)
      out[base + (unsigned)o * stride] = kInfPair;
  }
};

template <typename OutT>
struct FinalSink  // final: clamp to max_d2, narrow to the field's element type
{
  void* out;
  unsigned base, stride;
  unsigned clamp_pair;
  bool live;
  __device__ __forceinline__ void put(int o, unsigned v) const
  {
    if (!live)
      return;
    const unsigned i = base + (unsigned)o * stride;
    if (sizeof(OutT) == 1)
      static_cast<uint16_t*>(out)[i] = (uint16_t)((v & 0xffu) | ((v >> 8) & 0xff00u));
    else
      static_cast<uint32_t*>(out)[i] = v;
  }
  __device__ __forceinline__ void emit(int o, unsigned v) const
  {
    put(o, __vminu2(v, clamp_pair));
  }
  __device__ __forceinline__ void emit_inf(int o) const
  {
    put(o, clamp_pair);
  }
};

// ---- the march ----------------------------------------------------------------------------------------------------
// Outputs [oa, ob) of one line; reads inputs [oa - R, ob + R) clipped to [0, n_axis).  Every range test is on
// warp-uniform values (t0, oa, ob).  Inputs are fetched two blocks (16 rows) ahead: a block that turns out to be all
// INF costs a dozen instructions, so one block of look-ahead does not cover the memory latency.
// Three ways through a block of U input rows:
//   - no finite input and the window holds only INF: emit INF rows (nothing else to do)
//   - no finite input, window not yet drained: emit the finished rows, shift the window
//   - otherwise: all U x (2R+1) taps as one straight-line sequence; the shift by U registers is folded into the
//     destination registers of the last tap of every window entry, so the busy path has no register moves at all
// Step-granular variant for the plane kernel, whose finite rows are few (16 % of a sparse scene's rows) and scattered:
// the taps of a row are skipped unless some lane's input is finite; one block of look-ahead.
template <int R, typename Src, typename Sink>
__device__ __forceinline__ void march_steps(const Src& src, const Sink& sink, int oa, int ob, int n_axis)
{
  constexpr int W = 2 * R + 1, U = kBlockSteps, N = W + U - 1;
  unsigned acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
    acc[i] = kInfPair;
  const int t_end = min(ob + R, n_axis);
  typename Src::Raw raw[U];
  int t0 = max(oa - R, 0);
#pragma unroll
  for (int j = 0; j < U; ++j)
    raw[j] = (t0 + j < t_end) ? src.load(t0 + j) : Src::none();
  int clean = W;  // INF steps since the last finite input; the window holds only INF once it reaches 2R
  for (; t0 < ob + R; t0 += U)
  {
    unsigned f[U];
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < U; ++j)
    {
      f[j] = kInfPair;
      if (__any_sync(kFull, src.maybe_finite(raw[j])))
      {
        f[j] = src.convert(raw[j]);
        if (Src::kExactTest || __any_sync(kFull, f[j] != kInfPair))
          m |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j)  // next block's inputs are in flight while this one is processed
      raw[j] = (t0 + U + j < t_end) ? src.load(t0 + U + j) : Src::none();
    const int o0 = t0 - R;  // output row completed by step 0 of the block
    const bool inside = o0 >= oa && o0 + U <= ob;
    if (m == 0 && clean >= W)
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (inside || (o0 + j >= oa && o0 + j < ob))
          sink.emit_inf(o0 + j);
      continue;
    }
#pragma unroll
    for (int j = 0; j < U; ++j)
    {
      if ((m >> j) & 1u)
      {
#pragma unroll
        for (int k = 0; k < W; ++k)
          acc[j + k] = __viaddmin_u16x2(f[j], (unsigned)((k - R) * (k - R)) * 0x00010001u, acc[j + k]);
        clean = 0;
      }
      else
        ++clean;
      if (inside || (o0 + j >= oa && o0 + j < ob))  // has now seen every input within +-R of it
        sink.emit(o0 + j, acc[j]);
    }
#pragma unroll
    for (int k = 0; k < W - 1; ++k)
      acc[k] = acc[k + U];
#pragma unroll
    for (int k = W - 1; k < N; ++k)
      acc[k] = kInfPair;
  }
}

template <int R, typename Src, typename Sink>
__device__ __forceinline__ void march(const Src& src, const Sink& sink, int oa, int ob, int n_axis)
{
  constexpr int W = 2 * R + 1, U = kBlockSteps, N = W + U - 1;
  unsigned acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
    acc[i] = kInfPair;
  const int t_end = min(ob + R, n_axis);
  typename Src::Raw raw_a[U], raw_b[U];
  int t0 = max(oa - R, 0);  // rows before the array are INF: nothing to add, nothing to emit (their outputs are < oa)
#pragma unroll
  for (int j = 0; j < U; ++j)
  {
    raw_a[j] = (t0 + j < t_end) ? src.load(t0 + j) : Src::none();
    raw_b[j] = (t0 + U + j < t_end) ? src.load(t0 + U + j) : Src::none();
  }
  int drained = 1;  // the window holds only INF
  int quiet = 0;    // INF rows since the last finite one
  for (; t0 < ob + R; t0 += U)
  {
    unsigned f[U];
    bool lane_finite = false;
#pragma unroll
    for (int j = 0; j < U; ++j)
      lane_finite = lane_finite || src.maybe_finite(raw_a[j]);
    bool busy = __any_sync(kFull, lane_finite);
    if (busy)
    {
      lane_finite = false;
#pragma unroll
      for (int j = 0; j < U; ++j)
      {
        f[j] = src.convert(raw_a[j]);
        lane_finite = lane_finite || f[j] != kInfPair;
      }
      busy = Src::kExactTest || __any_sync(kFull, lane_finite);
    }
#pragma unroll
    for (int j = 0; j < U; ++j)  // rows t0 + 2U ... are now in flight behind the block after this one
    {
      raw_a[j] = raw_b[j];
      raw_b[j] = (t0 + 2 * U + j < t_end) ? src.load(t0 + 2 * U + j) : Src::none();
    }
    const int o0 = t0 - R;  // output row completed by step 0 of the block
    const bool inside = o0 >= oa && o0 + U <= ob;
    if (!busy)
    {
      if (drained)
      {
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (inside || (o0 + j >= oa && o0 + j < ob))
            sink.emit_inf(o0 + j);
        continue;
      }
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (inside || (o0 + j >= oa && o0 + j < ob))
          sink.emit(o0 + j, acc[j]);
      quiet += U;
      drained = quiet >= 2 * R;
    }
    else
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
      {
#pragma unroll
        for (int k = 0; k < W; ++k)
          acc[j + k] = __viaddmin_u16x2(f[j], (unsigned)((k - R) * (k - R)) * 0x00010001u, acc[j + k]);
      }
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (inside || (o0 + j >= oa && o0 + j < ob))  // has now seen every input within +-R of it
          sink.emit(o0 + j, acc[j]);
      quiet = 0;
      drained = 0;
    }
#pragma unroll
    for (int k = 0; k < W - 1; ++k)
      acc[k] = acc[k + U];
#pragma unroll
    for (int k = W - 1; k < N; ++k)
      acc[k] = kInfPair;
  }
}

// One warp per CTA: the block scheduler hands a new line group to whichever SM has a slot free, which is what evens
// out heavy (all rows finite) and light (all INF) groups.
// Both kernels work on a slab of voxel pairs [zp0, zp0 + slab) of every z row: the y and x passes of different z are
// independent, so the slabs of one rebuild run as separate launches on separate streams.  Their tails (the last,
// heaviest line groups of a launch leave most SM sub-partitions idle) then overlap with the next slab's work, and a
// slab's intermediate plane-kernel output (32 MB for a quarter of Field B) is still in L2 when its axis kernel reads it.
template <int R>
__global__ void __launch_bounds__(32) edt_plane_kernel(const uint32_t* __restrict__ bits, uint32_t* __restrict__ out,
                                                       unsigned n_lines, unsigned half, unsigned zp0, unsigned slab,
                                                       int ny, unsigned wpr, int reach, int max_d2, int seg_len)
{
  const unsigned line0 = blockIdx.x * 32u + threadIdx.x;
  const bool live = line0 < n_lines;
  const unsigned line = live ? line0 : n_lines - 1;  // surplus lanes shadow the last line (keeps the votes full)
  const unsigned x = line / slab, zp = zp0 + line % slab;
  const int z0 = 2 * (int)zp;
  const int wi = (z0 - 32) >> 5;  // -1 for the first 32 voxels of a row
  BitsSource src;
  src.bits = bits;
  src.row0 = x * (unsigned)ny * wpr + (unsigned)(wi + 1);
  src.wpr = wpr;
  src.o = z0 & 31;
  src.reach = reach;
  src.has0 = wi >= 0;
  src.has2 = wi + 2 < (int)wpr;
  PairSink sink;
  sink.out = out;
  sink.base = x * (unsigned)ny * half + zp;
  sink.stride = half;
  sink.bias_pair = (unsigned)(0x8000 - max_d2) * 0x00010001u;
  sink.live = live;
  const int oa = blockIdx.y * seg_len;
  march_steps<R>(src, sink, oa, min(oa + seg_len, ny), ny);
}

// TILED: a warp covers 4 y rows x 8 voxel pairs (16 z) instead of 32 consecutive pairs of one row - the same four
// 32-byte sectors per access, but a compact footprint touches fewer "finite" rows (needs half % 8 == 0, ny % 4 == 0).
template <int R, typename OutT, bool TILED>
__global__ void __launch_bounds__(32) edt_axis_kernel(const uint32_t* __restrict__ in, void* __restrict__ out,
                                                      unsigned n_lines, unsigned half, unsigned zp0, unsigned slab,
                                                      unsigned plane, int nx, int max_d2, int seg_len)
{
  // n_lines = ny * slab lines of this slab; a line is the pair (y, zp0 + zp), zp < slab
  unsigned y, zp;
  bool live;
  if (TILED)
  {
    const unsigned tiles_z = slab / 8;
    const unsigned ty = blockIdx.x / tiles_z, tz = blockIdx.x % tiles_z;
    y = 4 * ty + (threadIdx.x >> 3);
    zp = 8 * tz + (threadIdx.x & 7);
    live = true;
  }
  else
  {
    const unsigned line0 = blockIdx.x * 32u + threadIdx.x;
    live = line0 < n_lines;
    const unsigned line = live ? line0 : n_lines - 1;
    y = line / slab;
    zp = line % slab;
  }
  const unsigned q = y * half + zp0 + zp;
  PairSource src;
  src.in = in;
  src.base = q;
  src.stride = plane;
  FinalSink<OutT> sink;
  sink.out = out;
  sink.base = q;
  sink.stride = plane;
  sink.clamp_pair = (unsigned)max_d2 * 0x00010001u;
  sink.live = live;
  const int oa = blockIdx.y * seg_len;
  march<R>(src, sink, oa, min(oa + seg_len, nx), nx);
}

// ---- occupancy bytes -> bit rows ------------------------------------------------------------------------------------
__global__ void pack_bits_kernel(const uint8_t* __restrict__ occ, uint32_t* __restrict__ bits, int nz, int wpr,
                                 long n_words, bool invert)
{
  const long gw = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;  // one warp per output word
  if (gw >= n_words)
    return;
  const int lane = threadIdx.x & 31;
  const long row = gw / wpr;
  const int z = 32 * (int)(gw % wpr) + lane;
  const bool o = z < nz && ((occ[row * nz + z] != 0) != invert);
  const unsigned m = __ballot_sync(kFull, o);
  if (lane == 0)
    bits[gw] = m;
}

// nz % 16 == 0: a thread turns 16 occupancy bytes (one 16-byte load; occupancy bytes are 0 or 1) into one half-word
// of the bit row.  Block = (32, 8): threadIdx.y picks the row, threadIdx.x strides over the row's 16-voxel groups.
__global__ void pack_bits16_kernel(const uint4* __restrict__ occ, uint16_t* __restrict__ bits16, int groups, int wpr,
                                   long n_rows, bool invert)
{
  const long row = blockIdx.x * (long)blockDim.y + threadIdx.y;
  if (row >= n_rows)
    return;
  const unsigned flip = invert ? 0x01010101u : 0u;
  for (int h = threadIdx.x; h < groups; h += 32)
  {
    const uint4 v = occ[row * groups + h];
    const unsigned n0 = ((((v.x & 0x01010101u) ^ flip) * 0x01020408u) >> 24) & 0xfu;
    const unsigned n1 = ((((v.y & 0x01010101u) ^ flip) * 0x01020408u) >> 24) & 0xfu;
    const unsigned n2 = ((((v.z & 0x01010101u) ^ flip) * 0x01020408u) >> 24) & 0xfu;
    const unsigned n3 = ((((v.w & 0x01010101u) ^ flip) * 0x01020408u) >> 24) & 0xfu;
    bits16[row * (2 * wpr) + h] = (uint16_t)(n0 | (n1 << 4) | (n2 << 8) | (n3 << 12));
  }
}

inline unsigned blocks_for(long n, int threads)
{
  return (unsigned)((n + threads - 1) / threads);
}

int env_int(const char* name, int fallback)
{
  const char* s = getenv(name);
  return s && *s ? atoi(s) : fallback;
}

// Segments per line (grid.y): a segment re-reads a halo of R rows on either side, so long lines are only cut when the
// grid would otherwise be too small to fill the chip.
int pick_segments(long n_warps, int n_axis, int reach, const char* env)
{
  const int forced = env_int(env, 0);
  const int max_seg = std::max(1, n_axis / std::max(1, 2 * reach));  // a segment is at least as long as its halo
  if (forced > 0)
    return std::min(forced, max_seg);
  const long want = 148L * 32;  // ~2 waves of the 12-16 one-warp CTAs an SM holds (measured: Field B 0.259 -> 0.231 ms)
  return (int)std::min<long>(std::min(max_seg, 4), std::max<long>(1, (want + n_warps - 1) / n_warps));
}

template <int R, typename T>
int run_march(b200df_field* f, T* out)
{
  const int nz = f->grid.n[2], ny = f->grid.n[1], nx = f->grid.n[0];
  const int half = nz / 2, wpr = (nz + 31) / 32, reach = f->radius - 1;
  // z slabs (multiples of 8 pairs, so that the tiled footprint and the 32-byte sectors stay whole)
  int n_slabs = env_int("B200DF_EDT_SLABS", 0);
  if (n_slabs <= 0)
    n_slabs = f->n_voxels >= (8L << 20) ? 4 : 1;
  n_slabs = std::max(1, std::min(n_slabs, std::min(half / 8, (int)b200df_field::kEdtStreams)));
  const bool tiled_ok = half % 8 == 0 && ny % 4 == 0 && !env_int("B200DF_EDT_LINEAR", 0);
  static bool attr_done = false;
  if (!attr_done)
  {
    cudaFuncSetAttribute(edt_plane_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(edt_axis_kernel<R, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(edt_axis_kernel<R, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  // dynamic shared memory is only an occupancy knob here (experiments)
  const int smem_y = env_int("B200DF_EDT_SMEM_Y", 0), smem_x = env_int("B200DF_EDT_SMEM_X", 0);
  if (n_slabs > 1)
    B200DF_CUDA(cudaEventRecord(f->ev_fork, f->stream));
  int zp0 = 0;
  for (int s = 0; s < n_slabs; ++s)
  {
    int slab = n_slabs == 1 ? half : ((half / n_slabs + 7) / 8) * 8;
    if (s == n_slabs - 1 || zp0 + slab > half)
      slab = half - zp0;
    if (slab <= 0)
      break;
    cudaStream_t st = f->stream;
    if (n_slabs > 1)
    {
      st = f->edt_streams[s];
      B200DF_CUDA(cudaStreamWaitEvent(st, f->ev_fork, 0));
    }
    const long lines_y = (long)nx * slab, lines_x = (long)ny * slab;
    const long warps_y = (lines_y + 31) / 32, warps_x = (lines_x + 31) / 32;
    const int seg_y = pick_segments(warps_y * n_slabs, ny, reach, "B200DF_EDT_SEG_Y");
    const int seg_x = pick_segments(warps_x * n_slabs, nx, reach, "B200DF_EDT_SEG_X");
    const int len_y = (ny + seg_y - 1) / seg_y, len_x = (nx + seg_x - 1) / seg_x;
    edt_plane_kernel<R><<<dim3((unsigned)warps_y, (unsigned)((ny + len_y - 1) / len_y)), 32, smem_y, st>>>(
        f->bits, reinterpret_cast<uint32_t*>(f->tmp_a), (unsigned)lines_y, (unsigned)half, (unsigned)zp0,
        (unsigned)slab, ny, (unsigned)wpr, reach, f->max_d2, len_y);
    B200DF_LAUNCHED();
    const dim3 grid_x((unsigned)warps_x, (unsigned)((nx + len_x - 1) / len_x));
    const unsigned plane = (unsigned)ny * (unsigned)half;
    if (tiled_ok && slab % 8 == 0)
      edt_axis_kernel<R, T, true><<<grid_x, 32, smem_x, st>>>(reinterpret_cast<const uint32_t*>(f->tmp_a), out,
                                                              (unsigned)lines_x, (unsigned)half, (unsigned)zp0,
                                                              (unsigned)slab, plane, nx, f->max_d2, len_x);
    else
      edt_axis_kernel<R, T, false><<<grid_x, 32, smem_x, st>>>(reinterpret_cast<const uint32_t*>(f->tmp_a), out,
                                                               (unsigned)lines_x, (unsigned)half, (unsigned)zp0,
                                                               (unsigned)slab, plane, nx, f->max_d2, len_x);
    B200DF_LAUNCHED();
    if (n_slabs > 1)
    {
      B200DF_CUDA(cudaEventRecord(f->ev_join[s], st));
      B200DF_CUDA(cudaStreamWaitEvent(f->stream, f->ev_join[s], 0));
    }
    zp0 += slab;
  }
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}
}  // namespace

namespace b200df
{
bool edt_fast_applies(const b200df_field* f)
{
  return f->bits && f->grid.n[2] % 2 == 0 && f->radius >= 1 && f->radius <= 31 && f->n_voxels > 0;
}

// occupancy (or its complement) -> out (u8 when elem_size == 1, else u16); stream-ordered on f->stream
int edt_fast(b200df_field* f, bool invert, void* out)
{
  const int nz = f->grid.n[2], ny = f->grid.n[1], nx = f->grid.n[0];
  const int wpr = (nz + 31) / 32;
  const long n_rows = (long)nx * ny;
  if (!invert && f->bits_valid)
  {
    // the bit rows were kept current by reset / add_points: nothing to pack
  }
  else if (nz % 16 == 0)
  {
    pack_bits16_kernel<<<blocks_for(n_rows, 8), dim3(32, 8), 0, f->stream>>>(
        reinterpret_cast<const uint4*>(f->occ), reinterpret_cast<uint16_t*>(f->bits), nz / 16, wpr, n_rows, invert);
    B200DF_LAUNCHED();
  }
  else
  {
    const long n_words = n_rows * wpr;
    pack_bits_kernel<<<blocks_for(n_words * 32, 256), 256, 0, f->stream>>>(f->occ, f->bits, nz, wpr, n_words, invert);
    B200DF_LAUNCHED();
  }
  f->bits_valid = !invert;  // the bit rows now mirror occ (or its complement, for the negative field)
  const int reach = f->radius - 1;
  if (f->elem_size == 1)
  {
    uint8_t* o = static_cast<uint8_t*>(out);
    if (reach <= 12)
      return run_march<12, uint8_t>(f, o);
    return run_march<24, uint8_t>(f, o);  // max_d2 <= 255 means radius <= 16
  }
  uint16_t* o = static_cast<uint16_t*>(out);
  if (reach <= 12)
    return run_march<12, uint16_t>(f, o);
  if (reach <= 24)
    return run_march<24, uint16_t>(f, o);
  return run_march<30, uint16_t>(f, o);
}
}  // namespace b200df
