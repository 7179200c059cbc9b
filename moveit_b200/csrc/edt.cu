// Kernel 5: exact clamped squared EDT from bit-packed occupancy - the bit-row packing kernels, the dispatch, and the
// register-prefetch form of the two marching passes for grids the TMA form (edt_tma.cu) cannot take (nz not a multiple
// of 16: tensor-map strides must be multiples of 16 bytes).
//
// Replaces PropagationDistanceField::propagatePositive / propagateNegative (reference:
// moveit_core/distance_field/src/propagation_distance_field.cpp:389-505) for grids with an even nz and a clamp radius
// <= 31 cells; layout from voxel_grid.h:426-429 (z fastest).  d2(v) = min over obstacle voxels q of |v-q|^2, clamped
// to max_d2.  Any voxel whose true distance is below max_d2 has every axis offset <= reach = radius-1, so windows of
// +-reach per axis are exact after the final clamp (a tap at the radius itself adds >= max_d2).
//
//   plane kernel: a thread owns the voxel pair (x, z0, z0+1) and marches along y.  The input of a step is the pair's
//       squared z distance to the nearest obstacle of row (x, y), computed on the fly from the row's bit words (two bit
//       scans).  Output: the 2-D (y, z) squared distance as u16x2, values >= max_d2 normalised to INF.
//   axis kernel: the same march along x over the plane kernel's output, final clamp, u8 / u16 out.
//
// The march: input row t is added to the 2R+1 open outputs t-R..t+R.  The open outputs live in registers as a window of
// 2R + U values that is addressed statically inside a block of U steps and shifted by U registers between blocks.  A
// block whose inputs are all INF while the window holds nothing but INF (most of a sparse scene) emits INF rows and
// touches neither the taps nor the shift.  Long lines are cut into segments (grid.y) that re-read a halo of R rows on
// either side: more and shorter work items, which is what balances the SM sub-partitions.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "b200df_internal.cuh"

using namespace b200df;

namespace
{
constexpr unsigned kInfPair = 0x40004000u;  // leaves head-room for two additions of reach^2 <= 900 in 16 bits
constexpr unsigned kFull = 0xffffffffu;

// values >= c (per half) -> 0x4000.  Halves are < 0x8000, so adding 0x8000 - c raises bit 15 exactly for those.
__device__ __forceinline__ unsigned normalise_pair(unsigned v, unsigned bias_pair)
{
  const unsigned hit = ((v + bias_pair) >> 15) & 0x00010001u;
  const unsigned mask = hit * 0xffffu;
  return (v & ~mask) | (kInfPair & mask);
}

struct BitWords  // the three words of a bit row a voxel pair can reach
{
  unsigned a, b, c;
};

// ---- the march, second form: taps in pairs on both integer pipes ---------------------------------------------------
// VIADDMNMX runs on the ALU pipe only, whose issue rate is one warp instruction per two cycles and scheduler, and a
// march that is nothing but taps is bound by exactly that (ncu: ALU pipe 57 %, issue slots 55 %, math-pipe throttle the
// top stall after the fixed-latency wait).  Here the additions go to the otherwise idle FMA pipe and the minima take
// two candidates at once:
//     s_a(k) = f_a * one + k^2      IMAD  (one == 1 is a kernel argument so that ptxas cannot fold it into an IADD3)
//     acc[o] = min3(acc[o], s_a(|o - t_a|), s_b(|o - t_b|))      VIMNMX3.U16x2
// Two adjacent input rows (t, t + 1) are swept together from their centre outwards, so that every sum is used for the
// offsets +k and -k right after it was formed (three sums of the second row live at any time): per input row 24 IMAD
// + 25 VIMNMX3 instead of 49 VIADDMNMX - half the ALU-pipe cycles, the same number of issue slots.
__device__ __forceinline__ unsigned add_k2(unsigned f, unsigned one, unsigned k2pair)
{
  unsigned r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(f), "r"(one), "r"(k2pair));
  return r;
}

template <int R, int N, int J>
__device__ __forceinline__ void tap_pair(unsigned (&acc)[N], unsigned fa, unsigned fb, unsigned one)
{
  constexpr int c = J + R;  // window entry of row a's own position; row b's is c + 1
  unsigned b_prev = fb;
  unsigned b_cur = add_k2(fb, one, 0x00010001u);
  acc[c] = __vimin3_u16x2(acc[c], fa, b_cur);
#pragma unroll
  for (int k = 1; k <= R; ++k)
  {
    const unsigned a = add_k2(fa, one, (unsigned)(k * k) * 0x00010001u);
    if (k < R)
    {
      const unsigned b_next = add_k2(fb, one, (unsigned)((k + 1) * (k + 1)) * 0x00010001u);
      acc[c - k] = __vimin3_u16x2(acc[c - k], a, b_next);  // offset -(k + 1) from row b
      acc[c + k] = __vimin3_u16x2(acc[c + k], a, b_prev);  // offset k - 1 from row b
      b_prev = b_cur;
      b_cur = b_next;
    }
    else
    {
      acc[c - k] = __vminu2(acc[c - k], a);
      acc[c + k] = __vimin3_u16x2(acc[c + k], a, b_prev);
    }
  }
  acc[c + R + 1] = __vminu2(acc[c + R + 1], b_cur);
}

template <int R, int U, int N, int P>
struct TapPairs  // pairs P, P + 1, ... of a block, each behind its own warp-uniform test
{
  static __device__ __forceinline__ void run(unsigned (&acc)[N], const unsigned (&f)[U], unsigned busy, unsigned one)
  {
    if ((busy >> P) & 1u)
      tap_pair<R, N, 2 * P>(acc, f[2 * P], f[2 * P + 1], one);
    TapPairs<R, U, N, P + 1>::run(acc, f, busy, one);
  }
};
template <int R, int U, int N>
struct TapPairs<R, U, N, U / 2>
{
  static __device__ __forceinline__ void run(unsigned (&)[N], const unsigned (&)[U], unsigned, unsigned)
  {
  }
};

// Outputs [oa, ob) of one line from inputs [oa - R, ob + R) clipped to [0, n_axis).  A block is U input rows; its
// inputs were fetched one (plane kernel) or two (axis kernel) blocks earlier.  Everything that decides the path through
// a block is warp-uniform: which pairs of rows hold a finite value in some lane (one REDUX.OR of per-lane pair masks),
// whether the window holds anything but INF, whether all U outputs / all U prefetched rows lie inside the array.
template <int R, int U, int AHEAD, typename Src, typename Sink>
__device__ __forceinline__ void march2(Src& src, Sink& sink, int oa, int ob, int n_axis, unsigned one)
{
  constexpr int W = 2 * R + 1, N = W + U - 1;
  static_assert(U % 2 == 0 && (AHEAD == 1 || AHEAD == 2), "");
  unsigned acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
    acc[i] = kInfPair;
  const int t_end = min(ob + R, n_axis);
  int t0 = max(oa - R, 0);
  typename Src::Raw raw[AHEAD][U];
  src.seek(t0);
#pragma unroll
  for (int a = 0; a < AHEAD; ++a)
  {
#pragma unroll
    for (int j = 0; j < U; ++j)
      raw[a][j] = (t0 + a * U + j < t_end) ? src.load(j) : Src::none();
    src.advance(U);
  }
  sink.seek(t0 - R);
  int quiet = 2 * R;  // INF rows since the last finite one; the window holds only INF from 2R on
  for (; t0 < ob + R; t0 += U)
  {
    unsigned f[U];
    unsigned lane_busy = 0;
#pragma unroll
    for (int j = 0; j < U; j += 2)
      if (src.maybe_finite(raw[0][j]) || src.maybe_finite(raw[0][j + 1]))
        lane_busy |= 1u << (j / 2);
    const unsigned busy = __reduce_or_sync(kFull, lane_busy);
    if (busy)
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        f[j] = src.convert(raw[0][j]);
    }
    // rows t0 + AHEAD * U ... go in flight
    const int tn = t0 + AHEAD * U;
#pragma unroll
    for (int a = 0; a + 1 < AHEAD; ++a)
#pragma unroll
      for (int j = 0; j < U; ++j)
        raw[a][j] = raw[a + 1][j];
    if (tn + U <= t_end)
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        raw[AHEAD - 1][j] = src.load(j);
    }
    else
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        raw[AHEAD - 1][j] = (tn + j < t_end) ? src.load(j) : Src::none();
    }
    src.advance(U);
    const int o0 = t0 - R;  // output row completed by step 0 of the block
    const bool inside = o0 >= oa && o0 + U <= ob;
    const bool some = o0 + U > oa && o0 < ob;
    if (!busy && quiet >= 2 * R)
    {
      if (inside)
      {
#pragma unroll
        for (int j = 0; j < U; ++j)
          sink.emit_inf(j);
      }
      else if (some)
      {
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (o0 + j >= oa && o0 + j < ob)
            sink.emit_inf(j);
      }
      sink.advance(U);
      continue;
    }
    if (busy)
    {
      TapPairs<R, U, N, 0>::run(acc, f, busy, one);
      quiet = 0;
    }
    else
      quiet += U;
    if (inside)
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        sink.emit(j, acc[j]);
    }
    else if (some)
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (o0 + j >= oa && o0 + j < ob)
          sink.emit(j, acc[j]);
    }
    sink.advance(U);
#pragma unroll
    for (int k = 0; k < W - 1; ++k)
      acc[k] = acc[k + U];
#pragma unroll
    for (int k = W - 1; k < N; ++k)
      acc[k] = kInfPair;
  }
}

// sources / sinks of march2: a moving row pointer, rows addressed as pointer + j * stride inside a block
struct BitsSource2
{
  const uint32_t* p;  // word wi + 1 of the row of the current block's first step
  unsigned wpr;
  int o, reach;
  bool has0, has2;
  typedef BitWords Raw;
  __device__ __forceinline__ void seek(int t)
  {
    p += (size_t)t * wpr;
  }
  __device__ __forceinline__ void advance(int n)
  {
    p += (size_t)n * wpr;
  }
  __device__ __forceinline__ Raw load(int j) const
  {
    const uint32_t* br = p + (unsigned)j * wpr;
    Raw r;
    r.a = has0 ? __ldg(br - 1) : 0u;
    r.b = __ldg(br);
    r.c = has2 ? __ldg(br + 1) : 0u;
    return r;
  }
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
    const unsigned lo0 = __funnelshift_r(r.a, r.b, o), hi0 = __funnelshift_r(r.b, r.c, o);
    const unsigned above1 = hi0 >> 1;
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

struct PairSource2
{
  const uint32_t* p;
  unsigned stride;
  typedef unsigned Raw;
  __device__ __forceinline__ void seek(int t)
  {
    p += (size_t)t * stride;
  }
  __device__ __forceinline__ void advance(int n)
  {
    p += (size_t)n * stride;
  }
  __device__ __forceinline__ Raw load(int j) const
  {
    return __ldg(p + (unsigned)j * stride);
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

struct PairSink2
{
  uint32_t* p;  // may point before the array while the first outputs are still outside [oa, ob): never dereferenced then
  unsigned stride;
  unsigned bias_pair;
  bool live;
  __device__ __forceinline__ void seek(int o)
  {
    p += (ptrdiff_t)o * (ptrdiff_t)stride;
  }
  __device__ __forceinline__ void advance(int n)
  {
    p += (size_t)n * stride;
  }
  __device__ __forceinline__ void emit(int j, unsigned v) const
  {
    if (live)
      p[(unsigned)j * stride] = normalise_pair(v, bias_pair);
  }
  __device__ __forceinline__ void emit_inf(int j) const
  {
    if (live)
      p[(unsigned)j * stride] = kInfPair;
  }
};

template <typename OutT>
struct FinalSink2
{
  typedef typename std::conditional<sizeof(OutT) == 1, uint16_t, uint32_t>::type Store;
  Store* p;
  unsigned stride;
  unsigned clamp_pair;
  bool live;
  __device__ __forceinline__ void seek(int o)
  {
    p += (ptrdiff_t)o * (ptrdiff_t)stride;
  }
  __device__ __forceinline__ void advance(int n)
  {
    p += (size_t)n * stride;
  }
  __device__ __forceinline__ void put(int j, unsigned v) const
  {
    if (!live)
      return;
    if (sizeof(OutT) == 1)
      p[(unsigned)j * stride] = (Store)((v & 0xffu) | ((v >> 8) & 0xff00u));
    else
      p[(unsigned)j * stride] = (Store)v;
  }
  __device__ __forceinline__ void emit(int j, unsigned v) const
  {
    put(j, __vminu2(v, clamp_pair));
  }
  __device__ __forceinline__ void emit_inf(int j) const
  {
    put(j, clamp_pair);
  }
};

// One warp per CTA: the block scheduler hands a new line group to whichever SM has a slot free, which is what evens
// out heavy (all rows finite) and light (all INF) groups.
// Both kernels work on a slab of voxel pairs [zp0, zp0 + slab) of every z row: the y and x passes of different z are
// independent, so the slabs of one rebuild run as separate launches on separate streams.  Their tails (the last,
// heaviest line groups of a launch leave most SM sub-partitions idle) then overlap with the next slab's work, and a
// slab's intermediate plane-kernel output (32 MB for a quarter of Field B) is still in L2 when its axis kernel reads it.
template <int R, int U, int AHEAD>
__global__ void __launch_bounds__(32) edt_plane2_kernel(const uint32_t* __restrict__ bits, uint32_t* __restrict__ out,
                                                        unsigned n_lines, unsigned half, unsigned zp0, unsigned slab,
                                                        int ny, unsigned wpr, int reach, int max_d2, int seg_len,
                                                        unsigned one)
{
  const unsigned line0 = blockIdx.x * 32u + threadIdx.x;
  const bool live = line0 < n_lines;
  const unsigned line = live ? line0 : n_lines - 1;  // surplus lanes shadow the last line (keeps the votes full)
  const unsigned x = line / slab, zp = zp0 + line % slab;
  const int z0 = 2 * (int)zp;
  const int wi = (z0 - 32) >> 5;  // -1 for the first 32 voxels of a row
  BitsSource2 src;
  src.p = bits + ((size_t)x * (unsigned)ny * wpr + (unsigned)(wi + 1));
  src.wpr = wpr;
  src.o = z0 & 31;
  src.reach = reach;
  src.has0 = wi >= 0;
  src.has2 = wi + 2 < (int)wpr;
  PairSink2 sink;
  sink.p = out + ((size_t)x * (unsigned)ny * half + zp);
  sink.stride = half;
  sink.bias_pair = (unsigned)(0x8000 - max_d2) * 0x00010001u;
  sink.live = live;
  const int oa = blockIdx.y * seg_len;
  march2<R, U, AHEAD>(src, sink, oa, min(oa + seg_len, ny), ny, one);
}

template <int R, int U, int AHEAD, typename OutT, bool TILED>
__global__ void __launch_bounds__(32) edt_axis2_kernel(const uint32_t* __restrict__ in, void* __restrict__ out,
                                                       unsigned n_lines, unsigned half, unsigned zp0, unsigned slab,
                                                       unsigned plane, int nx, int max_d2, int seg_len, unsigned one)
{
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
  PairSource2 src;
  src.p = in + q;
  src.stride = plane;
  FinalSink2<OutT> sink;
  sink.p = static_cast<typename FinalSink2<OutT>::Store*>(out) + q;
  sink.stride = plane;
  sink.clamp_pair = (unsigned)max_d2 * 0x00010001u;
  sink.live = live;
  const int oa = blockIdx.y * seg_len;
  march2<R, U, AHEAD>(src, sink, oa, min(oa + seg_len, nx), nx, one);
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
  const int half = nz / 2, wpr = bits_wpr(nz), reach = f->radius - 1;
  // z slabs (multiples of 8 pairs, so that the tiled footprint and the 32-byte sectors stay whole)
  int n_slabs = env_int("B200DF_EDT_SLABS", 0);
  if (n_slabs <= 0)
    n_slabs = f->n_voxels >= (8L << 20) ? 4 : 1;
  n_slabs = std::max(1, std::min(n_slabs, std::min(half / 8, (int)b200df_field::kEdtStreams)));
  const bool tiled_ok = half % 8 == 0 && ny % 4 == 0 && !env_int("B200DF_EDT_LINEAR", 0);
  const int smem_y = 0, smem_x = 0;
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
    const dim3 grid_y((unsigned)warps_y, (unsigned)((ny + len_y - 1) / len_y));
    const dim3 grid_x((unsigned)warps_x, (unsigned)((nx + len_x - 1) / len_x));
    const unsigned plane = (unsigned)ny * (unsigned)half;
    uint32_t* mid = reinterpret_cast<uint32_t*>(f->tmp_a);
    const bool tiled = tiled_ok && slab % 8 == 0;
    {
#define B200DF_PLANE2(U, A)                                                                                         \
  edt_plane2_kernel<R, U, A><<<grid_y, 32, smem_y, st>>>(f->bits, mid, (unsigned)lines_y, (unsigned)half,           \
                                                         (unsigned)zp0, (unsigned)slab, ny, (unsigned)wpr, reach,   \
                                                         f->max_d2, len_y, 1u)
#define B200DF_AXIS2(U, A, TL)                                                                                      \
  edt_axis2_kernel<R, U, A, T, TL><<<grid_x, 32, smem_x, st>>>(mid, out, (unsigned)lines_x, (unsigned)half,         \
                                                               (unsigned)zp0, (unsigned)slab, plane, nx, f->max_d2, \
                                                               len_x, 1u)
      B200DF_PLANE2(8, 1);
      B200DF_LAUNCHED();
      if (tiled)
        B200DF_AXIS2(16, 1, true);
      else
        B200DF_AXIS2(16, 1, false);
      B200DF_LAUNCHED();
#undef B200DF_PLANE2
#undef B200DF_AXIS2
    }
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
  const int wpr = bits_wpr(nz);
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
  if (edt_tma_applies(f))
  {
    const int rc = edt_tma(f, out);
    if (rc != -1)
      return rc;
  }
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
