// Kernel 5, TMA form: exact clamped squared EDT from bit-packed occupancy, tiles moved by the tensor-memory accelerator.
//
// Replaces PropagationDistanceField::propagatePositive / propagateNegative (reference:
// moveit_core/distance_field/src/propagation_distance_field.cpp:389-505); layout from voxel_grid.h:426-429 (z fastest).
// Same two marching passes as edt.cu (plane kernel: z distances from the bit rows + march along y; axis kernel: march
// along x), restructured around three observations from the ncu profiles of the register-prefetch version
// (profiles/r02_edt_two_launch_full.txt: ALU pipe 57 %, issue 55 %, 2.5 warps per scheduler):
//
//  1. VIADDMNMX only runs on the ALU pipe (one warp instruction per two cycles and scheduler).  The taps are therefore
//     formed in pairs: the additions go to the FMA pipe as IMADs, and one VIMNMX3 folds two candidates into a window
//     entry.  Two adjacent input rows are swept together from their centre outwards, so every sum f - k^2 serves the
//     offsets +k and -k right after it was formed: per input row 24 IMAD + 25 VIMNMX3 instead of 49 VIADDMNMX.
//  2. The march works on w = BIAS + max(C - d^2, 0) (C = max_d2) and MAXIMISES: "nothing within the clamp radius" is the
//     smallest value, so the window starts at BIAS, anything at or beyond the clamp loses against it on its own (no
//     normalising pass over the emitted rows), rows that TMA zero-fills outside the array can never win, and the final
//     d^2 = C + BIAS - w is already clamped.  All halves stay inside (0, 0x8000): packed 32-bit IMADs never borrow.
//  3. Tiles instead of per-thread rows: a warp owns 4 lines x 8 voxel pairs; U rows of them are one TMA box
//     (cp.async.bulk.tensor.3d, 128 bytes per row) that lands in shared memory S blocks ahead of its use, and the U
//     finished rows leave as one TMA store.  A lane's element of row j sits at tile + 128 j + 4 lane: every LDS / STS
//     has an immediate offset, no address arithmetic, no bounds tests (TMA clips), no prefetch registers.  Blocks in
//     which nothing is within reach are written from a constant tile without touching the window.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "b200df_internal.cuh"

using namespace b200df;

namespace
{
#ifndef B200DF_EDT_S
#define B200DF_EDT_S 2
#endif
#ifndef B200DF_EDT_MINB
#define B200DF_EDT_MINB 16  // CTAs (= warps) per SM the plane kernel is compiled for: caps it at 128 registers
#endif
constexpr unsigned kBias = 0x2000u;
constexpr unsigned kBiasPair = 0x20002000u;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "B200DF_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra B200DF_DONE;\n"
      "bra B200DF_WAIT;\n"
      "B200DF_DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                   "r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src)
{
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned lds32(uint32_t addr)
{
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, unsigned v)
{
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, unsigned v)
{
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

// f * one - k^2 on both halves: an IMAD (one == 1 is a kernel argument, so ptxas cannot turn it into an ALU-pipe IADD3)
__device__ __forceinline__ unsigned sub_k2(unsigned f, unsigned one, unsigned k2)
{
  unsigned r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(f), "r"(one), "r"(0u - k2 * 0x00010001u));
  return r;
}

// Rows of lead-in before the first output block: at least R (an output is complete once the input R rows further on
// has been applied) and such that twice the lead is a multiple of the block length, which puts the first emitted
// block exactly on the segment start.
__host__ __device__ constexpr int lead_rows(int R, int U)
{
  int l = R;
  while ((2 * l) % U)
    ++l;
  return l;
}

// rows a (window entry J + LEAD) and b (J + LEAD + 1) swept together from the centre outwards
template <int R, int LEAD, int N, int J>
__device__ __forceinline__ void tap_pair(unsigned (&acc)[N], unsigned fa, unsigned fb, unsigned one)
{
  constexpr int c = J + LEAD;
  unsigned b_prev = fb;
  unsigned b_cur = sub_k2(fb, one, 1u);
  acc[c] = __vimax3_s16x2(acc[c], fa, b_cur);
#pragma unroll
  for (int k = 1; k <= R; ++k)
  {
    const unsigned a = sub_k2(fa, one, (unsigned)(k * k));
    if (k < R)
    {
      const unsigned b_next = sub_k2(fb, one, (unsigned)((k + 1) * (k + 1)));
      acc[c - k] = __vimax3_s16x2(acc[c - k], a, b_next);  // offset -(k + 1) from row b
      acc[c + k] = __vimax3_s16x2(acc[c + k], a, b_prev);  // offset k - 1 from row b
      b_prev = b_cur;
      b_cur = b_next;
    }
    else
    {
      acc[c - k] = __vmaxs2(acc[c - k], a);
      acc[c + k] = __vimax3_s16x2(acc[c + k], a, b_prev);
    }
  }
  acc[c + R + 1] = __vmaxs2(acc[c + R + 1], b_cur);
}

template <int R, int U, int LEAD, int N, int P>
struct TapPairs
{
  static __device__ __forceinline__ void run(unsigned (&acc)[N], const unsigned (&f)[U], unsigned busy, unsigned one)
  {
    if ((busy >> P) & 1u)
      tap_pair<R, LEAD, N, 2 * P>(acc, f[2 * P], f[2 * P + 1], one);
    TapPairs<R, U, LEAD, N, P + 1>::run(acc, f, busy, one);
  }
};
template <int R, int U, int LEAD, int N>
struct TapPairs<R, U, LEAD, N, U / 2>
{
  static __device__ __forceinline__ void run(unsigned (&)[N], const unsigned (&)[U], unsigned, unsigned)
  {
  }
};

// ---- the march over one tile column ------------------------------------------------------------------------------------
// Outputs [oa, ob) of the warp's 32 lines; oa is a multiple of U, ob is a multiple of U or the end of the axis (TMA
// clips).  Blocks are aligned: block i reads input rows [t0, t0 + U) with t0 = oa - LEAD + i U and completes output rows
// [t0 - LEAD, t0 - LEAD + U); LEAD = lead_rows(R, U).
template <int R, int U, typename Src, typename Sink>
__device__ __forceinline__ void march(Src& src, Sink& sink, int oa, int ob, int n_axis, unsigned one)
{
  constexpr int LEAD = lead_rows(R, U), N = LEAD + R + U;
  unsigned acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
    acc[i] = kBiasPair;
  int t0 = oa - LEAD;
  const int t_stop = ob + LEAD;  // first block that is not needed
  src.begin(t0, t_stop);
  constexpr int kDrain = LEAD + R;  // idle rows after which the window holds nothing but BIAS
  int quiet = kDrain;
  for (; t0 < t_stop; t0 += U)
  {
    src.wait();
    unsigned f[U];
    unsigned busy = 0;
#pragma unroll
    for (int j = 0; j < U; ++j)
      f[j] = src.get(j);
#pragma unroll
    for (int p = 0; p < U / 2; ++p)
      if (__any_sync(kFull, Src::in_reach(f[2 * p]) || Src::in_reach(f[2 * p + 1])))
        busy |= 1u << p;
    // rows outside the array hold no obstacle whatever the tile says (TMA zero-fills them)
    if (t0 < 0 || t0 + U > n_axis)
    {
#pragma unroll
      for (int p = 0; p < U / 2; ++p)
        if (t0 + 2 * p + 1 < 0 || t0 + 2 * p >= n_axis)
          busy &= ~(1u << p);
    }
    if (busy)
    {
#pragma unroll
      for (int j = 0; j < U; ++j)
        f[j] = src.convert(f[j], t0 + j >= 0 && t0 + j < n_axis);
    }
    src.release(t0, t_stop);  // every lane holds its rows in registers: the slot takes the block S steps ahead
    const int o0 = t0 - LEAD;
    const bool emit = o0 >= oa;
    if (!busy && quiet >= kDrain)
    {
      if (emit)
        sink.flush_inf(o0);
      continue;
    }
    if (busy)
    {
      TapPairs<R, U, LEAD, N, 0>::run(acc, f, busy, one);
      quiet = 0;
    }
    else
      quiet += U;
    if (emit)
    {
      sink.acquire();
#pragma unroll
      for (int j = 0; j < U; ++j)
        sink.put(j, acc[j]);
      sink.flush(o0);
    }
#pragma unroll
    for (int k = 0; k < N - U; ++k)
      acc[k] = acc[k + U];
#pragma unroll
    for (int k = N - U; k < N; ++k)
      acc[k] = kBiasPair;
  }
  sink.finish();
}

// ---- where finished rows go: an output tile in shared memory, one TMA store per block ----------------------------------
template <int U, int LANE_BYTES>
struct TileSink
{
  static constexpr int kRow = 32 * LANE_BYTES, kTile = U * kRow;
  const CUtensorMap* map;
  uint32_t tiles;     // two output tiles, then the constant tile
  uint32_t lane_off;  // lane * LANE_BYTES
  int c0, c1;         // tile coordinates in the two non-marching dimensions
  int cur = 0;
  int lane;
  __device__ __forceinline__ void init(unsigned inf_value)
  {
#pragma unroll
    for (int j = 0; j < U; ++j)
    {
      if (LANE_BYTES == 4)
        sts32(tiles + 2 * kTile + j * kRow + lane_off, inf_value);
      else
        sts16(tiles + 2 * kTile + j * kRow + lane_off, inf_value);
    }
    fence_async_smem();
    __syncwarp();
  }
  __device__ __forceinline__ void acquire()
  {
    // the store issued from this tile two blocks ago has read it
    if (lane == 0)
      tma_store_wait_read<1>();
    __syncwarp();
  }
  __device__ __forceinline__ void store(int j, unsigned v)
  {
    if (LANE_BYTES == 4)
      sts32(tiles + cur * kTile + j * kRow + lane_off, v);
    else
      sts16(tiles + cur * kTile + j * kRow + lane_off, v);
  }
  __device__ __forceinline__ void flush(int o0)
  {
    fence_async_smem();
    __syncwarp();
    if (lane == 0)
      tma_store_3d(map, c0, c1, o0, tiles + cur * kTile);
    cur ^= 1;
  }
  __device__ __forceinline__ void flush_inf(int o0)
  {
    if (lane == 0)
      tma_store_3d(map, c0, c1, o0, tiles + 2 * kTile);
  }
  __device__ __forceinline__ void finish()
  {
    if (lane == 0)
      tma_store_wait_read<0>();
    __syncwarp();
  }
};

template <int U>
struct MidSink : TileSink<U, 4>  // plane kernel: the window entries as they are
{
  __device__ __forceinline__ void put(int j, unsigned w)
  {
    this->store(j, w);
  }
};

template <int U, typename OutT>
struct FinalSink : TileSink<U, 2 * sizeof(OutT)>  // axis kernel: d^2 = C + BIAS - w, narrowed to the field's element type
{
  unsigned out_const;  // (C + BIAS) on both halves
  __device__ __forceinline__ void put(int j, unsigned w)
  {
    const unsigned v = out_const - w;
    if (sizeof(OutT) == 1)
      this->store(j, __byte_perm(v, 0u, 0x4420));  // bytes 0 and 2 -> one 16-bit pair
    else
      this->store(j, v);
  }
};

// ---- where input rows come from ------------------------------------------------------------------------------------------
// axis kernel: tiles of the plane kernel's output, S blocks in flight
template <int U, int S>
struct TileSource
{
  static constexpr int kTile = U * 128;
  const CUtensorMap* map;
  uint32_t tiles, bars, lane_off;
  int c0, c1, lane;
  int stage = 0;
  unsigned phase = 0;
  __device__ __forceinline__ void issue(int s, int t)
  {
    mbar_expect_tx(bars + 8 * s, kTile);
    tma_load_3d(tiles + s * kTile, map, c0, c1, t, bars + 8 * s);
  }
  __device__ __forceinline__ void begin(int t0, int t_stop)
  {
    if (lane == 0)
    {
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (t0 + s * U < t_stop)
          issue(s, t0 + s * U);
    }
  }
  __device__ __forceinline__ void wait()
  {
    mbar_wait(bars + 8 * stage, phase);
  }
  __device__ __forceinline__ unsigned get(int j) const
  {
    return lds32(tiles + stage * kTile + j * 128 + lane_off);
  }
  static __device__ __forceinline__ bool in_reach(unsigned w)
  {
    return w != kBiasPair;
  }
  __device__ __forceinline__ unsigned convert(unsigned w, bool) const
  {
    return w;
  }
  __device__ __forceinline__ void release(int t0, int t_stop)
  {
    // the votes of the caller consumed every lane's rows, so all reads of the slot have completed
    if (lane == 0 && t0 + S * U < t_stop)
      issue(stage, t0 + S * U);
    if (++stage == S)
    {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// plane kernel: tiles of the occupancy bit rows.  A warp's 4 x-rows x 8 voxel pairs need the bits [z_tile - 24,
// z_tile + 40) of every row: three words from w = (z_tile - 24) >> 5 on.  TMA wants the innermost box coordinate on a
// 16-byte boundary (tools/tma_probe.cu: an unaligned one raises "illegal instruction"), so the box row is the 8 words
// from w & ~3 on (-4 at the start of a row: TMA fills zeros, which is "no obstacle").  Tile layout [y row][x][8 words].
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <int U, int S>
struct BitTiles
{
  static constexpr int kTile = U * 128;
  const CUtensorMap* map;
  uint32_t tiles, bars;
  int c0, c1, lane;  // first word of the box (a multiple of 4), first x row
  uint32_t word_off;  // byte offset of the lane's three words inside its 32-byte box row
  int stage = 0;
  unsigned phase = 0;
  __device__ __forceinline__ void issue(int s, int t)
  {
    mbar_expect_tx(bars + 8 * s, kTile);
    tma_load_3d(tiles + s * kTile, map, c0, c1, t, bars + 8 * s);
  }
  __device__ __forceinline__ void begin(int t0, int t_stop)
  {
    if (lane == 0)
    {
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (t0 + s * U < t_stop)
          issue(s, t0 + s * U);
    }
  }
  __device__ __forceinline__ void wait()
  {
    mbar_wait(bars + 8 * stage, phase);
  }
  // bit j of the result: row j of the block holds an obstacle bit in some x row of the tile (each lane looks at one
  // (x, row) of the 4 x U, one ballot collects them)
  __device__ __forceinline__ unsigned rows_with_bits() const
  {
    unsigned rows = 0;
#pragma unroll
    for (int r = 0; r < U / 8; ++r)
    {
      // only the three words the tile's voxels can reach (the box row is wider for TMA's alignment rule)
      const uint32_t at = tiles + stage * kTile + (((lane & 7) + 8 * r) * 4 + (lane >> 3)) * 32 + word_off;
      const unsigned b = __ballot_sync(kFull, (lds32(at) | lds32(at + 4) | lds32(at + 8)) != 0u);
      rows |= ((b | (b >> 8) | (b >> 16) | (b >> 24)) & 0xffu) << (8 * r);
    }
    return rows;
  }
  __device__ __forceinline__ uint4 row(int j) const
  {
    const uint32_t at = tiles + stage * kTile + (j * 4 + (lane >> 3)) * 32 + word_off;
    return make_uint4(lds32(at), lds32(at + 4), lds32(at + 8), 0u);
  }
  __device__ __forceinline__ void release(int t0, int t_stop)
  {
    __syncwarp();
    if (lane == 0 && t0 + S * U < t_stop)
      issue(stage, t0 + S * U);
    if (++stage == S)
    {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// The voxel pair's value from the words of its row.  o = bit offset of z0 from the tile's first word: the window starts
// at the word of z_tile - 24 and z_tile is a multiple of 16, so o is 32 .. 62 and three words hold [z0 - 32, z0 + 32).
__device__ __forceinline__ unsigned pair_from_bits(const uint4& w, int o, int cap, unsigned w_here)
{
  const int s = o - 32;
  const unsigned lo = __funnelshift_r(w.x, w.y, s);  // bits [z0 - 32, z0): bit 31 is z0 - 1
  const unsigned hi = __funnelshift_r(w.y, w.z, s);  // bits [z0, z0 + 32): bit 0 is z0
  const unsigned above1 = hi >> 1;  // bits [z0 + 1, z0 + 32)
  const int up1 = above1 ? __ffs(above1) - 1 : 64;
  const int dn0 = lo ? __clz(lo) + 1 : 64;
  const bool occ0 = hi & 1u;
  const int up0 = occ0 ? 0 : up1 + 1;
  const int dn1 = occ0 ? 1 : dn0 + 1;
  // beyond the reach: cap = reach + 1 and cap^2 >= C, so the value is at or below BIAS and loses against the window
  const int d0 = min(min(up0, dn0), cap), d1 = min(min(up1, dn1), cap);
  return w_here - (unsigned)(d0 * d0) - ((unsigned)(d1 * d1) << 16);
}

template <int R, int U, int LEAD, int N, int P, typename Src>
struct BitTapPairs
{
  static __device__ __forceinline__ void run(unsigned (&acc)[N], const Src& src, unsigned rows, int o, int cap,
                                             unsigned w_here, unsigned one)
  {
    if ((rows >> (2 * P)) & 3u)
    {
      const unsigned fa = pair_from_bits(src.row(2 * P), o, cap, w_here);
      const unsigned fb = pair_from_bits(src.row(2 * P + 1), o, cap, w_here);
      tap_pair<R, LEAD, N, 2 * P>(acc, fa, fb, one);
    }
    BitTapPairs<R, U, LEAD, N, P + 1, Src>::run(acc, src, rows, o, cap, w_here, one);
  }
};
template <int R, int U, int LEAD, int N, typename Src>
struct BitTapPairs<R, U, LEAD, N, U / 2, Src>
{
  static __device__ __forceinline__ void run(unsigned (&)[N], const Src&, unsigned, int, int, unsigned, unsigned)
  {
  }
};

// ---- kernels ---------------------------------------------------------------------------------------------------------------
template <int R, int U, int S, typename OutT>
__global__ void __launch_bounds__(32)
    edt_axis_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                        int tiles_z, int zp0, int nx, int seg_len, int max_d2, unsigned one)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int LB = 2 * sizeof(OutT);
  const int lane = threadIdx.x;
  const int ty = blockIdx.x / tiles_z, tz = blockIdx.x % tiles_z;
  const uint32_t base = smem_u32(smem);
  TileSource<U, S> src;
  src.map = &in_map;
  src.tiles = base;
  src.bars = base + S * U * 128 + 3 * U * 32 * LB;
  src.lane_off = lane * 4;
  src.c0 = zp0 + 8 * tz;
  src.c1 = 4 * ty;
  src.lane = lane;
  FinalSink<U, OutT> sink;
  sink.map = &out_map;
  sink.tiles = base + S * U * 128;
  sink.lane_off = lane * LB;
  sink.c0 = src.c0;
  sink.c1 = src.c1;
  sink.lane = lane;
  sink.out_const = (unsigned)(max_d2 + (int)kBias) * 0x00010001u;
  if (lane == 0)
  {
#pragma unroll
    for (int s = 0; s < S; ++s)
      mbar_init(src.bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  sink.init(sizeof(OutT) == 1 ? (unsigned)(max_d2 | (max_d2 << 8)) : (unsigned)max_d2 * 0x00010001u);
  const int oa = blockIdx.y * seg_len;
  march<R, U>(src, sink, oa, min(oa + seg_len, nx), nx, one);
}

// The plane kernel runs the same loop on bit tiles: a row is turned into distances only inside a pair that holds
// obstacle bits (bits that all lie out of reach give values at or below BIAS, which change nothing).
template <int R, int U, int S>
__global__ void __launch_bounds__(32, R <= 24 ? B200DF_EDT_MINB : 8)
    edt_plane_tma_kernel(const __grid_constant__ CUtensorMap bits_map, const __grid_constant__ CUtensorMap out_map,
                         int tiles_z, int zp0, int ny, int reach, int max_d2, int seg_len, unsigned one)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x;
  const int tx = blockIdx.x / tiles_z, tz = blockIdx.x % tiles_z;
  const int zt = 2 * (zp0 + 8 * tz);   // first voxel of the tile
  const int w0 = (zt - 24) >> 5;       // first word of the tile's bit window (-1 at the start of a row)
  const int wa = w0 & ~3;              // first word of the TMA box
  const uint32_t base = smem_u32(smem);
  MidSink<U> sink;
  sink.map = &out_map;
  sink.tiles = base;
  sink.lane_off = lane * 4;
  sink.c0 = zp0 + 8 * tz;
  sink.c1 = 4 * tx;
  sink.lane = lane;
  BitTiles<U, S> src;
  src.map = &bits_map;
  src.tiles = base + 3 * U * 128;
  src.bars = src.tiles + S * U * 128;
  src.c0 = wa;
  src.word_off = (uint32_t)(w0 - wa) * 4u;
  src.c1 = 4 * tx;
  src.lane = lane;
  if (lane == 0)
  {
#pragma unroll
    for (int s = 0; s < S; ++s)
      mbar_init(src.bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  sink.init(kBiasPair);
  const int o = zt + 2 * (lane & 7) - 32 * w0;  // bit offset of the lane's z0 in the window
  const int cap = reach + 1;
  const unsigned w_here = (unsigned)(max_d2 + (int)kBias) * 0x00010001u;

  constexpr int LEAD = lead_rows(R, U), N = LEAD + R + U;
  unsigned acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
    acc[i] = kBiasPair;
  const int oa = blockIdx.y * seg_len, ob = min(oa + seg_len, ny);
  int t0 = oa - LEAD;
  const int t_stop = ob + LEAD;
  src.begin(t0, t_stop);
  constexpr int kDrain = LEAD + R;
  int quiet = kDrain;
  for (; t0 < t_stop; t0 += U)
  {
    src.wait();
    const unsigned rows = src.rows_with_bits();  // rows outside the array are zero-filled: no bits
    const int o0 = t0 - LEAD;
    const bool emit = o0 >= oa;
    if (!rows && quiet >= kDrain)
    {
      src.release(t0, t_stop);
      if (emit)
        sink.flush_inf(o0);
      continue;
    }
    if (rows)
    {
      BitTapPairs<R, U, LEAD, N, 0, BitTiles<U, S>>::run(acc, src, rows, o, cap, w_here, one);
      quiet = 0;
    }
    else
      quiet += U;
    src.release(t0, t_stop);
    if (emit)
    {
      sink.acquire();
#pragma unroll
      for (int j = 0; j < U; ++j)
        sink.put(j, acc[j]);
      sink.flush(o0);
    }
#pragma unroll
    for (int k = 0; k < N - U; ++k)
      acc[k] = acc[k + U];
#pragma unroll
    for (int k = N - U; k < N; ++k)
      acc[k] = kBiasPair;
  }
  sink.finish();
}

// ---- host side -------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 3-D map over an array of [d2][d1][d0] elements (d0 fastest) with the box {8, 4, U} in the ORDER (dim_a, dim_b, dim_c)
// given by the strides: dims[] / strides[] are listed fastest-box-dimension first.
bool make_map(CUtensorMap* m, void* base, int elem_bytes, const cuuint64_t dims[3], const cuuint64_t strides_bytes[2],
              int U)
{
  EncodeTiledFn fn = encode_fn();
  if (!fn)
    return false;
  const cuuint32_t box[3] = { 8u, 4u, (cuuint32_t)U };
  const cuuint32_t estr[3] = { 1u, 1u, 1u };
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
  return fn(m, dt, 3, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int env_int(const char* name, int fallback)
{
  const char* s = getenv(name);
  return s && *s ? atoi(s) : fallback;
}

template <int R, typename T, int U>
int run_u(b200df_field* f, T* out)
{
  constexpr int S = U >= 16 ? B200DF_EDT_S : 4, SB = U >= 24 ? 2 : U >= 16 ? 3 : 6;
  const int nz = f->grid.n[2], ny = f->grid.n[1], nx = f->grid.n[0];
  const int half = nz / 2, wpr = bits_wpr(nz), reach = f->radius - 1;
  CUtensorMap mid_store, mid_load, out_store, bits_map;
  {
    // bit rows [x][y][word]; box = 8 words x 4 x-rows x U y-rows
    EncodeTiledFn fn = encode_fn();
    const cuuint64_t dims[3] = { (cuuint64_t)wpr, (cuuint64_t)nx, (cuuint64_t)ny };
    const cuuint64_t str[2] = { (cuuint64_t)ny * wpr * 4u, (cuuint64_t)wpr * 4u };
    const cuuint32_t box[3] = { 8u, 4u, (cuuint32_t)U };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    if (!fn || fn(&bits_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, f->bits, dims, str, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
  }
  {
    // the plane kernel's tile is 8 pairs x 4 x-rows x U y-rows; the array is [x][y][pair]
    const cuuint64_t dims[3] = { (cuuint64_t)half, (cuuint64_t)nx, (cuuint64_t)ny };
    const cuuint64_t str[2] = { (cuuint64_t)ny * half * 4u, (cuuint64_t)half * 4u };
    if (!make_map(&mid_store, f->tmp_a, 4, dims, str, U))
      return -1;
  }
  {
    const cuuint64_t dims[3] = { (cuuint64_t)half, (cuuint64_t)ny, (cuuint64_t)nx };
    const cuuint64_t str[2] = { (cuuint64_t)half * 4u, (cuuint64_t)ny * half * 4u };
    if (!make_map(&mid_load, f->tmp_a, 4, dims, str, U))
      return -1;
    const cuuint64_t ostr[2] = { (cuuint64_t)half * 2u * sizeof(T), (cuuint64_t)ny * half * 2u * sizeof(T) };
    if (!make_map(&out_store, out, 2 * (int)sizeof(T), dims, ostr, U))
      return -1;
  }
  int n_slabs = env_int("B200DF_EDT_SLABS", 0);
  if (n_slabs <= 0)
    n_slabs = f->n_voxels >= (8L << 20) ? 4 : 1;
  n_slabs = std::max(1, std::min(n_slabs, std::min(half / 8, (int)b200df_field::kEdtStreams)));
  const int smem_plane = 3 * U * 128 + SB * U * 128 + 8 * SB;
  const int smem_axis = S * U * 128 + 3 * U * 32 * 2 * (int)sizeof(T) + 8 * S;
  const int tiles_x = (nx + 3) / 4, tiles_y = (ny + 3) / 4;
  auto seg_of = [&](int n_axis, const char* env, long warps) {
    int seg = env_int(env, 0);
    if (seg <= 0)
      seg = (int)std::min<long>(4, std::max<long>(1, (148L * 32 + warps - 1) / warps));
    seg = std::max(1, std::min(seg, std::max(1, n_axis / (2 * std::max(1, reach)))));
    int len = (n_axis + seg - 1) / seg;
    return ((len + U - 1) / U) * U;  // segment boundaries are block-aligned
  };
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
    const int tiles_z = slab / 8;
    const int len_y = seg_of(ny, "B200DF_EDT_SEG_Y", (long)tiles_x * (half / 8));
    const int len_x = seg_of(nx, "B200DF_EDT_SEG_X", (long)tiles_y * (half / 8));
    edt_plane_tma_kernel<R, U, SB><<<dim3((unsigned)(tiles_x * tiles_z), (unsigned)((ny + len_y - 1) / len_y)), 32,
                                     smem_plane, st>>>(bits_map, mid_store, tiles_z, zp0, ny, reach, f->max_d2, len_y,
                                                       1u);
    B200DF_LAUNCHED();
    edt_axis_tma_kernel<R, U, S, T><<<dim3((unsigned)(tiles_y * tiles_z), (unsigned)((nx + len_x - 1) / len_x)), 32,
                                      smem_axis, st>>>(mid_load, out_store, tiles_z, zp0, nx, len_x, f->max_d2, 1u);
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
// U = 16 rows per block: 4 register moves per row for the window shift (7 at U = 8, measured 0.191 -> 0.171 ms on
// Field B; U = 24 gives 0.170 with 163 registers in the plane kernel)
template <int R, typename T>
int run(b200df_field* f, T* out)
{
  return run_u<R, T, 16>(f, out);
}
}  // namespace

namespace b200df
{
// nz a multiple of 16 (whole 8-pair tiles, 16-byte row strides for every element size), clamp radius <= 31
bool edt_tma_applies(const b200df_field* f)
{
  return f->bits && f->grid.n[2] % 16 == 0 && f->radius >= 1 && f->radius <= 31 && f->n_voxels > 0 &&
         !env_int("B200DF_EDT_NO_TMA", 0) && encode_fn() != nullptr;
}

// bit rows (already packed by edt.cu) -> out (u8 when elem_size == 1, else u16); stream-ordered on f->stream.
// Returns -1 when a tensor map could not be encoded (the caller then takes the register-prefetch kernels).
int edt_tma(b200df_field* f, void* out)
{
  const int reach = f->radius - 1;
  if (f->elem_size == 1)
  {
    uint8_t* o = static_cast<uint8_t*>(out);
    if (reach <= 12)
      return run<12, uint8_t>(f, o);
    return run<24, uint8_t>(f, o);  // max_d2 <= 255 means radius <= 16
  }
  uint16_t* o = static_cast<uint16_t*>(out);
  if (reach <= 12)
    return run<12, uint16_t>(f, o);
  if (reach <= 24)
    return run<24, uint16_t>(f, o);
  return run<30, uint16_t>(f, o);
}
}  // namespace b200df
