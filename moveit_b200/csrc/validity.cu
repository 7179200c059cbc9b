// Batched state validity (kernels 1-4 fused): forward kinematics, collision-sphere posing, sphere-vs-field lookups
// and ACM-pruned intra-group sphere pairs, one thread per state, fp64 in the reference's operation order.
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
// Layout of the verdict kernel.  The links are swept once in model (DFS) order.  A verdict is a logical OR over
// sphere tests, so the intra-group pairs are folded into the same sweep: a link that is the lower-index side of an
// enabled pair keeps its posed sphere centres in shared memory ("stored" link); when the higher-index partner is
// swept, each of its freshly posed spheres is tested against the stored spheres of the partners whose bounding
// spheres pass the reference's prune.  Only stored links occupy shared memory ([item][component][thread], a thread
// only touches its own column), which is what lets 8-9 warps share an SM instead of 3.
//
// Compiled with --fmad=false so that mul/add round like the reference's non-FMA x86-64 build.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <map>
#include <memory>

#include "validity_common.cuh"

using namespace b200df;

namespace b200df
{
int launch_verdict_v1(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                      int q_dtype, long n, int flags, uint8_t* verdict, cudaStream_t stream);
}

namespace
{
constexpr int kBatch = 4;  // spheres posed / looked up / pair-tested together (independent chains for the fp64 pipe)

// redo_idx / redo_count: when set, the kernel re-evaluates the listed states only (the FAST path's ambiguous set); it is
// then launched with a fixed grid and every CTA strides over the list, whose length lives in device memory.
template <typename QT, typename FT, int B>
__global__ void __launch_bounds__(B) validity_kernel(GroupDev g, LookupDev wf, LookupDev sf, const QT* __restrict__ q,
                                                     long n, int flags, uint8_t* __restrict__ verdict,
                                                     const int* __restrict__ redo_idx,
                                                     const int* __restrict__ redo_count)
{
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const bool do_pairs = (flags & B200DF_CHECK_INTRA) != 0 && g.n_partners > 0;
  double* slots = smem;                                                           // [n_slots][12][B]
  double* stc = slots + (size_t)g.n_slots * 12 * B;                               // [n_store_spheres][3][B]
  double* stb = stc + (do_pairs ? (size_t)g.n_store_spheres * 3 * B : 0);         // [n_store_links][3][B]
  QT* qs = reinterpret_cast<QT*>(stb + (do_pairs ? (size_t)g.n_store_links * 3 * B : 0));  // [B][n_vars]
  const long total = redo_idx ? (long)*redo_count : n;
  for (long base = (long)blockIdx.x * B; base < total; base += (long)gridDim.x * B)
  {
  long state = base + tid;
  const bool active = state < total;
  if (redo_idx)
  {
    __syncthreads();  // the previous tile's readers are done with qs
    state = active ? redo_idx[state] : 0;
    for (int v = 0; v < g.n_vars; ++v)
      qs[tid * g.n_vars + v] = q[state * g.n_vars + v];
  }
  else
  {
    // the tile's joint values are one contiguous run of the batch: straight coalesced copy, rows stay rows
    const int count = B * g.n_vars;
    const long limit = (n - base) * g.n_vars;
    const QT* src = q + base * g.n_vars;
    for (int e = tid; e < count; e += B)
      qs[e] = (e < limit) ? src[e] : QT(0);
  }
  __syncthreads();
  const QT* my_q = qs + (size_t)tid * g.n_vars;
  double* my_slots = slots + tid;
  double* my_stc = stc + tid;
  double* my_stb = stb + tid;

  const bool do_robot = (flags & B200DF_CHECK_ROBOT) != 0;
  const bool do_self = (flags & B200DF_CHECK_SELF) != 0;
  bool hit = !active;  // padding lanes never contribute work
  double cur[12];

  for (int i = 0; i < g.n_links; ++i)
  {
    fk_step_sparse(g.links[i], my_q, my_slots, 1, B, cur);
    const int gs = g.link_gslot[i];
    if (gs < 0)
      continue;
    const DevGroupLink& gl = g.glinks[gs];
    const int np = do_pairs ? gl.partner_end - gl.partner_begin : 0;
    const bool stored = do_pairs && gl.store_link >= 0;
    const bool self_here = do_self && gl.self_enabled;
    double bx = 0.0, by = 0.0, bz = 0.0;
    if (np > 0 || stored)
    {
      iso_apply(cur, gl.bs[0], gl.bs[1], gl.bs[2], bx, by, bz);  // posed bounding-sphere centre (types.cpp:398-399)
      if (stored)
      {
        my_stb[(gl.store_link * 3 + 0) * B] = bx;
        my_stb[(gl.store_link * 3 + 1) * B] = by;
        my_stb[(gl.store_link * 3 + 2) * B] = bz;
      }
    }
    // partners are handled 32 at a time so that one bit mask per lane says which of them passed the prune
    const int n_chunks = np > 0 ? (np + 31) / 32 : 1;
    for (int c = 0; c < n_chunks; ++c)
    {
      const int p0 = gl.partner_begin + 32 * c;
      unsigned lane_mask = 0u;
      if (np > 0)
      {
        const int pc = min(32, gl.partner_end - p0);
#pragma unroll 2
        for (int b = 0; b < pc; ++b)
        {
          const DevPartner& pr = g.partners[p0 + b];
          const double dx = my_stb[(pr.store_link * 3 + 0) * B] - bx;
          const double dy = my_stb[(pr.store_link * 3 + 1) * B] - by;
          const double dz = my_stb[(pr.store_link * 3 + 2) * B] - bz;
          const double d2 = (dx * dx + dy * dy) + dz * dz;
          // doBoundingSpheresIntersect compares the SQUARED distance with the UN-squared radius sum (Q2);
          // tight2 additionally drops link pairs none of whose spheres can touch (never changes a verdict)
          const bool pass = !hit && (d2 < pr.radius_sum) && (d2 < pr.tight2);
          lane_mask |= (pass ? 1u : 0u) << b;
        }
      }
      const unsigned warp_mask = __reduce_or_sync(0xffffffffu, lane_mask);
      for (int s0 = gl.sphere_begin; s0 < gl.sphere_end; s0 += kBatch)
      {
        const int cnt = min(kBatch, gl.sphere_end - s0);  // warp-uniform
        double cx[kBatch], cy[kBatch], cz[kBatch];
        unsigned widx[kBatch], sidx[kBatch];
        bool win[kBatch], sin_[kBatch];
        int wv[kBatch], sv[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
        {
          if (k < cnt)
          {
            const double* rel = g.sph + 4 * (s0 + k);
            iso_apply(cur, rel[0], rel[1], rel[2], cx[k], cy[k], cz[k]);  // updatePose (types.cpp:393-396)
            if (c == 0)
            {
              // the voxel loads are issued here and consumed after the pair tests of this batch
              if (do_robot)
              {
                widx[k] = cell_index(wf, cx[k], cy[k], cz[k], win[k]);
                B200DF_CHECK_INDEX(widx[k], (long long)(wf.lim[0] + 2) * wf.stride1);
                wv[k] = (int)static_cast<const FT*>(wf.d2)[widx[k]];
              }
              if (self_here)
              {
                sidx[k] = cell_index(sf, cx[k], cy[k], cz[k], sin_[k]);
                B200DF_CHECK_INDEX(sidx[k], (long long)(sf.lim[0] + 2) * sf.stride1);
                sv[k] = (int)static_cast<const FT*>(sf.d2)[sidx[k]];
              }
              if (stored)
              {
                const int slot = gl.store_sphere + (s0 + k - gl.sphere_begin);
                my_stc[(slot * 3 + 0) * B] = cx[k];
                my_stc[(slot * 3 + 1) * B] = cy[k];
                my_stc[(slot * 3 + 2) * B] = cz[k];
              }
            }
          }
        }
        unsigned mm = warp_mask;
        const double* t_row = g.pair_T + gl.pair_const_begin + (long)(s0 - gl.sphere_begin) * gl.pair_const_stride;
        while (mm)
        {
          const int b = __ffs(mm) - 1;
          mm &= mm - 1;
          const DevPartner& pr = g.partners[p0 + b];
          const double* t = t_row + pr.const_offset;
          const double* pc = my_stc + (size_t)pr.store_sphere * 3 * B;
          const int stride = gl.pair_const_stride;
          bool h = false;
          for (int m = 0; m < pr.count; ++m)
          {
            const double px = pc[(m * 3 + 0) * B], py = pc[(m * 3 + 1) * B], pz = pc[(m * 3 + 2) * B];
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
            {
              if (k < cnt)
              {
                const double ex = cx[k] - px, ey = cy[k] - py, ez = cz[k] - pz;
                const double e2 = (ex * ex + ey * ey) + ez * ez;
                h |= e2 < t[k * stride + m];  // ||c1-c2|| < r1+r2 (:574-584), sqrt-free, bit-equivalent threshold
              }
            }
          }
          hit |= h && ((lane_mask >> b) & 1u);
        }
        if (c == 0)
        {
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
          {
            if (k < cnt)
            {
              const double radius = g.sph[4 * (s0 + k) + 3];
              const int thr = g.sph_thr[s0 + k];
              if (do_robot)
                hit |= voxel_hits<FT>(wf, widx[k], win[k], wv[k], radius, thr, g.tol, g.max_prop);
              if (self_here)
                hit |= voxel_hits<FT>(sf, sidx[k], sin_[k], sv[k], radius, thr, g.tol, g.max_prop);
            }
          }
        }
      }
    }
    if (__all_sync(0xffffffffu, hit))
      break;  // every state of this warp already has its verdict
  }
  if (active)
    verdict[state] = hit ? 1 : 0;
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

size_t verdict_smem_bytes(const GroupDev& g, int block, bool pairs, size_t q_elem)
{
  size_t d = (size_t)g.n_slots * 12;
  if (pairs)
    d += (size_t)g.n_store_spheres * 3 + (size_t)g.n_store_links * 3;
  return d * block * sizeof(double) + (size_t)g.n_vars * block * q_elem;
}

struct Redo  // the ambiguous-state list of the FAST path (device pointers), or nulls
{
  const int* idx = nullptr;
  const int* count = nullptr;
};

template <typename QT, typename FT, int B>
int launch_verdict_block(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, const void* q, long n, int flags,
                         uint8_t* verdict, size_t bytes, cudaStream_t stream, Redo redo)
{
  B200DF_CUDA(allow_dynamic_smem(validity_kernel<QT, FT, B>, bytes));
  // redo: a fixed grid (a few CTAs per SM) strides over the list; its length is only known on the device
  const unsigned grid = redo.idx ? (unsigned)std::min<long>((n + B - 1) / B, 148 * 8) : (unsigned)((n + B - 1) / B);
  validity_kernel<QT, FT, B><<<grid, B, bytes, stream>>>(g, make_lookup(wf), make_lookup(sf),
                                                         static_cast<const QT*>(q), n, flags, verdict, redo.idx,
                                                         redo.count);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}

template <typename QT, typename FT>
int launch_verdict(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, const void* q, long n, int flags,
                   uint8_t* verdict, cudaStream_t stream, Redo redo)
{
  const bool pairs = (flags & B200DF_CHECK_INTRA) && g.n_partners > 0;
  // 64 states per CTA when four or more such CTAs fit an SM, else one warp per CTA
  size_t bytes = verdict_smem_bytes(g, 64, pairs, sizeof(QT));
  if (bytes <= 56 * 1024)
    return launch_verdict_block<QT, FT, 64>(g, wf, sf, q, n, flags, verdict, bytes, stream, redo);
  bytes = verdict_smem_bytes(g, 32, pairs, sizeof(QT));
  if (bytes > 220 * 1024)
  {
    set_error("robot too large for the per-state shared-memory layout (%zu bytes for 32 states)", bytes);
    return B200DF_ERR_UNSUPPORTED;
  }
  return launch_verdict_block<QT, FT, 32>(g, wf, sf, q, n, flags, verdict, bytes, stream, redo);
}

int dispatch_verdict(const GroupDev& g, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                     int q_dtype, long n, int flags, uint8_t* verdict, cudaStream_t stream, Redo redo = Redo())
{
  if (q_dtype == B200DF_Q_F32)
  {
    if (field_elem == 1)
      return launch_verdict<float, uint8_t>(g, wf, sf, q, n, flags, verdict, stream, redo);
    return launch_verdict<float, uint16_t>(g, wf, sf, q, n, flags, verdict, stream, redo);
  }
  if (field_elem == 1)
    return launch_verdict<double, uint8_t>(g, wf, sf, q, n, flags, verdict, stream, redo);
  return launch_verdict<double, uint16_t>(g, wf, sf, q, n, flags, verdict, stream, redo);
}

constexpr int kPrecisionV1 = B200DF_PRECISION_DEBUG_SWEEP;        // the round-1 kernel, an independent cross-check
constexpr int kPrecisionFastRaw = B200DF_PRECISION_DEBUG_SCREEN;  // FAST screening pass only: 0 / 1 / 2 = undecided
constexpr int64_t kFastMinStates = 65536;
constexpr int kPrecisionSmall = B200DF_PRECISION_DEBUG_SMALL;     // the one-CTA-per-state kernel at any batch size

// Batches up to this many states run on the one-CTA-per-state kernel (validity_small.cu): below it the batched
// kernels leave most of the GPU idle and the call time is one thread's walk through a whole state.
// B200DF_SMALL_MAX_STATES overrides it (0 switches the small-batch kernel off); read on every call so that a test
// can flip it.
int64_t small_max_states()
{
  const char* e = getenv("B200DF_SMALL_MAX_STATES");
  return e ? (int64_t)atoll(e) : (int64_t)16384;
}

// amb_idx (n ints) / amb_count (1 int): device scratch of the FAST path; allocated stream-ordered when null
int run_verdict(const b200df_group* group, const FieldDev& wf_in, const FieldDev& sf_in, int elem, const void* q_dev,
                int q_dtype, int64_t n, int flags, int precision, uint8_t* verdict_dev, cudaStream_t stream,
                int* amb_idx = nullptr, int* amb_count = nullptr)
{
  if (precision == kPrecisionV1)
    return launch_verdict_v1(group->dev, wf_in, sf_in, elem, q_dev, q_dtype, n, flags, verdict_dev, stream);
  // Signed fields.  A sphere centre in a free cell has negative_distance_square 0, so the signed distance there IS the
  // unsigned one; in an occupied cell (d2 == 0) the signed distance is negative and the sphere collides whenever
  // max_propagation_distance > 0 and radius > tolerance - which is also what the unsigned integer test "0 < threshold"
  // says.  When every sphere of the group meets that condition the verdict kernels therefore ignore the negative
  // field (same verdicts, integer test, FAST usable); the debug sweep above keeps the full fp64 predicate.
  FieldDev wf = wf_in, sf = sf_in;
  if (group->neg_irrelevant)
  {
    wf.neg = nullptr;
    sf.neg = nullptr;
  }
  if (precision == kPrecisionSmall ||
      ((precision == B200DF_PRECISION_EXACT || precision == B200DF_PRECISION_FAST) && n <= small_max_states() &&
       small_supported(group)))
    return small_launch(group, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, stream);
  const bool raw = precision == kPrecisionFastRaw;
  // Robots whose per-state shared-memory column does not fit an SM even at 32 states per CTA (hundreds of stored
  // spheres) cannot use the one-thread-per-state kernels at all; the one-CTA-per-state kernel keeps one state's
  // transforms and centres only and serves them at any batch size.
  if (!raw && small_supported(group) &&
      verdict_smem_bytes(group->dev, 32, (flags & B200DF_CHECK_INTRA) && group->dev.n_partners > 0,
                         q_elem_size(q_dtype)) > 220 * 1024)
    return small_launch(group, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, stream);
  if (raw && !fast_supported(group, wf, sf, q_dtype, flags))
  {
    set_error("the FAST screening pass does not support this call (needs fp32 joint values, unsigned fields, a model "
              "that fits the constant-bank tables)");
    return B200DF_ERR_UNSUPPORTED;
  }
  // below ~64 k states the screening pass + redo launch cost more than they save (measured: 32 k states 156 us FAST vs
  // 92 us EXACT, 128 k states 225 vs 308); the verdicts are the same either way
  const bool fast_pays = raw || n >= kFastMinStates;
  if ((precision != B200DF_PRECISION_FAST && !raw) || !fast_pays || !fast_supported(group, wf, sf, q_dtype, flags))
    return dispatch_verdict(group->dev, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, stream);
  // FAST: fp32 screening pass, then the exact kernel over the states whose verdict could depend on rounding
  B200DF_REQUIRE(n < (1L << 31), "FAST batches are limited to 2^31 states per call");
  void* scratch = nullptr;
  if (!amb_idx)
  {
    keep_async_pool(group->device);
    B200DF_CUDA(cudaMallocAsync(&scratch, ((size_t)n + 1) * sizeof(int), stream));
    amb_count = static_cast<int*>(scratch);
    amb_idx = amb_count + 1;
  }
  int rc = fast_launch(group, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, amb_idx, amb_count, stream);
  if (!rc && !raw)
  {
    // the undecided states (a few hundred per million) go to the one-CTA-per-state kernel: the one-thread-per-state
    // kernel would spend a full ~85 us warp latency on a handful of warps
    if (small_supported(group))
      rc = small_launch(group, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, stream, amb_idx, amb_count);
    else
    {
      Redo redo;
      redo.idx = amb_idx;
      redo.count = amb_count;
      rc = dispatch_verdict(group->dev, wf, sf, elem, q_dev, q_dtype, n, flags, verdict_dev, stream, redo);
    }
  }
  if (scratch)
    cudaFreeAsync(scratch, stream);
  return rc;
}
}  // namespace

namespace
{
constexpr int64_t kChunkStates = 3 << 17;  // 393 216: 10 M page-locked states 7.01 ms (2^20: 7.14, 2^19: 7.08, 2^18: 7.16)

// B200DF_CHUNK_STATES overrides the pipeline's chunk size (tuning knob, read once)
int64_t chunk_states()
{
  static const int64_t v = [] {
    const char* e = getenv("B200DF_CHUNK_STATES");
    const long long x = e ? atoll(e) : 0;
    return x >= 1024 ? (int64_t)x : kChunkStates;
  }();
  return v;
}

// Small batches (the per-state isValid() callers of the reference: OMPL's StateValidityChecker, checkCollision on one
// RobotState) are latency-bound, not bandwidth-bound: two staged copies and their driver calls cost more than the
// kernel.  Up to kSmallStates states travel through one mapped pinned buffer instead: the kernel reads the joint
// values and writes the verdicts across PCIe itself, so the call is one launch and one stream synchronise.
constexpr int64_t kSmallStates = 2048;
constexpr size_t kSmallQBytes = 256 * 1024;

int ensure_ctx(b200df_group::StagePool::Ctx* c, size_t q_bytes, size_t v_bytes)
{
  for (int i = 0; i < kSlots; ++i)
  {
    if (!c->stream[i])
      B200DF_CUDA(cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking));
  }
  if (q_bytes > c->q_bytes)
  {
    for (int i = 0; i < kSlots; ++i)
    {
      cudaFree(c->q[i]);
      c->q[i] = nullptr;
      B200DF_CUDA(cudaMalloc(&c->q[i], q_bytes));
    }
    c->q_bytes = q_bytes;
  }
  if (v_bytes > c->v_bytes)
  {
    for (int i = 0; i < kSlots; ++i)
    {
      cudaFree(c->v[i]);
      c->v[i] = nullptr;
      B200DF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->v[i]), v_bytes));
      cudaFree(c->amb[i]);
      c->amb[i] = nullptr;
      B200DF_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->amb[i]), (v_bytes + 1) * sizeof(int)));
    }
    c->v_bytes = v_bytes;
  }
  return B200DF_OK;
}
}  // namespace

namespace b200df
{
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
    B200DF_REQUIRE(f->n_voxels < (1L << 31), "field too large for 32-bit voxel indices");
    // the verdict kernels treat out-of-bounds spheres as free, which is what the reference's predicate
    // (max_prop > dist with dist = the field's max_distance, collision_distance_field_types.cpp:216-222) says only
    // when the field propagates at least as far as the group looks (CollisionEnvDistanceField uses one value for both)
    B200DF_REQUIRE(f->desc.max_distance >= group->desc.max_propagation_distance,
                   "field max_distance is smaller than the group's max_propagation_distance");
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
}  // namespace b200df

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
  struct DestroyModel
  {
    void operator()(b200df_model* p) const
    {
      b200df_model_destroy(p);
    }
  };
  std::unique_ptr<b200df_model, DestroyModel> m(new b200df_model);
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
    l.jkind = 0;
    if (l.joint_type == B200DF_JOINT_REVOLUTE)
    {
      if (std::fabs(l.ax) == 1.0 && l.ay == 0.0 && l.az == 0.0)
        l.jkind = 1;
      else if (l.ax == 0.0 && std::fabs(l.ay) == 1.0 && l.az == 0.0)
        l.jkind = 2;
      else if (l.ax == 0.0 && l.ay == 0.0 && std::fabs(l.az) == 1.0)
        l.jkind = 3;
    }
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
  if (m->fk_group)
    b200df_group_destroy(m->fk_group);
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
  // every early return below releases what was built so far (device tables, staging pool, FAST / small-batch paths)
  struct Destroy
  {
    void operator()(b200df_group* p) const
    {
      b200df_group_destroy(p);
    }
  };
  std::unique_ptr<b200df_group, Destroy> g(new b200df_group);
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
    DevGroupLink gl{};
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
    gl.store_link = gl.store_sphere = -1;
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
  std::vector<double> pair_thr((size_t)U * U);
  for (int a = 0; a < U; ++a)
    for (int b = 0; b < U; ++b)
      pair_thr[(size_t)a * U + b] = sqrt_less_threshold(radii[a] + radii[b]);

  // ---- tables of the fused sweep.  A pair only matters when both links have spheres.
  const int G = (int)g->glinks.size();
  auto count_of = [&](int gs) { return g->glinks[gs].sphere_end - g->glinks[gs].sphere_begin; };
  // radius around the reference's bounding-sphere centre that really encloses the link's collision spheres
  // (the spheres may come from a link_body_decompositions override and stick out of the shape's bounding sphere)
  std::vector<double> enclosing(G, 0.0);
  for (int gs = 0; gs < G; ++gs)
  {
    const DevGroupLink& gl = g->glinks[gs];
    for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
    {
      const double* sp = &g->sph[(size_t)s * 4];
      const double dx = sp[0] - gl.bs[0], dy = sp[1] - gl.bs[1], dz = sp[2] - gl.bs[2];
      enclosing[gs] = std::max(enclosing[gs], std::sqrt(dx * dx + dy * dy + dz * dz) + sp[3]);
    }
  }
  std::vector<std::vector<int>> lower(G);  // lower-index partners of each link, ascending
  for (const DevPair& p : g->pairs)
    if (count_of(p.gi) > 0 && count_of(p.gj) > 0)
      lower[p.gj].push_back(p.gi);
  int n_store_links = 0, n_store_spheres = 0;
  for (int gs = 0; gs < G; ++gs)
  {
    bool is_lower = false;
    for (int h = gs + 1; h < G && !is_lower; ++h)
      is_lower = std::find(lower[h].begin(), lower[h].end(), gs) != lower[h].end();
    if (is_lower)
    {
      g->glinks[gs].store_link = n_store_links++;
      g->glinks[gs].store_sphere = n_store_spheres;
      n_store_spheres += count_of(gs);
    }
  }
  std::vector<DevPartner> partners;
  std::vector<double> pair_T;
  for (int gs = 0; gs < G; ++gs)
  {
    DevGroupLink& gl = g->glinks[gs];
    gl.partner_begin = (int)partners.size();
    int stride = 0;
    for (int lo : lower[gs])
    {
      DevPartner pr{};
      pr.store_link = g->glinks[lo].store_link;
      pr.store_sphere = g->glinks[lo].store_sphere;
      pr.count = count_of(lo);
      pr.const_offset = stride;
      pr.radius_sum = g->glinks[lo].bs[3] + gl.bs[3];
      const double reach = (enclosing[lo] + enclosing[gs]) * (1.0 + 1e-9) + 1e-9;  // slack >> any fp64 rounding here
      pr.tight2 = reach * reach;
      stride += pr.count;
      partners.push_back(pr);
    }
    gl.partner_end = (int)partners.size();
    gl.pair_const_begin = (int)pair_T.size();
    gl.pair_const_stride = stride;
    for (int s = gl.sphere_begin; s < gl.sphere_end && stride > 0; ++s)
      for (int lo : lower[gs])
        for (int m = g->glinks[lo].sphere_begin; m < g->glinks[lo].sphere_end; ++m)
          pair_T.push_back(pair_thr[(size_t)sph_cls[s] * U + sph_cls[m]]);
  }
  B200DF_REQUIRE(pair_T.size() < (size_t)1 << 31, "too many sphere pairs");

  DeviceGuard dg(g->device);
  GroupDev& dev = g->dev;
  dev.links = model->d_links;
  dev.n_links = L;
  dev.n_vars = model->n_vars;
  dev.n_slots = model->n_slots;
  dev.n_glinks = G;
  dev.n_spheres = (int)(g->sph.size() / 4);
  dev.n_pairs = (int)g->pairs.size();
  dev.n_cls = U;
  dev.n_store_links = n_store_links;
  dev.n_store_spheres = n_store_spheres;
  dev.n_partners = (int)partners.size();
  dev.n_pair_consts = (long)pair_T.size();
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
  g->partners = partners;
  g->sph_thr = sph_thr;
  g->neg_irrelevant = d->max_propagation_distance > 0.0;
  for (size_t i = 3; i < g->sph.size(); i += 4)
    if (!(g->sph[i] > d->collision_tolerance))
      g->neg_irrelevant = false;
  // gather form of the pair list for the gradient kernel
  std::vector<int> sph_glink(g->sph.size() / 4, 0), nbr_off(G + 1, 0);
  std::vector<int4> nbr;
  for (int gs = 0; gs < G; ++gs)
  {
    for (int s2 = g->glinks[gs].sphere_begin; s2 < g->glinks[gs].sphere_end; ++s2)
      sph_glink[s2] = gs;
    for (const DevPair& pr : g->pairs)  // already in (i asc, j asc) order
      if (pr.gj == gs)
        nbr.push_back(make_int4(pr.gi, 1, g->glinks[pr.gi].sphere_begin, g->glinks[pr.gi].sphere_end));
    for (const DevPair& pr : g->pairs)
      if (pr.gi == gs)
        nbr.push_back(make_int4(pr.gj, 0, g->glinks[pr.gj].sphere_begin, g->glinks[pr.gj].sphere_end));
    nbr_off[gs + 1] = (int)nbr.size();
  }
  if ((rc = upload(g.get(), partners, &dev.partners)))
    return rc;
  if ((rc = upload(g.get(), sph_glink, &dev.sph_glink)))
    return rc;
  if ((rc = upload(g.get(), nbr_off, &dev.nbr_off)))
    return rc;
  dev.n_nbr = (int)nbr.size();
  if ((rc = upload(g.get(), nbr, &dev.nbr)))
    return rc;
  if ((rc = upload(g.get(), pair_T, &dev.pair_T)))
    return rc;
  g->stage = new b200df_group::StagePool;
  if ((rc = small_create(g.get())))
    return rc;
  if ((rc = fast_create(g.get())))
  {
    delete g->stage;
    for (void* p : g->allocations)
      cudaFree(p);
    return rc;
  }
  *out = g.release();
  return B200DF_OK;
}

int b200df_group_destroy(b200df_group* g)
{
  if (!g)
    return B200DF_OK;
  DeviceGuard dg(g->device);
  delete g->stage;
  fast_destroy(g);
  small_destroy(g);
  cudaFree(g->d_link_fields);
  cudaFree(g->d_lf_enabled);
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

int b200df_check_route(const b200df_group* group, const b200df_field* world, const b200df_field* self, int q_dtype,
                       int64_t n, int flags, int precision, int32_t* route, int32_t* why_not_fast)
{
  B200DF_REQUIRE(group && route && why_not_fast && n >= 0, "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown joint-value type");
  B200DF_REQUIRE(precision == B200DF_PRECISION_EXACT || precision == B200DF_PRECISION_FAST, "unknown precision");
  *why_not_fast = 0;
  // the same decisions as run_verdict, in the same order
  const bool pairs = (flags & B200DF_CHECK_INTRA) && group->dev.n_partners > 0;
  if (n <= small_max_states() && small_supported(group))
  {
    *route = B200DF_ROUTE_SMALL;
    if (precision == B200DF_PRECISION_FAST)
      *why_not_fast = B200DF_NOFAST_BATCH;
    return B200DF_OK;
  }
  if (small_supported(group) && verdict_smem_bytes(group->dev, 32, pairs, q_elem_size(q_dtype)) > 220 * 1024)
  {
    *route = B200DF_ROUTE_SMALL;
    if (precision == B200DF_PRECISION_FAST)
      *why_not_fast = B200DF_NOFAST_SMEM;
    return B200DF_OK;
  }
  *route = B200DF_ROUTE_EXACT;
  if (precision != B200DF_PRECISION_FAST)
    return B200DF_OK;
  int why = 0;
  if (!fast_tables_ok(group))
    why |= B200DF_NOFAST_TABLES;
  const bool signed_world = (flags & B200DF_CHECK_ROBOT) && world && world->desc.propagate_negative;
  const bool signed_self = (flags & B200DF_CHECK_SELF) && self && self->desc.propagate_negative;
  if ((signed_world || signed_self) && !group->neg_irrelevant)
    why |= B200DF_NOFAST_SIGNED;
  if (n < kFastMinStates)
    why |= B200DF_NOFAST_BATCH;
  if (!why)
    *route = B200DF_ROUTE_FAST;
  *why_not_fast = why;
  return B200DF_OK;
}

int b200df_check_batch_device(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q_dev,
                              int q_dtype, int64_t n, int flags, int precision, uint8_t* verdict_dev, void* stream)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q_dev && verdict_dev)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  B200DF_REQUIRE(precision == B200DF_PRECISION_EXACT || precision == B200DF_PRECISION_FAST || precision == kPrecisionV1 ||
                     precision == kPrecisionFastRaw || precision == kPrecisionSmall,
                 "unknown precision");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  int elem;
  int rc = prepare_fields(group, world, self, flags, &wf, &sf, &elem);
  if (rc)
    return rc;
  DeviceGuard dg(group->device);
  return run_verdict(group, wf, sf, elem, q_dev, q_dtype, n, flags, precision, verdict_dev,
                     static_cast<cudaStream_t>(stream));
}

// fp64 -> fp32, round to nearest even (cvtpd2ps): exactly the conversion the kernels' own fp64 -> fp32 load performs.
// The loop is the host-side cost of the narrowing path below; with 256-bit converts a core streams ~8 GB/s of doubles.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void narrow_rows_avx2(float* dst, const double* src, size_t count)
{
  size_t k = 0;
  // the page-locked staging buffer is read next by the GPU's copy engine, not by this core: non-temporal stores skip
  // the read-for-ownership of every destination line (a third of the memory traffic of this loop)
  static const bool stream_stores = !(getenv("B200DF_NARROW_NT") && atoi(getenv("B200DF_NARROW_NT")) == 0);
  if (stream_stores && ((uintptr_t)dst & 31) == 0)
  {
    for (; k + 8 <= count; k += 8)
    {
      const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(src + k));
      const __m128 b = _mm256_cvtpd_ps(_mm256_loadu_pd(src + k + 4));
      _mm256_stream_ps(dst + k, _mm256_set_m128(b, a));
    }
    _mm_sfence();
  }
  for (; k + 8 <= count; k += 8)
  {
    const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(src + k));
    const __m128 b = _mm256_cvtpd_ps(_mm256_loadu_pd(src + k + 4));
    _mm_storeu_ps(dst + k, a);
    _mm_storeu_ps(dst + k + 4, b);
  }
  for (; k < count; ++k)
    dst[k] = (float)src[k];
}
#endif

static void narrow_rows(float* dst, const double* src, size_t count)
{
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2)
  {
    narrow_rows_avx2(dst, src, count);
    return;
  }
#endif
  for (size_t k = 0; k < count; ++k)
    dst[k] = (float)src[k];
}

// FAST on pageable fp64 rows (the shape a CollisionEnv caller has: RobotState variables are doubles in ordinary heap
// memory).  The screening pass works on the joint values rounded to fp32 anyway (the rounding is part of its error
// margins), so the rows are narrowed on the host WHILE they are staged into page-locked memory: half the staging
// writes and half the PCIe bytes (36 instead of 72 per Panda state).  Each of a few host threads runs an independent
// pipeline over its own chunks - narrow into its pinned buffer, H2D, screening kernel, D2H of the verdict bytes - so
// the conversion of one chunk overlaps the copies and kernels of the others.  Screening leaves 0 / 1 / 2 (undecided,
// ~0.03 % of the states); afterwards exactly those rows travel as fp64 and are decided by the exact kernel, like the
// device-side FAST path does.  Same verdicts as EXACT by the same argument.
static int check_batch_narrowed(const b200df_group* group, const FieldDev& wf, const FieldDev& sf, int elem,
                                const double* q, int64_t n, int flags, uint8_t* verdict,
                                b200df_group::StagePool::Ctx* ctx)
{
  typedef b200df_group::StagePool::Ctx Ctx;
  const int nv = group->dev.n_vars;
  const size_t row32 = (size_t)nv * sizeof(float);
  const int64_t cs = 1 << 18;
  const int64_t n_chunks = (n + cs - 1) / cs;
  const unsigned hw = std::thread::hardware_concurrency();
  // lanes: one per hardware thread of the host, at most kLanes (B200DF_NARROW_LANES overrides: measurement knob;
  // 16-core box, 10 M states: 4 lanes 4.95e8 states/s, 8 6.87e8, 12 7.05e8, 16 7.66e8)
  static const int lanes_env = getenv("B200DF_NARROW_LANES") ? atoi(getenv("B200DF_NARROW_LANES")) : 0;
  // the host's threads are shared with whoever else stages at the same time: other calls of this process (one per
  // device under b200df_multi) and, under torchrun, the other ranks of the node (LOCAL_WORLD_SIZE)
  static std::atomic<int> active_calls{ 0 };
  struct ActiveCall
  {
    std::atomic<int>& c;
    int now;
    explicit ActiveCall(std::atomic<int>& counter) : c(counter), now(++counter)
    {
    }
    ~ActiveCall()
    {
      --c;
    }
  } active(active_calls);
  static const int local_ranks = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
  const unsigned sharers = (unsigned)std::max(1, active.now) * (unsigned)local_ranks;
  const int64_t lanes_want = lanes_env > 0 ? lanes_env : (int64_t)std::max(2u, hw / sharers);
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(Ctx::kLanes, lanes_want), n_chunks));
  // lane buffers: [fp32 rows | verdict bytes], pinned and device
  const size_t v_at = ((size_t)cs * row32 + 255) & ~size_t(255);
  const size_t lane_bytes = v_at + (size_t)cs;
  for (int w = 0; w < T; ++w)
  {
    Ctx::Lane& l = ctx->lane[w];
    if (!l.st)
      B200DF_CUDA(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
    if (l.states < (size_t)cs || l.row < row32)
    {
      if (l.hq)
        cudaFreeHost(l.hq);
      cudaFree(l.dq);
      cudaFree(l.amb);
      l.hq = l.dq = nullptr;
      l.amb = nullptr;
      l.states = 0;
      B200DF_CUDA(cudaHostAlloc(&l.hq, lane_bytes, cudaHostAllocDefault));
      B200DF_CUDA(cudaMalloc(&l.dq, lane_bytes));
      B200DF_CUDA(cudaMalloc(&l.amb, ((size_t)cs + 1) * sizeof(int)));
      l.states = (size_t)cs;
      l.row = row32;
    }
  }
  struct Result
  {
    int rc = B200DF_OK;
    cudaError_t e = cudaSuccess;
    std::string msg;
  };
  std::vector<Result> res((size_t)T);
  auto work = [&](int w) {
    Result& r = res[(size_t)w];
    r.e = cudaSetDevice(group->device);
    Ctx::Lane& l = ctx->lane[w];
    float* hq = static_cast<float*>(l.hq);
    uint8_t* hv = static_cast<uint8_t*>(l.hq) + v_at;
    uint8_t* dv = static_cast<uint8_t*>(l.dq) + v_at;
    for (int64_t i = w; i < n_chunks && !r.rc && r.e == cudaSuccess; i += T)
    {
      const int64_t s0 = i * cs, m = std::min<int64_t>(cs, n - s0);
      const double* src = q + (size_t)s0 * nv;
      const size_t count = (size_t)m * nv;
      narrow_rows(hq, src, count);
      r.e = cudaMemcpyAsync(l.dq, hq, count * sizeof(float), cudaMemcpyHostToDevice, l.st);
      if (r.e != cudaSuccess)
        break;
      r.rc = run_verdict(group, wf, sf, elem, l.dq, B200DF_Q_F32, m, flags, kPrecisionFastRaw, dv, l.st, l.amb + 1, l.amb);
      if (r.rc)
      {
        r.msg = b200df_last_error();
        break;
      }
      r.e = cudaMemcpyAsync(hv, dv, (size_t)m, cudaMemcpyDeviceToHost, l.st);
      if (r.e == cudaSuccess)
        r.e = cudaStreamSynchronize(l.st);
      if (r.e == cudaSuccess)
        std::memcpy(verdict + s0, hv, (size_t)m);
    }
  };
  {
    std::vector<std::thread> workers;
    for (int w = 1; w < T; ++w)
      workers.emplace_back(work, w);
    work(0);
    for (std::thread& t : workers)
      t.join();
  }
  for (const Result& r : res)
  {
    if (r.rc)
    {
      set_error("%s", r.msg.c_str());
      return r.rc;
    }
    B200DF_CUDA(r.e);
  }
  // the undecided states: their unrounded fp64 rows through the exact kernel
  std::vector<int64_t> idx;
  {
    const uint64_t* w8 = reinterpret_cast<const uint64_t*>(verdict);
    const int64_t head = ((uintptr_t)verdict & 7) ? std::min<int64_t>(n, 8 - (int64_t)((uintptr_t)verdict & 7)) : 0;
    for (int64_t i = 0; i < head; ++i)
      if (verdict[i] == 2)
        idx.push_back(i);
    w8 = reinterpret_cast<const uint64_t*>(verdict + head);
    const int64_t words = (n - head) / 8;
    for (int64_t k = 0; k < words; ++k)
      if (w8[k] & 0x0202020202020202ull)
        for (int64_t i = head + 8 * k; i < head + 8 * k + 8; ++i)
          if (verdict[i] == 2)
            idx.push_back(i);
    for (int64_t i = head + 8 * words; i < n; ++i)
      if (verdict[i] == 2)
        idx.push_back(i);
  }
  if (idx.empty())
    return B200DF_OK;
  const int64_t k = (int64_t)idx.size();
  const size_t row64 = (size_t)nv * sizeof(double);
  const size_t need = (size_t)k * row64 + (size_t)k;
  cudaError_t e = b200df_group::StagePool::ensure_hstage(ctx, 0, need);
  if (e == cudaSuccess)
    e = b200df_group::StagePool::ensure_scratch(ctx, 0, need);
  if (e == cudaSuccess)
    e = b200df_group::StagePool::ensure_stream(ctx, 0);
  B200DF_CUDA(e);
  char* hs = static_cast<char*>(ctx->hstage[0]);
  char* ds = static_cast<char*>(ctx->scratch[0]);
  for (int64_t j = 0; j < k; ++j)
    std::memcpy(hs + (size_t)j * row64, q + (size_t)idx[(size_t)j] * nv, row64);
  B200DF_CUDA(cudaMemcpyAsync(ds, hs, (size_t)k * row64, cudaMemcpyHostToDevice, ctx->stream[0]));
  uint8_t* dv = reinterpret_cast<uint8_t*>(ds + (size_t)k * row64);
  int rc = run_verdict(group, wf, sf, elem, ds, B200DF_Q_F64, k, flags, B200DF_PRECISION_EXACT, dv, ctx->stream[0]);
  if (rc)
    return rc;
  uint8_t* hv = reinterpret_cast<uint8_t*>(hs + (size_t)k * row64);
  B200DF_CUDA(cudaMemcpyAsync(hv, dv, (size_t)k, cudaMemcpyDeviceToHost, ctx->stream[0]));
  B200DF_CUDA(cudaStreamSynchronize(ctx->stream[0]));
  for (int64_t j = 0; j < k; ++j)
    verdict[idx[(size_t)j]] = hv[j];
  return B200DF_OK;
}

int b200df_check_batch(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q, int q_dtype,
                       int64_t n, int flags, int precision, uint8_t* verdict)
{
  B200DF_REQUIRE(group && n >= 0 && (n == 0 || (q && verdict)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown q dtype");
  B200DF_REQUIRE(precision == B200DF_PRECISION_EXACT || precision == B200DF_PRECISION_FAST || precision == kPrecisionV1 ||
                     precision == kPrecisionFastRaw || precision == kPrecisionSmall,
                 "unknown precision");
  if (n == 0)
    return B200DF_OK;
  FieldDev wf, sf;
  int elem;
  int rc = prepare_fields(group, world, self, flags, &wf, &sf, &elem);
  if (rc)
    return rc;
  DeviceGuard dg(group->device);
  const size_t row = (size_t)group->dev.n_vars * q_elem_size(q_dtype);
  const int64_t chunk = std::min<int64_t>(n, chunk_states());
  b200df_group::StagePool::Ctx* ctx = group->stage->acquire();
  cudaError_t e = cudaSuccess;
  if (n <= kSmallStates && (size_t)n * row <= kSmallQBytes && small_batch_enabled() &&
      (precision == B200DF_PRECISION_EXACT || precision == B200DF_PRECISION_FAST))
  {
    // one launch of the small-batch exact kernel (FAST returns the same verdicts by construction; its screening pass
    // only pays off on batches that fill the GPU)
    e = b200df_group::StagePool::ensure_stream(ctx, 0);
    void* dpin = nullptr;
    if (e == cudaSuccess)
      e = b200df_group::StagePool::ensure_pin(ctx, kSmallQBytes + (size_t)kSmallStates, &dpin);
    if (e == cudaSuccess)
    {
      std::memcpy(ctx->pin, q, (size_t)n * row);
      uint8_t* hv = static_cast<uint8_t*>(ctx->pin) + kSmallQBytes;
      uint8_t* dv = static_cast<uint8_t*>(dpin) + kSmallQBytes;
      rc = run_verdict(group, wf, sf, elem, dpin, q_dtype, n, flags, precision, dv, ctx->stream[0]);
      e = cudaStreamSynchronize(ctx->stream[0]);
      if (!rc && e == cudaSuccess)
        std::memcpy(verdict, hv, (size_t)n);
    }
    group->stage->give_back(ctx);
    if (!rc && e != cudaSuccess)
    {
      set_error("small-batch check failed: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
    return rc;
  }
  if (precision == B200DF_PRECISION_FAST && q_dtype == B200DF_Q_F64 && n >= kFastMinStates && is_pageable_host(q) &&
      !getenv("B200DF_NO_NARROWING"))
  {
    FieldDev wn = wf, sn = sf;
    if (group->neg_irrelevant)
      wn.neg = sn.neg = nullptr;
    if (fast_supported(group, wn, sn, B200DF_Q_F32, flags))
    {
      rc = check_batch_narrowed(group, wf, sf, elem, static_cast<const double*>(q), n, flags, verdict, ctx);
      group->stage->give_back(ctx);
      return rc;
    }
  }
  // Pageable caller memory: a cudaMemcpyAsync on it goes through the driver's own staging at a few GB/s and holds this
  // thread (the D2H even until the chunk's kernel is done, which serialises the pipeline).  Such batches are staged
  // through page-locked buffers of the context instead: a few host threads fill slot k while the GPU works on the
  // other slots, and the verdicts of a chunk are handed over when its slot comes round again.
  if ((size_t)n * row >= (size_t(4) << 20) && (is_pageable_host(q) || is_pageable_host(verdict)))
  {
    const int64_t cs = std::min<int64_t>(n, 1 << 18);
    const size_t q_at = 0, v_at = ((size_t)cs * row + 255) & ~size_t(255);
    rc = ensure_ctx(ctx, (size_t)cs * row, (size_t)cs);
    for (int i = 0; i < kSlots && !rc && e == cudaSuccess; ++i)
      e = b200df_group::StagePool::ensure_hstage(ctx, i, v_at + (size_t)cs);
    const int64_t n_chunks = (n + cs - 1) / cs;
    auto drain = [&](int64_t i) {  // chunk i is done: its verdicts leave the staging buffer
      const int slot = (int)(i % kSlots);
      const cudaError_t es = cudaStreamSynchronize(ctx->stream[slot]);
      if (e == cudaSuccess)
        e = es;
      if (e == cudaSuccess)
        std::memcpy(verdict + i * cs, static_cast<const char*>(ctx->hstage[slot]) + v_at,
                    (size_t)std::min<int64_t>(cs, n - i * cs));
    };
    int64_t drained = 0;
    for (int64_t i = 0; i < n_chunks && !rc && e == cudaSuccess; ++i)
    {
      const int slot = (int)(i % kSlots);
      if (i >= kSlots)
        drain(drained++);  // == i - kSlots: frees this slot's staging buffer
      if (e != cudaSuccess)
        break;
      const int64_t s0 = i * cs, m = std::min<int64_t>(cs, n - s0);
      char* hs = static_cast<char*>(ctx->hstage[slot]);
      parallel_memcpy(hs + q_at, static_cast<const char*>(q) + (size_t)s0 * row, (size_t)m * row);
      cudaStream_t st = ctx->stream[slot];
      e = cudaMemcpyAsync(ctx->q[slot], hs + q_at, (size_t)m * row, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess)
        break;
      rc = run_verdict(group, wf, sf, elem, ctx->q[slot], q_dtype, m, flags, precision, ctx->v[slot], st,
                       ctx->amb[slot] + 1, ctx->amb[slot]);
      if (rc)
        break;
      e = cudaMemcpyAsync(hs + v_at, ctx->v[slot], (size_t)m, cudaMemcpyDeviceToHost, st);
    }
    while (!rc && e == cudaSuccess && drained < n_chunks)
      drain(drained++);
    for (int i = 0; i < kSlots; ++i)
      if (ctx->stream[i])
      {
        const cudaError_t es = cudaStreamSynchronize(ctx->stream[i]);
        if (e == cudaSuccess)
          e = es;
      }
    group->stage->give_back(ctx);
    if (!rc && e != cudaSuccess)
    {
      set_error("batched check failed while staging / running / reading back: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
    return rc;
  }
  rc = ensure_ctx(ctx, std::max<size_t>(1, (size_t)chunk * row), (size_t)chunk);
  int k = 0;
  for (int64_t s = 0; s < n && !rc && e == cudaSuccess; s += chunk, k = (k + 1) % kSlots)
  {
    const int64_t m = std::min<int64_t>(chunk, n - s);
    cudaStream_t st = ctx->stream[k];
    // the stream order protects the buffers of slot k: this H2D runs after the D2H of the chunk before last
    e = cudaMemcpyAsync(ctx->q[k], static_cast<const char*>(q) + (size_t)s * row, (size_t)m * row,
                        cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess)
      break;
    rc = run_verdict(group, wf, sf, elem, ctx->q[k], q_dtype, m, flags, precision, ctx->v[k], st, ctx->amb[k] + 1,
                     ctx->amb[k]);
    if (rc)
      break;
    e = cudaMemcpyAsync(verdict + s, ctx->v[k], (size_t)m, cudaMemcpyDeviceToHost, st);
  }
  for (int i = 0; i < kSlots; ++i)
  {
    const cudaError_t es = cudaStreamSynchronize(ctx->stream[i]);
    if (e == cudaSuccess)
      e = es;
  }
  group->stage->give_back(ctx);
  if (!rc && e != cudaSuccess)
  {
    set_error("batched check failed while staging / running / reading back: %s", cudaGetErrorString(e));
    rc = B200DF_ERR_CUDA;
  }
  return rc;
}

}  // extern "C"

namespace b200df
{
// keeps the stream-ordered allocator from returning memory to the OS between calls (once per device)
void keep_async_pool(int device)
{
  static std::mutex mu;
  static std::vector<char> done;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)done.size() <= device)
    done.resize(device + 1, 0);
  if (done[device])
    return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
  {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[device] = 1;
}
}  // namespace b200df
