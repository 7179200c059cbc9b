// Small-batch state validity: one CTA per state instead of one thread per state.
//
// The per-state callers of the reference (OMPL's StateValidityChecker::isValid, PlanningScene::isStateColliding,
// checkCollision on one RobotState) hand over one state - or a short path - at a time.  The batched kernel
// (validity.cu) walks a state on a single thread: right for a million states, but a lone warp of it takes ~85 us on a
// Panda because ~70 k dependent fp64 instructions run back to back.  Here the same arithmetic is spread over the
// threads of a CTA so that the critical path is the FK chain only:
//
//   prologue  thread i  : joint transform of link i (*JointModel::computeTransform, one sincos deep for all joints)
//   chain     12 lanes  : T_link = (T_parent * T_origin) * T_joint, one matrix entry per lane, links in DFS order
//                         (RobotState::updateLinkTransformsInternal, robot_state.cpp:718-755; every entry is the same
//                         left-to-right sum of products a scalar thread would evaluate => bit-identical transforms)
//   spheres   thread s  : pose sphere s, look it up in the world / self field (types.cpp:206-231, 390-408)
//   pairs     thread it : one (own sphere, lower partner link) item of getIntraGroupCollisions (:433-660): bounding
//                         sphere prune (Q2) then the partner's spheres against the host-resolved thresholds
//
// The verdict is a logical OR of the same predicates in the same arithmetic as validity.cu, so the two kernels agree
// bit for bit (tests/test_gpu_parity.py::test_small_batch_kernel_*).  Compiled with --fmad=false like validity.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "validity_common.cuh"

using namespace b200df;

namespace
{
struct SmallSphere
{
  int link;          // model link whose transform poses this sphere
  int self_enabled;  // the link's self-field check is on (self_collision_enabled_)
};

struct SmallItem  // one own sphere of the higher-index link against one enabled lower-index partner link
{
  int own;        // sphere index (group order)
  int up_gl;      // group-link slot of the own sphere's link
  int lo_gl;      // group-link slot of the partner
  int lo_begin;   // the partner's first sphere
  int lo_count;   // its sphere count
  int pad;
  long thr;       // offset of the thresholds of (own, partner sphere 0) in GroupDev::pair_T
  double radius_sum;  // doBoundingSpheresIntersect's un-squared sum (Q2)
  double tight2;      // exact cull, see DevPartner
};

struct SmallDev
{
  const SmallSphere* spheres;
  const SmallItem* items;
  int n_items;
};

template <typename QT, typename FT>
__global__ void __launch_bounds__(256) validity_small_kernel(GroupDev g, SmallDev t, LookupDev wf, LookupDev sf,
                                                             const QT* __restrict__ q, long n, int flags,
                                                             uint8_t* __restrict__ verdict,
                                                             const int* __restrict__ redo_idx,
                                                             const int* __restrict__ redo_count)
{
  extern __shared__ double sm[];
  const int L = g.n_links, S = g.n_spheres, G = g.n_glinks;
  double* T = sm;                                      // [L][12] global link transforms
  double* C = T + L * 12;                              // [S][3] posed sphere centres
  double* BC = C + S * 3;                              // [G][3] posed bounding-sphere centres
  double* staging = BC + G * 3;                        // warp_fk's joint transforms / origins / parents
  const int tid = threadIdx.x, nt = blockDim.x;
  const bool do_robot = (flags & B200DF_CHECK_ROBOT) != 0;
  const bool do_self = (flags & B200DF_CHECK_SELF) != 0;
  const bool do_pairs = (flags & B200DF_CHECK_INTRA) != 0 && t.n_items > 0;

  // redo_idx / redo_count: when set, only the listed states are evaluated (the FAST path's undecided set, whose length
  // lives in device memory; a few hundred states per million - far too few for the one-thread-per-state kernel)
  const long total = redo_idx ? (long)*redo_count : n;
  for (long item = blockIdx.x; item < total; item += gridDim.x)
  {
    const long state = redo_idx ? (long)redo_idx[item] : item;
    if (tid < 32)
      warp_fk(g.links, L, q + state * g.n_vars, tid, staging, T);
    __syncthreads();

    bool hit = false;
    for (int s = tid; s < S; s += nt)
    {
      const SmallSphere ss = t.spheres[s];
      const double* rel = g.sph + 4 * s;
      double cx, cy, cz;
      iso_apply(T + ss.link * 12, rel[0], rel[1], rel[2], cx, cy, cz);  // updatePose (types.cpp:393-396)
      C[s * 3 + 0] = cx;
      C[s * 3 + 1] = cy;
      C[s * 3 + 2] = cz;
      const double radius = rel[3];
      const int thr = g.sph_thr[s];
      if (do_robot)
      {
        bool in;
        const unsigned idx = cell_index(wf, cx, cy, cz, in);
        B200DF_CHECK_INDEX(idx, (long long)(wf.lim[0] + 2) * wf.stride1);
        const int d2 = (int)static_cast<const FT*>(wf.d2)[idx];
        hit |= voxel_hits<FT>(wf, idx, in, d2, radius, thr, g.tol, g.max_prop);
      }
      if (do_self && ss.self_enabled)
      {
        bool in;
        const unsigned idx = cell_index(sf, cx, cy, cz, in);
        B200DF_CHECK_INDEX(idx, (long long)(sf.lim[0] + 2) * sf.stride1);
        const int d2 = (int)static_cast<const FT*>(sf.d2)[idx];
        hit |= voxel_hits<FT>(sf, idx, in, d2, radius, thr, g.tol, g.max_prop);
      }
    }
    if (do_pairs)
      for (int gs = tid; gs < G; gs += nt)
      {
        const DevGroupLink& gl = g.glinks[gs];
        double bx, by, bz;
        iso_apply(T + gl.link * 12, gl.bs[0], gl.bs[1], gl.bs[2], bx, by, bz);  // types.cpp:398-399
        BC[gs * 3 + 0] = bx;
        BC[gs * 3 + 1] = by;
        BC[gs * 3 + 2] = bz;
      }
    int any = __syncthreads_or(hit ? 1 : 0);
    if (!any && do_pairs)
    {
      for (int it = tid; it < t.n_items; it += nt)
      {
        const SmallItem item = t.items[it];
        const double dx = BC[item.lo_gl * 3 + 0] - BC[item.up_gl * 3 + 0];
        const double dy = BC[item.lo_gl * 3 + 1] - BC[item.up_gl * 3 + 1];
        const double dz = BC[item.lo_gl * 3 + 2] - BC[item.up_gl * 3 + 2];
        const double d2 = (dx * dx + dy * dy) + dz * dz;
        if (!((d2 < item.radius_sum) && (d2 < item.tight2)))  // doBoundingSpheresIntersect (Q2) + exact cull
          continue;
        const double cx = C[item.own * 3 + 0], cy = C[item.own * 3 + 1], cz = C[item.own * 3 + 2];
        const double* thr = g.pair_T + item.thr;
        const double* pc = C + item.lo_begin * 3;
        for (int m = 0; m < item.lo_count; ++m)
        {
          const double ex = cx - pc[m * 3 + 0], ey = cy - pc[m * 3 + 1], ez = cz - pc[m * 3 + 2];
          const double e2 = (ex * ex + ey * ey) + ez * ez;
          hit |= e2 < thr[m];  // ||c1-c2|| < r1+r2 (:574-584) through the bit-equivalent squared threshold
        }
      }
      any = __syncthreads_or(hit ? 1 : 0);
    }
    if (tid == 0)
      verdict[state] = any ? 1 : 0;
  }
}

template <typename T>
int upload_table(std::vector<void*>& owned, const std::vector<T>& host, const T** dev)
{
  void* p = nullptr;
  B200DF_CUDA(cudaMalloc(&p, std::max<size_t>(1, host.size()) * sizeof(T)));
  owned.push_back(p);
  if (!host.empty())
    B200DF_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev = static_cast<const T*>(p);
  return B200DF_OK;
}
}  // namespace

struct b200df_group::SmallPath
{
  SmallDev dev{};
  std::vector<void*> owned;
  int threads = 32;
  size_t smem = 0;
};

namespace
{
template <typename QT, typename FT>
int launch_small(const b200df_group* g, const FieldDev& wf, const FieldDev& sf, const void* q, long n, int flags,
                 uint8_t* verdict, cudaStream_t stream, const int* redo_idx, const int* redo_count)
{
  const b200df_group::SmallPath& sp = *g->small;
  B200DF_CUDA(allow_dynamic_smem(validity_small_kernel<QT, FT>, sp.smem));
  const unsigned grid = (unsigned)std::min<long>(n, redo_idx ? 148L * 32 : 148L * 64);
  validity_small_kernel<QT, FT><<<grid, sp.threads, sp.smem, stream>>>(g->dev, sp.dev, make_lookup(wf), make_lookup(sf),
                                                                      static_cast<const QT*>(q), n, flags, verdict,
                                                                      redo_idx, redo_count);
  B200DF_LAUNCHED();
  B200DF_CUDA(cudaGetLastError());
  return B200DF_OK;
}
}  // namespace

namespace b200df
{
int small_create(b200df_group* g)
{
  std::unique_ptr<b200df_group::SmallPath> sp(new b200df_group::SmallPath);
  const int G = (int)g->glinks.size(), S = (int)(g->sph.size() / 4), L = g->model->n_links;
  std::vector<SmallSphere> spheres(S);
  std::vector<int> by_store_link(G, -1);
  for (int gs = 0; gs < G; ++gs)
  {
    const DevGroupLink& gl = g->glinks[gs];
    for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
      spheres[s] = SmallSphere{ gl.link, gl.self_enabled };
    if (gl.store_link >= 0)
      by_store_link[gl.store_link] = gs;
  }
  std::vector<SmallItem> items;
  for (int gs = 0; gs < G; ++gs)
  {
    const DevGroupLink& gl = g->glinks[gs];
    for (int s = gl.sphere_begin; s < gl.sphere_end; ++s)
      for (int p = gl.partner_begin; p < gl.partner_end; ++p)
      {
        const DevPartner& pr = g->partners[p];
        const int lo = by_store_link[pr.store_link];
        B200DF_REQUIRE(lo >= 0 && g->glinks[lo].sphere_end - g->glinks[lo].sphere_begin == pr.count,
                       "inconsistent partner tables");
        SmallItem it{};
        it.own = s;
        it.up_gl = gs;
        it.lo_gl = lo;
        it.lo_begin = g->glinks[lo].sphere_begin;
        it.lo_count = pr.count;
        it.thr = (long)gl.pair_const_begin + (long)(s - gl.sphere_begin) * gl.pair_const_stride + pr.const_offset;
        it.radius_sum = pr.radius_sum;
        it.tight2 = pr.tight2;
        items.push_back(it);
      }
  }
  DeviceGuard dg(g->device);
  int rc;
  if ((rc = upload_table(sp->owned, spheres, &sp->dev.spheres)))
    return rc;
  if ((rc = upload_table(sp->owned, items, &sp->dev.items)))
    return rc;
  sp->dev.n_items = (int)items.size();
  const int want = std::max(std::max(S, L), ((int)items.size() + 1) / 2);
  sp->threads = std::min(256, std::max(32, (want + 31) / 32 * 32));
  sp->smem = ((size_t)L * 12 + (size_t)S * 3 + (size_t)G * 3 + warp_fk_staging_doubles(L)) * sizeof(double);
  if (sp->smem > 200 * 1024)
    return B200DF_OK;  // no small-batch kernel for this robot: callers fall back to the batched kernel
  g->small = sp.release();
  return B200DF_OK;
}

void small_destroy(b200df_group* g)
{
  if (!g->small)
    return;
  for (void* p : g->small->owned)
    cudaFree(p);
  delete g->small;
  g->small = nullptr;
}

bool small_supported(const b200df_group* g)
{
  return g->small != nullptr;
}

int small_launch(const b200df_group* g, const FieldDev& wf, const FieldDev& sf, int field_elem, const void* q,
                 int q_dtype, long n, int flags, uint8_t* verdict, cudaStream_t stream, const int* redo_idx,
                 const int* redo_count)
{
  B200DF_REQUIRE(g->small, "this group has no small-batch kernel");
  if (q_dtype == B200DF_Q_F32)
    return field_elem == 1
               ? launch_small<float, uint8_t>(g, wf, sf, q, n, flags, verdict, stream, redo_idx, redo_count)
               : launch_small<float, uint16_t>(g, wf, sf, q, n, flags, verdict, stream, redo_idx, redo_count);
  return field_elem == 1
             ? launch_small<double, uint8_t>(g, wf, sf, q, n, flags, verdict, stream, redo_idx, redo_count)
             : launch_small<double, uint16_t>(g, wf, sf, q, n, flags, verdict, stream, redo_idx, redo_count);
}
}  // namespace b200df
