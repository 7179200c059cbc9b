// Batched motion (edge) validation: M straight-line motions between joint-space states, each discretised into K
// segments and every interior + end state checked by the fused validity kernels — the batched form of what a
// sampling-based planner does per edge through ompl::base::DiscreteMotionValidator::checkMotion calling MoveIt's
// StateValidityChecker::isValid (moveit_planners/ompl/ompl_interface/src/detail/state_validity_checker.cpp:76-121),
// and of PlanningScene::isPathValid's waypoint loop (planning_scene/src/planning_scene.cpp:2257-2312).
//
// Only the 2*M end states cross PCIe; the M*K interpolated states are generated on the device:
//   RevoluteJointModel::interpolate   robot_model/src/revolute_joint_model.cpp:125-148  (continuous: shortest arc + wrap)
//   PrismaticJointModel / bounded revolute: from + (to - from) * t
//   PlanarJointModel::interpolate     robot_model/src/planar_joint_model.cpp:180-205    (holonomic: x, y linear, theta shortest arc)
// Compiled with --fmad=false: the interpolation rounds like the reference's double arithmetic.
#include <algorithm>

#include "validity_common.cuh"

using namespace b200df;

namespace
{
constexpr double kPi = 3.14159265358979323846;  // boost::math::constants::pi<double>()

// state j (1..K) of edge e: t = j / K
template <typename QT>
__global__ void interpolate_edges_kernel(const double* __restrict__ qa, const double* __restrict__ qb, long m, int K,
                                         int n_vars, const uint8_t* __restrict__ angular, QT* __restrict__ out)
{
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // one thread per (state, variable)
  const long total = m * K * n_vars;
  if (i >= total)
    return;
  const int v = (int)(i % n_vars);
  const long s = i / n_vars;
  const long e = s / K;
  const int j = (int)(s - e * K) + 1;
  const double t = (double)j / (double)K;
  const double from = qa[e * n_vars + v], to = qb[e * n_vars + v];
  double r;
  double diff = to - from;
  if (!angular[v] || fabs(diff) <= kPi)
    r = from + diff * t;
  else
  {
    if (diff > 0.0)
      diff = 2.0 * kPi - diff;
    else
      diff = -2.0 * kPi - diff;
    r = from - diff * t;
    if (r > kPi)
      r -= 2.0 * kPi;
    else if (r < -kPi)
      r += 2.0 * kPi;
  }
  out[i] = (QT)r;
}

// first colliding sample of each edge (1-based), 0 when the whole motion is valid
__global__ void first_invalid_kernel(const uint8_t* __restrict__ verdict, long m, int K, int32_t* __restrict__ out)
{
  const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (e >= m)
    return;
  int first = 0;
  for (int j = 0; j < K; ++j)
    if (verdict[e * K + j])
    {
      first = j + 1;
      break;
    }
  out[e] = first;
}
}  // namespace

extern "C" {

int b200df_check_edges(const b200df_group* group, b200df_field* world, b200df_field* self, const double* qa,
                       const double* qb, int64_t m, int32_t segments, const uint8_t* var_is_angular, int flags,
                       int precision, int32_t* first_invalid)
{
  B200DF_REQUIRE(group && m >= 0 && segments >= 1 && (m == 0 || (qa && qb && first_invalid)), "bad arguments");
  if (m == 0)
    return B200DF_OK;
  const int nv = group->dev.n_vars;
  DeviceGuard dg(group->device);
  cudaStream_t stream;
  B200DF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  // the interpolated states stay fp64 (what the reference would check); FAST screens them in fp32 and re-runs the
  // undecided ones on the same fp64 values
  const bool f32 = false;
  const size_t elem = sizeof(double);
  const int64_t edges_per_chunk = std::max<int64_t>(1, (int64_t)(1 << 21) / segments);
  const int64_t chunk = std::min<int64_t>(m, edges_per_chunk);
  int rc = B200DF_OK;
  {
    ScratchBuffer da, db, dang, dq, dv, dout;
    da.stream = db.stream = dang.stream = dq.stream = dv.stream = dout.stream = stream;
    std::vector<uint8_t> ang(nv, 0);
    if (var_is_angular)
      ang.assign(var_is_angular, var_is_angular + nv);
    cudaError_t e = cudaMallocAsync(&da.p, (size_t)chunk * nv * sizeof(double), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&db.p, (size_t)chunk * nv * sizeof(double), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dang.p, std::max(nv, 1), stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dq.p, (size_t)chunk * segments * nv * elem, stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dv.p, (size_t)chunk * segments, stream);
    if (e == cudaSuccess)
      e = cudaMallocAsync(&dout.p, (size_t)chunk * sizeof(int32_t), stream);
    if (e == cudaSuccess && nv > 0)
      e = cudaMemcpyAsync(dang.p, ang.data(), nv, cudaMemcpyHostToDevice, stream);
    for (int64_t s = 0; s < m && !rc && e == cudaSuccess; s += chunk)
    {
      const int64_t mc = std::min<int64_t>(chunk, m - s);
      e = cudaMemcpyAsync(da.p, qa + s * nv, (size_t)mc * nv * sizeof(double), cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(db.p, qb + s * nv, (size_t)mc * nv * sizeof(double), cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess)
        break;
      const long total = mc * segments * nv;
      const unsigned blocks = (unsigned)((total + 255) / 256);
      if (f32)
        interpolate_edges_kernel<float><<<blocks, 256, 0, stream>>>(
            static_cast<const double*>(da.p), static_cast<const double*>(db.p), mc, segments, nv,
            static_cast<const uint8_t*>(dang.p), static_cast<float*>(dq.p));
      else
        interpolate_edges_kernel<double><<<blocks, 256, 0, stream>>>(
            static_cast<const double*>(da.p), static_cast<const double*>(db.p), mc, segments, nv,
            static_cast<const uint8_t*>(dang.p), static_cast<double*>(dq.p));
      B200DF_LAUNCHED();
      rc = b200df_check_batch_device(group, world, self, dq.p, f32 ? B200DF_Q_F32 : B200DF_Q_F64, mc * segments, flags,
                                     precision, static_cast<uint8_t*>(dv.p), stream);
      if (rc)
        break;
      first_invalid_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, stream>>>(static_cast<const uint8_t*>(dv.p), mc,
                                                                             segments, static_cast<int32_t*>(dout.p));
      B200DF_LAUNCHED();
      e = cudaGetLastError();
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(first_invalid + s, dout.p, (size_t)mc * sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);  // the next chunk reuses the buffers
    }
    if (!rc && e != cudaSuccess)
    {
      set_error("edge check failed while staging / running / reading back: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
  }
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
  return rc;
}

}  // extern "C"
