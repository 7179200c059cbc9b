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
#include <cfloat>
#include <cstring>

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
  const uint8_t kind = angular[v];
  if (kind == 3)
    return;  // a quaternion component: written by the thread of the quaternion's first variable
  if (kind == 2)
  {
    // FloatingJointModel::interpolate, rotation part (floating_joint_model.cpp:137-157): slerp on |dot| (the
    // reference's long-angle branch tests the absolute value for < 0 and is dead; reproduced by leaving it out)
    const double* f = qa + e * n_vars + v;
    const double* g = qb + e * n_vars + v;
    const double dq = fabs(((f[0] * g[0] + f[1] * g[1]) + f[2] * g[2]) + f[3] * g[3]);
    const double theta = (dq + DBL_EPSILON >= 1.0) ? 0.0 : acos(dq);
    if (theta > DBL_EPSILON)
    {
      const double d = 1.0 / sin(theta);
      const double s0 = sin((1.0 - t) * theta);
      const double s1 = sin(t * theta);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        out[i + c] = (QT)((f[c] * s0 + g[c] * s1) * d);
    }
    else
    {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        out[i + c] = (QT)f[c];
    }
    return;
  }
  const double from = qa[e * n_vars + v], to = qb[e * n_vars + v];
  double r;
  double diff = to - from;
  if (kind != 1 || fabs(diff) <= kPi)
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

constexpr size_t kSmallEdgeBytes = 256 * 1024;  // end points + results of a call that takes the mapped pinned staging

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
  typedef b200df_group::StagePool Pool;
  // the interpolated states stay fp64 (what the reference would check); FAST screens them in fp32 and re-runs the
  // undecided ones on the same fp64 values
  const int64_t edges_per_chunk = std::max<int64_t>(1, (int64_t)(1 << 21) / segments);
  const int64_t chunk = std::min<int64_t>(m, edges_per_chunk);
  const size_t end_bytes = (size_t)chunk * nv * sizeof(double);
  std::vector<uint8_t> ang(std::max(nv, 1), 0);
  if (var_is_angular)
  {
    // 0 linear, 1 shortest arc, 2 followed by exactly three 3s = the x y z w variables of one floating joint
    for (int v = 0; v < nv; ++v)
    {
      const int k = var_is_angular[v];
      B200DF_REQUIRE(k >= 0 && k <= 3, "var_is_angular values must be 0..3");
      if (k == 2)
      {
        B200DF_REQUIRE(v + 3 < nv && var_is_angular[v + 1] == 3 && var_is_angular[v + 2] == 3 &&
                           var_is_angular[v + 3] == 3 && (v + 4 >= nv || var_is_angular[v + 4] != 3),
                       "a quaternion is marked 2 3 3 3 (four variables inside the row)");
        v += 3;
      }
      else
        B200DF_REQUIRE(k != 3, "a 3 in var_is_angular must follow a 2 (quaternion x marks the start)");
    }
    std::copy(var_is_angular, var_is_angular + nv, ang.begin());
  }
  Pool::Ctx* ctx = group->stage->acquire();
  cudaError_t e = Pool::ensure_stream(ctx, 0);
  int rc = B200DF_OK;
  auto finish = [&]() {
    group->stage->give_back(ctx);
    if (!rc && e != cudaSuccess)
    {
      set_error("edge check failed while staging / running / reading back: %s", cudaGetErrorString(e));
      rc = B200DF_ERR_CUDA;
    }
    return rc;
  };
  if (e != cudaSuccess)
    return finish();
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  // scratch slots: two pipeline slots of {states, verdicts, A, B, out}, then the angular flags
  enum { kStates = 0, kVerdict = 1, kA = 2, kB = 3, kOut = 4, kPerSlot = 5, kAng = 10 };
  auto ensure_slot = [&](int k) {
    cudaError_t r = Pool::ensure_stream(ctx, k);
    if (r == cudaSuccess)
      r = Pool::ensure_scratch(ctx, kPerSlot * k + kStates, (size_t)chunk * segments * nv * sizeof(double));
    if (r == cudaSuccess)
      r = Pool::ensure_scratch(ctx, kPerSlot * k + kVerdict, (size_t)chunk * segments);
    return r;
  };
  e = ensure_slot(0);
  if (e != cudaSuccess)
    return finish();

  // interpolate -> verdicts -> first invalid sample, all on slot k's stream
  auto run_chunk = [&](int k, const double* da, const double* db, const uint8_t* dang, int64_t mc, int32_t* dout) {
    cudaStream_t stream = ctx->stream[k];
    double* dq = static_cast<double*>(ctx->scratch[kPerSlot * k + kStates]);
    uint8_t* dv = static_cast<uint8_t*>(ctx->scratch[kPerSlot * k + kVerdict]);
    const long total = mc * segments * nv;
    interpolate_edges_kernel<double><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(da, db, mc, segments, nv,
                                                                                          dang, dq);
    B200DF_LAUNCHED();
    rc = b200df_check_batch_device(group, world, self, dq, B200DF_Q_F64, mc * segments, flags, precision, dv, stream);
    if (rc)
      return;
    first_invalid_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, stream>>>(dv, mc, segments, dout);
    B200DF_LAUNCHED();
    e = cudaGetLastError();
  };

  // Small calls (a MotionValidator checks one edge at a time): end points in, first-invalid indices out through one
  // mapped pinned buffer - three launches and one synchronise, no staged copies.
  const size_t o_b = up(end_bytes), o_ang = 2 * o_b, o_out = o_ang + up(ang.size()),
               total_pin = o_out + up((size_t)m * sizeof(int32_t));
  if (m == chunk && total_pin <= kSmallEdgeBytes && small_batch_enabled())
  {
    void* dpin = nullptr;
    e = Pool::ensure_pin(ctx, total_pin, &dpin);
    if (e != cudaSuccess)
      return finish();
    char* hp = static_cast<char*>(ctx->pin);
    char* dp = static_cast<char*>(dpin);
    std::memcpy(hp, qa, end_bytes);
    std::memcpy(hp + o_b, qb, end_bytes);
    std::memcpy(hp + o_ang, ang.data(), ang.size());
    run_chunk(0, reinterpret_cast<const double*>(dp), reinterpret_cast<const double*>(dp + o_b),
              reinterpret_cast<const uint8_t*>(dp + o_ang), m, reinterpret_cast<int32_t*>(dp + o_out));
    const cudaError_t es = cudaStreamSynchronize(ctx->stream[0]);
    if (e == cudaSuccess)
      e = es;
    if (!rc && e == cudaSuccess)
      std::memcpy(first_invalid, hp + o_out, (size_t)m * sizeof(int32_t));
    return finish();
  }

  // Large calls: two slots on two streams, software-pipelined so that the end points of chunk i+1 are staged and its
  // kernels queued while chunk i computes; only then is chunk i's result read back (a copy into pageable memory
  // holds the host until that chunk is done).
  const int64_t n_chunks = (m + chunk - 1) / chunk;
  const int n_slots = n_chunks > 1 ? 2 : 1;
  for (int k = 0; k < n_slots && e == cudaSuccess; ++k)
  {
    e = ensure_slot(k);
    if (e == cudaSuccess)
      e = Pool::ensure_scratch(ctx, kPerSlot * k + kA, end_bytes);
    if (e == cudaSuccess)
      e = Pool::ensure_scratch(ctx, kPerSlot * k + kB, end_bytes);
    if (e == cudaSuccess)
      e = Pool::ensure_scratch(ctx, kPerSlot * k + kOut, (size_t)chunk * sizeof(int32_t));
  }
  if (e == cudaSuccess)
    e = Pool::ensure_scratch(ctx, kAng, ang.size());
  if (e == cudaSuccess)
    e = cudaMemcpy(ctx->scratch[kAng], ang.data(), ang.size(), cudaMemcpyHostToDevice);
  // end points in pageable memory are staged through page-locked buffers of the context by a few host threads (the
  // driver's own staging of a pageable H2D moves ~10 GB/s and holds this thread)
  const bool stage_in = end_bytes >= (size_t(2) << 20) && (is_pageable_host(qa) || is_pageable_host(qb));
  for (int k = 0; k < n_slots && stage_in && e == cudaSuccess; ++k)
    e = Pool::ensure_hstage(ctx, k, 2 * up(end_bytes));
  auto issue = [&](int64_t i) {  // stage + queue chunk i on slot i % 2
    const int k = (int)(i % 2);
    const int64_t s0 = i * chunk, mc = std::min<int64_t>(chunk, m - s0);
    void** buf = ctx->scratch + kPerSlot * k;
    const size_t bytes = (size_t)mc * nv * sizeof(double);
    const void *src_a = qa + s0 * nv, *src_b = qb + s0 * nv;
    if (stage_in)
    {
      char* hs = static_cast<char*>(ctx->hstage[k]);  // free: chunk i - 2 was drained before this call
      parallel_memcpy(hs, src_a, bytes);
      parallel_memcpy(hs + up(end_bytes), src_b, bytes);
      src_a = hs;
      src_b = hs + up(end_bytes);
    }
    e = cudaMemcpyAsync(buf[kA], src_a, bytes, cudaMemcpyHostToDevice, ctx->stream[k]);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(buf[kB], src_b, bytes, cudaMemcpyHostToDevice, ctx->stream[k]);
    if (e == cudaSuccess)
      run_chunk(k, static_cast<const double*>(buf[kA]), static_cast<const double*>(buf[kB]),
                static_cast<const uint8_t*>(ctx->scratch[kAng]), mc, static_cast<int32_t*>(buf[kOut]));
  };
  if (e == cudaSuccess)
    issue(0);
  for (int64_t i = 0; i < n_chunks && !rc && e == cudaSuccess; ++i)
  {
    if (i + 1 < n_chunks)
      issue(i + 1);  // slot (i+1) % 2 was drained when chunk i-1 was read back
    if (rc || e != cudaSuccess)
      break;
    const int k = (int)(i % 2);
    const int64_t s0 = i * chunk, mc = std::min<int64_t>(chunk, m - s0);
    e = cudaMemcpyAsync(first_invalid + s0, ctx->scratch[kPerSlot * k + kOut], (size_t)mc * sizeof(int32_t),
                        cudaMemcpyDeviceToHost, ctx->stream[k]);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(ctx->stream[k]);
  }
  for (int k = 0; k < n_slots; ++k)
  {
    const cudaError_t es = cudaStreamSynchronize(ctx->stream[k]);
    if (e == cudaSuccess)
      e = es;
  }
  return finish();
}

}  // extern "C"
