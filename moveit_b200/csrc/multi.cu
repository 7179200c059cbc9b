// Single-process multi-GPU: one (model, group) replicated on several devices of this process, fields replicated by peer
// copy, a batch cut into contiguous shards - one per device, each driven by its own host thread - and the verdicts
// written into the caller's single buffer.
//
// MoveIt is one process: planner threads share one const PlanningScene
// (moveit_ros/planning/planning_components_tools/src/evaluate_collision_checking_speed.cpp:44-57,121-124;
// planning_scene/test/test_multi_threaded.cpp), so the 8 GPUs of a box have to be reachable from one CollisionEnv
// without torchrun.  The state batch is the only thing that is sharded; the one exchange of the path is the field
// replication after a rebuild (cudaMemcpyPeerAsync over NVLink, b200df_field_copy_from) - no collective.
#include <algorithm>
#include <memory>
#include <thread>
#include <vector>

#include "b200df_internal.cuh"

using namespace b200df;

struct b200df_multi
{
  std::vector<int> devices;
  std::vector<b200df_model*> models;
  std::vector<b200df_group*> groups;
  // replicas of the fields (owned); slot 0 = world, 1 = self
  std::vector<b200df_field*> fields[2];
  int n_vars = 0;
  b200df_field_desc field_desc[2];
  bool have_field[2] = { false, false };
  std::mutex mu;  // guards field replication against concurrent checks
};

namespace
{
// error strings are thread-local: carry a worker's message over to the calling thread
struct WorkerResult
{
  int rc = B200DF_OK;
  std::string error;
};

template <typename F>
int for_each_device(const b200df_multi* m, F&& f)
{
  const int n = (int)m->devices.size();
  std::vector<WorkerResult> res((size_t)n);
  auto run = [&](int i) {
    ThreadDeviceScope scope(m->devices[(size_t)i]);
    int rc = B200DF_OK;
    if (scope.status != cudaSuccess)
    {
      set_error("cudaSetDevice(%d) failed: %s", m->devices[(size_t)i], cudaGetErrorString(scope.status));
      rc = B200DF_ERR_CUDA;
    }
    else
      rc = f(i);
    res[(size_t)i].rc = rc;
    if (rc)
      res[(size_t)i].error = b200df_last_error();
  };
  std::vector<std::thread> workers;
  for (int i = 1; i < n; ++i)
    workers.emplace_back(run, i);
  run(0);
  for (std::thread& t : workers)
    t.join();
  for (int i = 0; i < n; ++i)
    if (res[(size_t)i].rc)
    {
      set_error("device %d: %s", m->devices[(size_t)i], res[(size_t)i].error.c_str());
      return res[(size_t)i].rc;
    }
  return B200DF_OK;
}
}  // namespace

extern "C" {

int b200df_multi_destroy(b200df_multi* m)
{
  if (!m)
    return B200DF_OK;
  for (int s = 0; s < 2; ++s)
    for (b200df_field* f : m->fields[s])
      b200df_field_destroy(f);
  for (b200df_group* g : m->groups)
    b200df_group_destroy(g);
  for (b200df_model* mo : m->models)
    b200df_model_destroy(mo);
  delete m;
  return B200DF_OK;
}

int b200df_multi_create(const b200df_model_desc* model, const b200df_group_desc* group, const int32_t* devices,
                        int32_t n_devices, b200df_multi** out)
{
  B200DF_REQUIRE(model && group && devices && out && n_devices >= 1, "bad arguments");
  int n_dev = 0;
  B200DF_CUDA(cudaGetDeviceCount(&n_dev));
  for (int i = 0; i < n_devices; ++i)
  {
    B200DF_REQUIRE(devices[i] >= 0 && devices[i] < n_dev, "device index out of range");
    for (int j = 0; j < i; ++j)
      B200DF_REQUIRE(devices[i] != devices[j], "a device is listed twice");
  }
  struct Destroy
  {
    void operator()(b200df_multi* p) const
    {
      b200df_multi_destroy(p);
    }
  };
  std::unique_ptr<b200df_multi, Destroy> m(new b200df_multi);
  m->devices.assign(devices, devices + n_devices);
  m->models.assign((size_t)n_devices, nullptr);
  m->groups.assign((size_t)n_devices, nullptr);
  m->n_vars = model->n_vars;
  int rc = B200DF_OK;
  for (int i = 0; i < n_devices && !rc; ++i)
  {
    ThreadDeviceScope scope(devices[i]);
    B200DF_CUDA(scope.status);
    // direct NVLink copies between the replicas (ignore "already enabled" / unsupported pairs: the copy then stages)
    for (int j = 0; j < n_devices && !rc; ++j)
      if (j != i)
      {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can)
          if (cudaDeviceEnablePeerAccess(devices[j], 0) != cudaSuccess)
            cudaGetLastError();
      }
    if (!rc)
      rc = b200df_model_create(model, &m->models[(size_t)i]);
    if (!rc)
      rc = b200df_group_create(m->models[(size_t)i], group, &m->groups[(size_t)i]);
  }
  if (rc)
    return rc;
  *out = m.release();
  return B200DF_OK;
}

int b200df_multi_device_count(const b200df_multi* m, int32_t* n)
{
  B200DF_REQUIRE(m && n, "null argument");
  *n = (int32_t)m->devices.size();
  return B200DF_OK;
}

int b200df_multi_set_field(b200df_multi* m, int which, b200df_field* src)
{
  B200DF_REQUIRE(m && (which == 0 || which == 1), "bad arguments");
  std::lock_guard<std::mutex> lk(m->mu);
  if (!src)
  {
    for (b200df_field* f : m->fields[which])
      b200df_field_destroy(f);
    m->fields[which].clear();
    m->have_field[which] = false;
    return B200DF_OK;
  }
  b200df_field_desc d;
  int rc = b200df_field_get_desc(src, &d);
  if (rc)
    return rc;
  const bool same = m->have_field[which] && std::equal(d.size, d.size + 3, m->field_desc[which].size) &&
                    std::equal(d.origin, d.origin + 3, m->field_desc[which].origin) &&
                    d.resolution == m->field_desc[which].resolution &&
                    d.max_distance == m->field_desc[which].max_distance &&
                    d.propagate_negative == m->field_desc[which].propagate_negative;
  if (!same)
  {
    for (b200df_field* f : m->fields[which])
      b200df_field_destroy(f);
    m->fields[which].assign(m->devices.size(), nullptr);
    for (size_t i = 0; i < m->devices.size() && !rc; ++i)
    {
      ThreadDeviceScope scope(m->devices[i]);
      B200DF_CUDA(scope.status);
      rc = b200df_field_create(&d, &m->fields[which][i]);
    }
    if (rc)
      return rc;
    m->field_desc[which] = d;
    m->have_field[which] = true;
  }
  // build once on the source's device, then fan out: every replica pulls occupancy + distances over NVLink in its own
  // thread (b200df_field_copy_from rebuilds the source first if it is dirty; the source's lock serialises that)
  rc = b200df_field_rebuild(src, nullptr);
  if (rc)
    return rc;
  return for_each_device(m, [&](int i) { return b200df_field_copy_from(m->fields[which][(size_t)i], src); });
}

int b200df_multi_check_batch(const b200df_multi* m, const void* q, int q_dtype, int64_t n, int flags, int precision,
                             uint8_t* verdict)
{
  B200DF_REQUIRE(m && n >= 0 && (n == 0 || (q && verdict)), "bad arguments");
  B200DF_REQUIRE(q_dtype == B200DF_Q_F32 || q_dtype == B200DF_Q_F64, "unknown joint-value type");
  B200DF_REQUIRE(!(flags & B200DF_CHECK_ROBOT) || m->have_field[0], "B200DF_CHECK_ROBOT needs b200df_multi_set_field(0)");
  B200DF_REQUIRE(!(flags & B200DF_CHECK_SELF) || m->have_field[1], "B200DF_CHECK_SELF needs b200df_multi_set_field(1)");
  if (n == 0)
    return B200DF_OK;
  const int G = (int)m->devices.size();
  const int64_t per = (n + G - 1) / G;  // contiguous N/G shard per device
  const size_t row = (size_t)m->n_vars * (q_dtype == B200DF_Q_F32 ? sizeof(float) : sizeof(double));
  std::lock_guard<std::mutex> lk(const_cast<b200df_multi*>(m)->mu);
  return for_each_device(m, [&](int i) -> int {
    const int64_t a = std::min<int64_t>(n, (int64_t)i * per), b = std::min<int64_t>(n, a + per);
    if (a >= b)
      return B200DF_OK;
    return b200df_check_batch(m->groups[(size_t)i], m->have_field[0] ? m->fields[0][(size_t)i] : nullptr,
                              m->have_field[1] ? m->fields[1][(size_t)i] : nullptr,
                              static_cast<const char*>(q) + (size_t)a * row, q_dtype, b - a, flags, precision, verdict + a);
  });
}

int b200df_multi_check_batch_device(const b200df_multi* m, const void* const* q_dev, int q_dtype, const int64_t* n,
                                    int flags, int precision, uint8_t* const* verdict_dev, void* const* streams)
{
  B200DF_REQUIRE(m && q_dev && n && verdict_dev, "null argument");
  int rc = B200DF_OK;
  for (size_t i = 0; i < m->devices.size() && !rc; ++i)
  {
    if (n[i] <= 0)
      continue;
    ThreadDeviceScope scope(m->devices[i]);
    B200DF_CUDA(scope.status);
    rc = b200df_check_batch_device(m->groups[i], m->have_field[0] ? m->fields[0][i] : nullptr,
                                   m->have_field[1] ? m->fields[1][i] : nullptr, q_dev[i], q_dtype, n[i], flags,
                                   precision, verdict_dev[i], streams ? streams[i] : nullptr);
  }
  return rc;
}

int b200df_multi_field(const b200df_multi* m, int which, int32_t device_slot, b200df_field** out)
{
  B200DF_REQUIRE(m && out && (which == 0 || which == 1) && device_slot >= 0 &&
                     device_slot < (int32_t)m->devices.size() && m->have_field[which],
                 "bad arguments");
  *out = m->fields[which][(size_t)device_slot];
  return B200DF_OK;
}

}  // extern "C"
