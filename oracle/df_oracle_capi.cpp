// ORACLE — TEST INFRASTRUCTURE ONLY (see df_oracle.hpp).  Flat C API over the CPU restatement so that
// tests/ (ctypes) and bench.py's cpu_baseline / --impl reference legs can drive it.
#include "df_oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace orc;

namespace
{
struct ModelHandle
{
  std::shared_ptr<RobotModel> model;
};
struct EnvHandle
{
  std::shared_ptr<RobotModel> model;
  std::unique_ptr<CollisionEnvDistanceField> env;
};

Iso isoFrom12(const double* p)  // row-major 3x4
{
  Iso t;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      t.m[i][j] = p[4 * i + j];
  return t;
}

std::vector<V3> toPoints(const double* xyz, long n)
{
  std::vector<V3> pts((size_t)n);
  for (long i = 0; i < n; ++i)
    pts[i] = V3{ xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] };
  return pts;
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- fields
void* orc_field_create(double sx, double sy, double sz, double res, double ox, double oy, double oz, double max_dist,
                       int propagate_negative)
{
  return new PropagationDistanceField(sx, sy, sz, res, ox, oy, oz, max_dist, propagate_negative != 0);
}
void orc_field_destroy(void* f)
{
  delete static_cast<PropagationDistanceField*>(f);
}
void orc_field_dims(void* f, int* out4)
{
  auto* df = static_cast<PropagationDistanceField*>(f);
  out4[0] = df->g_.num_cells[0];
  out4[1] = df->g_.num_cells[1];
  out4[2] = df->g_.num_cells[2];
  out4[3] = df->max_distance_sq_;
}
void orc_field_add_points(void* f, const double* xyz, long n)
{
  static_cast<PropagationDistanceField*>(f)->addPointsToField(toPoints(xyz, n));
}
void orc_field_remove_points(void* f, const double* xyz, long n)
{
  static_cast<PropagationDistanceField*>(f)->removePointsFromField(toPoints(xyz, n));
}
void orc_field_update_points(void* f, const double* old_xyz, long n_old, const double* new_xyz, long n_new)
{
  static_cast<PropagationDistanceField*>(f)->updatePointsInField(toPoints(old_xyz, n_old), toPoints(new_xyz, n_new));
}
void orc_field_reset(void* f)
{
  static_cast<PropagationDistanceField*>(f)->reset();
}
void orc_field_get_d2(void* f, int32_t* out, int negative)
{
  auto* df = static_cast<PropagationDistanceField*>(f);
  for (size_t i = 0; i < df->cells_.size(); ++i)
    out[i] = negative ? df->cells_[i].negative_distance_square : df->cells_[i].distance_square;
}
double orc_field_distance_world(void* f, double x, double y, double z)
{
  return static_cast<PropagationDistanceField*>(f)->getDistance(x, y, z);
}
double orc_field_distance_cell(void* f, int x, int y, int z)
{
  return static_cast<PropagationDistanceField*>(f)->getDistance(x, y, z);
}
// out5 = dist, gx, gy, gz, in_bounds
void orc_field_distance_gradient(void* f, double x, double y, double z, double* out5)
{
  bool in_bounds;
  out5[0] = static_cast<PropagationDistanceField*>(f)->getDistanceGradient(x, y, z, out5[1], out5[2], out5[3], in_bounds);
  out5[4] = in_bounds ? 1.0 : 0.0;
}
int orc_field_nearest_cell(void* f, int x, int y, int z, double* dist, int* pos3)
{
  I3 pos;
  int r = static_cast<PropagationDistanceField*>(f)->getNearestCell(x, y, z, *dist, pos);
  pos3[0] = pos.x;
  pos3[1] = pos.y;
  pos3[2] = pos.z;
  return r;
}
int orc_field_world_to_grid(void* f, double x, double y, double z, int* out3)
{
  return static_cast<PropagationDistanceField*>(f)->g_.worldToGrid(x, y, z, out3[0], out3[1], out3[2]) ? 1 : 0;
}
void orc_field_grid_to_world(void* f, int x, int y, int z, double* out3)
{
  auto* df = static_cast<PropagationDistanceField*>(f);
  out3[0] = df->g_.getLocationFromCell(0, x);
  out3[1] = df->g_.getLocationFromCell(1, y);
  out3[2] = df->g_.getLocationFromCell(2, z);
}
double orc_field_uninitialized_distance(void* f)
{
  return static_cast<PropagationDistanceField*>(f)->getUninitializedDistance();
}
long orc_field_pack_occupancy(void* f, uint8_t* out, long cap)
{
  std::vector<uint8_t> v = static_cast<PropagationDistanceField*>(f)->packOccupancy();
  if ((long)v.size() <= cap)
    std::memcpy(out, v.data(), v.size());
  return (long)v.size();
}
void orc_field_unpack_occupancy(void* f, const uint8_t* bytes, long n)
{
  static_cast<PropagationDistanceField*>(f)->unpackOccupancy(bytes, (size_t)n);
}
// writeToStream into a caller buffer (two-pass: returns the byte count; copies when it fits)
long orc_field_write_stream(void* f, uint8_t* out, long cap)
{
  std::vector<uint8_t> s;
  if (!static_cast<PropagationDistanceField*>(f)->writeToStream(s))
    return -1;
  if ((long)s.size() <= cap)
    std::memcpy(out, s.data(), s.size());
  return (long)s.size();
}
// PropagationDistanceField(std::istream&, max_distance, propagate_negative) (:64-74); NULL when readFromStream fails
void* orc_field_from_stream(const uint8_t* bytes, long n, double max_distance, int propagate_negative)
{
  auto* df = new PropagationDistanceField(1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, max_distance, propagate_negative != 0);
  if (!df->readFromStream(bytes, (size_t)n))
  {
    delete df;
    return nullptr;
  }
  return df;
}
void orc_field_geometry(void* f, double* out7)
{
  auto* df = static_cast<PropagationDistanceField*>(f);
  out7[0] = df->resolution_;
  for (int i = 0; i < 3; ++i)
  {
    out7[1 + i] = df->size_[i];
    out7[4 + i] = df->origin_[i];
  }
}
void orc_field_add_voxels(void* f, const int32_t* ijk, long n)
{
  auto* df = static_cast<PropagationDistanceField*>(f);
  std::vector<I3> v;
  for (long i = 0; i < n; ++i)
    if (df->g_.isCellValid(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]) &&
        df->cell(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]).distance_square > 0)
      v.push_back(I3{ ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2] });
  df->addNewObstacleVoxels(v);
}

// Replaces the propagated d^2 arrays by the EXACT clamped squared EDT of the field's current zero set (F4): the
// second oracle of SURVEY.md 8(c).  Occupancy (the zero set) is untouched, so later add/remove calls continue
// from the same obstacle voxels.
void orc_field_make_exact(void* f)
{
  auto* df = static_cast<PropagationDistanceField*>(f);
  const int nx = df->g_.num_cells[0], ny = df->g_.num_cells[1], nz = df->g_.num_cells[2];
  const size_t n = df->cells_.size();
  std::vector<uint8_t> occ(n);
  std::vector<int32_t> out(n);
  for (size_t i = 0; i < n; ++i)
    occ[i] = df->cells_[i].distance_square == 0;
  edt_exact(occ.data(), nx, ny, nz, df->max_distance_sq_, out.data());
  for (size_t i = 0; i < n; ++i)
    df->cells_[i].distance_square = out[i];
  if (df->propagate_negative_)
  {
    for (size_t i = 0; i < n; ++i)
      occ[i] = !occ[i];
    edt_exact(occ.data(), nx, ny, nz, df->max_distance_sq_, out.data());
    for (size_t i = 0; i < n; ++i)
      df->cells_[i].negative_distance_square = out[i];
  }
}

void orc_edt_exact(const uint8_t* occ, int nx, int ny, int nz, int max_d2, int32_t* out)
{
  edt_exact(occ, nx, ny, nz, max_d2, out);
}
void orc_edt_brute(const uint8_t* occ, int nx, int ny, int nz, int max_d2, int32_t* out)
{
  edt_brute(occ, nx, ny, nz, max_d2, out);
}

// ---------------------------------------------------------------- shapes
// dims3: sphere r,-,- ; cylinder r,length,- ; box x,y,z.  pose12 row-major 3x4.  Returns the number of
// points; writes at most cap of them.
long orc_shape_points(int type, const double* dims3, const double* pose12, double scale, double padding,
                      double resolution, double* out, long cap)
{
  Body b;
  b.type = type;
  b.dims[0] = dims3[0];
  b.dims[1] = dims3[1];
  b.dims[2] = dims3[2];
  b.pose = isoFrom12(pose12);
  b.scale = scale;
  b.padding = padding;
  b.update();
  std::vector<V3> pts;
  findInternalPointsConvex(b, resolution, pts);
  for (long i = 0; i < (long)pts.size() && i < cap; ++i)
  {
    out[3 * i] = pts[i].x;
    out[3 * i + 1] = pts[i].y;
    out[3 * i + 2] = pts[i].z;
  }
  return (long)pts.size();
}

long orc_octree_points(double resolution, const double* leaves4, int n, double* out, long cap)
{
  std::vector<V3> pts;
  octreeLeavesToPoints(resolution, leaves4, n, pts);
  for (long i = 0; i < (long)pts.size() && i < cap; ++i)
  {
    out[3 * i] = pts[i].x;
    out[3 * i + 1] = pts[i].y;
    out[3 * i + 2] = pts[i].z;
  }
  return (long)pts.size();
}

// BodyDecomposition(shapes, poses, resolution, padding): spheres (x,y,z,r), points, merged bounding sphere.
void orc_body_decomposition(int n_shapes, const int* types, const double* dims3, const double* poses12,
                            double resolution, double padding, double* spheres4, long sphere_cap, long* n_spheres,
                            double* points3, long point_cap, long* n_points, double* bsphere4)
{
  std::vector<Body> bodies((size_t)n_shapes);
  for (int i = 0; i < n_shapes; ++i)
  {
    bodies[i].type = types[i];
    for (int k = 0; k < 3; ++k)
      bodies[i].dims[k] = dims3[3 * i + k];
    bodies[i].pose = isoFrom12(poses12 + 12 * i);
  }
  BodyDecomposition bd;
  bd.init(bodies, resolution, padding);
  *n_spheres = (long)bd.collision_spheres.size();
  *n_points = (long)bd.relative_collision_points.size();
  for (long i = 0; i < *n_spheres && i < sphere_cap; ++i)
  {
    spheres4[4 * i] = bd.collision_spheres[i].relative_vec.x;
    spheres4[4 * i + 1] = bd.collision_spheres[i].relative_vec.y;
    spheres4[4 * i + 2] = bd.collision_spheres[i].relative_vec.z;
    spheres4[4 * i + 3] = bd.collision_spheres[i].radius;
  }
  for (long i = 0; i < *n_points && i < point_cap; ++i)
  {
    points3[3 * i] = bd.relative_collision_points[i].x;
    points3[3 * i + 1] = bd.relative_collision_points[i].y;
    points3[3 * i + 2] = bd.relative_collision_points[i].z;
  }
  bsphere4[0] = bd.relative_bounding_sphere.center.x;
  bsphere4[1] = bd.relative_bounding_sphere.center.y;
  bsphere4[2] = bd.relative_bounding_sphere.center.z;
  bsphere4[3] = bd.relative_bounding_sphere.radius;
}

// ---------------------------------------------------------------- robot model
void* orc_model_create(int L, const int* parent, const int* joint_type, const double* axis3, const double* origin12,
                       const int* origin_is_identity, const int* var_index, const int* mimic_var,
                       const double* mimic_mult, const double* mimic_offset, int nvar)
{
  auto m = std::make_shared<RobotModel>();
  m->links.resize(L);
  m->geom.resize(L);
  m->nvar = nvar;
  for (int i = 0; i < L; ++i)
  {
    LinkDesc& l = m->links[i];
    l.parent = parent[i];
    l.joint_type = joint_type[i];
    for (int k = 0; k < 3; ++k)
      l.axis[k] = axis3[3 * i + k];
    l.origin = isoFrom12(origin12 + 12 * i);
    l.origin_is_identity = origin_is_identity[i];
    l.var_index = var_index[i];
    l.mimic_var = mimic_var[i];
    l.mimic_mult = mimic_mult[i];
    l.mimic_offset = mimic_offset[i];
  }
  m->finalize();
  return new ModelHandle{ m };
}
void orc_model_destroy(void* h)
{
  delete static_cast<ModelHandle*>(h);
}
void orc_model_set_link_geometry(void* h, int link, int has_geometry, const double* spheres4, int ns,
                                 const double* points3, long np, const double* bsphere4)
{
  LinkGeom& g = static_cast<ModelHandle*>(h)->model->geom[link];
  g.has_geometry = has_geometry;
  g.spheres.clear();
  for (int i = 0; i < ns; ++i)
    g.spheres.push_back(
        CollisionSphere{ V3{ spheres4[4 * i], spheres4[4 * i + 1], spheres4[4 * i + 2] }, spheres4[4 * i + 3] });
  g.points = toPoints(points3, np);
  g.bsphere.center = V3{ bsphere4[0], bsphere4[1], bsphere4[2] };
  g.bsphere.radius = bsphere4[3];
}
void orc_model_fk(void* h, const double* q, double* out12)
{
  const RobotModel& m = *static_cast<ModelHandle*>(h)->model;
  std::vector<Iso> T(m.links.size());
  m.fk(q, T.data());
  for (size_t i = 0; i < T.size(); ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c)
        out12[12 * i + 4 * r + c] = T[i].m[r][c];
}

// ---------------------------------------------------------------- collision env
void* orc_env_create(void* model, const double* size3, const double* origin_center3, int use_signed, double resolution,
                     double collision_tolerance, double max_propagation_distance)
{
  EnvParams p;
  for (int i = 0; i < 3; ++i)
  {
    p.size[i] = size3[i];
    p.origin[i] = origin_center3[i];
  }
  p.use_signed = use_signed != 0;
  p.resolution = resolution;
  p.collision_tolerance = collision_tolerance;
  p.max_propagation_distance = max_propagation_distance;
  auto* h = new EnvHandle;
  h->model = static_cast<ModelHandle*>(model)->model;
  h->env.reset(new CollisionEnvDistanceField(h->model, p));
  return h;
}
void orc_env_destroy(void* h)
{
  delete static_cast<EnvHandle*>(h);
}
void* orc_env_world_field(void* h)  // borrowed
{
  return static_cast<EnvHandle*>(h)->env->world_field_.get();
}
void* orc_env_self_field(void* h)  // borrowed, may be null
{
  auto& d = static_cast<EnvHandle*>(h)->env->dfce_;
  return d ? d->distance_field.get() : nullptr;
}
void orc_env_set_group(void* h, const int* group_links, int n, const int* acm_entry_LxL, const double* self_state)
{
  std::vector<int> g(group_links, group_links + n);
  static_cast<EnvHandle*>(h)->env->setGroup(g, acm_entry_LxL, self_state);
}
void orc_env_build_link_fields(void* h)
{
  static_cast<EnvHandle*>(h)->env->buildLinkFields();
}
int orc_env_group_sphere_count(void* h)
{
  auto& env = *static_cast<EnvHandle*>(h)->env;
  int s = 0;
  for (size_t i = 0; i < env.dfce_->link_indices.size(); ++i)
    if (env.dfce_->link_has_geometry[i])
      s += (int)env.model_->geom[env.dfce_->link_indices[i]].spheres.size();
  return s;
}
// resolved ACM tables of the current group: self_enabled[n], intra_enabled[n*n]
void orc_env_group_tables(void* h, uint8_t* self_enabled, uint8_t* intra_enabled)
{
  auto& d = *static_cast<EnvHandle*>(h)->env->dfce_;
  const size_t n = d.link_indices.size();
  for (size_t i = 0; i < n; ++i)
  {
    self_enabled[i] = d.self_collision_enabled[i] ? 1 : 0;
    for (size_t j = 0; j < n; ++j)
      intra_enabled[i * n + j] = d.intra_group_collision_enabled[i][j] ? 1 : 0;
  }
}

// posed sphere centres of the group for one state, flattened over links with geometry [S][3]
void orc_env_sphere_centers(void* h, const double* q, double* out)
{
  auto& env = *static_cast<EnvHandle*>(h)->env;
  GroupStateRepresentation gsr;
  env.poseGroup(q, gsr);
  size_t k = 0;
  for (size_t i = 0; i < gsr.sphere_centers.size(); ++i)
  {
    if (!env.dfce_->link_has_geometry[i])
      continue;
    for (const V3& c : gsr.sphere_centers[i])
    {
      out[k++] = c.x;
      out[k++] = c.y;
      out[k++] = c.z;
    }
  }
}

// mode: bit0 robot-vs-world, bit1 self (self field, then intra group).  Evaluation order = PlanningScene
// (planning_scene.cpp:492-503) when both; mode 4 = CollisionEnvDistanceField::checkCollision order (:1441-1464).
// q is [N][nvar] doubles.  n_threads contiguous shards (evaluate_collision_checking_speed.cpp:121-124 shape).
// Returns wall seconds spent in the checking loop.
double orc_env_check_batch(void* h, const double* q, long N, int mode, uint8_t* verdict, int n_threads)
{
  auto& env = *static_cast<EnvHandle*>(h)->env;
  const int nvar = env.model_->nvar;
  if (n_threads < 1)
    n_threads = 1;
  auto work = [&](long lo, long hi) {
    for (long s = lo; s < hi; ++s)
    {
      const double* qs = q + s * nvar;
      bool c;
      if (mode == 4)
        c = env.checkCollision(qs);
      else
        c = env.checkSceneOrder(qs, (mode & 1) != 0, (mode & 2) != 0);
      verdict[s] = c ? 1 : 0;
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads == 1)
    work(0, N);
  else
  {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
    {
      long lo = N * t / n_threads, hi = N * (t + 1) / n_threads;
      th.emplace_back(work, lo, hi);
    }
    for (auto& t : th)
      t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// getCollisionGradients for N states.  flags: bit0 self (link fields + self field), bit1 intra, bit2 env.
// Outputs flattened over group spheres: dist[N][S], grad[N][S][3], type[N][S], centers[N][S][3] (nullable),
// closest[N][n_group_links] (GradientInfo::closest_distance, nullable).
double orc_env_gradients_batch(void* h, const double* q, long N, int flags, double* dist, double* grad, int32_t* type,
                               double* centers, double* closest)
{
  auto& env = *static_cast<EnvHandle*>(h)->env;
  const int nvar = env.model_->nvar;
  const int S = orc_env_group_sphere_count(h);
  const size_t n = env.dfce_->link_indices.size();
  auto t0 = std::chrono::steady_clock::now();
  GroupStateRepresentation gsr;
  for (long s = 0; s < N; ++s)
  {
    env.getCollisionGradients(q + s * nvar, gsr, (flags & 1) != 0, (flags & 2) != 0, (flags & 4) != 0);
    size_t k = 0;
    for (size_t i = 0; i < n; ++i)
    {
      if (closest)
        closest[s * n + i] = env.dfce_->link_has_geometry[i] ? gsr.gradients[i].closest_distance : DBL_MAX;
      if (!env.dfce_->link_has_geometry[i])
        continue;
      const GradientInfo& gi = gsr.gradients[i];
      for (size_t j = 0; j < gi.distances.size(); ++j, ++k)
      {
        dist[s * S + k] = gi.distances[j];
        grad[(s * S + k) * 3 + 0] = gi.gradients[j].x;
        grad[(s * S + k) * 3 + 1] = gi.gradients[j].y;
        grad[(s * S + k) * 3 + 2] = gi.gradients[j].z;
        type[s * S + k] = gi.types[j];
        if (centers)
        {
          centers[(s * S + k) * 3 + 0] = gi.sphere_locations[j].x;
          centers[(s * S + k) * 3 + 1] = gi.sphere_locations[j].y;
          centers[(s * S + k) * 3 + 2] = gi.sphere_locations[j].z;
        }
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// checkCollision with req.contacts == true for one state.  out: rows of (link1, body2, sphere, x, y, z) as doubles;
// body2 -1 = "self", -2 = "environment".  Returns the contact count (may exceed cap; only cap rows are written);
// *collision receives res.collision.
long orc_env_check_contacts(void* h, const double* q, long max_contacts, long max_contacts_per_pair, double* out,
                            long cap, int* collision)
{
  auto& env = *static_cast<EnvHandle*>(h)->env;
  auto res = env.checkCollisionContacts(q, (size_t)max_contacts, (size_t)max_contacts_per_pair);
  *collision = res.collision ? 1 : 0;
  long k = 0;
  for (const auto& c : res.contacts)
  {
    if (k < cap)
    {
      double* r = out + 6 * k;
      r[0] = c.link1;
      r[1] = c.body2;
      r[2] = c.sphere;
      r[3] = c.pos.x;
      r[4] = c.pos.y;
      r[5] = c.pos.z;
    }
    ++k;
  }
  return k;
}

int orc_hardware_concurrency()
{
  return (int)std::thread::hardware_concurrency();
}

}  // extern "C"
