// Test / demo driver of the C++ host mirror (collision_env_b200.hpp): reads a scenario written by
// moveit_b200/scenario_io.py, builds RobotModel + World + AllowedCollisionMatrix through the reference-shaped API,
// allocates the env through CollisionDetectorAllocatorB200 and writes verdicts / contacts / gradients that
// tests/test_host_mirror.py compares with the CPU oracle.  Everything numeric happens on the GPU via libb200df.so.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "b200_config.hpp"
#include "collision_env_b200.hpp"
#include "distance_field_b200.hpp"

using namespace collision_detection_b200;

namespace
{
Isometry3d readPose(std::istream& in)
{
  Isometry3d p;
  for (double& v : p.m)
    in >> v;
  return p;
}

template <typename T>
void writeFile(const std::string& path, const std::vector<T>& v)
{
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}
}  // namespace

// distance_field/test/test_distance_field.cpp:882-928 (TestReadWrite) through PropagationDistanceFieldB200, plus the
// files the Python-side test cross-reads: host_driver --field-stream <prefix> [<stream written elsewhere>]
static int fieldStreamTest(const std::string& prefix, const char* foreign)
{
  using distance_field_b200::PropagationDistanceFieldB200;
  auto require = [](bool ok, const char* what) {
    if (!ok)
      throw std::runtime_error(std::string("field stream test: ") + what);
  };
  auto dump = [](const std::string& path, const std::vector<std::int32_t>& v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(std::int32_t)));
  };
  PropagationDistanceFieldB200 small_df(1.0, 1.0, 1.0, 0.1, 0.0, 0.0, 0.0, 0.3);
  small_df.addPointsToField({ Vector3d{ 0.1, 0.0, 0.0 }, Vector3d{ 0.0, 0.1, 0.2 }, Vector3d{ 0.4, 0.0, 0.0 } });
  {
    std::ofstream sf(prefix + ".small.df", std::ios::out | std::ios::binary);
    require(small_df.writeToStream(sf), "writeToStream(small)");
  }
  {
    std::ifstream si(prefix + ".small.df", std::ios::in | std::ios::binary);
    PropagationDistanceFieldB200 small_df2(si, 0.25, false);
    require(small_df2.valid() && small_df2.getXNumCells() == 10, "small field read back");
    require(small_df.distanceSquares() == small_df2.distanceSquares(), "small field distances differ after the round trip");
    // cell and world queries agree with the gradient query's distance
    double gx, gy, gz;
    bool inb;
    require(small_df2.getDistance(1, 0, 0) == 0.0 && small_df2.getDistance(0.1, 0.0, 0.0) == 0.0, "obstacle cell");
    require(small_df2.getDistance(5, 5, 5) == small_df2.getDistanceGradient(0.5, 0.5, 0.5, gx, gy, gz, inb) && inb,
            "cell vs gradient query");
    require(small_df2.getDistance(1000.0, 0.0, 0.0) == small_df2.getUninitializedDistance(), "out of bounds distance");
  }
  dump(prefix + ".small.d2.i32", small_df.distanceSquares());

  PropagationDistanceFieldB200 df(3.0, 3.0, 4.0, 0.02, 0.0, 0.0, 0.0, 0.25, false);
  shapes::Shape sphere;
  sphere.type = shapes::SPHERE;
  sphere.dims[0] = 0.5;
  Isometry3d pose;
  pose.m[3] = pose.m[7] = pose.m[11] = 0.5;
  df.addShapeToField(&sphere, pose);
  {
    std::ofstream f(prefix + ".big.df", std::ios::out | std::ios::binary);
    require(df.writeToStream(f), "writeToStream(big)");
  }
  std::ifstream i(prefix + ".big.df", std::ios::in | std::ios::binary);
  PropagationDistanceFieldB200 df2(i, 0.25, false);
  require(df2.valid() && df.distanceSquares() == df2.distanceSquares(), "big field distances differ after the round trip");
  PropagationDistanceFieldB200 dfx(i, 0.25, false);  // "we shouldn't segfault if we start with an old ifstream"
  require(!dfx.valid(), "an exhausted stream must not yield a field");
  std::ifstream i2(prefix + ".big.df", std::ios::in | std::ios::binary);
  PropagationDistanceFieldB200 df3(i2, 0.25 + 0.02, false);
  require(df3.valid() && df.distanceSquares() != df3.distanceSquares(), "a larger max distance must change the field");
  dump(prefix + ".big.d2.i32", df.distanceSquares());
  // moveShapeInField == remove + add (test_distance_field.cpp:601-649)
  Isometry3d moved = pose;
  moved.m[3] = moved.m[7] = moved.m[11] = 0.7;
  df.moveShapeInField(&sphere, pose, moved);
  PropagationDistanceFieldB200 fresh(3.0, 3.0, 4.0, 0.02, 0.0, 0.0, 0.0, 0.25, false);
  fresh.addShapeToField(&sphere, moved);
  require(df.distanceSquares() == fresh.distanceSquares(), "moved shape differs from a fresh add");
  if (foreign)
  {
    std::ifstream fi(foreign, std::ios::in | std::ios::binary);
    PropagationDistanceFieldB200 other(fi, 0.25, false);
    require(other.valid(), "could not read the foreign stream");
    dump(prefix + ".foreign.d2.i32", other.distanceSquares());
  }
  std::printf("field stream test ok\n");
  return 0;
}

struct Scenario
{
  std::string robot_name;
  std::shared_ptr<core::RobotModel> model;
  CollisionEnvB200::LinkSpheres overrides;
  int use_acm = 0;
  std::shared_ptr<AllowedCollisionMatrix> acm;
  double size[3], center[3], res = 0.02, max_dist = 0.25, tol = 0.0;
  int use_signed = 0;
  WorldPtr world;
  std::vector<std::string> object_ids;
  std::vector<core::AttachedBody> attached;
  std::string group_name;
  int n_single = 0, n_contacts = 0, n_grad = 0, mutate = 0;
  std::size_t max_contacts = 1, max_per_pair = 1;
};

static Scenario parseScenario(const char* path)
{
  Scenario sc;
  std::string& robot_name = sc.robot_name;
  CollisionEnvB200::LinkSpheres& overrides = sc.overrides;
  int& use_acm = sc.use_acm;
  double(&size)[3] = sc.size;
  double(&center)[3] = sc.center;
  double &res = sc.res, &max_dist = sc.max_dist, &tol = sc.tol;
  int& use_signed = sc.use_signed;
  std::vector<std::string>& object_ids = sc.object_ids;
  std::vector<core::AttachedBody>& attached = sc.attached;
  std::string& group_name = sc.group_name;
  int &n_single = sc.n_single, &n_contacts = sc.n_contacts, &n_grad = sc.n_grad, &mutate = sc.mutate;
  std::size_t &max_contacts = sc.max_contacts, &max_per_pair = sc.max_per_pair;
  {
    std::ifstream in(path);
    if (!in)
      throw std::runtime_error("cannot open scenario");
    in.precision(17);
    std::string tok;
    int n_links = 0;
    in >> tok >> robot_name >> n_links;
    auto model = std::make_shared<core::RobotModel>(robot_name);
    sc.model = model;
    for (int i = 0; i < n_links; ++i)
    {
      core::LinkModel l;
      int jt, n_shapes, n_spheres;
      in >> tok >> l.name >> l.parent >> l.joint_name >> jt >> l.axis[0] >> l.axis[1] >> l.axis[2];
      l.joint_type = static_cast<core::JointType>(jt);
      l.joint_origin = readPose(in);
      in >> l.mimic_joint >> l.mimic_factor >> l.mimic_offset;
      if (l.mimic_joint == "-")
        l.mimic_joint.clear();
      in >> n_shapes;
      for (int s = 0; s < n_shapes; ++s)
      {
        shapes::Shape sh;
        int t;
        in >> t >> sh.dims[0] >> sh.dims[1] >> sh.dims[2];
        sh.type = static_cast<shapes::ShapeType>(t);
        l.shapes.push_back(sh);
        l.shape_origins.push_back(readPose(in));
      }
      in >> n_spheres;
      if (n_spheres >= 0)
      {
        std::vector<CollisionSphere>& v = overrides[l.name];
        for (int s = 0; s < n_spheres; ++s)
        {
          CollisionSphere cs;
          in >> cs.relative_vec_.x >> cs.relative_vec_.y >> cs.relative_vec_.z >> cs.radius_;
          v.push_back(cs);
        }
      }
      model->addLink(l);
    }
    int n_groups = 0;
    in >> tok >> n_groups;
    for (int g = 0; g < n_groups; ++g)
    {
      std::string name;
      int nj;
      in >> name >> nj;
      std::vector<std::string> joints(nj);
      for (std::string& j : joints)
        in >> j;
      model->addJointModelGroup(name, joints);
    }
    // default ACM of a PlanningScene: every pair NEVER, SRDF disable_collisions pairs ALWAYS
    int n_disabled = 0;
    in >> tok >> use_acm >> n_disabled;
    sc.acm = std::make_shared<AllowedCollisionMatrix>(model->getLinkModelNames(), false);
    AllowedCollisionMatrix& acm = *sc.acm;
    for (int k = 0; k < n_disabled; ++k)
    {
      std::string a, b;
      in >> a >> b;
      acm.setEntry(a, b, true);
    }
    in >> tok >> size[0] >> size[1] >> size[2] >> center[0] >> center[1] >> center[2] >> res >> max_dist >>
        use_signed >> tol;
    int n_objects = 0;
    in >> tok >> n_objects;
    auto world = std::make_shared<World>();
    sc.world = world;
    for (int k = 0; k < n_objects; ++k)
    {
      std::string id;
      int t;
      auto sh = std::make_shared<shapes::Shape>();
      in >> id >> t >> sh->dims[0] >> sh->dims[1] >> sh->dims[2];
      sh->type = static_cast<shapes::ShapeType>(t);
      const Isometry3d pose = readPose(in);
      world->addToObject(id, sh, pose);
      object_ids.push_back(id);
    }
    int n_attached = 0;
    in >> tok >> n_attached;
    for (int k = 0; k < n_attached; ++k)
    {
      core::AttachedBody b;
      int ns = 0, nt = 0;
      in >> b.name >> b.link_name >> ns;
      for (int s = 0; s < ns; ++s)
      {
        shapes::Shape sh;
        int t;
        in >> t >> sh.dims[0] >> sh.dims[1] >> sh.dims[2];
        sh.type = static_cast<shapes::ShapeType>(t);
        b.shapes.push_back(sh);
        b.attach_trans.push_back(readPose(in));
      }
      in >> nt;
      for (int t = 0; t < nt; ++t)
      {
        std::string name;
        in >> name;
        b.touch_links.insert(name);
      }
      attached.push_back(b);
    }
    in >> tok >> group_name >> n_single >> n_contacts >> max_contacts >> max_per_pair >> n_grad >> mutate;
    if (group_name == "-")
      group_name.clear();
    if (!in)
      throw std::runtime_error("scenario parse error");

  }
  return sc;
}

// Reads every row of a raw float64 file as one state of the model.
static std::vector<double> readStates(const char* path, int nvar, std::int64_t* n)
{
  std::ifstream sf(path, std::ios::binary | std::ios::ate);
  if (!sf)
    throw std::runtime_error("cannot open the state file");
  const std::streamsize bytes = sf.tellg();
  sf.seekg(0);
  *n = bytes / (std::streamsize)(sizeof(double) * nvar);
  std::vector<double> q((std::size_t)*n * nvar);
  sf.read(reinterpret_cast<char*>(q.data()), bytes);
  return q;
}

static void appendGsr(const GroupStateRepresentation& gsr, std::vector<double>& out)
{
  for (const GradientInfo& gi : gsr.gradients_)
    for (std::size_t s = 0; s < gi.distances.size(); ++s)
      out.insert(out.end(), { gi.distances[s], gi.gradients[s].x, gi.gradients[s].y, gi.gradients[s].z,
                              (double)gi.types[s], gi.sphere_locations[s].x, gi.sphere_locations[s].y,
                              gi.sphere_locations[s].z });
}

// host_driver --chomp <scenario> <states> <prefix> <waypoints per trajectory>
// The consumer side of ChompOptimizer (moveit_planners/chomp/chomp_motion_planner/src/chomp_optimizer.cpp) on the
// hybrid env CHOMP casts the active collision env to (:81-88):
//   initialize() :93-106    getCollisionGradients(req, res, state, &acm, gsr)   - the ACM enters the gsr's cache entry
//   performForwardKinematics() :895-965, per waypoint: getCollisionGradients(req, res, state, nullptr, gsr), then the
//       loop over gsr->gradients_ that reads sphere_locations, distances, sphere_radii, gradients and flags the point
//       as colliding when distance - radius < radius
// written out per waypoint; then the same trajectories through ONE getCollisionGradientsBatch call reusing the gsr,
// which has to give the same numbers.
static int chompTest(const char* scen_path, const char* states_path, const std::string& prefix, int waypoints)
{
  Scenario sc = parseScenario(scen_path);
  std::int64_t n = 0;
  const int nvar = sc.model->getVariableCount();
  std::vector<double> q = readStates(states_path, nvar, &n);
  const Vector3d origin{ sc.center[0], sc.center[1], sc.center[2] };
  auto hy = std::make_shared<CollisionEnvHybridB200>(sc.model, sc.world, sc.overrides, sc.size[0], sc.size[1], sc.size[2],
                                                     origin, sc.use_signed != 0, sc.res, sc.tol, sc.max_dist);
  // what ChompOptimizer does with the planning scene's env: a dynamic_cast to the hybrid class
  const CollisionEnvNoMeshBackend* base = hy.get();
  const CollisionEnvHybridB200* hy_env = dynamic_cast<const CollisionEnvHybridB200*>(base);
  if (!hy_env)
    throw std::runtime_error("dynamic_cast to the hybrid env failed");
  CollisionRequest req;
  req.group_name = sc.group_name;
  core::RobotState state(sc.model);
  GroupStateRepresentationPtr gsr;
  CollisionResult res;
  hy_env->getCollisionGradients(req, res, state, &*sc.acm, gsr);  // ChompOptimizer::initialize
  if (!gsr)
    throw std::runtime_error(hy_env->getCollisionWorldDistanceField()->lastError());
  std::size_t num_collision_points = 0;
  for (const GradientInfo& gi : gsr->gradients_)
    num_collision_points += gi.gradients.size();
  std::vector<double> loop_out, batch_out;
  std::vector<std::uint8_t> in_collision;
  const auto t0 = std::chrono::steady_clock::now();
  for (std::int64_t i = 0; i < n; ++i)  // performForwardKinematics over every waypoint of every trajectory
  {
    state.setVariablePositions(&q[(std::size_t)i * nvar]);
    CollisionResult r;
    hy_env->getCollisionGradients(req, r, state, nullptr, gsr);
    if (!gsr)
      throw std::runtime_error(hy_env->getCollisionWorldDistanceField()->lastError());
    bool colliding = false;
    for (const GradientInfo& info : gsr->gradients_)
      for (std::size_t k = 0; k < info.sphere_locations.size(); ++k)
        if (info.distances[k] - info.sphere_radii[k] < info.sphere_radii[k])
          colliding = true;
    in_collision.push_back(colliding ? 1 : 0);
    appendGsr(*gsr, loop_out);
  }
  const double us_loop = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  const auto t1 = std::chrono::steady_clock::now();
  std::vector<GroupStateRepresentationPtr> all;
  if (!hy_env->getCollisionGradientsBatch(req, q.data(), n, nullptr, all, nullptr, gsr))
    throw std::runtime_error(hy_env->getCollisionWorldDistanceField()->lastError());
  const double us_batch = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count();
  for (const GroupStateRepresentationPtr& g : all)
    appendGsr(*g, batch_out);
  if (batch_out != loop_out)
    throw std::runtime_error("batched gradients differ from the per-waypoint loop");
  if (loop_out.size() != (std::size_t)n * num_collision_points * 8)
    throw std::runtime_error("collision point count changed between waypoints");
  writeFile(prefix + ".chomp.f64", loop_out);
  writeFile(prefix + ".chomp_collision.u8", in_collision);
  std::printf("chomp test ok: %lld trajectories x %d waypoints, %zu collision points, per-waypoint loop %.1f us/waypoint, "
              "one batched call %.2f us/waypoint, link-field term %d\n",
              (long long)(n / std::max(1, waypoints)), waypoints, num_collision_points, us_loop / std::max<std::int64_t>(n, 1),
              us_batch / std::max<std::int64_t>(n, 1), (int)hy_env->getCollisionWorldDistanceField()->linkFieldTerm());
  return 0;
}

// host_driver --contract
// The error contract of the CollisionEnv boundary (collision_env.h: void methods, no exceptions, results only through
// CollisionResult): an env that cannot work - no CUDA device (this is what the CPU-only test box sees) or a field the
// engine refuses (resolution 0) - must construct without throwing and answer every call FAIL-CLOSED.
static int contractTest()
{
  auto model = std::make_shared<core::RobotModel>("contract_bot");
  core::LinkModel base;
  base.name = "base";
  base.joint_name = "fixed_base";
  base.joint_type = core::FIXED;
  shapes::Shape sp;
  sp.type = shapes::SPHERE;
  sp.dims[0] = 0.04;
  base.shapes.push_back(sp);
  base.shape_origins.push_back(Isometry3d::Identity());
  model->addLink(base);
  core::LinkModel arm;
  arm.name = "arm";
  arm.parent = 0;
  arm.joint_name = "j1";
  arm.joint_type = core::REVOLUTE;
  arm.shapes.push_back(sp);
  arm.shape_origins.push_back(Isometry3d::Identity());
  model->addLink(arm);
  auto world = std::make_shared<World>();
  auto box = std::make_shared<shapes::Shape>();
  box->type = shapes::BOX;
  box->dims[0] = box->dims[1] = box->dims[2] = 0.02;
  world->addToObject("box", box, Isometry3d::Identity());
  int n_dev = 0;
  b200df_device_count(&n_dev);
  // with a device the env is made to fail through a resolution the engine refuses; without one it fails by itself
  const double resolution = n_dev > 0 ? 0.0 : 0.02;
  auto check = [](bool ok, const char* what) {
    if (!ok)
      throw std::runtime_error(std::string("contract test: ") + what);
  };
  CollisionEnvB200Ptr env;
  try
  {
    env = std::make_shared<CollisionEnvB200>(model, world, CollisionEnvB200::LinkSpheres(), 0.1, 0.1, 0.2, Vector3d(),
                                             false, resolution);
  }
  catch (...)
  {
    check(false, "the constructor threw");
  }
  check(!env->valid() && !env->lastError().empty(), "a failed env must say so");
  AllowedCollisionMatrix acm(model->getLinkModelNames(), false);
  core::RobotState st(model);
  CollisionRequest req;
  try
  {
    const std::size_t before = env->errorCount();
    CollisionResult a, b, c, d, e, f, g;
    env->checkCollision(req, a, st);
    env->checkCollision(req, b, st, acm);
    env->checkSelfCollision(req, c, st);
    env->checkSelfCollision(req, d, st, acm);
    env->checkRobotCollision(req, e, st);
    env->checkRobotCollision(req, f, st, acm);
    env->getAllCollisions(req, g, st, &acm);
    check(a.collision && b.collision && c.collision && d.collision && e.collision && f.collision && g.collision,
          "a check that could not run must report a collision (fail-closed)");
    check(env->errorCount() >= before + 7, "every failed call is counted");
    std::vector<double> q(3 * (std::size_t)model->getVariableCount(), 0.0);
    std::vector<std::uint8_t> v(3, 0);
    check(!env->checkCollisionBatch(req, q.data(), 3, &acm, B200DF_CHECK_ROBOT, st, v.data()), "batch must report failure");
    check(v[0] == 1 && v[1] == 1 && v[2] == 1, "batch verdicts must be fail-closed");
    std::vector<std::int32_t> first(1, 0);
    check(!env->checkMotionBatch(req, q.data(), q.data(), 1, 4, &acm, st, first.data()) && first[0] == 1,
          "motion batch must be fail-closed");
    std::vector<std::size_t> invalid;
    check(!env->isPathValid(req, q.data(), 3, &acm, st, &invalid) && invalid.size() == 3, "isPathValid must be false");
    GroupStateRepresentationPtr gsr;
    CollisionResult r;
    env->getCollisionGradients(req, r, st, &acm, gsr);
    check(!gsr, "failed gradients leave a null gsr");
    check(env->distanceRobot(st) == 0.0, "failed distance query answers 0 (touching)");
    check(env->rebuildWorldField() < 0.f, "failed rebuild is signalled");
    // world changes and setWorld on a failed env are absorbed as well
    world->addToObject("box2", box, Isometry3d::Identity());
    world->removeObject("box");
    env->setWorld(std::make_shared<World>());
    CollisionEnvHybridB200 hy(model, world, CollisionEnvB200::LinkSpheres(), 0.1, 0.1, 0.2, Vector3d(), false, resolution);
    CollisionResult h;
    hy.checkCollisionDistanceField(req, h, st, acm);
    check(h.collision, "hybrid forwards the fail-closed verdict");
    auto alloc = CollisionDetectorAllocatorB200::create();
    CollisionEnvB200Ptr copy = alloc->allocateEnv(env, world);
    check(copy && (n_dev > 0 || !copy->valid()), "allocateEnv(orig, world) must not throw");
  }
  catch (const std::runtime_error&)
  {
    throw;
  }
  catch (...)
  {
    check(false, "a public method threw");
  }
  std::printf("contract test ok: %zu errors recorded, last: %s\n", env->errorCount(), env->lastError().c_str());
  return 0;
}

// host_driver --octree <nodes.txt> <prefix>
// An <octomap> world object (shapes::OcTree, planning_scene.cpp:1390-1412) through the env's World observer:
// nodes.txt lists every node of the tree in begin_tree() order as "x y z size occupied leaf".  Written out: the world
// field's d^2 with the reference env's all-nodes ingestion (Q9), with occupied leaves only, and after the object was
// removed again.
static int octreeTest(const char* nodes_path, const std::string& prefix)
{
  auto nodes = std::make_shared<shapes::OcTreeNodes>();
  {
    std::ifstream in(nodes_path);
    double x, y, z, size;
    int occ, leaf;
    while (in >> x >> y >> z >> size >> occ >> leaf)
    {
      nodes->xyzs.insert(nodes->xyzs.end(), { x, y, z, size });
      nodes->occupied.push_back((std::uint8_t)occ);
      nodes->leaf.push_back((std::uint8_t)leaf);
    }
  }
  auto model = std::make_shared<core::RobotModel>("octree_bot");
  core::LinkModel base;
  base.name = "base";
  base.joint_name = "fixed_base";
  base.joint_type = core::FIXED;
  model->addLink(base);
  auto dump = [&](const CollisionEnvB200& env, const std::string& name) {
    distance_field_b200::PropagationDistanceFieldB200 view(env.getDistanceField());
    std::vector<std::int32_t> d2 = view.distanceSquares();
    std::ofstream f(prefix + name, std::ios::binary);
    f.write(reinterpret_cast<const char*>(d2.data()), (std::streamsize)(d2.size() * sizeof(std::int32_t)));
  };
  auto tree = std::make_shared<shapes::Shape>();
  tree->type = shapes::OCTREE;
  tree->octree = nodes;
  for (int mode = 0; mode < 2; ++mode)
  {
    auto world = std::make_shared<World>();
    CollisionEnvB200 env(model, world, CollisionEnvB200::LinkSpheres(), 1.0, 1.0, 1.0, Vector3d{ 0.5, 0.5, 0.5 }, false,
                         0.02);
    if (!env.valid())
      throw std::runtime_error(env.lastError());
    env.setOctreeOccupiedLeavesOnly(mode == 1);
    world->addToObject("<octomap>", tree, Isometry3d::Identity());
    dump(env, mode == 0 ? ".q9.d2.i32" : ".leaves.d2.i32");
    // a second env allocated on the already populated world sees the same tree (notifyObserverAllObjects(CREATE))
    if (mode == 0)
    {
      CollisionEnvB200 late(model, world, CollisionEnvB200::LinkSpheres(), 1.0, 1.0, 1.0, Vector3d{ 0.5, 0.5, 0.5 },
                            false, 0.02);
      dump(late, ".late.d2.i32");
    }
    world->removeObject("<octomap>");
    dump(env, mode == 0 ? ".q9_removed.d2.i32" : ".leaves_removed.d2.i32");
    if (env.errorCount() != 0)
      throw std::runtime_error(env.lastError());
  }
  std::printf("octree test ok: %zu nodes\n", nodes->occupied.size());
  return 0;
}

// host_driver --config <params.yaml> <prefix> [namespace]
// Reads the plugin's parameter namespace (b200_config.hpp; link_spheres in the format of
// collision_distance_field_ros_helpers.h:46-120), writes what it understood, and allocates an env from it on a small
// robot so that the values can be seen to arrive (field dimensions, sphere counts).
static int configTest(const char* yaml_path, const std::string& prefix, const char* ns)
{
  std::ifstream in(yaml_path);
  if (!in)
    throw std::runtime_error("cannot open the parameter file");
  const B200DistanceFieldConfig c = configFromYaml(in, ns ? ns : "");
  std::ofstream out(prefix + ".config.txt");
  out.precision(17);
  out << "size " << c.size_x << " " << c.size_y << " " << c.size_z << "\n";
  out << "origin " << c.origin.x << " " << c.origin.y << " " << c.origin.z << "\n";
  out << "resolution " << c.resolution << "\n";
  out << "collision_tolerance " << c.collision_tolerance << "\n";
  out << "max_propogation_distance " << c.max_propogation_distance << "\n";
  out << "use_signed_distance_field " << (c.use_signed_distance_field ? 1 : 0) << "\n";
  out << "padding " << c.padding << "\n";
  out << "octree_occupied_leaves_only " << (c.octree_occupied_leaves_only ? 1 : 0) << "\n";
  out << "devices";
  for (int d : c.devices)
    out << " " << d;
  out << "\n";
  for (const auto& kv : c.link_spheres)
  {
    out << "link " << kv.first << " " << kv.second.size();
    for (const CollisionSphere& s : kv.second)
      out << " " << s.relative_vec_.x << " " << s.relative_vec_.y << " " << s.relative_vec_.z << " " << s.radius_;
    out << "\n";
  }
  for (const std::string& w : c.warnings)
    out << "warning " << w << "\n";
  int n_dev = 0;
  b200df_device_count(&n_dev);
  if (n_dev > 0)
  {
    auto model = std::make_shared<core::RobotModel>("config_bot");
    for (int i = 0; i < 2; ++i)
    {
      core::LinkModel l;
      l.name = i == 0 ? "base" : "arm";
      l.joint_name = i == 0 ? "fixed_base" : "j1";
      l.parent = i - 1;
      l.joint_type = i == 0 ? core::FIXED : core::REVOLUTE;
      shapes::Shape box;
      box.type = shapes::BOX;
      box.dims[0] = box.dims[1] = 0.05;
      box.dims[2] = 0.3;
      l.shapes.push_back(box);
      l.shape_origins.push_back(Isometry3d::Identity());
      model->addLink(l);
    }
    B200DistanceFieldConfig one = c;
    one.devices.clear();  // the test box has one GPU
    ConfiguredAllocatorB200 alloc(one);
    CollisionEnvB200Ptr env = alloc.allocateEnv(std::make_shared<World>(), model);
    if (!env->valid())
      throw std::runtime_error(env->lastError());
    distance_field_b200::PropagationDistanceFieldB200 view(env->getDistanceField());
    out << "field_cells " << view.getXNumCells() << " " << view.getYNumCells() << " " << view.getZNumCells() << "\n";
    out << "field_max_distance " << view.getUninitializedDistance() << "\n";
    CollisionRequest req;
    core::RobotState st(model);
    std::vector<GroupStateRepresentationPtr> gsr;
    if (!env->getCollisionGradientsBatch(req, st.getVariablePositions(), 1, nullptr, gsr) || gsr.empty())
      throw std::runtime_error(env->lastError());
    for (std::size_t l = 0; l < gsr[0]->link_names_.size(); ++l)
      out << "env_spheres " << gsr[0]->link_names_[l] << " " << gsr[0]->gradients_[l].sphere_radii.size() << "\n";
  }
  std::printf("config test ok: %zu link_spheres entries, %zu warnings\n", c.link_spheres.size(), c.warnings.size());
  return 0;
}

int main(int argc, char** argv)
{
  if (argc >= 3 && std::string(argv[1]) == "--field-stream")
  {
    try
    {
      return fieldStreamTest(argv[2], argc > 3 ? argv[3] : nullptr);
    }
    catch (const std::exception& ex)
    {
      std::fprintf(stderr, "host_driver failed: %s\n", ex.what());
      return 1;
    }
  }
  if (argc >= 2 && (std::string(argv[1]) == "--contract" || std::string(argv[1]) == "--chomp" ||
                    std::string(argv[1]) == "--octree" || std::string(argv[1]) == "--config"))
  {
    try
    {
      if (std::string(argv[1]) == "--contract")
        return contractTest();
      if (std::string(argv[1]) == "--config")
      {
        if (argc < 4)
          throw std::runtime_error("usage: host_driver --config <params.yaml> <prefix> [namespace]");
        return configTest(argv[2], argv[3], argc > 4 ? argv[4] : nullptr);
      }
      if (std::string(argv[1]) == "--octree")
      {
        if (argc < 4)
          throw std::runtime_error("usage: host_driver --octree <nodes.txt> <prefix>");
        return octreeTest(argv[2], argv[3]);
      }
      if (argc < 6)
        throw std::runtime_error("usage: host_driver --chomp <scenario> <states.f64> <prefix> <waypoints>");
      return chompTest(argv[2], argv[3], argv[4], std::atoi(argv[5]));
    }
    catch (const std::exception& ex)
    {
      std::fprintf(stderr, "host_driver failed: %s\n", ex.what());
      return 1;
    }
  }
  if (argc < 4)
  {
    std::fprintf(stderr, "usage: host_driver <scenario.txt> <states.f64> <out_prefix>\n");
    return 2;
  }
  try
  {
    Scenario sc = parseScenario(argv[1]);
    const std::string& robot_name = sc.robot_name;
    std::shared_ptr<core::RobotModel> model = sc.model;
    CollisionEnvB200::LinkSpheres& overrides = sc.overrides;
    const int use_acm = sc.use_acm;
    AllowedCollisionMatrix& acm = *sc.acm;
    double(&size)[3] = sc.size;
    double(&center)[3] = sc.center;
    const double res = sc.res, max_dist = sc.max_dist, tol = sc.tol;
    const int use_signed = sc.use_signed;
    WorldPtr world = sc.world;
    std::vector<std::string>& object_ids = sc.object_ids;
    std::vector<core::AttachedBody>& attached = sc.attached;
    std::string group_name = sc.group_name;
    const int n_single = sc.n_single, n_contacts = sc.n_contacts, n_grad = sc.n_grad, mutate = sc.mutate;
    const std::size_t max_contacts = sc.max_contacts, max_per_pair = sc.max_per_pair;

    // states
    std::ifstream sf(argv[2], std::ios::binary | std::ios::ate);
    const std::streamsize bytes = sf.tellg();
    sf.seekg(0);
    const int nvar = model->getVariableCount();
    const std::int64_t n = bytes / (std::streamsize)(sizeof(double) * nvar);
    std::vector<double> q((std::size_t)n * nvar);
    sf.read(reinterpret_cast<char*>(q.data()), bytes);

    // the env is created the way PlanningScene does it: through the allocator, on an already populated world
    auto alloc = CollisionDetectorAllocatorB200::create();
    const Vector3d origin{ center[0], center[1], center[2] };
    CollisionEnvB200Ptr env = std::make_shared<CollisionEnvB200>(model, world, overrides, size[0], size[1], size[2],
                                                                 origin, use_signed != 0, res, tol, max_dist);
    if (const char* devs = std::getenv("B200DF_DRIVER_DEVICES"))  // "0,1": shard every batch over these devices
    {
      std::vector<int> list;
      std::stringstream ss(devs);
      std::string item;
      while (std::getline(ss, item, ','))
        list.push_back(std::atoi(item.c_str()));
      env->setDevices(list, /*min_states=*/1);
    }
    const std::string prefix = argv[3];
    const AllowedCollisionMatrix* acm_ptr = use_acm ? &acm : nullptr;
    CollisionRequest req;
    req.group_name = group_name;
    core::RobotState ref_state(model);  // all-zero scene state poses the non-group links
    for (const core::AttachedBody& b : attached)
      ref_state.attachBody(b);
    std::vector<std::uint8_t> v(n);
    if (const char* reps_env = std::getenv("B200DF_DRIVER_BENCH_REPS"))
    {
      // CollisionEnvB200::checkCollisionBatch at scale: the states of the file (pageable fp64 rows, the shape a planner
      // holds) through the C++ env, all three checks; one warm-up call builds the cache entry
      const int reps = std::max(1, std::atoi(reps_env));
      const int all = B200DF_CHECK_SELF | B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT;
      bool ok = env->checkCollisionBatch(req, q.data(), n, acm_ptr, all, ref_state, v.data());
      double best = 1e300;
      for (int k = 0; k < reps; ++k)
      {
        const auto t0 = std::chrono::steady_clock::now();
        ok = env->checkCollisionBatch(req, q.data(), n, acm_ptr, all, ref_state, v.data()) && ok;
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
      }
      std::int64_t colliding = 0;
      for (std::uint8_t x : v)
        colliding += x;
      std::printf("bench_batch ok=%d states=%lld best_s=%.6f states_per_s=%.6e colliding=%lld errors=%zu\n", ok ? 1 : 0,
                  (long long)n, best, (double)n / best, (long long)colliding, env->errorCount());
      writeFile(prefix + ".all.u8", v);
      return ok ? 0 : 1;
    }
    env->checkCollisionBatch(req, q.data(), n, acm_ptr, B200DF_CHECK_ROBOT, ref_state, v.data());
    writeFile(prefix + ".robot.u8", v);
    env->checkCollisionBatch(req, q.data(), n, acm_ptr, B200DF_CHECK_SELF | B200DF_CHECK_INTRA, ref_state, v.data());
    writeFile(prefix + ".self.u8", v);
    env->checkCollisionBatch(req, q.data(), n, acm_ptr, B200DF_CHECK_SELF | B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT,
                             ref_state, v.data());
    writeFile(prefix + ".all.u8", v);

    // single-state CollisionEnv API
    std::vector<std::uint8_t> single;
    for (std::int64_t i = 0; i < std::min<std::int64_t>(n_single, n); ++i)
    {
      core::RobotState st(ref_state);
      st.setVariablePositions(&q[(std::size_t)i * nvar]);
      CollisionResult r1, r2, r3;
      if (acm_ptr)
      {
        env->checkRobotCollision(req, r1, st, acm);
        env->checkSelfCollision(req, r2, st, acm);
        env->checkCollision(req, r3, st, acm);
      }
      else
      {
        env->checkRobotCollision(req, r1, st);
        env->checkSelfCollision(req, r2, st);
        env->checkCollision(req, r3, st);
      }
      single.push_back((r1.collision ? 1 : 0) | (r2.collision ? 2 : 0) | (r3.collision ? 4 : 0));
    }
    writeFile(prefix + ".single.u8", single);

    // concurrent callers on one shared env (planning_scene/test/test_multi_threaded.cpp: every thread repeats
    // checkCollision on its own state and must keep getting the answer a lone caller got)
    {
      const int n_threads = (int)std::min<std::int64_t>(4, (std::int64_t)single.size());
      const int trials = 200;
      std::vector<int> bad(n_threads > 0 ? n_threads : 1, 0);
      std::vector<std::thread> threads;
      for (int t = 0; t < n_threads; ++t)
        threads.emplace_back([&, t] {
          core::RobotState st(ref_state);
          st.setVariablePositions(&q[(std::size_t)t * nvar]);
          for (int k = 0; k < trials; ++k)
          {
            CollisionResult r;
            if (acm_ptr)
              env->checkCollision(req, r, st, acm);
            else
              env->checkCollision(req, r, st);
            if (r.collision != ((single[t] & 4) != 0))
              ++bad[t];
          }
        });
      for (std::thread& th : threads)
        th.join();
      for (int t = 0; t < n_threads; ++t)
        if (bad[t])
          throw std::runtime_error("concurrent checkCollision disagrees with the single-threaded answer");
    }

    // per-call cost of the single-state CollisionEnv API (what an OMPL StateValidityChecker pays per isValid())
    // (one fixed state: a planner keeps the joints outside the group still, so the cached group entry stays valid)
    double us_plain = 0.0, us_contacts = 0.0, us_grad = 0.0;
    std::size_t regenerated = 0;
    if (!single.empty())
    {
      const int reps = 300;
      const std::size_t gen0 = env->cacheGenerations();
      CollisionRequest creq = req;
      creq.contacts = true;
      creq.max_contacts = 10;
      creq.max_contacts_per_pair = 2;
      for (int pass = 0; pass < 2; ++pass)
      {
        const CollisionRequest& r = pass ? creq : req;
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < reps; ++k)
        {
          core::RobotState st(ref_state);
          st.setVariablePositions(&q[0]);
          CollisionResult res;
          if (acm_ptr)
            env->checkCollision(r, res, st, acm);
          else
            env->checkCollision(r, res, st);
        }
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        (pass ? us_contacts : us_plain) = us;
      }
      {
        // one waypoint's getCollisionGradients, the call CHOMP makes per trajectory point (chomp_optimizer.cpp:924)
        core::RobotState st(ref_state);
        st.setVariablePositions(&q[0]);
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < reps; ++k)
        {
          CollisionResult res;
          GroupStateRepresentationPtr gsr;
          env->getCollisionGradients(req, res, st, acm_ptr, gsr);
        }
        us_grad = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
      }
      regenerated = env->cacheGenerations() - gen0;
    }

    // contacts
    {
      std::ofstream cf(prefix + ".contacts.txt");
      cf.precision(17);
      CollisionRequest creq = req;
      creq.contacts = true;
      creq.max_contacts = max_contacts;
      creq.max_contacts_per_pair = max_per_pair;
      for (std::int64_t i = 0; i < std::min<std::int64_t>(n_contacts, n); ++i)
      {
        core::RobotState st(ref_state);
        st.setVariablePositions(&q[(std::size_t)i * nvar]);
        CollisionResult r;
        if (acm_ptr)
          env->checkCollision(creq, r, st, acm);
        else
          env->checkCollision(creq, r, st);
        cf << "state " << i << " " << (r.collision ? 1 : 0) << " " << r.contact_count << "\n";
        for (const auto& kv : r.contacts)
          for (const Contact& c : kv.second)
            cf << "contact " << c.body_name_1 << " " << c.body_name_2 << " " << c.pos.x << " " << c.pos.y << " "
               << c.pos.z << " " << (int)c.body_type_1 << " " << (int)c.body_type_2 << "\n";
      }
    }

    // gradients (CHOMP entry point), flattened per state: S x (dist, gx, gy, gz, type, cx, cy, cz)
    {
      std::vector<double> out;
      for (std::int64_t i = 0; i < std::min<std::int64_t>(n_grad, n); ++i)
      {
        core::RobotState st(ref_state);
        st.setVariablePositions(&q[(std::size_t)i * nvar]);
        CollisionResult r;
        GroupStateRepresentationPtr gsr;
        // the reference entry is built from the first state it sees; pose the non-group links with the scene state
        std::vector<GroupStateRepresentationPtr> one;
        env->getCollisionGradientsBatch(req, st.getVariablePositions(), 1, acm_ptr, one, &ref_state);
        gsr = one[0];
        for (const GradientInfo& gi : gsr->gradients_)
          for (std::size_t s = 0; s < gi.distances.size(); ++s)
            out.insert(out.end(), { gi.distances[s], gi.gradients[s].x, gi.gradients[s].y, gi.gradients[s].z,
                                    (double)gi.types[s], gi.sphere_locations[s].x, gi.sphere_locations[s].y,
                                    gi.sphere_locations[s].z });
      }
      writeFile(prefix + ".grad.f64", out);
    }

    // batched motion validation + isPathValid: consecutive states of the file are taken as motion end points
    if (n >= 2)
    {
      const std::int64_t m = n / 2;
      std::vector<double> from((std::size_t)m * nvar), to((std::size_t)m * nvar);
      for (std::int64_t i = 0; i < m; ++i)
      {
        std::copy(&q[(std::size_t)(2 * i) * nvar], &q[(std::size_t)(2 * i) * nvar] + nvar, &from[(std::size_t)i * nvar]);
        std::copy(&q[(std::size_t)(2 * i + 1) * nvar], &q[(std::size_t)(2 * i + 1) * nvar] + nvar,
                  &to[(std::size_t)i * nvar]);
      }
      std::vector<std::int32_t> first((std::size_t)m);
      env->checkMotionBatch(req, from.data(), to.data(), m, 8, acm_ptr, ref_state, first.data());
      writeFile(prefix + ".motion.i32", first);
      std::vector<std::size_t> invalid;
      const bool ok = env->isPathValid(req, q.data(), std::min<std::int64_t>(n, 500), acm_ptr, ref_state, &invalid);
      std::vector<std::int32_t> inv(invalid.begin(), invalid.end());
      inv.push_back(ok ? 1 : 0);
      writeFile(prefix + ".path.i32", inv);
    }

    // world mutation through the observer: move the first object by +0.1 m in x, destroy the second
    if (mutate && object_ids.size() >= 2)
    {
      World::ObjectConstPtr o = world->getObject(object_ids[0]);
      Isometry3d moved = o->shape_poses_[0];
      moved.m[3] += 0.1;
      world->moveShapeInObject(object_ids[0], 0, moved);
      world->removeObject(object_ids[1]);
      env->checkCollisionBatch(req, q.data(), n, acm_ptr, B200DF_CHECK_ROBOT, ref_state, v.data());
      writeFile(prefix + ".robot2.u8", v);
      // a second env allocated from the first on the same world must agree (allocateEnv(orig, world))
      CollisionEnvB200Ptr copy = alloc->allocateEnv(env, world);
      std::vector<std::uint8_t> v2(n);
      copy->checkCollisionBatch(req, q.data(), n, acm_ptr, B200DF_CHECK_ROBOT, ref_state, v2.data());
      if (v2 != v)
        throw std::runtime_error("env copied onto the same world disagrees with the original");
    }
    // getDistanceField() (.h:200-203) through the DistanceField API mirror: the world field as a non-owning view
    {
      distance_field_b200::PropagationDistanceFieldB200 view(env->getDistanceField());
      if (view.getXNumCells() <= 0 || view.getResolution() != res || view.getUninitializedDistance() != max_dist)
        throw std::runtime_error("getDistanceField view reports the wrong geometry");
      double gx, gy, gz;
      bool inb;
      const double dc = view.getDistanceGradient(center[0], center[1], center[2], gx, gy, gz, inb);
      if (!inb || dc != view.getDistance(center[0], center[1], center[2]))
        throw std::runtime_error("getDistanceField view: cell query and gradient query disagree");
    }
    std::printf("host_driver ok: %s, %lld states, detector %s, %lld kernel launches, checkCollision %.1f us/call "
                "(%.1f us with contacts), getCollisionGradients %.1f us/call; %zu cache regenerations in the timed calls\n",
                robot_name.c_str(), (long long)n, alloc->getName().c_str(), (long long)b200df_launch_count(), us_plain,
                us_contacts, us_grad, regenerated);
    return 0;
  }
  catch (const std::exception& ex)
  {
    std::fprintf(stderr, "host_driver failed: %s\n", ex.what());
    return 1;
  }
}
