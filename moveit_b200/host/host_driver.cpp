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
  if (argc < 4)
  {
    std::fprintf(stderr, "usage: host_driver <scenario.txt> <states.f64> <out_prefix>\n");
    return 2;
  }
  try
  {
    std::ifstream in(argv[1]);
    if (!in)
      throw std::runtime_error("cannot open scenario");
    in.precision(17);
    std::string tok, robot_name;
    int n_links = 0;
    in >> tok >> robot_name >> n_links;
    auto model = std::make_shared<core::RobotModel>(robot_name);
    CollisionEnvB200::LinkSpheres overrides;
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
    int use_acm = 0, n_disabled = 0;
    in >> tok >> use_acm >> n_disabled;
    AllowedCollisionMatrix acm(model->getLinkModelNames(), false);
    for (int k = 0; k < n_disabled; ++k)
    {
      std::string a, b;
      in >> a >> b;
      acm.setEntry(a, b, true);
    }
    double size[3], center[3], res, max_dist, tol;
    int use_signed;
    in >> tok >> size[0] >> size[1] >> size[2] >> center[0] >> center[1] >> center[2] >> res >> max_dist >>
        use_signed >> tol;
    int n_objects = 0;
    in >> tok >> n_objects;
    auto world = std::make_shared<World>();
    std::vector<std::string> object_ids;
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
    std::vector<core::AttachedBody> attached;
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
    std::string group_name;
    int n_single = 0, n_contacts = 0, n_grad = 0, mutate = 0;
    std::size_t max_contacts = 1, max_per_pair = 1;
    in >> tok >> group_name >> n_single >> n_contacts >> max_contacts >> max_per_pair >> n_grad >> mutate;
    if (group_name == "-")
      group_name.clear();
    if (!in)
      throw std::runtime_error("scenario parse error");

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
    const std::string prefix = argv[3];
    const AllowedCollisionMatrix* acm_ptr = use_acm ? &acm : nullptr;
    CollisionRequest req;
    req.group_name = group_name;
    core::RobotState ref_state(model);  // all-zero scene state poses the non-group links
    for (const core::AttachedBody& b : attached)
      ref_state.attachBody(b);
    std::vector<std::uint8_t> v(n);
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
