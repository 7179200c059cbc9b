#include <moveit/collision_detection_b200/collision_env_b200.h>

#include <ros/console.h>

#include <functional>

namespace collision_detection
{
#define B200_COMMA ,
namespace ad = b200_adapters;
namespace mirror = collision_detection_b200;

namespace
{
const char* const LOGNAME = "collision_detection.b200";

CollisionEnvB200::Config defaultConfig(double padding, double scale)
{
  CollisionEnvB200::Config c;  // collision_env_distance_field.h:48-54 defaults
  c.padding = padding;
  c.scale = scale;
  return c;
}

// the mirror records every failure (and has already answered fail-closed); this surfaces it where MoveIt users look
void report(const mirror::CollisionEnvB200& env, std::size_t errors_before)
{
  if (env.errorCount() != errors_before)
    ROS_ERROR_NAMED(LOGNAME, "%s", env.lastError().c_str());
}
}  // namespace

CollisionEnvB200::CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, double padding, double scale)
  : CollisionEnv(robot_model, padding, scale)
{
  construct(defaultConfig(padding, scale));
}

CollisionEnvB200::CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                                   double padding, double scale)
  : CollisionEnv(robot_model, world, padding, scale)
{
  construct(defaultConfig(padding, scale));
}

CollisionEnvB200::CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                                   const Config& config)
  : CollisionEnv(robot_model, world, config.padding, config.scale)
{
  construct(config);
}

CollisionEnvB200::CollisionEnvB200(const CollisionEnvB200& other, const WorldPtr& world) : CollisionEnv(other, world)
{
  construct(other.config_);
}

CollisionEnvB200::~CollisionEnvB200()
{
  getWorld()->removeObserver(observer_handle_);
}

void CollisionEnvB200::construct(const Config& config)
{
  config_ = config;
  // mesh links: MoveIt's own sphere decomposition becomes a link_body_decompositions override unless the
  // configuration already names the link (collision_env_distance_field.cpp:1085-1088)
  mirror::CollisionEnvB200::LinkSpheres mesh_spheres;
  mirror_model_ = ad::toMirror(*getRobotModel(), config_.resolution, config_.padding, mesh_spheres);
  mirror::CollisionEnvB200::LinkSpheres link_spheres = config_.link_spheres;
  link_spheres.insert(mesh_spheres.begin(), mesh_spheres.end());
  mirror_world_ = std::make_shared<mirror::World>();
  impl_.reset(new mirror::CollisionEnvB200(mirror_model_, mirror_world_, link_spheres, config_.size_x, config_.size_y,
                                           config_.size_z, config_.origin, config_.use_signed_distance_field,
                                           config_.resolution, config_.collision_tolerance,
                                           config_.max_propogation_distance, config_.padding, config_.scale));
  impl_->setOctreeOccupiedLeavesOnly(config_.octree_occupied_leaves_only);
  if (!config_.devices.empty())
    impl_->setDevices(config_.devices);
  report(*impl_, 0);
  // collision_env_distance_field.cpp:88-92: observe the world, then replay what it already holds
  observer_handle_ = getWorld()->addObserver(
      std::bind(&CollisionEnvB200::notifyObjectChange, this, std::placeholders::_1, std::placeholders::_2));
  getWorld()->notifyObserverAllObjects(observer_handle_, World::CREATE);
}

void CollisionEnvB200::setWorld(const WorldPtr& world)
{
  if (world == getWorld())
    return;
  // collision_env_distance_field.cpp:1702-1721: drop the old observer, start from an empty field, replay
  getWorld()->removeObserver(observer_handle_);
  CollisionEnv::setWorld(world);
  for (const std::string& id : mirror_world_->getObjectIds())
    mirror_world_->removeObject(id);
  observer_handle_ = getWorld()->addObserver(
      std::bind(&CollisionEnvB200::notifyObjectChange, this, std::placeholders::_1, std::placeholders::_2));
  getWorld()->notifyObserverAllObjects(observer_handle_, World::CREATE);
}

// collision_env_distance_field.cpp:1723-1747.  The mirror's own observer does the field bookkeeping (which points
// leave, which arrive, quirk Q8); here the MoveIt object is re-expressed in the mirror world.
void CollisionEnvB200::notifyObjectChange(const ObjectConstPtr& obj, World::Action action)
{
  const std::size_t errors = impl_->errorCount();
  if (mirror_world_->getObject(obj->id_))
    mirror_world_->removeObject(obj->id_);
  if (action != World::DESTROY)
  {
    for (std::size_t i = 0; i < obj->shapes_.size(); ++i)
    {
      // world shapes are rasterised with the default padding 0.01 (collision_distance_field_types.h:227, quirk Q10)
      auto shape = std::make_shared<mirror::shapes::Shape>(ad::toMirror(obj->shapes_[i], config_.resolution, 0.01));
      mirror_world_->addToObject(obj->id_, shape, ad::toMirror(obj->global_shape_poses_[i]));
    }
  }
  report(*impl_, errors);
}

mirror::core::RobotState CollisionEnvB200::mirrorState(const moveit::core::RobotState& state) const
{
  return ad::toMirror(state, mirror_model_, config_.resolution);
}

#define B200_FORWARD(method, acm_arg)                                                                                \
  const std::size_t errors = impl_->errorCount();                                                                    \
  mirror::CollisionResult r;                                                                                         \
  impl_->method(ad::toMirror(req), r, mirrorState(state) acm_arg);                                                    \
  ad::mergeResult(r, res);                                                                                           \
  report(*impl_, errors)

void CollisionEnvB200::checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                          const moveit::core::RobotState& state) const
{
  B200_FORWARD(checkSelfCollision, );
}

void CollisionEnvB200::checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                          const moveit::core::RobotState& state,
                                          const AllowedCollisionMatrix& acm) const
{
  const mirror::AllowedCollisionMatrix m = ad::toMirror(acm);
  B200_FORWARD(checkSelfCollision, B200_COMMA m);
}

void CollisionEnvB200::checkCollision(const CollisionRequest& req, CollisionResult& res,
                                      const moveit::core::RobotState& state) const
{
  B200_FORWARD(checkCollision, );
}

void CollisionEnvB200::checkCollision(const CollisionRequest& req, CollisionResult& res,
                                      const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
{
  const mirror::AllowedCollisionMatrix m = ad::toMirror(acm);
  B200_FORWARD(checkCollision, B200_COMMA m);
}

void CollisionEnvB200::checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                           const moveit::core::RobotState& state) const
{
  B200_FORWARD(checkRobotCollision, );
}

void CollisionEnvB200::checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                           const moveit::core::RobotState& state,
                                           const AllowedCollisionMatrix& acm) const
{
  const mirror::AllowedCollisionMatrix m = ad::toMirror(acm);
  B200_FORWARD(checkRobotCollision, B200_COMMA m);
}

void CollisionEnvB200::checkRobotCollision(const CollisionRequest& /*req*/, CollisionResult& /*res*/,
                                           const moveit::core::RobotState& /*state1*/,
                                           const moveit::core::RobotState& /*state2*/,
                                           const AllowedCollisionMatrix& /*acm*/) const
{
  ROS_ERROR_NAMED(LOGNAME, "Continuous collision checking not implemented");
}

void CollisionEnvB200::checkRobotCollision(const CollisionRequest& /*req*/, CollisionResult& /*res*/,
                                           const moveit::core::RobotState& /*state1*/,
                                           const moveit::core::RobotState& /*state2*/) const
{
  ROS_ERROR_NAMED(LOGNAME, "Continuous collision checking not implemented");
}

// stubs in the reference (collision_env_distance_field.h:108-123, 178-198); here the smallest sphere-to-obstacle
// clearance of the state is reported
void CollisionEnvB200::distanceSelf(const DistanceRequest& /*req*/, DistanceResult& /*res*/,
                                    const moveit::core::RobotState& /*state*/) const
{
  ROS_ERROR_NAMED(LOGNAME, "Collision distance to self not implemented yet.");
}

void CollisionEnvB200::distanceRobot(const DistanceRequest& req, DistanceResult& res,
                                     const moveit::core::RobotState& state) const
{
  const std::size_t errors = impl_->errorCount();
  res.minimum_distance.distance = impl_->distanceRobot(mirrorState(state), req.group_name);
  res.collision = res.minimum_distance.distance < 0.0;
  report(*impl_, errors);
}

void CollisionEnvB200::getCollisionGradients(const CollisionRequest& req, CollisionResult& res,
                                             const moveit::core::RobotState& state, const AllowedCollisionMatrix* acm,
                                             GroupStateRepresentationPtr& gsr) const
{
  const std::size_t errors = impl_->errorCount();
  mirror::GroupStateRepresentationPtr mg;
  {
    std::lock_guard<std::mutex> lk(gsr_lock_);
    if (gsr)
      mg = gsr_[gsr.get()];
  }
  mirror::AllowedCollisionMatrix m;
  if (acm)
    m = ad::toMirror(*acm);
  mirror::CollisionResult r;
  impl_->getCollisionGradients(ad::toMirror(req), r, mirrorState(state), acm ? &m : nullptr, mg);
  if (!gsr)
    gsr = std::make_shared<GroupStateRepresentation>();
  if (mg)
  {
    ad::fillGradients(*mg, *gsr);
    std::lock_guard<std::mutex> lk(gsr_lock_);
    if (gsr_.size() > 64)  // representations CHOMP has dropped
      gsr_.clear();
    gsr_[gsr.get()] = mg;
  }
  ad::mergeResult(r, res);
  report(*impl_, errors);
}

void CollisionEnvB200::getAllCollisions(const CollisionRequest& req, CollisionResult& res,
                                        const moveit::core::RobotState& state, const AllowedCollisionMatrix* acm,
                                        GroupStateRepresentationPtr& /*gsr*/) const
{
  const std::size_t errors = impl_->errorCount();
  mirror::AllowedCollisionMatrix m;
  if (acm)
    m = ad::toMirror(*acm);
  mirror::CollisionResult r;
  impl_->getAllCollisions(ad::toMirror(req), r, mirrorState(state), acm ? &m : nullptr);
  ad::mergeResult(r, res);
  report(*impl_, errors);
}

bool CollisionEnvB200::checkCollisionBatch(const CollisionRequest& req, const double* states, std::int64_t n,
                                           const AllowedCollisionMatrix* acm,
                                           const moveit::core::RobotState& reference_state,
                                           std::uint8_t* verdicts) const
{
  const std::size_t errors = impl_->errorCount();
  mirror::AllowedCollisionMatrix m;
  if (acm)
    m = ad::toMirror(*acm);
  const mirror::core::RobotState ref = mirrorState(reference_state);
  const bool ok = impl_->checkCollisionBatch(ad::toMirror(req), states, n, acm ? &m : nullptr,
                                             B200DF_CHECK_SELF | B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT, ref,
                                             verdicts);
  report(*impl_, errors);
  return ok;
}

bool CollisionEnvB200::checkMotionBatch(const CollisionRequest& req, const double* from, const double* to,
                                        std::int64_t m, int segments, const AllowedCollisionMatrix* acm,
                                        const moveit::core::RobotState& reference_state,
                                        std::int32_t* first_invalid) const
{
  const std::size_t errors = impl_->errorCount();
  mirror::AllowedCollisionMatrix ma;
  if (acm)
    ma = ad::toMirror(*acm);
  const mirror::core::RobotState ref = mirrorState(reference_state);
  const bool ok = impl_->checkMotionBatch(ad::toMirror(req), from, to, m, segments, acm ? &ma : nullptr, ref,
                                          first_invalid);
  report(*impl_, errors);
  return ok;
}
}  // namespace collision_detection
