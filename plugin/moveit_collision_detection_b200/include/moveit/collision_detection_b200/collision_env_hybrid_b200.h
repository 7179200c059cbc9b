// CollisionEnvHybrid with the distance-field member on the GPU.
//
// Reference: moveit_core/collision_distance_field/include/moveit/collision_distance_field/collision_env_hybrid.h:50-151
// and src/collision_env_hybrid.cpp:47-75 (constructors), :77-172 (the *DistanceField checks), :174-179 (setWorld),
// :181-230 (getCollisionGradients / getAllCollisions).  An FCL env answers the CollisionEnv virtuals (meshes, exact
// contacts); every *DistanceField method and getCollisionGradients is forwarded to cenv_distance_, which here is a
// CollisionEnvB200.  ChompOptimizer (chomp_optimizer.cpp:75-90, 895-965) reaches the path through this class only.
#pragma once

#include <moveit/collision_detection_fcl/collision_env_fcl.h>
#include <moveit/collision_detection_b200/collision_env_b200.h>

namespace collision_detection
{
class CollisionEnvHybridB200 : public CollisionEnvFCL
{
public:
  explicit CollisionEnvHybridB200(const moveit::core::RobotModelConstPtr& robot_model, double padding = 0.0,
                                  double scale = 1.0)
    : CollisionEnvFCL(robot_model, padding, scale), cenv_distance_(new CollisionEnvB200(robot_model, getWorld(), padding, scale))
  {
  }
  CollisionEnvHybridB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                         double padding = 0.0, double scale = 1.0)
    : CollisionEnvFCL(robot_model, world, padding, scale)
    , cenv_distance_(new CollisionEnvB200(robot_model, getWorld(), padding, scale))
  {
  }
  CollisionEnvHybridB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                         const CollisionEnvB200::Config& config)
    : CollisionEnvFCL(robot_model, world, config.padding, config.scale)
    , cenv_distance_(new CollisionEnvB200(robot_model, getWorld(), config))
  {
  }
  CollisionEnvHybridB200(const CollisionEnvHybridB200& other, const WorldPtr& world)
    : CollisionEnvFCL(other, world), cenv_distance_(new CollisionEnvB200(*other.cenv_distance_, world))
  {
  }

  void checkSelfCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                       const moveit::core::RobotState& state) const
  {
    cenv_distance_->checkSelfCollision(req, res, state);
  }
  void checkSelfCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                       const moveit::core::RobotState& state, GroupStateRepresentationPtr& gsr) const
  {
    // the gsr overloads of the reference hand the representation to the caller: getCollisionGradients fills the same
    // structure from the same cache entry
    cenv_distance_->checkSelfCollision(req, res, state);
    cenv_distance_->getCollisionGradients(req, res, state, nullptr, gsr);
  }
  void checkSelfCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                       const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
  {
    cenv_distance_->checkSelfCollision(req, res, state, acm);
  }
  void checkSelfCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                       const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm,
                                       GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkSelfCollision(req, res, state, acm);
    cenv_distance_->getCollisionGradients(req, res, state, &acm, gsr);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state) const
  {
    cenv_distance_->checkCollision(req, res, state);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state, GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkCollision(req, res, state);
    cenv_distance_->getCollisionGradients(req, res, state, nullptr, gsr);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
  {
    cenv_distance_->checkCollision(req, res, state, acm);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm,
                                   GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkCollision(req, res, state, acm);
    cenv_distance_->getCollisionGradients(req, res, state, &acm, gsr);
  }
  void checkRobotCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                        const moveit::core::RobotState& state) const
  {
    cenv_distance_->checkRobotCollision(req, res, state);
  }
  void checkRobotCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                        const moveit::core::RobotState& state, GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkRobotCollision(req, res, state);
    cenv_distance_->getCollisionGradients(req, res, state, nullptr, gsr);
  }
  void checkRobotCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                        const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
  {
    cenv_distance_->checkRobotCollision(req, res, state, acm);
  }
  void checkRobotCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                        const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm,
                                        GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkRobotCollision(req, res, state, acm);
    cenv_distance_->getCollisionGradients(req, res, state, &acm, gsr);
  }

  void setWorld(const WorldPtr& world) override  // collision_env_hybrid.cpp:174-179
  {
    if (world == getWorld())
      return;
    cenv_distance_->setWorld(world);
    CollisionEnvFCL::setWorld(world);
  }

  void getCollisionGradients(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                             const AllowedCollisionMatrix* acm, GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->getCollisionGradients(req, res, state, acm, gsr);
  }
  void getAllCollisions(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                        const AllowedCollisionMatrix* acm, GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->getAllCollisions(req, res, state, acm, gsr);
  }
  const CollisionEnvB200ConstPtr getCollisionRobotDistanceField() const
  {
    return cenv_distance_;
  }
  const CollisionEnvB200ConstPtr getCollisionWorldDistanceField() const
  {
    return cenv_distance_;
  }

protected:
  CollisionEnvB200Ptr cenv_distance_;
};
}  // namespace collision_detection
