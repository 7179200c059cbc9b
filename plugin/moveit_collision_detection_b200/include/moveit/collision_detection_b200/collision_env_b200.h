// collision_detection::CollisionEnvB200 - MoveIt's distance-field collision environment on an NVIDIA B200.
//
// A CollisionEnv (moveit_core/collision_detection/include/moveit/collision_detection/collision_env.h:54-307) that
// answers like CollisionEnvDistanceField (moveit_core/collision_distance_field/include/moveit/
// collision_distance_field/collision_env_distance_field.h:57-322) and adds batched entry points.  It holds the host
// mirror collision_detection_b200::CollisionEnvB200 (moveit_b200/host/collision_env_b200.hpp, the class
// tests/test_host_mirror.py checks against the CPU oracle), which calls libb200df.so through include/b200df.h.
#pragma once

#include <moveit/collision_detection/collision_env.h>
#include <moveit/collision_distance_field/collision_common_distance_field.h>
#include <moveit/collision_distance_field/collision_env_distance_field.h>  // DEFAULT_* constants
#include <moveit/collision_detection_b200/moveit_adapters.h>

#include <b200_config.hpp>  // moveit_b200/host

#include <mutex>

namespace collision_detection
{
MOVEIT_CLASS_FORWARD(CollisionEnvB200);

class CollisionEnvB200 : public CollisionEnv
{
public:
  using Config = collision_detection_b200::B200DistanceFieldConfig;

  // the three constructors CollisionDetectorAllocatorTemplate uses (collision_detector_allocator.h:73-88)
  explicit CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, double padding = 0.0,
                            double scale = 1.0);
  CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world, double padding = 0.0,
                   double scale = 1.0);
  CollisionEnvB200(const CollisionEnvB200& other, const WorldPtr& world);
  // the configured form (field size / origin / resolution / max distance / signed flag / link_spheres / devices),
  // collision_env_distance_field.h:62-84
  CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world, const Config& config);
  ~CollisionEnvB200() override;

  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                          const moveit::core::RobotState& state) const override;
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                          const AllowedCollisionMatrix& acm) const override;
  void checkCollision(const CollisionRequest& req, CollisionResult& res,
                      const moveit::core::RobotState& state) const override;
  void checkCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                      const AllowedCollisionMatrix& acm) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                           const moveit::core::RobotState& state) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                           const AllowedCollisionMatrix& acm) const override;
  // continuous checks: "not implemented" is logged, like collision_env_distance_field.cpp:1521-1534
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                           const moveit::core::RobotState& state2, const AllowedCollisionMatrix& acm) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                           const moveit::core::RobotState& state2) const override;
  void distanceSelf(const DistanceRequest& req, DistanceResult& res,
                    const moveit::core::RobotState& state) const override;
  void distanceRobot(const DistanceRequest& req, DistanceResult& res,
                     const moveit::core::RobotState& state) const override;
  void setWorld(const WorldPtr& world) override;

  // what CollisionEnvHybrid forwards to its distance-field member (collision_env_hybrid.cpp:174-230)
  void getCollisionGradients(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                             const AllowedCollisionMatrix* acm, GroupStateRepresentationPtr& gsr) const;
  void getAllCollisions(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                        const AllowedCollisionMatrix* acm, GroupStateRepresentationPtr& gsr) const;

  // NEW: n states (rows of RobotState::getVariablePositions()) in one launch.  verdicts[i] is what
  // checkCollision(req, res, state_i, acm) would leave in res.collision; reference_state supplies what is not in the
  // rows (attached bodies).  False = the call failed and every verdict is 1 (fail-closed).
  bool checkCollisionBatch(const CollisionRequest& req, const double* states, std::int64_t n,
                           const AllowedCollisionMatrix* acm, const moveit::core::RobotState& reference_state,
                           std::uint8_t* verdicts) const;
  // NEW: m motions from[i] -> to[i], `segments` interpolated states each (JointModel::interpolate arithmetic on the
  // device); first_invalid[i] = index of the first colliding interpolated state, -1 = the motion is free
  bool checkMotionBatch(const CollisionRequest& req, const double* from, const double* to, std::int64_t m,
                        int segments, const AllowedCollisionMatrix* acm,
                        const moveit::core::RobotState& reference_state, std::int32_t* first_invalid) const;

  const collision_detection_b200::CollisionEnvB200& mirror() const
  {
    return *impl_;
  }

private:
  void construct(const Config& config);
  void notifyObjectChange(const ObjectConstPtr& obj, World::Action action);
  collision_detection_b200::core::RobotState mirrorState(const moveit::core::RobotState& state) const;

  Config config_;
  collision_detection_b200::core::RobotModelConstPtr mirror_model_;
  collision_detection_b200::WorldPtr mirror_world_;
  std::unique_ptr<collision_detection_b200::CollisionEnvB200> impl_;
  World::ObserverHandle observer_handle_;
  // the mirror GroupStateRepresentation behind each MoveIt one handed out (its dfce_ keeps the ACM of the first call,
  // chomp_optimizer.cpp:102 vs :924)
  mutable std::mutex gsr_lock_;
  mutable std::map<const GroupStateRepresentation*, collision_detection_b200::GroupStateRepresentationPtr> gsr_;
};
}  // namespace collision_detection
