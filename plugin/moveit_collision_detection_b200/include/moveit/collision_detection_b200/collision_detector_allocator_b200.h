// Allocators for CollisionEnvB200 (and the hybrid env CHOMP asks for), next to CollisionDetectorAllocatorFCL /
// Bullet / DistanceField / Hybrid.  Reference: moveit_core/collision_detection/include/moveit/collision_detection/
// collision_detector_allocator.h:46-98, collision_distance_field/include/moveit/collision_distance_field/
// collision_detector_allocator_distance_field.h, collision_detector_allocator_hybrid.h.
#pragma once

#include <moveit/collision_detection/collision_detector_allocator.h>
#include <moveit/collision_detection_b200/collision_env_b200.h>
#include <moveit/collision_detection_b200/collision_env_hybrid_b200.h>

namespace collision_detection
{
class CollisionDetectorAllocatorB200
  : public CollisionDetectorAllocatorTemplate<CollisionEnvB200, CollisionDetectorAllocatorB200>
{
public:
  const std::string& getName() const override
  {
    static const std::string NAME("B200DistanceField");
    return NAME;
  }
};

// The same allocator carrying a configuration read from the plugin's parameter namespace: allocateEnv passes the
// configured field size / origin / resolution / max distance / signed flag / link_spheres / devices where the template
// above passes the defaults of collision_env_distance_field.h:48-54.
class CollisionDetectorAllocatorB200Configured : public CollisionDetectorAllocator
{
public:
  explicit CollisionDetectorAllocatorB200Configured(const CollisionEnvB200::Config& config) : config_(config)
  {
  }
  const std::string& getName() const override
  {
    static const std::string NAME("B200DistanceField");
    return NAME;
  }
  CollisionEnvPtr allocateEnv(const WorldPtr& world, const moveit::core::RobotModelConstPtr& robot_model) const override
  {
    return CollisionEnvPtr(new CollisionEnvB200(robot_model, world, config_));
  }
  CollisionEnvPtr allocateEnv(const CollisionEnvConstPtr& orig, const WorldPtr& world) const override
  {
    return CollisionEnvPtr(new CollisionEnvB200(dynamic_cast<const CollisionEnvB200&>(*orig), world));
  }
  CollisionEnvPtr allocateEnv(const moveit::core::RobotModelConstPtr& robot_model) const override
  {
    return CollisionEnvPtr(new CollisionEnvB200(robot_model, WorldPtr(new World()), config_));
  }
  static CollisionDetectorAllocatorPtr create(const CollisionEnvB200::Config& config)
  {
    return CollisionDetectorAllocatorPtr(new CollisionDetectorAllocatorB200Configured(config));
  }

private:
  CollisionEnvB200::Config config_;
};

// what chomp_interface / chomp_optimizer.cpp:75-90 look for by type: an env that is FCL for the CollisionEnv virtuals
// and a distance field for getCollisionGradients
class CollisionDetectorAllocatorHybridB200
  : public CollisionDetectorAllocatorTemplate<CollisionEnvHybridB200, CollisionDetectorAllocatorHybridB200>
{
public:
  const std::string& getName() const override
  {
    static const std::string NAME("HYBRID_B200");
    return NAME;
  }
};
}  // namespace collision_detection
