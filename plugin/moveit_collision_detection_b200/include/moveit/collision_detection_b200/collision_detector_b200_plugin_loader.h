// pluginlib entry: mirrors moveit_core/collision_detection_fcl/include/moveit/collision_detection_fcl/
// collision_detector_fcl_plugin_loader.h:42-52 (contract: collision_plugin.h:78-93).
#pragma once

#include <moveit/collision_detection/collision_plugin.h>
#include <moveit/collision_detection_b200/collision_detector_allocator_b200.h>

namespace collision_detection
{
class CollisionDetectorB200PluginLoader : public CollisionPlugin
{
public:
  bool initialize(const planning_scene::PlanningScenePtr& scene, bool exclusive) const override;
};
}  // namespace collision_detection
