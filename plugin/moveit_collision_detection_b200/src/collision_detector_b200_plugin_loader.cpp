// Mirrors moveit_core/collision_detection_fcl/src/collision_detector_fcl_plugin_loader.cpp:42-49; additionally reads
// the plugin's configuration (field size, origin, resolution, max distance, signed flag, link_spheres in the format of
// moveit_experimental/collision_distance_field_ros/.../collision_distance_field_ros_helpers.h:46-120, devices) from
// the private namespace "~b200_distance_field" when it is there.
#include <moveit/collision_detection_b200/collision_detector_b200_plugin_loader.h>
#include <moveit/planning_scene/planning_scene.h>
#include <pluginlib/class_list_macros.hpp>
#include <ros/ros.h>

#include <sstream>

namespace collision_detection
{
bool CollisionDetectorB200PluginLoader::initialize(const planning_scene::PlanningScenePtr& scene, bool exclusive) const
{
  ros::NodeHandle nh("~");
  std::string yaml;
  // the namespace as one YAML document (rosparam dump form); the same reader the host tests drive
  // (collision_detection_b200::configFromYaml, tests/test_host_mirror.py)
  if (nh.getParam("b200_distance_field_yaml", yaml) && !yaml.empty())
  {
    std::istringstream in(yaml);
    const CollisionEnvB200::Config config = collision_detection_b200::configFromYaml(in);
    for (const std::string& w : config.warnings)
      ROS_WARN_NAMED("collision_detection.b200", "%s", w.c_str());
    scene->setActiveCollisionDetector(CollisionDetectorAllocatorB200Configured::create(config), exclusive);
  }
  else
    scene->setActiveCollisionDetector(CollisionDetectorAllocatorB200::create(), exclusive);
  return true;
}
}  // namespace collision_detection

PLUGINLIB_EXPORT_CLASS(collision_detection::CollisionDetectorB200PluginLoader, collision_detection::CollisionPlugin)
