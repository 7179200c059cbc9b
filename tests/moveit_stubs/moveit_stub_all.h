// Declarations-only stand-ins for the MoveIt / Eigen / geometric_shapes / octomap / ROS headers the plugin sources
// include, so that `g++ -fsyntax-only` can check plugin/moveit_collision_detection_b200 in an image without MoveIt.
// Every declaration repeats the member the reference declares (file:line under /root/reference/moveit_core unless
// noted); nothing here is linked or run in the product.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#define MOVEIT_CLASS_FORWARD(C)                                                                                        \
  class C;                                                                                                             \
  typedef std::shared_ptr<C> C##Ptr;                                                                                   \
  typedef std::shared_ptr<const C> C##ConstPtr;

namespace Eigen
{
struct Vector3d
{
  Vector3d();
  Vector3d(double x, double y, double z);
  double x() const;
  double y() const;
  double z() const;
  double operator[](int i) const;
};
struct Isometry3d
{
  double operator()(int row, int col) const;
};
}  // namespace Eigen
namespace EigenSTL
{
typedef std::vector<Eigen::Isometry3d> vector_Isometry3d;
typedef std::vector<Eigen::Vector3d> vector_Vector3d;
}  // namespace EigenSTL

namespace octomap  // octomap/OcTree.h, OcTreeBaseImpl.h (tree_iterator)
{
struct point3d
{
  float x() const;
  float y() const;
  float z() const;
};
struct OcTreeNode
{
};
class OcTree
{
public:
  class tree_iterator
  {
  public:
    point3d getCoordinate() const;
    double getSize() const;
    bool isLeaf() const;
    const OcTreeNode& operator*() const;
    tree_iterator& operator++();
    bool operator!=(const tree_iterator& o) const;
  };
  tree_iterator begin_tree(unsigned char max_depth = 0) const;
  tree_iterator end_tree() const;
  bool isNodeOccupied(const OcTreeNode& n) const;
};
}  // namespace octomap

namespace shapes  // geometric_shapes/shapes.h
{
enum ShapeType
{
  UNKNOWN_SHAPE,
  SPHERE,
  CYLINDER,
  CONE,
  BOX,
  PLANE,
  MESH,
  OCTREE
};
class Shape
{
public:
  virtual ~Shape();
  ShapeType type;
};
typedef std::shared_ptr<const Shape> ShapeConstPtr;
class Sphere : public Shape
{
public:
  double radius;
};
class Cylinder : public Shape
{
public:
  double length, radius;
};
class Box : public Shape
{
public:
  double size[3];
};
class OcTree : public Shape
{
public:
  std::shared_ptr<const octomap::OcTree> octree;
};
}  // namespace shapes

namespace moveit
{
namespace core
{
class LinkModel;
class JointModelGroup;
class JointModel  // robot_model/include/moveit/robot_model/joint_model.h:136-390
{
public:
  enum JointType
  {
    UNKNOWN,
    REVOLUTE,
    PRISMATIC,
    PLANAR,
    FLOATING,
    FIXED
  };
  virtual ~JointModel();
  const std::string& getName() const;
  JointType getType() const;
  std::size_t getVariableCount() const;
  int getFirstVariableIndex() const;
  const JointModel* getMimic() const;
  double getMimicOffset() const;
  double getMimicFactor() const;
};
class RevoluteJointModel : public JointModel  // revolute_joint_model.h:73-82
{
public:
  bool isContinuous() const;
  const Eigen::Vector3d& getAxis() const;
};
class PrismaticJointModel : public JointModel  // prismatic_joint_model.h:71
{
public:
  const Eigen::Vector3d& getAxis() const;
};
class LinkModel  // link_model.h:86-175
{
public:
  const std::string& getName() const;
  int getLinkIndex() const;
  const JointModel* getParentJointModel() const;
  const LinkModel* getParentLinkModel() const;
  const Eigen::Isometry3d& getJointOriginTransform() const;
  const EigenSTL::vector_Isometry3d& getCollisionOriginTransforms() const;
  const std::vector<shapes::ShapeConstPtr>& getShapes() const;
};
class JointModelGroup  // joint_model_group.h:150
{
public:
  const std::string& getName() const;
  const std::vector<std::string>& getJointModelNames() const;
};
class RobotModel  // robot_model.h:166-402
{
public:
  const std::string& getName() const;
  const std::vector<const LinkModel*>& getLinkModels() const;
  const std::vector<const JointModelGroup*>& getJointModelGroups() const;
};
typedef std::shared_ptr<const RobotModel> RobotModelConstPtr;
class AttachedBody  // robot_state/include/moveit/robot_state/attached_body.h:92-131
{
public:
  const std::string& getName() const;
  const std::string& getAttachedLinkName() const;
  const std::vector<shapes::ShapeConstPtr>& getShapes() const;
  const std::set<std::string>& getTouchLinks() const;
  const EigenSTL::vector_Isometry3d& getShapePosesInLinkFrame() const;
};
class RobotState  // robot_state.h:1678
{
public:
  const double* getVariablePositions() const;
  void getAttachedBodies(std::vector<const AttachedBody*>& attached_bodies) const;
};
}  // namespace core
}  // namespace moveit

namespace collision_detection
{
namespace BodyTypes  // collision_detection/include/moveit/collision_detection/collision_common.h:56-70
{
enum Type
{
  ROBOT_LINK,
  ROBOT_ATTACHED,
  WORLD_OBJECT
};
}
using BodyType = BodyTypes::Type;
struct Contact  // collision_common.h:79-120
{
  Eigen::Vector3d pos;
  Eigen::Vector3d normal;
  double depth;
  std::string body_name_1;
  BodyType body_type_1;
  std::string body_name_2;
  BodyType body_type_2;
};
struct CollisionRequest  // collision_common.h:146-196
{
  std::string group_name;
  bool distance;
  bool cost;
  bool contacts;
  std::size_t max_contacts;
  std::size_t max_contacts_per_pair;
  bool verbose;
};
struct CollisionResult  // collision_common.h:350-390
{
  typedef std::map<std::pair<std::string, std::string>, std::vector<Contact>> ContactMap;
  bool collision;
  double distance;
  std::size_t contact_count;
  ContactMap contacts;
};
struct DistanceRequest  // collision_common.h:216-262
{
  std::string group_name;
};
struct DistanceResultsData  // collision_common.h:269-280
{
  double distance;
};
struct DistanceResult  // collision_common.h:320-347
{
  bool collision;
  DistanceResultsData minimum_distance;
};
namespace AllowedCollision  // collision_matrix.h:51-68
{
enum Type
{
  NEVER,
  ALWAYS,
  CONDITIONAL
};
}
class AllowedCollisionMatrix  // collision_matrix.h:102, 185
{
public:
  bool getEntry(const std::string& name1, const std::string& name2, AllowedCollision::Type& allowed_collision) const;
  void getAllEntryNames(std::vector<std::string>& names) const;
};
MOVEIT_CLASS_FORWARD(World);
class World  // collision_detection/include/moveit/collision_detection/world.h:79-318
{
public:
  struct Object
  {
    std::string id_;
    std::vector<shapes::ShapeConstPtr> shapes_;
    EigenSTL::vector_Isometry3d shape_poses_;
    EigenSTL::vector_Isometry3d global_shape_poses_;
  };
  typedef std::shared_ptr<const Object> ObjectConstPtr;
  enum ActionBits
  {
    UNINITIALIZED = 0,
    CREATE = 1,
    DESTROY = 2,
    MOVE_SHAPE = 4,
    ADD_SHAPE = 8,
    REMOVE_SHAPE = 16
  };
  class Action
  {
  public:
    Action();
    Action(int v);
    operator ActionBits() const;
  };
  class ObserverHandle
  {
  public:
    ObserverHandle();
  };
  using ObserverCallbackFn = std::function<void(const ObjectConstPtr&, Action)>;  // boost::function in the reference
  World();
  ObserverHandle addObserver(const ObserverCallbackFn& callback);
  void removeObserver(const ObserverHandle observer_handle);
  void notifyObserverAllObjects(const ObserverHandle observer_handle, Action action) const;
};
MOVEIT_CLASS_FORWARD(CollisionEnv);
class CollisionEnv  // collision_detection/include/moveit/collision_detection/collision_env.h:54-307
{
public:
  using ObjectConstPtr = World::ObjectConstPtr;
  CollisionEnv() = delete;
  CollisionEnv(const moveit::core::RobotModelConstPtr& model, double padding = 0.0, double scale = 1.0);
  CollisionEnv(const moveit::core::RobotModelConstPtr& model, const WorldPtr& world, double padding = 0.0,
               double scale = 1.0);
  CollisionEnv(const CollisionEnv& other, const WorldPtr& world);
  virtual ~CollisionEnv();
  virtual void checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                  const moveit::core::RobotState& state) const = 0;
  virtual void checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                  const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const = 0;
  virtual void checkCollision(const CollisionRequest& req, CollisionResult& res,
                              const moveit::core::RobotState& state) const;
  virtual void checkCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                              const AllowedCollisionMatrix& acm) const;
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state) const = 0;
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const = 0;
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state1, const moveit::core::RobotState& state2,
                                   const AllowedCollisionMatrix& acm) const = 0;
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state1,
                                   const moveit::core::RobotState& state2) const = 0;
  virtual void distanceSelf(const DistanceRequest& req, DistanceResult& res,
                            const moveit::core::RobotState& state) const = 0;
  virtual void distanceRobot(const DistanceRequest& req, DistanceResult& res,
                             const moveit::core::RobotState& state) const = 0;
  virtual void setWorld(const WorldPtr& world);
  const WorldPtr& getWorld();
  const WorldConstPtr& getWorld() const;
  const moveit::core::RobotModelConstPtr& getRobotModel() const;
};
MOVEIT_CLASS_FORWARD(CollisionDetectorAllocator);
class CollisionDetectorAllocator  // collision_detector_allocator.h:46-70
{
public:
  virtual ~CollisionDetectorAllocator();
  virtual const std::string& getName() const = 0;
  virtual CollisionEnvPtr allocateEnv(const WorldPtr& world,
                                      const moveit::core::RobotModelConstPtr& robot_model) const = 0;
  virtual CollisionEnvPtr allocateEnv(const CollisionEnvConstPtr& orig, const WorldPtr& world) const = 0;
  virtual CollisionEnvPtr allocateEnv(const moveit::core::RobotModelConstPtr& robot_model) const = 0;
};
template <class CollisionEnvType, class CollisionDetectorAllocatorType>
class CollisionDetectorAllocatorTemplate : public CollisionDetectorAllocator  // :73-98
{
public:
  CollisionEnvPtr allocateEnv(const WorldPtr& world, const moveit::core::RobotModelConstPtr& robot_model) const override
  {
    return CollisionEnvPtr(new CollisionEnvType(robot_model, world));
  }
  CollisionEnvPtr allocateEnv(const CollisionEnvConstPtr& orig, const WorldPtr& world) const override
  {
    return CollisionEnvPtr(new CollisionEnvType(dynamic_cast<const CollisionEnvType&>(*orig), world));
  }
  CollisionEnvPtr allocateEnv(const moveit::core::RobotModelConstPtr& robot_model) const override
  {
    return CollisionEnvPtr(new CollisionEnvType(robot_model));
  }
  static CollisionDetectorAllocatorPtr create()
  {
    return CollisionDetectorAllocatorPtr(new CollisionDetectorAllocatorType());
  }
};

// collision_distance_field/include/moveit/collision_distance_field/collision_distance_field_types.h:57-262
enum CollisionType
{
  NONE = 0,
  SELF = 1,
  INTRA = 2,
  ENVIRONMENT = 3,
};
struct CollisionSphere
{
  Eigen::Vector3d relative_vec_;
  double radius_;
};
struct GradientInfo
{
  double closest_distance;
  bool collision;
  EigenSTL::vector_Vector3d sphere_locations;
  std::vector<double> distances;
  EigenSTL::vector_Vector3d gradients;
  std::vector<CollisionType> types;
  std::vector<double> sphere_radii;
  std::string joint_name;
};
class BodyDecomposition
{
public:
  BodyDecomposition(const shapes::ShapeConstPtr& shape, double resolution, double padding = 0.01);
  BodyDecomposition(const std::vector<shapes::ShapeConstPtr>& shapes, const EigenSTL::vector_Isometry3d& poses,
                    double resolution, double padding);
  const std::vector<CollisionSphere>& getCollisionSpheres() const;
  const EigenSTL::vector_Vector3d& getCollisionPoints() const;
};
// collision_common_distance_field.h:120-167
struct GroupStateRepresentation
{
  std::vector<GradientInfo> gradients_;
};
typedef std::shared_ptr<GroupStateRepresentation> GroupStateRepresentationPtr;

class CollisionEnvFCL : public CollisionEnv  // collision_detection_fcl/include/.../collision_env_fcl.h:50-130
{
public:
  CollisionEnvFCL(const moveit::core::RobotModelConstPtr& model, double padding = 0.0, double scale = 1.0);
  CollisionEnvFCL(const moveit::core::RobotModelConstPtr& model, const WorldPtr& world, double padding = 0.0,
                  double scale = 1.0);
  CollisionEnvFCL(const CollisionEnvFCL& other, const WorldPtr& world);
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                          const moveit::core::RobotState& state) const override;
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                          const AllowedCollisionMatrix& acm) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                           const moveit::core::RobotState& state) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                           const AllowedCollisionMatrix& acm) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                           const moveit::core::RobotState& state2, const AllowedCollisionMatrix& acm) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                           const moveit::core::RobotState& state2) const override;
  void distanceSelf(const DistanceRequest& req, DistanceResult& res,
                    const moveit::core::RobotState& state) const override;
  void distanceRobot(const DistanceRequest& req, DistanceResult& res,
                     const moveit::core::RobotState& state) const override;
  void setWorld(const WorldPtr& world) override;
};
}  // namespace collision_detection

namespace planning_scene  // planning_scene/include/moveit/planning_scene/planning_scene.h (setActiveCollisionDetector)
{
class PlanningScene
{
public:
  void setActiveCollisionDetector(const collision_detection::CollisionDetectorAllocatorPtr& allocator,
                                  bool exclusive = false);
};
typedef std::shared_ptr<PlanningScene> PlanningScenePtr;
}  // namespace planning_scene

namespace collision_detection
{
class CollisionPlugin  // collision_detection/include/moveit/collision_detection/collision_plugin.h:78-93
{
public:
  CollisionPlugin();
  virtual ~CollisionPlugin();
  virtual bool initialize(const planning_scene::PlanningScenePtr& scene, bool exclusive) const = 0;
};
}  // namespace collision_detection

namespace ros
{
class NodeHandle
{
public:
  explicit NodeHandle(const std::string& ns = std::string());
  bool getParam(const std::string& key, std::string& s) const;
};
}  // namespace ros
inline void ros_log_stub(const char*, const char*, ...)
{
}
#define ROS_ERROR_NAMED(name, ...) ros_log_stub(name, __VA_ARGS__)
#define ROS_WARN_NAMED(name, ...) ros_log_stub(name, __VA_ARGS__)
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type)                                                            \
  static_assert(std::is_base_of<base_class_type, class_type>::value, "plugin class must derive from its base");
