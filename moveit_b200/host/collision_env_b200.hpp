// CollisionEnvB200 — host-side C++ mirror of collision_detection::CollisionEnvDistanceField over the b200df C ABI.
//
// This is the code a MoveIt maintainer drops behind the plugin loader: same class and method
// names, argument meaning and result conventions as the reference, every GPU interaction through include/b200df.h.
// MoveIt's own headers (Eigen, RobotModel/RobotState, World, pluginlib ...) are not in this image, so the types the
// reference takes from moveit_core are small stand-ins with the same member names.  Swapping them for the real
// headers is mechanical; nothing below depends on more than the members declared here.
//
// Reference (paths relative to /root/reference/moveit_core):
//   collision_distance_field/include/moveit/collision_distance_field/collision_env_distance_field.h:48-303
//   collision_distance_field/src/collision_env_distance_field.cpp   (cited per method)
//   collision_distance_field/include/.../collision_common_distance_field.h:56-167   DistanceFieldCacheEntry, GSR
//   collision_distance_field/include/.../collision_distance_field_types.h:57-105      CollisionType, GradientInfo
//   collision_detection/include/moveit/collision_detection/collision_common.h:146-390 request / result / contact
//   collision_detection/src/collision_matrix.cpp:113-145                              AllowedCollisionMatrix
//   collision_detection/include/moveit/collision_detection/world.h:305-318            World observers
//   collision_detection/include/.../collision_detector_allocator.h:46-98              allocator template
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200df.h"

namespace collision_detection_b200
{
// ------------------------------------------------------------------------------------------------------------
// stand-ins for moveit_core / Eigen / geometric_shapes types
// ------------------------------------------------------------------------------------------------------------
struct Vector3d
{
  double x = 0.0, y = 0.0, z = 0.0;
};

struct Isometry3d  // row-major 3x4, implied last row 0 0 0 1
{
  double m[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
  static Isometry3d Identity()
  {
    return Isometry3d();
  }
  bool isIdentity() const
  {
    static const double id[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
    return std::memcmp(m, id, sizeof(m)) == 0;
  }
  // Isometry3d * Vector3d the way Eigen evaluates it: t + ((r0*x + r1*y) + r2*z)
  Vector3d operator*(const Vector3d& p) const
  {
    Vector3d r;
    r.x = m[3] + ((m[0] * p.x + m[1] * p.y) + m[2] * p.z);
    r.y = m[7] + ((m[4] * p.x + m[5] * p.y) + m[6] * p.z);
    r.z = m[11] + ((m[8] * p.x + m[9] * p.y) + m[10] * p.z);
    return r;
  }
};

namespace shapes
{
enum ShapeType  // geometric_shapes/shapes.h values
{
  SPHERE = B200DF_SHAPE_SPHERE,
  CYLINDER = B200DF_SHAPE_CYLINDER,
  BOX = B200DF_SHAPE_BOX,
  OCTREE = 8,        // shapes::OcTree: the <octomap> world object PlanningScene adds (planning_scene.cpp:1390-1412)
  POINT_CLOUD = 100, // not a geometric_shapes type: a body-frame point list (how the plugin hands over shapes it
                     // decomposed on the host with MoveIt's own BodyDecomposition, e.g. meshes)
};
// What the path reads of an octomap::OcTree: every node of the tree in begin_tree() order (all depths, inner nodes
// included), each with its centre, edge length, occupancy (isNodeOccupied) and whether it is a leaf.
struct OcTreeNodes
{
  std::vector<double> xyzs;             // [n][4] centre + size
  std::vector<std::uint8_t> occupied;   // [n]
  std::vector<std::uint8_t> leaf;       // [n]
};
struct Shape
{
  ShapeType type = BOX;
  double dims[3] = { 0, 0, 0 };  // sphere: radius; cylinder: radius, length; box: x, y, z
  std::shared_ptr<const OcTreeNodes> octree;           // OCTREE
  std::shared_ptr<const std::vector<double>> points;   // POINT_CLOUD: [n][3], body frame
};
using ShapeConstPtr = std::shared_ptr<const Shape>;
}  // namespace shapes

namespace core
{
enum JointType
{
  FIXED = B200DF_JOINT_FIXED,
  REVOLUTE = B200DF_JOINT_REVOLUTE,
  PRISMATIC = B200DF_JOINT_PRISMATIC,
  PLANAR = B200DF_JOINT_PLANAR,
  FLOATING = B200DF_JOINT_FLOATING,
};

struct LinkModel  // a link together with its parent joint, as RobotModel stores them in DFS order
{
  std::string name, joint_name;
  int parent = -1;
  JointType joint_type = FIXED;
  bool continuous = false;  // RevoluteJointModel::isContinuous(): interpolation takes the shortest arc
  double axis[3] = { 0, 0, 1 };
  Isometry3d joint_origin;
  std::string mimic_joint;  // empty: none
  double mimic_factor = 1.0, mimic_offset = 0.0;
  std::vector<shapes::Shape> shapes;
  std::vector<Isometry3d> shape_origins;
  int first_variable = -1;
};

class RobotModel
{
public:
  explicit RobotModel(std::string name) : name_(std::move(name))
  {
  }
  // links must be added parent before child (RobotModel::buildModel order, robot_model.cpp:783-838)
  int addLink(const LinkModel& l)
  {
    static const int nv[5] = { 0, 1, 1, 3, 7 };
    LinkModel c = l;
    if (c.parent >= (int)links_.size())
      throw std::invalid_argument("parent link must be added before its child");
    c.first_variable = nv[c.joint_type] ? variable_count_ : -1;
    variable_count_ += nv[c.joint_type];
    link_index_[c.name] = (int)links_.size();
    joint_index_[c.joint_name] = (int)links_.size();
    links_.push_back(c);
    return (int)links_.size() - 1;
  }
  void addJointModelGroup(const std::string& group, const std::vector<std::string>& joints)
  {
    groups_[group] = joints;
  }
  bool hasJointModelGroup(const std::string& g) const
  {
    return groups_.count(g) != 0;
  }
  const std::vector<LinkModel>& getLinkModels() const
  {
    return links_;
  }
  std::vector<std::string> getLinkModelNames() const
  {
    std::vector<std::string> n;
    for (const LinkModel& l : links_)
      n.push_back(l.name);
    return n;
  }
  int getLinkIndex(const std::string& n) const
  {
    auto it = link_index_.find(n);
    return it == link_index_.end() ? -1 : it->second;
  }
  int getJointIndex(const std::string& n) const
  {
    auto it = joint_index_.find(n);
    return it == joint_index_.end() ? -1 : it->second;
  }
  int getVariableCount() const
  {
    return variable_count_;
  }
  // JointModelGroup::getUpdatedLinkModelNames (joint_model_group.cpp:223-247): every link moved by a group joint
  std::vector<int> getUpdatedLinkModelIndices(const std::string& group) const
  {
    const std::vector<std::string>& joints = groups_.at(group);
    std::set<int> moved;
    for (int i = 0; i < (int)links_.size(); ++i)
      if (std::find(joints.begin(), joints.end(), links_[i].joint_name) != joints.end() ||
          (links_[i].parent >= 0 && moved.count(links_[i].parent)))
        moved.insert(i);
    return std::vector<int>(moved.begin(), moved.end());
  }
  // the variables the group's active joints own (generateDistanceFieldCacheEntry :881-896 "updated_map")
  std::set<int> getGroupVariableIndices(const std::string& group) const
  {
    static const int nv[5] = { 0, 1, 1, 3, 7 };
    std::set<int> v;
    for (const LinkModel& l : links_)
    {
      const bool in = group.empty() || std::find(groups_.at(group).begin(), groups_.at(group).end(), l.joint_name) !=
                                           groups_.at(group).end();
      if (in && l.mimic_joint.empty())
        for (int k = 0; k < nv[l.joint_type]; ++k)
          v.insert(l.first_variable + k);
    }
    return v;
  }
  const std::string& getName() const
  {
    return name_;
  }

private:
  std::string name_;
  std::vector<LinkModel> links_;
  std::map<std::string, int> link_index_, joint_index_;
  std::map<std::string, std::vector<std::string>> groups_;
  int variable_count_ = 0;
};
using RobotModelConstPtr = std::shared_ptr<const RobotModel>;

struct AttachedBody  // moveit::core::AttachedBody: shapes rigidly attached to a link
{
  std::string name, link_name;
  std::vector<shapes::Shape> shapes;
  std::vector<Isometry3d> attach_trans;  // shape poses in the link frame
  std::set<std::string> touch_links;
};

class RobotState  // the slice of moveit::core::RobotState the collision env reads: variables + attached bodies
{
public:
  void attachBody(const AttachedBody& body)
  {
    attached_[body.name] = body;
  }
  bool clearAttachedBody(const std::string& name)
  {
    return attached_.erase(name) != 0;
  }
  void getAttachedBodies(std::vector<const AttachedBody*>& out) const  // name order, like RobotState's std::map
  {
    out.clear();
    for (const auto& kv : attached_)
      out.push_back(&kv.second);
  }
  std::string attachedSignature() const  // identifies the attached geometry (compareCacheEntryToState :1296-1338)
  {
    std::string s;
    for (const auto& kv : attached_)
    {
      const AttachedBody& b = kv.second;
      s += b.name + "@" + b.link_name + ":";
      for (std::size_t i = 0; i < b.shapes.size(); ++i)
      {
        s.append(reinterpret_cast<const char*>(&b.shapes[i]), sizeof(shapes::Shape));
        s.append(reinterpret_cast<const char*>(b.attach_trans[i].m), sizeof(b.attach_trans[i].m));
      }
      for (const std::string& t : b.touch_links)
        s += "|" + t;
      s += ";";
    }
    return s;
  }
  explicit RobotState(const RobotModelConstPtr& model) : model_(model), position_(model->getVariableCount(), 0.0)
  {
  }
  void setVariablePositions(const double* p)
  {
    std::copy(p, p + position_.size(), position_.begin());
  }
  const double* getVariablePositions() const
  {
    return position_.data();
  }
  double getVariablePosition(int i) const
  {
    return position_[i];
  }
  std::size_t getVariableCount() const
  {
    return position_.size();
  }
  const RobotModelConstPtr& getRobotModel() const
  {
    return model_;
  }

private:
  RobotModelConstPtr model_;
  std::vector<double> position_;
  std::map<std::string, AttachedBody> attached_;
};
}  // namespace core

namespace AllowedCollision
{
enum Type
{
  NEVER,
  ALWAYS,
  CONDITIONAL,
};
}

class AllowedCollisionMatrix  // collision_matrix.cpp:46-64, 113-145
{
public:
  AllowedCollisionMatrix() = default;
  AllowedCollisionMatrix(const std::vector<std::string>& names, bool allowed)
  {
    for (std::size_t i = 0; i < names.size(); ++i)
      for (std::size_t j = i; j < names.size(); ++j)
        setEntry(names[i], names[j], allowed);
  }
  void setEntry(const std::string& a, const std::string& b, bool allowed)
  {
    entries_[a][b] = entries_[b][a] = allowed ? AllowedCollision::ALWAYS : AllowedCollision::NEVER;
  }
  bool getEntry(const std::string& a, const std::string& b, AllowedCollision::Type& t) const
  {
    auto i = entries_.find(a);
    if (i == entries_.end())
      return false;
    auto j = i->second.find(b);
    if (j == i->second.end())
      return false;
    t = j->second;
    return true;
  }
  bool operator==(const AllowedCollisionMatrix& o) const
  {
    return entries_ == o.entries_;
  }
  std::size_t getSize() const
  {
    return entries_.size();
  }

private:
  std::map<std::string, std::map<std::string, AllowedCollision::Type>> entries_;
};

class World  // collision_detection/world.h: objects + synchronous observers
{
public:
  enum ActionBits
  {
    UNINITIALIZED = 0,
    CREATE = 1,
    DESTROY = 2,
    MOVE_SHAPE = 4,
    ADD_SHAPE = 8,
    REMOVE_SHAPE = 16,
  };
  using Action = int;
  struct Object
  {
    std::string id_;
    std::vector<shapes::ShapeConstPtr> shapes_;
    std::vector<Isometry3d> shape_poses_;  // global shape transforms
  };
  using ObjectConstPtr = std::shared_ptr<const Object>;
  using ObserverCallbackFn = std::function<void(const ObjectConstPtr&, Action)>;
  using ObserverHandle = int;

  void addToObject(const std::string& id, const shapes::ShapeConstPtr& shape, const Isometry3d& pose)
  {
    std::shared_ptr<Object>& o = objects_[id];
    Action a = ADD_SHAPE;
    if (!o)
    {
      o = std::make_shared<Object>();
      o->id_ = id;
      a |= CREATE;
    }
    else
      o = std::make_shared<Object>(*o);  // copy on write like World::ensureUnique
    o->shapes_.push_back(shape);
    o->shape_poses_.push_back(pose);
    notify(o, a);
  }
  bool moveShapeInObject(const std::string& id, std::size_t shape_index, const Isometry3d& pose)
  {
    auto it = objects_.find(id);
    if (it == objects_.end() || shape_index >= it->second->shapes_.size())
      return false;
    it->second = std::make_shared<Object>(*it->second);
    it->second->shape_poses_[shape_index] = pose;
    notify(it->second, MOVE_SHAPE);
    return true;
  }
  bool removeObject(const std::string& id)
  {
    auto it = objects_.find(id);
    if (it == objects_.end())
      return false;
    ObjectConstPtr o = it->second;
    objects_.erase(it);
    notify(o, DESTROY);
    return true;
  }
  ObjectConstPtr getObject(const std::string& id) const
  {
    auto it = objects_.find(id);
    return it == objects_.end() ? ObjectConstPtr() : ObjectConstPtr(it->second);
  }
  std::vector<std::string> getObjectIds() const  // world.h: sorted, like the std::map the reference keeps
  {
    std::vector<std::string> ids;
    for (const auto& kv : objects_)
      ids.push_back(kv.first);
    return ids;
  }
  ObserverHandle addObserver(const ObserverCallbackFn& cb)
  {
    observers_[++next_handle_] = cb;
    return next_handle_;
  }
  void removeObserver(ObserverHandle h)
  {
    observers_.erase(h);
  }
  void notifyObserverAllObjects(ObserverHandle h, Action a) const
  {
    auto it = observers_.find(h);
    if (it != observers_.end())
      for (const auto& kv : objects_)
        it->second(kv.second, a);
  }

private:
  void notify(const ObjectConstPtr& o, Action a)
  {
    for (auto& kv : observers_)
      kv.second(o, a);
  }
  std::map<std::string, std::shared_ptr<Object>> objects_;
  std::map<ObserverHandle, ObserverCallbackFn> observers_;
  ObserverHandle next_handle_ = 0;
};
using WorldPtr = std::shared_ptr<World>;

namespace BodyTypes
{
enum Type
{
  ROBOT_LINK,
  ROBOT_ATTACHED,
  WORLD_OBJECT,
};
}

struct Contact  // collision_common.h:79-120; the distance-field env fills pos and the body names/types only
{
  Vector3d pos;
  std::string body_name_1, body_name_2;
  BodyTypes::Type body_type_1 = BodyTypes::ROBOT_LINK, body_type_2 = BodyTypes::WORLD_OBJECT;
};

struct CollisionRequest  // collision_common.h:146-196
{
  std::string group_name;
  bool distance = false;
  bool cost = false;
  bool contacts = false;
  std::size_t max_contacts = 1;
  std::size_t max_contacts_per_pair = 1;
  bool verbose = false;
};

struct CollisionResult  // collision_common.h:350-390
{
  using ContactMap = std::map<std::pair<std::string, std::string>, std::vector<Contact>>;
  bool collision = false;
  double distance = std::numeric_limits<double>::max();
  std::size_t contact_count = 0;
  ContactMap contacts;
  void clear()
  {
    collision = false;
    distance = std::numeric_limits<double>::max();
    contact_count = 0;
    contacts.clear();
  }
};

struct CollisionSphere  // collision_distance_field_types.h:66-76
{
  Vector3d relative_vec_;
  double radius_ = 0.0;
};

enum CollisionType  // collision_distance_field_types.h:57-63
{
  NONE = B200DF_TYPE_NONE,
  SELF = B200DF_TYPE_SELF,
  INTRA = B200DF_TYPE_INTRA,
  ENVIRONMENT = B200DF_TYPE_ENVIRONMENT,
};

struct GradientInfo  // collision_distance_field_types.h:78-105
{
  double closest_distance = DBL_MAX;
  bool collision = false;
  std::vector<Vector3d> sphere_locations;
  std::vector<double> distances;
  std::vector<Vector3d> gradients;
  std::vector<CollisionType> types;
  std::vector<double> sphere_radii;
  std::string joint_name;
};

struct GroupStateRepresentation  // collision_common_distance_field.h:120-167 (what CHOMP reads of it)
{
  std::vector<std::string> link_names_;
  std::vector<GradientInfo> gradients_;
  // dfce_ (collision_common_distance_field.h:139): the cache entry this representation was generated from.  A call
  // that is handed a non-null gsr keeps using it (updateGroupStateRepresentationState, .cpp:1118-1170) - which is how
  // ChompOptimizer keeps the ACM of its first call while passing acm = nullptr for every waypoint
  // (chomp_optimizer.cpp:102 vs :924).  Opaque here: only CollisionEnvB200 looks inside.
  std::shared_ptr<const void> dfce_;
};
using GroupStateRepresentationPtr = std::shared_ptr<GroupStateRepresentation>;

static const double DEFAULT_SIZE_X = 3.0;  // collision_env_distance_field.h:48-54
static const double DEFAULT_SIZE_Y = 3.0;
static const double DEFAULT_SIZE_Z = 4.0;
static const bool DEFAULT_USE_SIGNED_DISTANCE_FIELD = false;
static const double DEFAULT_RESOLUTION = .02;
static const double DEFAULT_COLLISION_TOLERANCE = 0.0;
static const double DEFAULT_MAX_PROPOGATION_DISTANCE = .25;
static const double EPSILON = 0.001f;  // collision_env_distance_field.cpp:54 (a float literal stored in a double, Q16)

class B200Error : public std::runtime_error
{
public:
  explicit B200Error(const std::string& what) : std::runtime_error(what)
  {
  }
};

inline void checkStatus(int rc, const char* what)
{
  if (rc != B200DF_OK)
    throw B200Error(std::string(what) + ": " + b200df_last_error());
}

// page-locked host buffer (b200df_host_alloc): batch results are copied back at PCIe speed
template <typename T>
class PinnedBuffer
{
public:
  // Page-locked above kPinThreshold bytes (large result copies then run at PCIe speed); plain heap memory below it:
  // the C ABI stages small calls through its own mapped buffer, and cudaHostAlloc costs more than such a call.
  static constexpr std::size_t kPinThreshold = 1 << 20;
  explicit PinnedBuffer(std::size_t n) : n_(n), pinned_(n * sizeof(T) > kPinThreshold)
  {
    if (pinned_)
    {
      void* p = nullptr;
      checkStatus(b200df_host_alloc(n * sizeof(T), &p), "b200df_host_alloc");
      p_ = static_cast<T*>(p);
    }
    else
      p_ = static_cast<T*>(::operator new(std::max<std::size_t>(n, 1) * sizeof(T)));
  }
  ~PinnedBuffer()
  {
    if (pinned_)
      b200df_host_free(p_);
    else
      ::operator delete(p_);
  }
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  T* data()
  {
    return p_;
  }
  const T& operator[](std::size_t i) const
  {
    return p_[i];
  }
  std::size_t size() const
  {
    return n_;
  }

private:
  T* p_ = nullptr;
  std::size_t n_ = 0;
  bool pinned_ = false;
};

// ------------------------------------------------------------------------------------------------------------
// CollisionEnvB200
// ------------------------------------------------------------------------------------------------------------
class CollisionEnvB200
{
public:
  using LinkSpheres = std::map<std::string, std::vector<CollisionSphere>>;

  // (robot_model) constructor, collision_env_distance_field.h:62-72 / .cpp:58-75: a private empty world, padding
  // forced to 0 (Q18)
  explicit CollisionEnvB200(const core::RobotModelConstPtr& robot_model, const LinkSpheres& link_body_decompositions = {},
                            double size_x = DEFAULT_SIZE_X, double size_y = DEFAULT_SIZE_Y,
                            double size_z = DEFAULT_SIZE_Z, const Vector3d& origin = Vector3d(),
                            bool use_signed_distance_field = DEFAULT_USE_SIGNED_DISTANCE_FIELD,
                            double resolution = DEFAULT_RESOLUTION,
                            double collision_tolerance = DEFAULT_COLLISION_TOLERANCE,
                            double max_propogation_distance = DEFAULT_MAX_PROPOGATION_DISTANCE, double /*padding*/ = 0.0,
                            double /*scale*/ = 1.0)
    : CollisionEnvB200(robot_model, std::make_shared<World>(), link_body_decompositions, size_x, size_y, size_z, origin,
                       use_signed_distance_field, resolution, collision_tolerance, max_propogation_distance, 0.0, 1.0)
  {
  }

  // (robot_model, world) constructor, .h:74-84 / .cpp:77-93
  CollisionEnvB200(const core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                   const LinkSpheres& link_body_decompositions = {}, double size_x = DEFAULT_SIZE_X,
                   double size_y = DEFAULT_SIZE_Y, double size_z = DEFAULT_SIZE_Z, const Vector3d& origin = Vector3d(),
                   bool use_signed_distance_field = DEFAULT_USE_SIGNED_DISTANCE_FIELD,
                   double resolution = DEFAULT_RESOLUTION, double collision_tolerance = DEFAULT_COLLISION_TOLERANCE,
                   double max_propogation_distance = DEFAULT_MAX_PROPOGATION_DISTANCE, double padding = 0.0,
                   double /*scale*/ = 1.0)
    : robot_model_(robot_model), world_(world), link_padding_(padding)
  {
    size_[0] = size_x;
    size_[1] = size_y;
    size_[2] = size_z;
    origin_ = origin;
    construct(link_body_decompositions, use_signed_distance_field, resolution, collision_tolerance,
              max_propogation_distance);
  }

  // copy onto another world, .h:86 / .cpp:95-116 (the allocator's allocateEnv(orig, world))
  CollisionEnvB200(const CollisionEnvB200& other, const WorldPtr& world)
    : robot_model_(other.robot_model_), world_(world), link_padding_(other.link_padding_)
  {
    std::copy(other.size_, other.size_ + 3, size_);
    origin_ = other.origin_;
    octree_occupied_leaves_only_ = other.octree_occupied_leaves_only_;
    devices_ = other.devices_;
    multi_min_states_ = other.multi_min_states_;
    construct(other.link_spheres_, other.use_signed_distance_field_, other.resolution_, other.collision_tolerance_,
              other.max_propogation_distance_);
  }

  CollisionEnvB200(const CollisionEnvB200&) = delete;
  CollisionEnvB200& operator=(const CollisionEnvB200&) = delete;

  ~CollisionEnvB200()  // .cpp:118-121
  {
    world_->removeObserver(observer_handle_);
    dfce_.reset();
    model_cache_.clear();
    base_model_.reset();
    b200df_field_destroy(world_field_);
  }

  // ---- single-state API of CollisionEnv (each call is a batch of one) -----------------------------------
  // Error contract (collision_env.h has no error channel: void methods, results through CollisionResult): no
  // exception ever leaves a public method.  A failing call (CUDA error, lost device, inconsistent model) is recorded
  // (lastError(), errorCount(), one line on stderr) and answers FAIL-CLOSED: res.collision = true, a batch verdict of
  // all ones, first_invalid = 1 - a planner then treats the state as invalid instead of driving through an obstacle
  // it could not test.
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state) const
  {
    if (!guard("checkSelfCollision", [&] { checkSelfCollisionHelper(req, res, state, nullptr); }))  // .cpp:238-244
      failClosed(res);
  }
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                          const AllowedCollisionMatrix& acm) const
  {
    if (!guard("checkSelfCollision", [&] { checkSelfCollisionHelper(req, res, state, &acm); }))  // .cpp:254-261
      failClosed(res);
  }
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state) const
  {
    // .cpp:1466-1487 (no self field generated, Q13)
    if (!guard("checkRobotCollision", [&] { run(req, res, state, nullptr, B200DF_CHECK_ROBOT, false); }))
      failClosed(res);
  }
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                           const AllowedCollisionMatrix& acm) const
  {
    if (!guard("checkRobotCollision", [&] { run(req, res, state, &acm, B200DF_CHECK_ROBOT, true); }))  // .cpp:1500-1519
      failClosed(res);
  }
  // continuous check: unimplemented in the reference too (.cpp:1521-1534: logs an error, leaves the result alone)
  void checkRobotCollision(const CollisionRequest&, CollisionResult&, const core::RobotState&, const core::RobotState&,
                           const AllowedCollisionMatrix&) const
  {
    recordError("checkRobotCollision(state1, state2)", "Continuous collision checking not implemented");
  }
  void checkCollision(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state) const
  {
    const int all = B200DF_CHECK_SELF | B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT;
    if (!guard("checkCollision", [&] { run(req, res, state, nullptr, all, true); }))  // .cpp:1401-1416
      failClosed(res);
  }
  void checkCollision(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                      const AllowedCollisionMatrix& acm) const
  {
    const int all = B200DF_CHECK_SELF | B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT;
    if (!guard("checkCollision", [&] { run(req, res, state, &acm, all, true); }))  // .cpp:1441-1464
      failClosed(res);
  }
  std::size_t cacheGenerations() const
  {
    std::lock_guard<std::mutex> lk(update_cache_lock_);
    return cache_generations_;
  }
  // distanceRobot / distanceSelf are stubs in the reference (.h:108-123,178-198); defined here as the minimum over
  // the group's spheres of (world-field distance - radius)
  // (a failing call answers 0.0 = "touching", the conservative value, and records the error)
  double distanceRobot(const core::RobotState& state, const std::string& group_name = "") const
  {
    double d = 0.0;
    if (!guard("distanceRobot", [&] {
          std::shared_ptr<const CacheEntry> e = getDistanceFieldCacheEntry(group_name, state, nullptr, false);
          checkStatus(b200df_distance_batch(e->group, world_field_, state.getVariablePositions(), B200DF_Q_F64, 1, &d),
                      "b200df_distance_batch");
        }))
      d = 0.0;
    return d;
  }

  // getCollisionGradients .cpp:1536-1557 — the CHOMP entry point
  void getCollisionGradients(const CollisionRequest& req, CollisionResult& /*res*/, const core::RobotState& state,
                             const AllowedCollisionMatrix* acm, GroupStateRepresentationPtr& gsr) const
  {
    std::vector<GroupStateRepresentationPtr> one;
    if (getCollisionGradientsBatch(req, state.getVariablePositions(), 1, acm, one, &state, gsr) && !one.empty())
      gsr = one[0];
    else
      gsr.reset();  // the reference has no failure path here; a null gsr is what a caller can test for
  }

  // getAllCollisions .cpp:1559-1578: self field, intra group and environment are all visited (each stops on its own
  // at max_contacts); with req.contacts == false this is the logical OR the verdict kernel computes
  void getAllCollisions(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                        const AllowedCollisionMatrix* acm) const
  {
    if (!guard("getAllCollisions", [&] { getAllCollisionsImpl(req, res, state, acm); }))
      failClosed(res);
  }

private:
  void getAllCollisionsImpl(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                            const AllowedCollisionMatrix* acm) const
  {
    std::shared_ptr<const CacheEntry> e = getDistanceFieldCacheEntry(req.group_name, state, acm, true);
    const int flags = B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT | (e->self_field ? B200DF_CHECK_SELF : 0);
    if (!req.contacts)
    {
      std::uint8_t v = 0;
      checkStatus(b200df_check_batch(e->group, world_field_, e->self_field, state.getVariablePositions(), B200DF_Q_F64,
                                     1, flags, B200DF_PRECISION_EXACT, &v),
                  "b200df_check_batch");
      if (v)
        res.collision = true;
      return;
    }
    contactsFromSpheres(req, res, state, *e, flags, /*stop_when_done=*/false);
  }

public:
  // ---- batched entry points (new): N states, row-major [n][variable count] -------------------------------
  // verdict[i] = res.collision of the corresponding single-state call with req.contacts == false.
  // which: B200DF_CHECK_ROBOT = checkRobotCollision, SELF|INTRA = checkSelfCollision, all three = checkCollision.
  // The non-group links are posed by reference_state (the scene's current state) for the self field, exactly
  // like the cache entry of the reference, which is built from the first state it sees and then reused while the
  // non-group joints stay within EPSILON.
  // Returns false (and fills verdict with ones: fail-closed) when the call could not be carried out.
  bool checkCollisionBatch(const CollisionRequest& req, const double* states, std::int64_t n,
                           const AllowedCollisionMatrix* acm, int which, const core::RobotState& reference_state,
                           std::uint8_t* verdict) const
  {
    const bool ok = guard("checkCollisionBatch", [&] {
      const bool need_self = (which & B200DF_CHECK_SELF) != 0;
      std::shared_ptr<const CacheEntry> e = getDistanceFieldCacheEntry(req.group_name, reference_state, acm, need_self);
      int flags = which;
      if (!e->self_field)
        flags &= ~B200DF_CHECK_SELF;
      // FAST: fp32 screening + exact re-run of the undecided states = the EXACT verdicts at about half the time
      if (devices_.size() > 1 && n >= multi_min_states_)
      {
        // one process, several GPUs: contiguous N/G shards, one host thread per device inside the library
        b200df_multi* multi = multiFor(*e);
        checkStatus(b200df_multi_check_batch(multi, states, B200DF_Q_F64, n, flags, B200DF_PRECISION_FAST, verdict),
                    "b200df_multi_check_batch");
        return;
      }
      checkStatus(b200df_check_batch(e->group, world_field_, e->self_field, states, B200DF_Q_F64, n, flags,
                                     B200DF_PRECISION_FAST, verdict),
                  "b200df_check_batch");
    });
    if (!ok && verdict && n > 0)
      std::fill(verdict, verdict + n, std::uint8_t(1));
    return ok;
  }

  // Which kernels checkCollisionBatch would run for a batch of n states (nothing is launched): B200DF_ROUTE_* and, when
  // the FAST screening pass is not used, the B200DF_NOFAST_* reasons - a robot beyond the FAST tables silently runs
  // the exact kernels (same verdicts, 2-4x slower), and this is where a caller sees it.
  bool checkCollisionBatchRoute(const CollisionRequest& req, std::int64_t n, const AllowedCollisionMatrix* acm, int which,
                                const core::RobotState& reference_state, int& route, int& why_not_fast) const
  {
    route = B200DF_ROUTE_EXACT;
    why_not_fast = 0;
    return guard("checkCollisionBatchRoute", [&] {
      std::shared_ptr<const CacheEntry> e =
          getDistanceFieldCacheEntry(req.group_name, reference_state, acm, (which & B200DF_CHECK_SELF) != 0);
      int flags = which;
      if (!e->self_field)
        flags &= ~B200DF_CHECK_SELF;
      std::int32_t r = 0, w = 0;
      checkStatus(b200df_check_route(e->group, world_field_, e->self_field, B200DF_Q_F64, n, flags, B200DF_PRECISION_FAST,
                                     &r, &w),
                  "b200df_check_route");
      route = r;
      why_not_fast = w;
    });
  }

  // PlanningScene::isPathValid's waypoint loop (planning_scene.cpp:2257-2312), collision part: every waypoint in one
  // launch; invalid_index receives the colliding waypoints like the reference (nullptr: only the verdict).
  bool isPathValid(const CollisionRequest& req, const double* waypoints, std::int64_t n,
                   const AllowedCollisionMatrix* acm, const core::RobotState& reference_state,
                   std::vector<std::size_t>* invalid_index = nullptr) const
  {
    std::vector<std::uint8_t> v((std::size_t)n);
    checkCollisionBatch(req, waypoints, n, acm, B200DF_CHECK_SELF | B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT,
                        reference_state, v.data());
    if (invalid_index)
      invalid_index->clear();
    bool result = true;
    for (std::int64_t i = 0; i < n; ++i)
      if (v[(std::size_t)i])
      {
        result = false;
        if (!invalid_index)
          break;
        invalid_index->push_back((std::size_t)i);
      }
    return result;
  }

  // m motions from[i] -> to[i], each discretised into `segments` pieces and checked on the device: the batched form
  // of ompl::base::DiscreteMotionValidator::checkMotion over StateValidityChecker::isValid
  // (ompl_interface/src/detail/state_validity_checker.cpp:76-121).  first_invalid[i]: 0 = valid, else the first
  // colliding sample j (state at t = j/segments); OMPL's lastValid fraction is (j-1)/segments.
  bool checkMotionBatch(const CollisionRequest& req, const double* from, const double* to, std::int64_t m,
                        int segments, const AllowedCollisionMatrix* acm, const core::RobotState& reference_state,
                        std::int32_t* first_invalid, int precision = B200DF_PRECISION_FAST) const
  {
    const bool ok = guard("checkMotionBatch", [&] {
      checkMotionBatchImpl(req, from, to, m, segments, acm, reference_state, first_invalid, precision);
    });
    if (!ok && first_invalid && m > 0)
      std::fill(first_invalid, first_invalid + m, std::int32_t(1));  // fail-closed: invalid from the first sample on
    return ok;
  }

private:
  void checkMotionBatchImpl(const CollisionRequest& req, const double* from, const double* to, std::int64_t m,
                            int segments, const AllowedCollisionMatrix* acm, const core::RobotState& reference_state,
                            std::int32_t* first_invalid, int precision) const
  {
    std::shared_ptr<const CacheEntry> e = getDistanceFieldCacheEntry(req.group_name, reference_state, acm, true);
    int flags = B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT | (e->self_field ? B200DF_CHECK_SELF : 0);
    // JointModel::interpolate per joint type: continuous revolute joints and the planar joint's theta take the
    // shortest arc, a floating joint's quaternion (variables 3..6) is slerped
    std::vector<std::uint8_t> angular((std::size_t)robot_model_->getVariableCount(), 0);
    for (const core::LinkModel& l : robot_model_->getLinkModels())
    {
      if (l.joint_type == core::REVOLUTE && l.continuous)
        angular[(std::size_t)l.first_variable] = 1;
      if (l.joint_type == core::PLANAR)
        angular[(std::size_t)l.first_variable + 2] = 1;
      if (l.joint_type == core::FLOATING)
      {
        angular[(std::size_t)l.first_variable + 3] = 2;
        for (int k = 4; k < 7; ++k)
          angular[(std::size_t)l.first_variable + k] = 3;
      }
    }
    checkStatus(b200df_check_edges(e->group, world_field_, e->self_field, from, to, m, segments, angular.data(), flags,
                                   precision, first_invalid),
                "b200df_check_edges");
  }

public:
  // Returns false (out cleared) when the call could not be carried out.  gradientTerms() tells which terms of
  // getCollisionGradients the last successfully generated group entry computes.
  // reuse: a representation from an earlier call; when given (and generated by this env) its cache entry - group,
  // ACM tables, self field, link fields - serves the call and req.group_name / acm are not looked at, like the
  // reference's `if (!gsr) generate... else updateGroupStateRepresentationState` (.cpp:1541-1549).
  bool getCollisionGradientsBatch(const CollisionRequest& req, const double* states, std::int64_t n,
                                  const AllowedCollisionMatrix* acm, std::vector<GroupStateRepresentationPtr>& out,
                                  const core::RobotState* reference_state = nullptr,
                                  const GroupStateRepresentationPtr& reuse = GroupStateRepresentationPtr()) const
  {
    out.clear();
    const bool ok = guard("getCollisionGradientsBatch",
                          [&] { gradientsBatchImpl(req, states, n, acm, out, reference_state, reuse); });
    if (!ok)
      out.clear();
    return ok;
  }

  // The per-link PosedDistanceField term of getSelfProximityGradients (.cpp:395-418) needs one bit per link with
  // geometry in the kernel's masks: groups with more than B200DF_MAX_LINK_FIELDS such links get gradients WITHOUT
  // that term (self field, intra group and environment terms are unaffected).  Not silent: this returns
  // LINK_FIELDS_UNSUPPORTED for such a group and a warning is recorded once per generated entry.
  enum LinkFieldTerm
  {
    LINK_FIELDS_APPLIED = 0,
    LINK_FIELDS_NOT_REQUESTED = 1,  // null / empty ACM: the reference skips the term as well (.cpp:388-392)
    LINK_FIELDS_UNSUPPORTED = 2,
  };
  LinkFieldTerm linkFieldTerm() const
  {
    std::lock_guard<std::mutex> lk(update_cache_lock_);
    return dfce_ ? dfce_->link_field_term : LINK_FIELDS_NOT_REQUESTED;
  }

private:
  void gradientsBatchImpl(const CollisionRequest& req, const double* states, std::int64_t n,
                          const AllowedCollisionMatrix* acm, std::vector<GroupStateRepresentationPtr>& out,
                          const core::RobotState* reference_state, const GroupStateRepresentationPtr& reuse) const
  {
    std::shared_ptr<const CacheEntry> e;
    if (reuse && reuse->dfce_)
    {
      std::lock_guard<std::mutex> lk(update_cache_lock_);
      if (live_entries_.count(reuse->dfce_.get()))  // generated by THIS env (an entry of another env is ignored)
        e = std::static_pointer_cast<const CacheEntry>(reuse->dfce_);
    }
    if (!e)
    {
      core::RobotState ref(robot_model_);
      if (reference_state)
        ref = *reference_state;
      else if (n > 0)
        ref.setVariablePositions(states);
      e = getDistanceFieldCacheEntry(req.group_name, ref, acm, true);
    }
    const std::size_t S = e->sphere_radii.size();
    PinnedBuffer<double> dist(n * S), grad(n * S * 3), cen(n * S * 3);
    PinnedBuffer<std::uint8_t> type(n * S);
    int flags = B200DF_CHECK_INTRA | B200DF_CHECK_ROBOT | (e->self_field ? B200DF_CHECK_SELF : 0) |
                (e->link_fields.empty() ? 0 : B200DF_CHECK_LINK_FIELDS);
    checkStatus(b200df_gradients_batch(e->group, world_field_, e->self_field, states, B200DF_Q_F64, n, flags,
                                       dist.data(), grad.data(), type.data(), cen.data()),
                "b200df_gradients_batch");
    out.clear();
    for (std::int64_t i = 0; i < n; ++i)
    {
      auto gsr = std::make_shared<GroupStateRepresentation>();
      gsr->dfce_ = e;
      gsr->link_names_ = e->link_names;
      gsr->gradients_.resize(e->link_names.size());
      for (std::size_t l = 0; l < e->link_names.size(); ++l)
      {
        GradientInfo& gi = gsr->gradients_[l];
        gi.joint_name = e->joint_names[l];
        for (int s = e->sphere_begin[l]; s < e->sphere_begin[l + 1]; ++s)
        {
          const std::size_t k = i * S + s;
          gi.sphere_radii.push_back(e->sphere_radii[s]);
          gi.sphere_locations.push_back(Vector3d{ cen[3 * k], cen[3 * k + 1], cen[3 * k + 2] });
          gi.distances.push_back(dist[k]);
          gi.gradients.push_back(Vector3d{ grad[3 * k], grad[3 * k + 1], grad[3 * k + 2] });
          gi.types.push_back(static_cast<CollisionType>(type[k]));
          gi.closest_distance = std::min(gi.closest_distance, dist[k]);
        }
      }
      out.push_back(gsr);
    }
  }

public:
  void setWorld(const WorldPtr& world)  // .cpp:1702-1721
  {
    if (world == world_)
      return;
    world_->removeObserver(observer_handle_);
    world_ = world;
    {
      std::lock_guard<std::mutex> lk(update_cache_lock_world_);
      world_in_sync_ = false;  // resyncWorldLocked() below (through the observer) rebuilds the field from scratch
    }
    observer_handle_ = world_->addObserver([this](const World::ObjectConstPtr& o, World::Action a) {
      notifyObjectChange(o, a);
    });
    std::lock_guard<std::mutex> lk(update_cache_lock_world_);
    if (!valid_)
      return;
    try
    {
      resyncWorldLocked();
    }
    catch (const std::exception& ex)
    {
      recordError("setWorld", ex.what());  // world_in_sync_ stays false: checks retry, then answer fail-closed
    }
  }
  const WorldPtr& getWorld() const
  {
    return world_;
  }
  b200df_field* getDistanceField() const  // .h:200-203 (the world field)
  {
    return world_field_;
  }
  float rebuildWorldField() const  // forces the pending EDT rebuild; returns its device time in ms (-1 on failure)
  {
    float ms = 0.f;
    if (!guard("rebuildWorldField", [&] { checkStatus(b200df_field_rebuild(world_field_, &ms), "b200df_field_rebuild"); }))
      return -1.f;
    return ms;
  }
  // see changeShapeInWorldField: false (default) = the reference env's all-nodes octree ingestion (Q9), true =
  // occupied leaves only.  Takes effect for objects added afterwards.
  void setOctreeOccupiedLeavesOnly(bool on)
  {
    std::lock_guard<std::mutex> lk(update_cache_lock_world_);
    octree_occupied_leaves_only_ = on;
  }
  // The CUDA devices checkCollisionBatch shards large batches over (empty / one entry: the device the env was created
  // on).  MoveIt is one process (evaluate_collision_checking_speed.cpp:44-57): the replicas live in this process, the
  // world field is re-replicated over NVLink after every change, nothing else is exchanged.
  static constexpr std::int64_t kMultiMinStates = 1 << 18;
  void setDevices(const std::vector<int>& devices, std::int64_t min_states = kMultiMinStates)
  {
    std::lock_guard<std::mutex> lk(update_cache_lock_);
    devices_ = devices;
    multi_min_states_ = min_states;
    dfce_.reset();  // the next call regenerates the entry (and its replicas) for the new device list
  }
  const std::vector<int>& getDevices() const
  {
    return devices_;
  }
  bool valid() const  // false: construction failed (no CUDA device, ...); every check then answers fail-closed
  {
    return valid_;
  }
  std::string lastError() const  // a copy: the string is written by concurrent const callers
  {
    std::lock_guard<std::mutex> lk(error_lock_);
    return last_error_;
  }
  std::size_t errorCount() const
  {
    std::lock_guard<std::mutex> lk(error_lock_);
    return error_count_;
  }

private:
  // The robot as the kernels see it: the RobotModel's links followed by one fixed pseudo-link per attached shape
  // (an attached shape is posed by link_transform * attach_trans exactly like a fixed joint: robot_state.cpp
  // updateAttachedBodies / getAttachedBodySphereDecomposition collision_common_distance_field.cpp:105-117).
  struct ModelEntry
  {
    std::vector<core::LinkModel> links;
    std::vector<int> attached_body;  // per link: -1 = robot link, else index into attached_*
    std::vector<std::string> attached_names, attached_links;
    std::vector<std::set<std::string>> attached_touch;
    std::vector<double> points, spheres_flat, bsphere;
    std::vector<int64_t> point_offset;
    std::vector<int32_t> sphere_offset;
    // the remaining arrays of the b200df_model_desc the model was created from (kept for the per-device replicas)
    std::vector<int32_t> d_parent, d_jtype, d_ident, d_var, d_mimic_var, d_has_geom;
    std::vector<double> d_axis, d_origin, d_mmult, d_moff;
    int n_vars = 0;
    void fillDesc(b200df_model_desc& d) const
    {
      d = b200df_model_desc{};
      d.n_links = (int32_t)d_parent.size();
      d.n_vars = n_vars;
      d.parent = d_parent.data();
      d.joint_type = d_jtype.data();
      d.axis = d_axis.data();
      d.origin = d_origin.data();
      d.origin_is_identity = d_ident.data();
      d.var_index = d_var.data();
      d.mimic_var = d_mimic_var.data();
      d.mimic_mult = d_mmult.data();
      d.mimic_offset = d_moff.data();
      d.has_geometry = d_has_geom.data();
      d.sphere_offset = sphere_offset.data();
      d.spheres = spheres_flat.data();
      d.bsphere = bsphere.data();
      d.point_offset = point_offset.data();
      d.points = points.data();
    }
    b200df_model* model = nullptr;
    ~ModelEntry()
    {
      b200df_model_destroy(model);
    }
  };

  struct CacheEntry  // DistanceFieldCacheEntry, collision_common_distance_field.h:56-118
  {
    std::shared_ptr<const ModelEntry> model;
    std::string attached_signature;
    std::vector<BodyTypes::Type> geom_body_type;  // ROBOT_LINK / ROBOT_ATTACHED per link with geometry
    std::string group_name;
    bool has_acm = false;
    AllowedCollisionMatrix acm;
    std::vector<double> state_values;
    std::vector<int> state_check_indices;
    std::vector<int> link_indices;  // group links in model order
    std::vector<std::string> link_names, joint_names;  // links WITH geometry, in group order (sphere layout)
    std::vector<int> sphere_begin;
    std::vector<double> sphere_radii;
    std::vector<int> geom_model_index;           // model link index of each link with geometry
    std::vector<std::uint8_t> geom_self_enabled;  // self_collision_enabled_ of those links
    std::vector<std::uint8_t> geom_intra;         // intra_group_collision_enabled_ between them, [g][g]
    b200df_group* group = nullptr;
    b200df_field* self_field = nullptr;
    std::vector<b200df_field*> link_fields;  // GroupStateRepresentation::link_distance_fields_, per geometry link
    LinkFieldTerm link_field_term = LINK_FIELDS_NOT_REQUESTED;
    // single-process multi-GPU replicas of (model, group, self field, world field), created on first use
    std::vector<int32_t> d_group_links;
    std::vector<std::uint8_t> d_self_enabled, d_intra;
    mutable std::mutex multi_lock;
    mutable b200df_multi* multi = nullptr;
    mutable std::size_t multi_world_generation = 0;
    ~CacheEntry()
    {
      b200df_multi_destroy(multi);
      b200df_group_destroy(group);
      b200df_field_destroy(self_field);
      for (b200df_field* f : link_fields)
        b200df_field_destroy(f);
    }
  };

  b200df_multi* multiFor(const CacheEntry& e) const
  {
    std::lock_guard<std::mutex> lk(e.multi_lock);
    if (!e.multi)
    {
      b200df_model_desc md;
      e.model->fillDesc(md);
      b200df_group_desc gd{};
      gd.n_group_links = (int32_t)e.d_group_links.size();
      gd.group_links = e.d_group_links.data();
      gd.self_enabled = e.d_self_enabled.data();
      gd.intra_enabled = e.d_intra.data();
      gd.resolution = resolution_;
      gd.collision_tolerance = collision_tolerance_;
      gd.max_propagation_distance = max_propogation_distance_;
      std::vector<int32_t> dev(devices_.begin(), devices_.end());
      checkStatus(b200df_multi_create(&md, &gd, dev.data(), (int32_t)dev.size(), &e.multi), "b200df_multi_create");
      if (e.self_field)
        checkStatus(b200df_multi_set_field(e.multi, 1, e.self_field), "b200df_multi_set_field(self)");
      e.multi_world_generation = 0;
    }
    std::size_t gen;
    {
      std::lock_guard<std::mutex> wl(update_cache_lock_world_);
      gen = world_generation_;
    }
    if (e.multi_world_generation != gen)
    {
      checkStatus(b200df_multi_set_field(e.multi, 0, world_field_), "b200df_multi_set_field(world)");
      e.multi_world_generation = gen;
    }
    return e.multi;
  }

  // ---- error handling ---------------------------------------------------------------------------------
  void recordError(const char* where, const std::string& what, bool count_as_error = true) const
  {
    {
      std::lock_guard<std::mutex> lk(error_lock_);
      last_error_ = std::string(where) + ": " + what;
      if (count_as_error)
        ++error_count_;
    }
    std::fprintf(stderr, "[b200df] %s %s: %s\n", count_as_error ? "ERROR" : "WARN", where, what.c_str());  // ROS_ERROR_NAMED
  }
  static void failClosed(CollisionResult& res)
  {
    res.collision = true;
  }
  template <typename F>
  bool guard(const char* where, F&& f) const noexcept
  {
    try
    {
      if (!valid_)
        throw B200Error("the environment failed to initialise (" + init_error_ + ")");
      {
        // a world change that could not be applied leaves the field out of date: rebuild it from the World before
        // answering, or answer fail-closed
        std::lock_guard<std::mutex> lk(update_cache_lock_world_);
        if (!world_in_sync_)
          const_cast<CollisionEnvB200*>(this)->resyncWorldLocked();
      }
      f();
      return true;
    }
    catch (const std::exception& ex)
    {
      recordError(where, ex.what());
    }
    catch (...)
    {
      recordError(where, "unknown exception");
    }
    return false;
  }

  void construct(const LinkSpheres& link_body_decompositions, bool use_signed_distance_field, double resolution,
                 double collision_tolerance, double max_propogation_distance) noexcept
  {
    try
    {
      initialize(link_body_decompositions, use_signed_distance_field, resolution, collision_tolerance,
                 max_propogation_distance);
      world_field_ = newField();  // generateDistanceFieldCacheEntryWorld :1800-1816
      valid_ = true;
    }
    catch (const std::exception& ex)
    {
      init_error_ = ex.what();
      recordError("CollisionEnvB200", init_error_);
    }
    observer_handle_ = world_->addObserver([this](const World::ObjectConstPtr& o, World::Action a) {
      notifyObjectChange(o, a);
    });
    if (valid_)
    {
      std::lock_guard<std::mutex> lk(update_cache_lock_world_);
      try
      {
        resyncWorldLocked();
      }
      catch (const std::exception& ex)
      {
        recordError("CollisionEnvB200", ex.what());  // world_in_sync_ stays false: the first check retries
      }
    }
  }

  // reset the world field and add every object of the World again (notifyObserverAllObjects(CREATE), .cpp:89-92).
  // Caller holds update_cache_lock_world_.  Throws on failure, leaving world_in_sync_ false.
  void resyncWorldLocked()
  {
    world_in_sync_ = false;
    checkStatus(b200df_field_reset(world_field_), "b200df_field_reset");
    posed_objects_.clear();
    for (const std::string& id : world_->getObjectIds())
    {
      World::ObjectConstPtr now = world_->getObject(id);
      if (!now)
        continue;
      for (std::size_t i = 0; i < now->shapes_.size(); ++i)
        addShapeToWorldField(*now->shapes_[i], now->shape_poses_[i]);
      posed_objects_[id] = now;
    }
    world_in_sync_ = true;
    ++world_generation_;
  }

  // initialize .cpp:123-161 + addLinkBodyDecompositions .cpp:1065-1093
  void initialize(const LinkSpheres& link_body_decompositions, bool use_signed_distance_field, double resolution,
                  double collision_tolerance, double max_propogation_distance)
  {
    use_signed_distance_field_ = use_signed_distance_field;
    resolution_ = resolution;
    collision_tolerance_ = collision_tolerance;
    max_propogation_distance_ = max_propogation_distance;
    link_spheres_ = link_body_decompositions;
    base_model_ = buildModelEntry({});
  }

  std::shared_ptr<const ModelEntry> modelFor(const core::RobotState& state) const
  {
    std::vector<const core::AttachedBody*> bodies;
    state.getAttachedBodies(bodies);
    if (bodies.empty())
      return base_model_;
    const std::string sig = state.attachedSignature();
    std::lock_guard<std::mutex> lk(model_cache_lock_);
    auto it = model_cache_.find(sig);
    if (it != model_cache_.end())
      return it->second;
    std::shared_ptr<const ModelEntry> me = buildModelEntry(bodies);
    model_cache_[sig] = me;
    return me;
  }

  std::shared_ptr<const ModelEntry> buildModelEntry(const std::vector<const core::AttachedBody*>& bodies) const
  {
    auto me = std::make_shared<ModelEntry>();
    me->links = robot_model_->getLinkModels();
    me->attached_body.assign(me->links.size(), -1);
    for (const core::AttachedBody* b : bodies)
    {
      const int parent = robot_model_->getLinkIndex(b->link_name);
      if (parent < 0)
        throw B200Error("attached body " + b->name + " refers to unknown link " + b->link_name);
      me->attached_names.push_back(b->name);
      me->attached_links.push_back(b->link_name);
      me->attached_touch.push_back(b->touch_links);
      for (std::size_t s = 0; s < b->shapes.size(); ++s)
      {
        core::LinkModel l;
        l.name = b->name;
        l.joint_name = b->name + "/attach" + std::to_string(s);
        l.parent = parent;
        l.joint_type = core::FIXED;
        l.joint_origin = b->attach_trans[s];
        l.shapes.push_back(b->shapes[s]);
        l.shape_origins.push_back(Isometry3d::Identity());
        me->links.push_back(l);
        me->attached_body.push_back((int)me->attached_names.size() - 1);
      }
    }
    const std::vector<core::LinkModel>& links = me->links;
    const LinkSpheres& link_body_decompositions = link_spheres_;

    const int L = (int)links.size();
    std::vector<int32_t> parent(L), jtype(L), ident(L), var(L), mimic_var(L, -1), has_geom(L), sph_off(L + 1, 0);
    std::vector<int64_t> pt_off(L + 1, 0);
    std::vector<double> axis(3 * L), origin(12 * L), mmult(L, 1.0), moff(L, 0.0), spheres, points, bsphere(4 * L, 0.0);
    for (int i = 0; i < L; ++i)
    {
      const core::LinkModel& l = links[i];
      parent[i] = l.parent;
      jtype[i] = l.joint_type;
      std::copy(l.axis, l.axis + 3, &axis[3 * i]);
      std::copy(l.joint_origin.m, l.joint_origin.m + 12, &origin[12 * i]);
      ident[i] = l.joint_origin.isIdentity() ? 1 : 0;
      var[i] = l.first_variable;
      if (!l.mimic_joint.empty())
      {
        const int src = robot_model_->getJointIndex(l.mimic_joint);
        if (src < 0)
          throw B200Error("unknown mimic joint " + l.mimic_joint);
        mimic_var[i] = links[src].first_variable;
        mmult[i] = l.mimic_factor;
        moff[i] = l.mimic_offset;
      }
      has_geom[i] = l.shapes.empty() ? 0 : 1;
      if (has_geom[i])
      {
        // BodyDecomposition(shapes, poses, resolution, getLinkPadding(link)) .cpp:1078-1083
        std::vector<int32_t> types;
        std::vector<double> dims, poses;
        for (std::size_t s = 0; s < l.shapes.size(); ++s)
        {
          types.push_back(l.shapes[s].type);
          dims.insert(dims.end(), l.shapes[s].dims, l.shapes[s].dims + 3);
          poses.insert(poses.end(), l.shape_origins[s].m, l.shape_origins[s].m + 12);
        }
        // attached shapes: getBodyDecompositionCacheEntry -> BodyDecomposition(shape, resolution), padding 0.01
        const double padding = me->attached_body[i] >= 0 ? 0.01 : link_padding_;
        int64_t ns = 0, np = 0;
        checkStatus(b200df_body_decomposition((int32_t)types.size(), types.data(), dims.data(), poses.data(),
                                              resolution_, padding, nullptr, 0, &ns, nullptr, 0, &np, &bsphere[4 * i]),
                    "b200df_body_decomposition");
        std::vector<double> sph(std::max<int64_t>(ns, 1) * 4), pts(std::max<int64_t>(np, 1) * 3);
        checkStatus(b200df_body_decomposition((int32_t)types.size(), types.data(), dims.data(), poses.data(),
                                              resolution_, padding, sph.data(), ns, &ns, pts.data(), np, &np,
                                              &bsphere[4 * i]),
                    "b200df_body_decomposition");
        sph.resize(ns * 4);
        pts.resize(np * 3);
        auto over = link_body_decompositions.find(l.name);  // replaceCollisionSpheres .cpp:1085-1088
        if (me->attached_body[i] < 0 && over != link_body_decompositions.end())
        {
          sph.clear();
          for (const CollisionSphere& cs : over->second)
            sph.insert(sph.end(), { cs.relative_vec_.x, cs.relative_vec_.y, cs.relative_vec_.z, cs.radius_ });
        }
        spheres.insert(spheres.end(), sph.begin(), sph.end());
        points.insert(points.end(), pts.begin(), pts.end());
      }
      sph_off[i + 1] = (int32_t)(spheres.size() / 4);
      pt_off[i + 1] = (int64_t)(points.size() / 3);
    }
    me->points = points;
    me->point_offset = pt_off;
    me->sphere_offset = sph_off;
    me->spheres_flat = spheres;
    me->bsphere = bsphere;
    me->d_parent = parent;
    me->d_jtype = jtype;
    me->d_ident = ident;
    me->d_var = var;
    me->d_mimic_var = mimic_var;
    me->d_has_geom = has_geom;
    me->d_axis = axis;
    me->d_origin = origin;
    me->d_mmult = mmult;
    me->d_moff = moff;
    me->n_vars = robot_model_->getVariableCount();
    b200df_model_desc d;
    me->fillDesc(d);
    checkStatus(b200df_model_create(&d, &me->model), "b200df_model_create");
    return me;
  }

  b200df_field* newField() const  // PropagationDistanceField(size, resolution, origin - size/2, ...) .cpp:946-948, Q14
  {
    b200df_field_desc fd{};
    const double o[3] = { origin_.x, origin_.y, origin_.z };
    for (int i = 0; i < 3; ++i)
    {
      fd.size[i] = size_[i];
      fd.origin[i] = o[i] - 0.5 * size_[i];
    }
    fd.resolution = resolution_;
    fd.max_distance = max_propogation_distance_;
    fd.propagate_negative = use_signed_distance_field_ ? 1 : 0;
    b200df_field* f = nullptr;
    checkStatus(b200df_field_create(&fd, &f), "b200df_field_create");
    return f;
  }

  // notifyObjectChange .cpp:1723-1747 + updateDistanceObject :1749-1798.  Every shape goes through the
  // BodyDecomposition route with the default padding 0.01 (collision_distance_field_types.h:227, Q10).
  void notifyObjectChange(const World::ObjectConstPtr& obj, World::Action action) noexcept
  {
    std::lock_guard<std::mutex> lk(update_cache_lock_world_);
    if (!valid_)
      return;
    try
    {
      if (!world_in_sync_)
        resyncWorldLocked();  // includes this change: the World already holds it
      else
        applyObjectChangeLocked(obj, action);
    }
    catch (const std::exception& ex)
    {
      world_in_sync_ = false;  // checks answer fail-closed until a resync succeeds
      recordError("notifyObjectChange", ex.what());
    }
  }

  // The points a world shape contributes (updateDistanceObject .cpp:1749-1798), added (value = true) or taken out
  // again (removePointsFromField of the stored decomposition, :1731-1743).
  //  - OCTREE: PosedBodyPointDecomposition(octree) (types.cpp:355-366) lists the centre of EVERY node begin_tree()
  //    visits - inner nodes and free leaves included, no pose applied (Q9).  That is the default here too (parity).
  //    setOctreeOccupiedLeavesOnly(true) switches to what DistanceField::addOcTreeToField (distance_field.cpp:254-287)
  //    does for the same tree: occupied leaves only, leaves bigger than a voxel expanded to a lattice (on the GPU).
  //  - primitives: BodyDecomposition in the body frame with the default padding 0.01, then posed (Q10), on the GPU.
  //  - POINT_CLOUD: body-frame points posed by the shape transform (PosedBodyPointDecomposition::updatePose).
  void changeShapeInWorldField(const shapes::Shape& shape, const Isometry3d& pose, bool add)
  {
    if (shape.type == shapes::OCTREE)
    {
      if (!shape.octree)
        return;
      const shapes::OcTreeNodes& t = *shape.octree;
      const std::size_t n = t.occupied.size();
      if (octree_occupied_leaves_only_)
      {
        std::vector<double> leaves;
        for (std::size_t i = 0; i < n; ++i)
          if (t.leaf[i] && t.occupied[i])
            leaves.insert(leaves.end(), &t.xyzs[4 * i], &t.xyzs[4 * i] + 4);
        checkStatus(add ? b200df_field_add_octree_leaves(world_field_, leaves.data(), (int64_t)(leaves.size() / 4)) :
                          b200df_field_remove_octree_leaves(world_field_, leaves.data(), (int64_t)(leaves.size() / 4)),
                    "b200df_field_*_octree_leaves");
        return;
      }
      std::vector<double> pts(3 * n);
      for (std::size_t i = 0; i < n; ++i)
        std::copy(&t.xyzs[4 * i], &t.xyzs[4 * i] + 3, &pts[3 * i]);
      checkStatus(add ? b200df_field_add_points(world_field_, pts.data(), (int64_t)n) :
                        b200df_field_remove_points(world_field_, pts.data(), (int64_t)n),
                  "b200df_field_*_points");
      return;
    }
    if (shape.type == shapes::POINT_CLOUD)
    {
      if (!shape.points)
        return;
      const std::vector<double>& b = *shape.points;
      std::vector<double> pts(b.size());
      for (std::size_t i = 0; i + 2 < b.size(); i += 3)
      {
        const Vector3d w = pose * Vector3d{ b[i], b[i + 1], b[i + 2] };
        pts[i] = w.x;
        pts[i + 1] = w.y;
        pts[i + 2] = w.z;
      }
      checkStatus(add ? b200df_field_add_points(world_field_, pts.data(), (int64_t)(pts.size() / 3)) :
                        b200df_field_remove_points(world_field_, pts.data(), (int64_t)(pts.size() / 3)),
                  "b200df_field_*_points");
      return;
    }
    checkStatus(add ? b200df_field_add_posed_body(world_field_, shape.type, shape.dims, pose.m, 0.01) :
                      b200df_field_remove_posed_body(world_field_, shape.type, shape.dims, pose.m, 0.01),
                "b200df_field_*_posed_body");
  }
  void addShapeToWorldField(const shapes::Shape& shape, const Isometry3d& pose)
  {
    changeShapeInWorldField(shape, pose, true);
  }

  void applyObjectChangeLocked(const World::ObjectConstPtr& obj, World::Action action)
  {
    ++world_generation_;
    auto cur = posed_objects_.find(obj->id_);
    if (cur != posed_objects_.end() && (action == World::DESTROY || (action & (World::MOVE_SHAPE | World::REMOVE_SHAPE))))
      for (std::size_t i = 0; i < cur->second->shapes_.size(); ++i)
        changeShapeInWorldField(*cur->second->shapes_[i], cur->second->shape_poses_[i], false);
    World::ObjectConstPtr now = world_->getObject(obj->id_);
    if (now)
    {
      if (action != World::DESTROY)
      {
        // CREATE / ADD_SHAPE re-add every shape of the object like the reference (add_points holds all of them)
        for (std::size_t i = 0; i < now->shapes_.size(); ++i)
          addShapeToWorldField(*now->shapes_[i], now->shape_poses_[i]);
      }
      posed_objects_[obj->id_] = now;
    }
    else
      posed_objects_.erase(obj->id_);
  }

  // getDistanceFieldCacheEntry .cpp:206-236 (one cached entry, regenerated when group / ACM / non-group joints
  // change) + generateDistanceFieldCacheEntry :728-975
  std::shared_ptr<const CacheEntry> getDistanceFieldCacheEntry(const std::string& group_name,
                                                                const core::RobotState& state,
                                                                const AllowedCollisionMatrix* acm,
                                                                bool generate_distance_field) const
  {
    std::lock_guard<std::mutex> lk(update_cache_lock_);
    std::shared_ptr<const CacheEntry> cur = dfce_;
    if (cur && cur->group_name == group_name && cur->attached_signature == state.attachedSignature() &&
        compareCacheEntryToState(*cur, state) &&
        compareCacheEntryToAllowedCollisionMatrix(*cur, acm) && (!generate_distance_field || cur->self_field ||
                                                                  !hasNonGroupGeometry(*cur)))
      return cur;
    ++cache_generations_;
    dfce_ = generateDistanceFieldCacheEntry(group_name, state, acm, generate_distance_field);
    for (auto it = live_entries_.begin(); it != live_entries_.end();)  // forget entries nobody holds any more
      it = it->second.expired() ? live_entries_.erase(it) : std::next(it);
    live_entries_[dfce_.get()] = dfce_;
    return dfce_;
  }

  bool compareCacheEntryToState(const CacheEntry& e, const core::RobotState& state) const  // .cpp:1270-1310
  {
    for (int i : e.state_check_indices)
      if (std::fabs(e.state_values[i] - state.getVariablePosition(i)) > EPSILON)
        return false;
    return true;
  }
  static bool compareCacheEntryToAllowedCollisionMatrix(const CacheEntry& e, const AllowedCollisionMatrix* acm)
  {
    if (!acm)
      return !e.has_acm;
    return e.has_acm && e.acm == *acm;  // .cpp:1341-1389 compares the entries name by name
  }
  bool hasNonGroupGeometry(const CacheEntry& e) const
  {
    const std::vector<core::LinkModel>& links = e.model->links;
    std::set<int> in(e.link_indices.begin(), e.link_indices.end());
    for (int i = 0; i < (int)links.size(); ++i)
      if (!links[i].shapes.empty() && !in.count(i) && !(e.model->attached_body[i] >= 0 && in.count(links[i].parent)))
        return true;
    return false;
  }

  std::shared_ptr<CacheEntry> generateDistanceFieldCacheEntry(const std::string& group_name,
                                                              const core::RobotState& state,
                                                              const AllowedCollisionMatrix* acm,
                                                              bool generate_distance_field) const
  {
    auto e = std::make_shared<CacheEntry>();
    e->group_name = group_name;
    e->model = modelFor(state);
    e->attached_signature = state.attachedSignature();
    const ModelEntry& me = *e->model;
    const std::vector<core::LinkModel>& links = me.links;
    const int n_robot_links = (int)robot_model_->getLinkModels().size();
    if (group_name.empty())  // :736-739
      for (int i = 0; i < n_robot_links; ++i)
        e->link_indices.push_back(i);
    else
      e->link_indices = robot_model_->getUpdatedLinkModelIndices(group_name);
    // attached bodies of the group's links follow the links (dfce->attached_body_names_, :806-812).  The reference
    // only fills that list inside `if (acm)`: without an ACM a body attached to a group link takes part in no test at
    // all (neither self field, nor intra group, nor environment) - a grasped object would otherwise always collide
    // with the gripper link it overlaps.
    std::set<int> group_attached;
    {
      std::set<int> in_group(e->link_indices.begin(), e->link_indices.end());
      for (int i = n_robot_links; i < (int)links.size(); ++i)
        if (in_group.count(links[i].parent))
        {
          group_attached.insert(i);
          if (acm)
            e->link_indices.push_back(i);
        }
    }
    if (acm)
    {
      e->has_acm = true;
      e->acm = *acm;
    }
    const int n = (int)e->link_indices.size();
    std::vector<std::uint8_t> self_enabled(n, 1), intra(n * n, 1);
    e->sphere_begin.push_back(0);
    for (int i = 0; i < n; ++i)
    {
      const int mi = e->link_indices[i];
      const core::LinkModel& li = links[mi];
      if (li.shapes.empty())  // :846-852
      {
        self_enabled[i] = 0;
        std::fill(intra.begin() + i * n, intra.begin() + (i + 1) * n, 0);
        continue;
      }
      e->link_names.push_back(li.name);
      e->joint_names.push_back(me.attached_body[mi] >= 0 ? links[li.parent].joint_name : li.joint_name);
      e->geom_body_type.push_back(me.attached_body[mi] >= 0 ? BodyTypes::ROBOT_ATTACHED : BodyTypes::ROBOT_LINK);
      for (int s = me.sphere_offset[mi]; s < me.sphere_offset[mi + 1]; ++s)
        e->sphere_radii.push_back(me.spheres_flat[4 * s + 3]);
      e->sphere_begin.push_back((int)e->sphere_radii.size());
      for (int j = i + 1; j < n; ++j)  // the shapes of one attached body are one body: never tested against each other
        if (me.attached_body[mi] >= 0 && me.attached_body[e->link_indices[j]] == me.attached_body[mi])
          intra[i * n + j] = 0;
      if (!acm)
        continue;  // :840-844 all enabled (touch links are only looked at inside the acm branch)
      AllowedCollision::Type t;
      if (acm->getEntry(li.name, li.name, t) && t == AllowedCollision::ALWAYS)  // :785-790, :857-861
        self_enabled[i] = 0;
      for (int j = i + 1; j < n; ++j)
      {
        const int mj = e->link_indices[j];
        const bool i_att = me.attached_body[mi] >= 0, j_att = me.attached_body[mj] >= 0;
        if (!i_att && !j_att)  // link / link :793-805; only explicit ALWAYS entries disable a pair (Q12)
        {
          if (acm->getEntry(li.name, links[mj].name, t) && t == AllowedCollision::ALWAYS)
            intra[i * n + j] = 0;
        }
        else if (!i_att && j_att)
        {
          // link / attached body :806-834: the reference only visits the bodies attached to THIS link; the ACM entry
          // and the touch links of a body are ignored for every other link
          if (links[mj].parent == mi)
          {
            const int b = me.attached_body[mj];
            if ((acm->getEntry(li.name, links[mj].name, t) && t == AllowedCollision::ALWAYS) ||
                me.attached_touch[b].count(li.name))
              intra[i * n + j] = 0;
          }
        }
        else if (i_att && j_att)  // attached / attached :862-868
        {
          if (acm->getEntry(li.name, links[mj].name, t) && t == AllowedCollision::ALWAYS)
            intra[i * n + j] = 0;
        }
      }
    }
    {
      std::vector<int> pos;  // position in the group list of each link with geometry
      for (int i = 0; i < n; ++i)
        if (!links[e->link_indices[i]].shapes.empty())
          pos.push_back(i);
      const std::size_t g = pos.size();
      e->geom_intra.assign(g * g, 0);
      for (std::size_t a = 0; a < g; ++a)
      {
        e->geom_model_index.push_back(e->link_indices[pos[a]]);
        e->geom_self_enabled.push_back(self_enabled[pos[a]]);
        for (std::size_t b = a + 1; b < g; ++b)
          e->geom_intra[a * g + b] = intra[pos[a] * n + pos[b]];
      }
    }
    // :898-908 — the non-group variables that invalidate the entry when they move
    const std::set<int> group_vars = robot_model_->getGroupVariableIndices(group_name);
    for (int v = 0; v < (int)state.getVariableCount(); ++v)
    {
      e->state_values.push_back(state.getVariablePosition(v));
      if (!group_vars.count(v))
        e->state_check_indices.push_back(v);
    }
    b200df_group_desc gd{};
    std::vector<int32_t> gl(e->link_indices.begin(), e->link_indices.end());
    gd.n_group_links = n;
    gd.group_links = gl.data();
    gd.self_enabled = self_enabled.data();
    gd.intra_enabled = intra.data();
    gd.resolution = resolution_;
    gd.collision_tolerance = collision_tolerance_;
    gd.max_propagation_distance = max_propogation_distance_;
    checkStatus(b200df_group_create(me.model, &gd, &e->group), "b200df_group_create");
    e->d_group_links = gl;
    e->d_self_enabled = self_enabled;
    e->d_intra = intra;
    if (acm && acm->getSize() > 0)
      generateLinkDistanceFields(*e);
    if (generate_distance_field)  // :910-973: the links outside the group, posed by this state, as one point set
    {
      std::set<int> in(e->link_indices.begin(), e->link_indices.end());
      std::vector<double> T(links.size() * 12);
      checkStatus(b200df_model_fk(me.model, state.getVariablePositions(), B200DF_Q_F64, 1, T.data()), "b200df_model_fk");
      std::vector<double> all_points;
      bool any = false;
      for (int li = 0; li < (int)links.size(); ++li)
      {
        // :925-945: links outside the group and the bodies attached to THOSE links; a body attached to a group link
        // is never part of the self field
        if (links[li].shapes.empty() || in.count(li) || group_attached.count(li))
          continue;
        any = true;
        Isometry3d pose;
        std::copy(&T[12 * li], &T[12 * li] + 12, pose.m);
        for (int64_t p = me.point_offset[li]; p < me.point_offset[li + 1]; ++p)
        {
          // PosedBodyPointDecomposition::updatePose, collision_distance_field_types.cpp:368-379
          const Vector3d w = pose * Vector3d{ me.points[3 * p], me.points[3 * p + 1], me.points[3 * p + 2] };
          all_points.insert(all_points.end(), { w.x, w.y, w.z });
        }
      }
      if (any)
      {
        e->self_field = newField();
        checkStatus(b200df_field_add_points(e->self_field, all_points.data(), (int64_t)(all_points.size() / 3)),
                    "b200df_field_add_points");
      }
    }
    return e;
  }

  // the per-link PosedDistanceFields of getGroupStateRepresentation .cpp:1187-1210 and the ACM gate of
  // getSelfProximityGradients :388-404 (a pair whose entry exists and is not NEVER is skipped)
  void generateLinkDistanceFields(CacheEntry& e) const
  {
    const std::size_t G = e.geom_model_index.size();
    if (G == 0)
      return;
    if (G > B200DF_MAX_LINK_FIELDS)
    {
      e.link_field_term = LINK_FIELDS_UNSUPPORTED;
      recordError("generateLinkDistanceFields",
                  "group '" + e.group_name + "' has " + std::to_string(G) + " links with geometry (more than " +
                      std::to_string(B200DF_MAX_LINK_FIELDS) +
                      "): gradients are computed WITHOUT the per-link distance-field term of getSelfProximityGradients",
                  /*count_as_error=*/false);
      return;
    }
    e.link_field_term = LINK_FIELDS_APPLIED;
    const ModelEntry& me = *e.model;
    std::vector<std::uint8_t> enabled(G * G, 0);
    for (std::size_t a = 0; a < G; ++a)
    {
      const int li = e.geom_model_index[a];
      const double diameter = 2 * me.bsphere[4 * li + 3];
      b200df_field_desc fd{};
      for (int k = 0; k < 3; ++k)
      {
        fd.size[k] = diameter;
        fd.origin[k] = me.bsphere[4 * li + k] - 0.5 * diameter;
      }
      fd.resolution = resolution_;
      fd.max_distance = max_propogation_distance_;
      fd.propagate_negative = use_signed_distance_field_ ? 1 : 0;
      b200df_field* f = nullptr;
      checkStatus(b200df_field_create(&fd, &f), "b200df_field_create");
      e.link_fields.push_back(f);
      const int64_t np = me.point_offset[li + 1] - me.point_offset[li];
      if (np > 0)
        checkStatus(b200df_field_add_points(f, &me.points[3 * me.point_offset[li]], np), "b200df_field_add_points");
      for (std::size_t b = 0; b < G; ++b)
      {
        AllowedCollision::Type t;
        if (a != b && !(e.acm.getEntry(e.link_names[a], e.link_names[b], t) && t != AllowedCollision::NEVER))
          enabled[a * G + b] = 1;
      }
    }
    checkStatus(b200df_group_set_link_fields(e.group, e.link_fields.data(), enabled.data()),
                "b200df_group_set_link_fields");
  }

  void checkSelfCollisionHelper(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                                const AllowedCollisionMatrix* acm) const  // .cpp:182-204
  {
    run(req, res, state, acm, B200DF_CHECK_SELF | B200DF_CHECK_INTRA, true);
  }

  // one state through the batched kernels.  req.contacts == false: the verdict kernel.  req.contacts == true: the
  // posed sphere centres and per-sphere field distances come from the GPU and the contact bookkeeping of
  // getSelfCollisions :277-356 / getIntraGroupCollisions :433-660 / getEnvironmentCollisions :1580-1662 is replayed
  // on them in the reference's order (self field, intra group, environment for checkCollision).
  void run(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
           const AllowedCollisionMatrix* acm, int which, bool generate_distance_field) const
  {
    std::shared_ptr<const CacheEntry> e =
        getDistanceFieldCacheEntry(req.group_name, state, acm, generate_distance_field && (which & B200DF_CHECK_SELF));
    int flags = which;
    if (!e->self_field)
      flags &= ~B200DF_CHECK_SELF;
    if (!req.contacts)
    {
      std::uint8_t v = 0;
      checkStatus(b200df_check_batch(e->group, world_field_, e->self_field, state.getVariablePositions(), B200DF_Q_F64,
                                     1, flags, B200DF_PRECISION_EXACT, &v),
                  "b200df_check_batch");
      if (v)
        res.collision = true;
      return;
    }
    contactsFromSpheres(req, res, state, *e, flags);
  }

  void contactsFromSpheres(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                           const CacheEntry& e, int flags, bool stop_when_done = true) const
  {
    const std::size_t S = e.sphere_radii.size();
    std::vector<double> cen(S * 3), dist(S), grad(S * 3);
    std::vector<std::uint8_t> type(S);
    auto field_pass = [&](int flag, b200df_field* world, b200df_field* self) {
      checkStatus(b200df_gradients_batch(e.group, world, self, state.getVariablePositions(), B200DF_Q_F64, 1, flag,
                                         dist.data(), grad.data(), type.data(), cen.data()),
                  "b200df_gradients_batch");
    };
    // per-sphere predicate of getCollisionSphereCollision (types.cpp:233-267): the gradient pass leaves DBL_MAX
    // where the field distance is not below max_propogation_distance
    auto sphere_hits = [&](std::size_t s) { return dist[s] != DBL_MAX && e.sphere_radii[s] - dist[s] > collision_tolerance_; };
    auto field_contacts = [&](const char* other, BodyTypes::Type other_type, const std::vector<std::uint8_t>* enabled) {
      for (std::size_t l = 0; l < e.link_names.size(); ++l)
      {
        if (enabled && !(*enabled)[l])
          continue;
        const std::size_t allowed = std::min(req.max_contacts_per_pair, req.max_contacts - res.contact_count);
        std::size_t found = 0;
        if (allowed == 0)  // getCollisionSphereCollision with num_coll == 0 reports the collision without a contact
          for (int s = e.sphere_begin[l]; s < e.sphere_begin[l + 1]; ++s)
            if (sphere_hits(s))
            {
              res.collision = true;
              break;
            }
        for (int s = e.sphere_begin[l]; s < e.sphere_begin[l + 1] && found < allowed; ++s)
        {
          if (!sphere_hits(s))
            continue;
          Contact con;
          con.pos = Vector3d{ cen[3 * s], cen[3 * s + 1], cen[3 * s + 2] };
          con.body_name_1 = e.link_names[l];
          con.body_type_1 = e.geom_body_type[l];
          con.body_name_2 = other;
          con.body_type_2 = other_type;
          res.contacts[std::make_pair(con.body_name_1, con.body_name_2)].push_back(con);
          res.contact_count++;
          res.collision = true;
          ++found;
        }
        if (res.contact_count >= req.max_contacts)
          return true;
      }
      return res.contact_count >= req.max_contacts;
    };
    bool done = false;
    if ((flags & B200DF_CHECK_SELF) && e.self_field)
    {
      field_pass(B200DF_CHECK_SELF, nullptr, e.self_field);
      done = field_contacts("self", BodyTypes::ROBOT_LINK, &e.geom_self_enabled);
    }
    else
      field_pass(0, nullptr, nullptr);  // posed centres only
    if (!(done && stop_when_done) && (flags & B200DF_CHECK_INTRA))
      done = intraContacts(req, res, state, e, cen);
    if (!(done && stop_when_done) && (flags & B200DF_CHECK_ROBOT))
    {
      field_pass(B200DF_CHECK_ROBOT, world_field_, nullptr);
      field_contacts("environment", BodyTypes::WORLD_OBJECT, nullptr);
    }
  }

  // getIntraGroupCollisions .cpp:433-660 with req.contacts == true, replayed on the exact GPU-posed sphere centres;
  // the bounding-sphere prune (doBoundingSpheresIntersect, types.cpp:410-420, Q2) uses GPU link transforms
  bool intraContacts(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                     const CacheEntry& e, const std::vector<double>& cen) const
  {
    const std::size_t n = e.link_names.size();
    const ModelEntry& me = *e.model;
    std::vector<double> T(me.links.size() * 12);
    checkStatus(b200df_model_fk(me.model, state.getVariablePositions(), B200DF_Q_F64, 1, T.data()), "b200df_model_fk");
    std::vector<Vector3d> bcen(n);
    for (std::size_t i = 0; i < n; ++i)
    {
      const int li = e.geom_model_index[i];
      Isometry3d pose;
      std::copy(&T[12 * li], &T[12 * li] + 12, pose.m);
      bcen[i] = pose * Vector3d{ me.bsphere[4 * li], me.bsphere[4 * li + 1], me.bsphere[4 * li + 2] };
    }
    for (std::size_t i = 0; i < n; ++i)
      for (std::size_t j = i + 1; j < n; ++j)
      {
        if (!e.geom_intra[i * n + j])
          continue;
        const double bx = bcen[i].x - bcen[j].x, by = bcen[i].y - bcen[j].y, bz = bcen[i].z - bcen[j].z;
        const double p_dist = (bx * bx + by * by) + bz * bz;  // squaredNorm vs the un-squared radius sum (Q2)
        if (!(p_dist < me.bsphere[4 * e.geom_model_index[i] + 3] + me.bsphere[4 * e.geom_model_index[j] + 3]))
          continue;
        auto key = std::make_pair(e.link_names[i], e.link_names[j]);
        std::size_t num_pair = res.contacts.count(key) ? res.contacts[key].size() : 0;
        for (int k = e.sphere_begin[i]; k < e.sphere_begin[i + 1] && num_pair < req.max_contacts_per_pair; ++k)
          for (int l = e.sphere_begin[j]; l < e.sphere_begin[j + 1] && num_pair < req.max_contacts_per_pair; ++l)
          {
            const double gx = cen[3 * k] - cen[3 * l], gy = cen[3 * k + 1] - cen[3 * l + 1],
                         gz = cen[3 * k + 2] - cen[3 * l + 2];
            const double d = std::sqrt((gx * gx + gy * gy) + gz * gz);
            if (d < e.sphere_radii[k] + e.sphere_radii[l])
            {
              Contact con;
              con.pos = Vector3d{ cen[3 * k], cen[3 * k + 1], cen[3 * k + 2] };
              con.body_name_1 = e.link_names[i];
              con.body_name_2 = e.link_names[j];
              con.body_type_1 = e.geom_body_type[i];
              con.body_type_2 = e.geom_body_type[j];
              res.collision = true;
              res.contact_count++;
              res.contacts[key].push_back(con);
              ++num_pair;
              if (res.contact_count >= req.max_contacts)
                return true;
            }
          }
      }
    return false;
  }

  core::RobotModelConstPtr robot_model_;
  WorldPtr world_;
  double size_[3] = { DEFAULT_SIZE_X, DEFAULT_SIZE_Y, DEFAULT_SIZE_Z };
  Vector3d origin_;
  bool use_signed_distance_field_ = DEFAULT_USE_SIGNED_DISTANCE_FIELD;
  double resolution_ = DEFAULT_RESOLUTION;
  double collision_tolerance_ = DEFAULT_COLLISION_TOLERANCE;
  double max_propogation_distance_ = DEFAULT_MAX_PROPOGATION_DISTANCE;
  double link_padding_ = 0.0;
  LinkSpheres link_spheres_;
  std::shared_ptr<const ModelEntry> base_model_;  // the robot without attached bodies
  mutable std::map<std::string, std::shared_ptr<const ModelEntry>> model_cache_;  // by RobotState::attachedSignature()
  mutable std::mutex model_cache_lock_;
  b200df_field* world_field_ = nullptr;
  std::map<std::string, World::ObjectConstPtr> posed_objects_;  // posed_body_point_decompositions_ of the world entry
  World::ObserverHandle observer_handle_ = 0;
  mutable std::mutex update_cache_lock_, update_cache_lock_world_;  // .h:295,302
  mutable std::shared_ptr<const CacheEntry> dfce_;
  mutable std::map<const void*, std::weak_ptr<const CacheEntry>> live_entries_;  // entries handed out through a gsr
  mutable std::size_t cache_generations_ = 0;  // how often the entry had to be regenerated (guarded by update_cache_lock_)
  bool valid_ = false;             // construction succeeded
  std::string init_error_;
  bool world_in_sync_ = false;     // guarded by update_cache_lock_world_
  bool octree_occupied_leaves_only_ = false;
  std::size_t world_generation_ = 1;  // bumped by every world change (guarded by update_cache_lock_world_)
  std::vector<int> devices_;
  std::int64_t multi_min_states_ = kMultiMinStates;
  mutable std::mutex error_lock_;  // last_error_ / error_count_ are written by concurrent const callers
  mutable std::string last_error_;
  mutable std::size_t error_count_ = 0;
};
using CollisionEnvB200Ptr = std::shared_ptr<CollisionEnvB200>;

// ------------------------------------------------------------------------------------------------------------
// CollisionEnvHybridB200 - collision_env_hybrid.h:50-151 / .cpp:47-179: the class CHOMP dynamic_casts the active
// collision env to (chomp_optimizer.cpp:81-88).  The reference derives it from CollisionEnvFCL (the mesh-exact sibling
// backend answers the CollisionEnv virtuals) and forwards every *DistanceField method and getCollisionGradients to a
// CollisionEnvDistanceField member built on the same World.  Here the base is a template parameter: the plugin
// instantiates it with collision_detection::CollisionEnvFCL (moveit_b200/plugin/), this image with the stand-in
// below, which has the constructors / getWorld / setWorld the hybrid uses and nothing else.
// ------------------------------------------------------------------------------------------------------------
class CollisionEnvNoMeshBackend
{
public:
  explicit CollisionEnvNoMeshBackend(const core::RobotModelConstPtr& model, double = 0.0, double = 1.0)
    : robot_model_(model), world_(std::make_shared<World>())
  {
  }
  CollisionEnvNoMeshBackend(const core::RobotModelConstPtr& model, const WorldPtr& world, double = 0.0, double = 1.0)
    : robot_model_(model), world_(world)
  {
  }
  CollisionEnvNoMeshBackend(const CollisionEnvNoMeshBackend& other, const WorldPtr& world)
    : robot_model_(other.robot_model_), world_(world)
  {
  }
  virtual ~CollisionEnvNoMeshBackend() = default;
  const WorldPtr& getWorld() const
  {
    return world_;
  }
  virtual void setWorld(const WorldPtr& world)
  {
    world_ = world;
  }

protected:
  core::RobotModelConstPtr robot_model_;
  WorldPtr world_;
};

template <class MeshBackend = CollisionEnvNoMeshBackend>
class CollisionEnvHybridB200T : public MeshBackend
{
public:
  using LinkSpheres = CollisionEnvB200::LinkSpheres;

  explicit CollisionEnvHybridB200T(const core::RobotModelConstPtr& robot_model,
                                   const LinkSpheres& link_body_decompositions = {}, double size_x = DEFAULT_SIZE_X,
                                   double size_y = DEFAULT_SIZE_Y, double size_z = DEFAULT_SIZE_Z,
                                   const Vector3d& origin = Vector3d(),
                                   bool use_signed_distance_field = DEFAULT_USE_SIGNED_DISTANCE_FIELD,
                                   double resolution = DEFAULT_RESOLUTION,
                                   double collision_tolerance = DEFAULT_COLLISION_TOLERANCE,
                                   double max_propogation_distance = DEFAULT_MAX_PROPOGATION_DISTANCE,
                                   double padding = 0.0, double scale = 1.0)
    : MeshBackend(robot_model)
    , cenv_distance_(new CollisionEnvB200(robot_model, this->getWorld(), link_body_decompositions, size_x, size_y,
                                          size_z, origin, use_signed_distance_field, resolution, collision_tolerance,
                                          max_propogation_distance, padding, scale))
  {
  }
  CollisionEnvHybridB200T(const core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                          const LinkSpheres& link_body_decompositions = {}, double size_x = DEFAULT_SIZE_X,
                          double size_y = DEFAULT_SIZE_Y, double size_z = DEFAULT_SIZE_Z,
                          const Vector3d& origin = Vector3d(),
                          bool use_signed_distance_field = DEFAULT_USE_SIGNED_DISTANCE_FIELD,
                          double resolution = DEFAULT_RESOLUTION,
                          double collision_tolerance = DEFAULT_COLLISION_TOLERANCE,
                          double max_propogation_distance = DEFAULT_MAX_PROPOGATION_DISTANCE, double padding = 0.0,
                          double scale = 1.0)
    : MeshBackend(robot_model, world, padding, scale)
    , cenv_distance_(new CollisionEnvB200(robot_model, this->getWorld(), link_body_decompositions, size_x, size_y,
                                          size_z, origin, use_signed_distance_field, resolution, collision_tolerance,
                                          max_propogation_distance, padding, scale))
  {
  }
  CollisionEnvHybridB200T(const CollisionEnvHybridB200T& other, const WorldPtr& world)
    : MeshBackend(other, world), cenv_distance_(new CollisionEnvB200(*other.getCollisionWorldDistanceField(), world))
  {
  }

  // .cpp:77-172 - every overload forwards; the gsr overloads hand back the representation of the call
  void checkSelfCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                       const core::RobotState& state) const
  {
    cenv_distance_->checkSelfCollision(req, res, state);
  }
  void checkSelfCollisionDistanceField(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                                       const AllowedCollisionMatrix& acm) const
  {
    cenv_distance_->checkSelfCollision(req, res, state, acm);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                   const core::RobotState& state) const
  {
    cenv_distance_->checkCollision(req, res, state);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                                   const AllowedCollisionMatrix& acm) const
  {
    cenv_distance_->checkCollision(req, res, state, acm);
  }
  void checkRobotCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                        const core::RobotState& state) const
  {
    cenv_distance_->checkRobotCollision(req, res, state);
  }
  void checkRobotCollisionDistanceField(const CollisionRequest& req, CollisionResult& res,
                                        const core::RobotState& state, const AllowedCollisionMatrix& acm) const
  {
    cenv_distance_->checkRobotCollision(req, res, state, acm);
  }
  // the gsr variants (.cpp:84-90, 100-107, 115-121, 131-138, 146-152, 161-168): same verdict; gsr receives the
  // representation (posed spheres, distances, gradients) of this state, generated from / reusing its cache entry
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                                   const AllowedCollisionMatrix& acm, GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkCollision(req, res, state, acm);
    cenv_distance_->getCollisionGradients(req, res, state, &acm, gsr);
  }
  void checkCollisionDistanceField(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                                   GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->checkCollision(req, res, state);
    cenv_distance_->getCollisionGradients(req, res, state, nullptr, gsr);
  }

  void setWorld(const WorldPtr& world) override  // .cpp:174-181
  {
    if (world == this->getWorld())
      return;
    cenv_distance_->setWorld(world);
    MeshBackend::setWorld(world);
  }

  // .cpp:183-188 - the CHOMP entry point
  void getCollisionGradients(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                             const AllowedCollisionMatrix* acm, GroupStateRepresentationPtr& gsr) const
  {
    cenv_distance_->getCollisionGradients(req, res, state, acm, gsr);
  }
  // all waypoints of all candidate trajectories in one launch (the batched form of ChompOptimizer's loop)
  bool getCollisionGradientsBatch(const CollisionRequest& req, const double* states, std::int64_t n,
                                  const AllowedCollisionMatrix* acm, std::vector<GroupStateRepresentationPtr>& out,
                                  const core::RobotState* reference_state = nullptr,
                                  const GroupStateRepresentationPtr& reuse = GroupStateRepresentationPtr()) const
  {
    return cenv_distance_->getCollisionGradientsBatch(req, states, n, acm, out, reference_state, reuse);
  }
  void getAllCollisions(const CollisionRequest& req, CollisionResult& res, const core::RobotState& state,
                        const AllowedCollisionMatrix* acm) const
  {
    cenv_distance_->getAllCollisions(req, res, state, acm);
  }
  const std::shared_ptr<const CollisionEnvB200> getCollisionWorldDistanceField() const  // collision_env_hybrid.h:141-144
  {
    return cenv_distance_;
  }

protected:
  std::shared_ptr<CollisionEnvB200> cenv_distance_;
};
using CollisionEnvHybridB200 = CollisionEnvHybridB200T<>;

// CollisionDetectorAllocatorHybrid (collision_detector_allocator_hybrid.h): NAME "HYBRID" in the reference
class CollisionDetectorAllocatorHybridB200
{
public:
  static const std::string& NAME()
  {
    static const std::string name = "HYBRID_B200";
    return name;
  }
  const std::string& getName() const
  {
    return NAME();
  }
  std::shared_ptr<CollisionEnvHybridB200> allocateEnv(const WorldPtr& world,
                                                       const core::RobotModelConstPtr& robot_model) const
  {
    return std::make_shared<CollisionEnvHybridB200>(robot_model, world);
  }
  std::shared_ptr<CollisionEnvHybridB200> allocateEnv(const std::shared_ptr<const CollisionEnvHybridB200>& orig,
                                                       const WorldPtr& world) const
  {
    return std::make_shared<CollisionEnvHybridB200>(*orig, world);
  }
  std::shared_ptr<CollisionEnvHybridB200> allocateEnv(const core::RobotModelConstPtr& robot_model) const
  {
    return std::make_shared<CollisionEnvHybridB200>(robot_model);
  }
  static std::shared_ptr<CollisionDetectorAllocatorHybridB200> create()
  {
    return std::make_shared<CollisionDetectorAllocatorHybridB200>();
  }
};

// CollisionDetectorAllocatorTemplate<CollisionEnvB200, CollisionDetectorAllocatorB200>, collision_detector_allocator.h:73-98
class CollisionDetectorAllocatorB200
{
public:
  static const std::string& NAME()
  {
    static const std::string name = "B200DistanceField";
    return name;
  }
  const std::string& getName() const
  {
    return NAME();
  }
  CollisionEnvB200Ptr allocateEnv(const WorldPtr& world, const core::RobotModelConstPtr& robot_model) const
  {
    return std::make_shared<CollisionEnvB200>(robot_model, world);
  }
  CollisionEnvB200Ptr allocateEnv(const std::shared_ptr<const CollisionEnvB200>& orig, const WorldPtr& world) const
  {
    return std::make_shared<CollisionEnvB200>(*orig, world);
  }
  CollisionEnvB200Ptr allocateEnv(const core::RobotModelConstPtr& robot_model) const
  {
    return std::make_shared<CollisionEnvB200>(robot_model);
  }
  static std::shared_ptr<CollisionDetectorAllocatorB200> create()
  {
    return std::make_shared<CollisionDetectorAllocatorB200>();
  }
};
}  // namespace collision_detection_b200
