// MoveIt types -> the types of the tested host mirror (moveit_b200/host/collision_env_b200.hpp).
//
// The plugin is an adapter: everything that decides a verdict (cache entries, ACM flags, world-object bookkeeping,
// contact replay, batching, the fail-closed error contract) is the mirror's code, which tests/test_host_mirror.py
// drives through host_driver.cpp against the CPU oracle.  These functions only re-express MoveIt objects in the
// mirror's stand-in types; they are written against the member names of the reference headers cited per function, and
// tests/test_plugin_sources.py compiles them against tests/moveit_stubs/ (the same declarations, bodies omitted).
#pragma once

#include <moveit/collision_detection/collision_common.h>
#include <moveit/collision_detection/collision_matrix.h>
#include <moveit/collision_detection/world.h>
#include <moveit/robot_model/robot_model.h>
#include <moveit/robot_state/robot_state.h>
#include <moveit/collision_distance_field/collision_distance_field_types.h>
#include <geometric_shapes/shapes.h>
#include <octomap/octomap.h>

#include <collision_env_b200.hpp>  // moveit_b200/host

namespace collision_detection
{
namespace b200_adapters
{
namespace mirror = collision_detection_b200;

// Eigen::Isometry3d (column-major 4x4) -> row-major 3x4
inline mirror::Isometry3d toMirror(const Eigen::Isometry3d& t)
{
  mirror::Isometry3d r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      r.m[4 * i + j] = t(i, j);
  return r;
}

inline mirror::Vector3d toMirror(const Eigen::Vector3d& v)
{
  mirror::Vector3d r;
  r.x = v.x();
  r.y = v.y();
  r.z = v.z();
  return r;
}

inline Eigen::Vector3d fromMirror(const mirror::Vector3d& v)
{
  return Eigen::Vector3d(v.x, v.y, v.z);
}

// octomap::OcTree -> every node in begin_tree() order (collision_distance_field_types.cpp:355-366 iterates exactly this,
// quirk Q9; the occupied-leaves reading of distance_field.cpp:254-279 is a flag on the mirror env)
inline std::shared_ptr<const mirror::shapes::OcTreeNodes> toMirror(const octomap::OcTree& tree)
{
  auto nodes = std::make_shared<mirror::shapes::OcTreeNodes>();
  for (octomap::OcTree::tree_iterator it = tree.begin_tree(); it != tree.end_tree(); ++it)
  {
    const octomap::point3d c = it.getCoordinate();
    nodes->xyzs.push_back(c.x());
    nodes->xyzs.push_back(c.y());
    nodes->xyzs.push_back(c.z());
    nodes->xyzs.push_back(it.getSize());
    nodes->occupied.push_back(tree.isNodeOccupied(*it) ? 1 : 0);
    nodes->leaf.push_back(it.isLeaf() ? 1 : 0);
  }
  return nodes;
}

// geometric_shapes/shapes.h.  Spheres, cylinders and boxes go to the GPU rasteriser as they are; every other shape
// (mesh, cone, plane) is decomposed on the host by MoveIt's own BodyDecomposition (collision_distance_field_types.h:
// 197-262) and handed over as its body-frame point list.
inline mirror::shapes::Shape toMirror(const shapes::ShapeConstPtr& shape, double resolution, double padding)
{
  mirror::shapes::Shape s;
  switch (shape->type)
  {
    case shapes::SPHERE:
      s.type = mirror::shapes::SPHERE;
      s.dims[0] = static_cast<const shapes::Sphere*>(shape.get())->radius;
      break;
    case shapes::CYLINDER:
      s.type = mirror::shapes::CYLINDER;
      s.dims[0] = static_cast<const shapes::Cylinder*>(shape.get())->radius;
      s.dims[1] = static_cast<const shapes::Cylinder*>(shape.get())->length;
      break;
    case shapes::BOX:
      s.type = mirror::shapes::BOX;
      for (int k = 0; k < 3; ++k)
        s.dims[k] = static_cast<const shapes::Box*>(shape.get())->size[k];
      break;
    case shapes::OCTREE:
      s.type = mirror::shapes::OCTREE;
      s.octree = toMirror(*static_cast<const shapes::OcTree*>(shape.get())->octree);
      break;
    default:
    {
      s.type = mirror::shapes::POINT_CLOUD;
      const BodyDecomposition bd(shape, resolution, padding);
      auto pts = std::make_shared<std::vector<double>>();
      for (const Eigen::Vector3d& p : bd.getCollisionPoints())
      {
        pts->push_back(p.x());
        pts->push_back(p.y());
        pts->push_back(p.z());
      }
      s.points = pts;
      break;
    }
  }
  return s;
}

inline mirror::core::JointType toMirror(moveit::core::JointModel::JointType t)
{
  switch (t)
  {
    case moveit::core::JointModel::REVOLUTE:
      return mirror::core::REVOLUTE;
    case moveit::core::JointModel::PRISMATIC:
      return mirror::core::PRISMATIC;
    case moveit::core::JointModel::PLANAR:
      return mirror::core::PLANAR;
    case moveit::core::JointModel::FLOATING:
      return mirror::core::FLOATING;
    default:
      return mirror::core::FIXED;
  }
}

// RobotModel::getLinkModels() is in DFS pre-order (robot_model.cpp:783-838), the order the mirror and the kernels walk.
// Mesh links additionally get MoveIt's sphere decomposition as an override (link_body_decompositions,
// collision_env_distance_field.cpp:1085-1088), because the engine derives spheres only from primitive shapes.
inline mirror::core::RobotModelConstPtr toMirror(const moveit::core::RobotModel& model, double resolution,
                                                 double padding,
                                                 mirror::CollisionEnvB200::LinkSpheres& mesh_link_spheres)
{
  auto out = std::make_shared<mirror::core::RobotModel>(model.getName());
  for (const moveit::core::LinkModel* l : model.getLinkModels())
  {
    const moveit::core::JointModel* j = l->getParentJointModel();
    mirror::core::LinkModel m;
    m.name = l->getName();
    m.joint_name = j->getName();
    m.parent = l->getParentLinkModel() ? l->getParentLinkModel()->getLinkIndex() : -1;
    m.joint_type = toMirror(j->getType());
    if (j->getType() == moveit::core::JointModel::REVOLUTE)
    {
      const auto* r = static_cast<const moveit::core::RevoluteJointModel*>(j);
      m.continuous = r->isContinuous();
      for (int k = 0; k < 3; ++k)
        m.axis[k] = r->getAxis()[k];
    }
    else if (j->getType() == moveit::core::JointModel::PRISMATIC)
    {
      const auto* p = static_cast<const moveit::core::PrismaticJointModel*>(j);
      for (int k = 0; k < 3; ++k)
        m.axis[k] = p->getAxis()[k];
    }
    m.joint_origin = toMirror(l->getJointOriginTransform());
    if (j->getMimic())
    {
      m.mimic_joint = j->getMimic()->getName();
      m.mimic_factor = j->getMimicFactor();
      m.mimic_offset = j->getMimicOffset();
    }
    m.first_variable = j->getVariableCount() ? j->getFirstVariableIndex() : -1;
    bool primitive_only = true;
    for (std::size_t s = 0; s < l->getShapes().size(); ++s)
    {
      m.shapes.push_back(toMirror(l->getShapes()[s], resolution, padding));
      m.shape_origins.push_back(toMirror(l->getCollisionOriginTransforms()[s]));
      primitive_only = primitive_only && m.shapes.back().type != mirror::shapes::POINT_CLOUD;
    }
    if (!primitive_only)
    {
      const BodyDecomposition bd(l->getShapes(), l->getCollisionOriginTransforms(), resolution, padding);
      std::vector<mirror::CollisionSphere>& spheres = mesh_link_spheres[l->getName()];
      for (const CollisionSphere& cs : bd.getCollisionSpheres())
      {
        mirror::CollisionSphere ms;
        ms.relative_vec_ = toMirror(cs.relative_vec_);
        ms.radius_ = cs.radius_;
        spheres.push_back(ms);
      }
    }
    out->addLink(m);
  }
  for (const moveit::core::JointModelGroup* g : model.getJointModelGroups())
    out->addJointModelGroup(g->getName(), g->getJointModelNames());
  return out;
}

// RobotState::getVariablePositions() + attached bodies (robot_state.h:1678, attached_body.h:92-131)
inline mirror::core::RobotState toMirror(const moveit::core::RobotState& state,
                                         const mirror::core::RobotModelConstPtr& model, double resolution)
{
  mirror::core::RobotState out(model);
  out.setVariablePositions(state.getVariablePositions());
  std::vector<const moveit::core::AttachedBody*> bodies;
  state.getAttachedBodies(bodies);
  for (const moveit::core::AttachedBody* b : bodies)
  {
    mirror::core::AttachedBody mb;
    mb.name = b->getName();
    mb.link_name = b->getAttachedLinkName();
    for (std::size_t s = 0; s < b->getShapes().size(); ++s)
    {
      // 0.01: the default padding getAttachedBodyPointDecomposition uses (collision_common_distance_field.cpp:79)
      mb.shapes.push_back(toMirror(b->getShapes()[s], resolution, 0.01));
      mb.attach_trans.push_back(toMirror(b->getShapePosesInLinkFrame()[s]));
    }
    mb.touch_links = b->getTouchLinks();
    out.attachBody(mb);
  }
  return out;
}

// collision_matrix.h:102,185.  Only explicit entries matter to the distance-field env (quirk Q12).
inline mirror::AllowedCollisionMatrix toMirror(const AllowedCollisionMatrix& acm)
{
  mirror::AllowedCollisionMatrix out;
  std::vector<std::string> names;
  acm.getAllEntryNames(names);
  for (std::size_t i = 0; i < names.size(); ++i)
    for (std::size_t j = i; j < names.size(); ++j)
    {
      AllowedCollision::Type t;
      if (acm.getEntry(names[i], names[j], t) && t != AllowedCollision::CONDITIONAL)
        out.setEntry(names[i], names[j], t == AllowedCollision::ALWAYS);
    }
  return out;
}

inline mirror::CollisionRequest toMirror(const CollisionRequest& req)
{
  mirror::CollisionRequest r;
  r.group_name = req.group_name;
  r.distance = req.distance;
  r.cost = req.cost;
  r.contacts = req.contacts;
  r.max_contacts = req.max_contacts;
  r.max_contacts_per_pair = req.max_contacts_per_pair;
  r.verbose = req.verbose;
  return r;
}

inline BodyType fromMirror(mirror::BodyTypes::Type t)
{
  return t == mirror::BodyTypes::ROBOT_LINK ? BodyTypes::ROBOT_LINK :
         t == mirror::BodyTypes::ROBOT_ATTACHED ? BodyTypes::ROBOT_ATTACHED : BodyTypes::WORLD_OBJECT;
}

// results are merged the way the reference's methods write into a caller-owned CollisionResult: collision is only
// ever raised, contacts are appended
inline void mergeResult(const mirror::CollisionResult& in, CollisionResult& res)
{
  res.collision = res.collision || in.collision;
  res.contact_count += in.contact_count;
  for (const auto& kv : in.contacts)
  {
    std::vector<Contact>& dst = res.contacts[kv.first];
    for (const mirror::Contact& c : kv.second)
    {
      Contact o;
      o.pos = fromMirror(c.pos);
      o.body_name_1 = c.body_name_1;
      o.body_name_2 = c.body_name_2;
      o.body_type_1 = fromMirror(c.body_type_1);
      o.body_type_2 = fromMirror(c.body_type_2);
      dst.push_back(o);
    }
  }
}

// GroupStateRepresentation as CHOMP reads it (collision_common_distance_field.h:120-167, chomp_optimizer.cpp:924-961)
inline void fillGradients(const mirror::GroupStateRepresentation& in, GroupStateRepresentation& out)
{
  out.gradients_.resize(in.gradients_.size());
  for (std::size_t i = 0; i < in.gradients_.size(); ++i)
  {
    const mirror::GradientInfo& g = in.gradients_[i];
    GradientInfo& o = out.gradients_[i];
    o.closest_distance = g.closest_distance;
    o.collision = g.collision;
    o.joint_name = g.joint_name;
    o.distances = g.distances;
    o.sphere_radii = g.sphere_radii;
    o.types.resize(g.types.size());
    o.gradients.resize(g.gradients.size());
    o.sphere_locations.resize(g.sphere_locations.size());
    for (std::size_t k = 0; k < g.types.size(); ++k)
    {
      o.types[k] = static_cast<CollisionType>(g.types[k]);
      o.gradients[k] = fromMirror(g.gradients[k]);
      o.sphere_locations[k] = fromMirror(g.sphere_locations[k]);
    }
  }
}
}  // namespace b200_adapters
}  // namespace collision_detection
