// Plugin configuration of the B200 distance-field collision detector.
//
// The reference has no parameters for this path: field size / origin / resolution / tolerance / propagation distance
// are constructor arguments of CollisionEnvDistanceField with compile-time defaults
// (collision_distance_field/include/moveit/collision_distance_field/collision_env_distance_field.h:48-65), and the
// allocator template (collision_detector_allocator.h:73-88) always passes the defaults.  The only configuration format
// that exists for it is the `link_spheres` parameter of moveit_experimental
// (collision_distance_field_ros/include/collision_distance_field_ros/collision_distance_field_ros_helpers.h:46-120):
//
//   link_spheres:
//     - link: panda_link0
//       spheres:
//         - {x: 0.0, y: 0.0, z: 0.05, radius: 0.08}
//         - {x: 0.0, y: 0.0, z: 0.15, radius: 0.06}
//     - link: panda_link8
//       spheres: none
//
// This header defines the parameter namespace the B200 plugin reads (same keys whether they come from the ROS
// parameter server - moveit_b200/plugin/src/collision_detector_b200_plugin_loader.cpp - or from a YAML file), keeps the
// link_spheres format and its warnings ("All link spheres must have link", a sphere without x / y / z / radius is
// skipped, `spheres: none` = a link without spheres), and carries a reader for the YAML subset such files use, so the
// format is tested in this image without ROS.
#pragma once

#include <cctype>
#include <cstdlib>
#include <istream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "collision_env_b200.hpp"

namespace collision_detection_b200
{
struct B200DistanceFieldConfig
{
  // collision_env_distance_field.h:48-54 defaults
  double size_x = DEFAULT_SIZE_X, size_y = DEFAULT_SIZE_Y, size_z = DEFAULT_SIZE_Z;
  Vector3d origin;  // the field CENTRE (Q14)
  bool use_signed_distance_field = DEFAULT_USE_SIGNED_DISTANCE_FIELD;
  double resolution = DEFAULT_RESOLUTION;
  double collision_tolerance = DEFAULT_COLLISION_TOLERANCE;
  double max_propogation_distance = DEFAULT_MAX_PROPOGATION_DISTANCE;
  double padding = 0.0, scale = 1.0;
  // engine-side switches
  bool octree_occupied_leaves_only = false;  // false = the reference env's all-nodes octree ingestion (Q9)
  std::vector<int> devices;                  // CUDA devices batched checks are sharded over; empty = the current one
  CollisionEnvB200::LinkSpheres link_spheres;
  std::vector<std::string> warnings;  // what the reference would ROS_WARN about while reading link_spheres
};

// ---- a value tree for the YAML subset: block mappings, block sequences, flow mappings / sequences, plain scalars ----
struct ParamValue
{
  enum Kind
  {
    SCALAR,
    MAP,
    LIST
  } kind = SCALAR;
  std::string scalar;
  std::vector<std::pair<std::string, ParamValue>> map;
  std::vector<ParamValue> list;
  const ParamValue* find(const std::string& key) const
  {
    for (const auto& kv : map)
      if (kv.first == key)
        return &kv.second;
    return nullptr;
  }
  bool asDouble(double& out) const
  {
    if (kind != SCALAR || scalar.empty())
      return false;
    char* end = nullptr;
    out = std::strtod(scalar.c_str(), &end);
    return end && *end == '\0';
  }
  bool asBool(bool& out) const
  {
    if (kind != SCALAR)
      return false;
    std::string s;
    for (char c : scalar)
      s.push_back((char)std::tolower((unsigned char)c));
    if (s == "true" || s == "yes" || s == "on" || s == "1")
      out = true;
    else if (s == "false" || s == "no" || s == "off" || s == "0")
      out = false;
    else
      return false;
    return true;
  }
};

namespace param_yaml
{
struct Line
{
  int indent;
  std::string text;
};

inline std::string trim(const std::string& s)
{
  std::size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a]))
    ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1]))
    --b;
  return s.substr(a, b - a);
}

inline std::string unquote(const std::string& s)
{
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
    return s.substr(1, s.size() - 2);
  return s;
}

// flow value: {a: 1, b: {c: 2}}, [1, 2, 3] or a scalar
inline ParamValue parseFlow(const std::string& s, std::size_t& i)
{
  auto skip = [&] {
    while (i < s.size() && std::isspace((unsigned char)s[i]))
      ++i;
  };
  ParamValue v;
  skip();
  if (i < s.size() && s[i] == '{')
  {
    v.kind = ParamValue::MAP;
    ++i;
    for (;;)
    {
      skip();
      if (i >= s.size() || s[i] == '}')
      {
        ++i;
        break;
      }
      std::size_t k0 = i;
      while (i < s.size() && s[i] != ':')
        ++i;
      const std::string key = unquote(trim(s.substr(k0, i - k0)));
      ++i;
      v.map.emplace_back(key, parseFlow(s, i));
      skip();
      if (i < s.size() && s[i] == ',')
        ++i;
    }
    return v;
  }
  if (i < s.size() && s[i] == '[')
  {
    v.kind = ParamValue::LIST;
    ++i;
    for (;;)
    {
      skip();
      if (i >= s.size() || s[i] == ']')
      {
        ++i;
        break;
      }
      v.list.push_back(parseFlow(s, i));
      skip();
      if (i < s.size() && s[i] == ',')
        ++i;
    }
    return v;
  }
  std::size_t a = i;
  while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']')
    ++i;
  v.scalar = unquote(trim(s.substr(a, i - a)));
  return v;
}

inline ParamValue parseInline(const std::string& text)
{
  std::size_t i = 0;
  return parseFlow(text, i);
}

ParamValue parseBlock(const std::vector<Line>& lines, std::size_t& i, int indent);

// "key: value" / "key:" at position i (already known to be a mapping entry of this indent)
inline void parseMapEntry(const std::vector<Line>& lines, std::size_t& i, int indent, const std::string& text,
                          ParamValue& into)
{
  const std::size_t colon = text.find(':');
  const std::string key = unquote(trim(text.substr(0, colon)));
  const std::string rest = trim(text.substr(colon + 1));
  ++i;
  if (!rest.empty())
  {
    into.map.emplace_back(key, parseInline(rest));
    return;
  }
  // nested block (a sequence may sit at the same indent as its key)
  if (i < lines.size() && (lines[i].indent > indent || (lines[i].indent == indent && lines[i].text[0] == '-')))
    into.map.emplace_back(key, parseBlock(lines, i, lines[i].indent));
  else
    into.map.emplace_back(key, ParamValue());
}

inline ParamValue parseBlock(const std::vector<Line>& lines, std::size_t& i, int indent)
{
  ParamValue v;
  if (i >= lines.size())
    return v;
  if (lines[i].text[0] == '-')
  {
    v.kind = ParamValue::LIST;
    while (i < lines.size() && lines[i].indent == indent && lines[i].text[0] == '-')
    {
      const std::string rest = trim(lines[i].text.substr(1));
      const std::size_t off = lines[i].text.find_first_not_of(' ', 1);  // column of the item's content after "- "
      const int inner = indent + (int)(off == std::string::npos ? 2 : off);
      if (rest.empty())
      {
        ++i;
        v.list.push_back(i < lines.size() && lines[i].indent > indent ? parseBlock(lines, i, lines[i].indent) :
                                                                        ParamValue());
      }
      else if (rest[0] == '{' || rest[0] == '[' || rest.find(':') == std::string::npos)
      {
        v.list.push_back(parseInline(rest));
        ++i;
      }
      else
      {
        // "- key: value" starts a mapping whose further keys are indented to the column of `key`
        ParamValue item;
        item.kind = ParamValue::MAP;
        parseMapEntry(lines, i, inner, rest, item);
        while (i < lines.size() && lines[i].indent == inner && lines[i].text[0] != '-')
          parseMapEntry(lines, i, inner, lines[i].text, item);
        v.list.push_back(item);
      }
    }
    return v;
  }
  v.kind = ParamValue::MAP;
  while (i < lines.size() && lines[i].indent == indent && lines[i].text[0] != '-')
    parseMapEntry(lines, i, indent, lines[i].text, v);
  return v;
}

inline ParamValue parse(std::istream& in)
{
  std::vector<Line> lines;
  std::string raw;
  while (std::getline(in, raw))
  {
    bool quoted = false;
    std::size_t cut = raw.size();
    for (std::size_t k = 0; k < raw.size(); ++k)
    {
      if (raw[k] == '"' || raw[k] == '\'')
        quoted = !quoted;
      if (raw[k] == '#' && !quoted && (k == 0 || std::isspace((unsigned char)raw[k - 1])))
      {
        cut = k;
        break;
      }
    }
    raw = raw.substr(0, cut);
    const std::string t = trim(raw);
    if (t.empty() || t == "---")
      continue;
    lines.push_back(Line{ (int)raw.find_first_not_of(' '), t });
  }
  std::size_t i = 0;
  return lines.empty() ? ParamValue() : parseBlock(lines, i, lines[0].indent);
}
}  // namespace param_yaml

// loadLinkBodySphereDecompositions (collision_distance_field_ros_helpers.h:46-120) over the value tree; returns false
// with the reference's reasons ("No parameter for link spheres", not an array, no entries)
inline bool loadLinkBodySphereDecompositions(const ParamValue& ns, CollisionEnvB200::LinkSpheres& link_body_spheres,
                                             std::vector<std::string>& warnings)
{
  const ParamValue* link_spheres = ns.find("link_spheres");
  if (!link_spheres)
    return false;  // ROS_INFO: No parameter for link spheres
  if (link_spheres->kind != ParamValue::LIST)
  {
    warnings.push_back("Link spheres not an array");
    return false;
  }
  if (link_spheres->list.empty())
  {
    warnings.push_back("Link spheres has no entries");
    return false;
  }
  for (const ParamValue& entry : link_spheres->list)
  {
    const ParamValue* link = entry.find("link");
    const ParamValue* spheres = entry.find("spheres");
    if (!link)
    {
      warnings.push_back("All link spheres must have link");
      continue;
    }
    if (!spheres)
    {
      warnings.push_back("All link spheres must have spheres");
      continue;
    }
    std::vector<CollisionSphere> coll_spheres;
    if (spheres->kind != ParamValue::LIST)
    {
      if (spheres->scalar == "none")
        link_body_spheres[link->scalar] = coll_spheres;  // a link without spheres
      else
        warnings.push_back("spheres of link " + link->scalar + " is neither an array nor 'none'");
      continue;
    }
    for (const ParamValue& s : spheres->list)
    {
      CollisionSphere cs;
      const char* keys[4] = { "x", "y", "z", "radius" };
      double* dst[4] = { &cs.relative_vec_.x, &cs.relative_vec_.y, &cs.relative_vec_.z, &cs.radius_ };
      bool ok = true;
      for (int k = 0; k < 4 && ok; ++k)
      {
        const ParamValue* v = s.find(keys[k]);
        if (!v || !v->asDouble(*dst[k]))
        {
          warnings.push_back(std::string("All spheres must specify a value for ") + keys[k]);
          ok = false;
        }
      }
      if (ok)
        coll_spheres.push_back(cs);
    }
    link_body_spheres[link->scalar] = coll_spheres;
  }
  return true;
}

// The plugin's parameter namespace (all keys optional; defaults = the reference's constructor defaults):
//   size_x, size_y, size_z, origin: {x, y, z} (field centre), resolution, collision_tolerance,
//   max_propogation_distance (the reference's spelling; max_propagation_distance is accepted too),
//   use_signed_distance_field, padding, scale, octree_occupied_leaves_only, devices: [0, 1, ...], link_spheres: [...]
inline B200DistanceFieldConfig configFromParams(const ParamValue& ns)
{
  B200DistanceFieldConfig c;
  auto number = [&](const char* key, double& dst) {
    const ParamValue* v = ns.find(key);
    if (v && !v->asDouble(dst))
      c.warnings.push_back(std::string(key) + " is not a number");
  };
  auto flag = [&](const char* key, bool& dst) {
    const ParamValue* v = ns.find(key);
    if (v && !v->asBool(dst))
      c.warnings.push_back(std::string(key) + " is not a boolean");
  };
  number("size_x", c.size_x);
  number("size_y", c.size_y);
  number("size_z", c.size_z);
  number("resolution", c.resolution);
  number("collision_tolerance", c.collision_tolerance);
  number("max_propagation_distance", c.max_propogation_distance);
  number("max_propogation_distance", c.max_propogation_distance);
  number("padding", c.padding);
  number("scale", c.scale);
  flag("use_signed_distance_field", c.use_signed_distance_field);
  flag("octree_occupied_leaves_only", c.octree_occupied_leaves_only);
  if (const ParamValue* o = ns.find("origin"))
  {
    const char* keys[3] = { "x", "y", "z" };
    double* dst[3] = { &c.origin.x, &c.origin.y, &c.origin.z };
    for (int k = 0; k < 3; ++k)
    {
      const ParamValue* v = o->kind == ParamValue::MAP ? o->find(keys[k]) :
                            (o->kind == ParamValue::LIST && o->list.size() == 3) ? &o->list[k] : nullptr;
      if (!v || !v->asDouble(*dst[k]))
        c.warnings.push_back(std::string("origin needs ") + keys[k]);
    }
  }
  if (const ParamValue* d = ns.find("devices"))
  {
    for (const ParamValue& v : d->list)
    {
      double x = 0;
      if (v.asDouble(x) && x >= 0 && x == (double)(int)x)
        c.devices.push_back((int)x);
      else
        c.warnings.push_back("devices must be a list of non-negative integers");
    }
  }
  if (!(c.resolution > 0.0))
  {
    c.warnings.push_back("resolution must be > 0; using the default");
    c.resolution = DEFAULT_RESOLUTION;
  }
  loadLinkBodySphereDecompositions(ns, c.link_spheres, c.warnings);
  return c;
}

inline B200DistanceFieldConfig configFromYaml(std::istream& in, const std::string& ns_name = "")
{
  const ParamValue root = param_yaml::parse(in);
  if (!ns_name.empty())
    if (const ParamValue* sub = root.find(ns_name))
      return configFromParams(*sub);
  return configFromParams(root);
}

// CollisionDetectorAllocator with a configuration: what the plugin loader installs after reading its namespace.
// allocateEnv passes the configured arguments where collision_detector_allocator.h:73-88 passes the defaults.
class ConfiguredAllocatorB200
{
public:
  explicit ConfiguredAllocatorB200(const B200DistanceFieldConfig& config = B200DistanceFieldConfig()) : config_(config)
  {
  }
  const std::string& getName() const
  {
    return CollisionDetectorAllocatorB200::NAME();
  }
  const B200DistanceFieldConfig& config() const
  {
    return config_;
  }
  CollisionEnvB200Ptr allocateEnv(const WorldPtr& world, const core::RobotModelConstPtr& robot_model) const
  {
    auto env = std::make_shared<CollisionEnvB200>(robot_model, world, config_.link_spheres, config_.size_x,
                                                  config_.size_y, config_.size_z, config_.origin,
                                                  config_.use_signed_distance_field, config_.resolution,
                                                  config_.collision_tolerance, config_.max_propogation_distance,
                                                  config_.padding, config_.scale);
    env->setOctreeOccupiedLeavesOnly(config_.octree_occupied_leaves_only);
    if (!config_.devices.empty())
      env->setDevices(config_.devices);
    return env;
  }
  CollisionEnvB200Ptr allocateEnv(const std::shared_ptr<const CollisionEnvB200>& orig, const WorldPtr& world) const
  {
    return std::make_shared<CollisionEnvB200>(*orig, world);
  }
  CollisionEnvB200Ptr allocateEnv(const core::RobotModelConstPtr& robot_model) const
  {
    return allocateEnv(std::make_shared<World>(), robot_model);
  }

private:
  B200DistanceFieldConfig config_;
};
}  // namespace collision_detection_b200
