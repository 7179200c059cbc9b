/* b200df — B200-native batched state-validity engine for MoveIt's distance-field collision path.
 *
 * Thin C ABI between host code (the C++ CollisionEnv mirror in moveit_b200/host/, the MoveIt plugin described in
 * INTEGRATION.md, ctypes in tests/bench) and the hand-written sm_100a kernels in moveit_b200/csrc/.
 * Plain pointers and sizes only.  Every call returns a status; nothing throws across this boundary; there is no
 * CPU fallback: without a CUDA device every compute entry point returns B200DF_ERR_CUDA.
 *
 * Each entry point names the reference interface it replaces (paths relative to /root/reference/moveit_core).
 *
 * Conventions
 *   - poses / transforms: row-major 3x4 doubles (rotation | translation), implied last row 0 0 0 1
 *   - voxel arrays: x-major, z fastest, index = x*(ny*nz) + y*nz + z   (voxel_grid.h:426-429)
 *   - "host" entry points take host pointers and include the H2D/D2H copies; "_device" entry points take
 *     device pointers plus a cudaStream_t (passed as void*) and never synchronise the device.
 */
#ifndef B200DF_H
#define B200DF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DF_VERSION 100

enum
{
  B200DF_OK = 0,
  B200DF_ERR_INVALID = 1,     /* bad argument / inconsistent handles */
  B200DF_ERR_CUDA = 2,        /* CUDA runtime failure or no device */
  B200DF_ERR_NOMEM = 3,
  B200DF_ERR_UNSUPPORTED = 4, /* reference leaves it unimplemented too (e.g. continuous checks) */
};

/* joint types, matching the *JointModel::computeTransform set (robot_model/src/<type>_joint_model.cpp) */
enum
{
  B200DF_JOINT_FIXED = 0,
  B200DF_JOINT_REVOLUTE = 1, /* also continuous */
  B200DF_JOINT_PRISMATIC = 2,
  B200DF_JOINT_PLANAR = 3,
  B200DF_JOINT_FLOATING = 4,
};

/* shapes::ShapeType values this path can rasterise (distance_field.cpp:196-219) */
enum
{
  B200DF_SHAPE_SPHERE = 1,   /* dims: radius */
  B200DF_SHAPE_CYLINDER = 2, /* dims: radius, length (axis = local z) */
  B200DF_SHAPE_BOX = 4,      /* dims: x, y, z */
};

/* which tests a batch runs; evaluation order is irrelevant for a yes/no verdict (logical OR) */
enum
{
  B200DF_CHECK_ROBOT = 1, /* group spheres vs world field      getEnvironmentCollisions  collision_env_distance_field.cpp:1580-1662 */
  B200DF_CHECK_SELF = 2,  /* group spheres vs self field       getSelfCollisions         :277-356 */
  B200DF_CHECK_INTRA = 4, /* sphere pairs inside the group     getIntraGroupCollisions   :433-660 */
  B200DF_CHECK_LINK_FIELDS = 8, /* gradients only: per-link PosedDistanceField term of getSelfProximityGradients
                                   :395-418 (needs b200df_group_set_link_fields) */
};

/* joint-value element type of a batch */
enum
{
  B200DF_Q_F32 = 0,
  B200DF_Q_F64 = 1,
};

/* arithmetic of the validity kernels */
enum
{
  B200DF_PRECISION_EXACT = 0, /* fp64 FK + fp64 voxel indexing, operation order of the reference */
  B200DF_PRECISION_FAST = 1,  /* fp32 fast path; states whose verdict could depend on rounding are re-run EXACT:
                                 same verdicts as EXACT (fp64 joint values are screened rounded to fp32 and re-run
                                 unrounded).  Needs a robot that fits the constant-bank tables
                                 (<= 40 links, 128 spheres, 512 link pairs) and, with signed fields, every sphere
                                 radius above the collision tolerance (the negative distances then cannot change a
                                 verdict); otherwise the call silently runs EXACT */
  /* test hooks, not for production use */
  B200DF_PRECISION_DEBUG_SWEEP = 2,  /* the independently written round-1 sweep kernel (cross-check of EXACT) */
  B200DF_PRECISION_DEBUG_SCREEN = 3, /* FAST's screening pass alone: verdict 0 / 1, or 2 = undecided */
  B200DF_PRECISION_DEBUG_SMALL = 4,  /* the one-CTA-per-state kernel that EXACT and FAST pick for small batches, at
                                        any batch size (same verdicts as EXACT) */
};

/* CollisionType of a gradient entry (collision_distance_field_types.h:57-63) */
enum
{
  B200DF_TYPE_NONE = 0,
  B200DF_TYPE_SELF = 1,
  B200DF_TYPE_INTRA = 2,
  B200DF_TYPE_ENVIRONMENT = 3,
};

/* which kernels a verdict call runs (b200df_check_route) */
enum
{
  B200DF_ROUTE_EXACT = 0, /* one thread per state, fp64 */
  B200DF_ROUTE_FAST = 1,  /* fp32 screening pass + exact re-run of the undecided states */
  B200DF_ROUTE_SMALL = 2, /* one CTA per state (small batches; robots beyond the batched kernels' shared memory) */
};
/* why a B200DF_PRECISION_FAST call does not take the FAST route (bit mask) */
enum
{
  B200DF_NOFAST_TABLES = 1, /* robot beyond the constant-bank tables: > 40 links, > 128 spheres or > 512 link pairs */
  B200DF_NOFAST_SIGNED = 2, /* signed field and a sphere radius <= collision tolerance: fp64 signed predicate needed */
  B200DF_NOFAST_BATCH = 4,  /* batch too small for the screening pass to pay (same verdicts either way) */
  B200DF_NOFAST_SMEM = 8,   /* robot too large for the one-thread-per-state kernels: one CTA per state */
};
#define B200DF_MAX_LINK_FIELDS 32 /* links with geometry b200df_group_set_link_fields accepts (kernel bit masks) */

typedef struct b200df_model b200df_model; /* immutable robot description on the device */
typedef struct b200df_group b200df_group; /* immutable (model, link subset, ACM tables, thresholds) */
typedef struct b200df_field b200df_field; /* PropagationDistanceField replacement (occupancy + exact EDT) */

/* ---- errors / device ----------------------------------------------------------------------------------- */
const char* b200df_last_error(void); /* thread-local message of the last failing call on this thread */
int b200df_version(void);
int b200df_device_count(int* count);
int b200df_set_device(int device); /* device used by handles created afterwards on this thread */

/* ---- robot model -----------------------------------------------------------------------------------------
 * Replaces the slice of moveit::core::RobotModel / LinkModel / JointModel that FK and the sphere decomposition
 * read: robot_state.cpp:718-755, revolute_joint_model.cpp:66-75,222-259, prismatic_joint_model.cpp:119-144,
 * fixed_joint_model.cpp:94-97, planar_joint_model.cpp:327-331, floating_joint_model.cpp:224-229, and the
 * BodyDecomposition per link built by CollisionEnvDistanceField::addLinkBodyDecompositions
 * (collision_env_distance_field.cpp:1065-1093).  Links are in DFS pre-order (parent index < child index). */
typedef struct b200df_model_desc
{
  int32_t n_links;
  int32_t n_vars; /* RobotState variable count; a batch row holds n_vars joint values */
  const int32_t* parent;             /* [L] parent link index, -1 for the root */
  const int32_t* joint_type;         /* [L] B200DF_JOINT_* of the parent joint */
  const double* axis;                /* [L][3] joint axis (normalised by the engine like setAxis does) */
  const double* origin;              /* [L][12] LinkModel::getJointOriginTransform */
  const int32_t* origin_is_identity; /* [L] LinkModel::jointOriginTransformIsIdentity */
  const int32_t* var_index;          /* [L] first variable of the parent joint, -1 if it has none */
  const int32_t* mimic_var;          /* [L] first variable of the mimicked joint, -1 if not a mimic joint */
  const double* mimic_mult;          /* [L] JointModel::getMimicFactor */
  const double* mimic_offset;        /* [L] JointModel::getMimicOffset */
  /* per-link BodyDecomposition */
  const int32_t* has_geometry;  /* [L] !LinkModel::getShapes().empty() */
  const int32_t* sphere_offset; /* [L+1] CSR offsets into spheres */
  const double* spheres;        /* [S][4] CollisionSphere relative_vec_ xyz + radius_ */
  const double* bsphere;        /* [L][4] BodyDecomposition::getRelativeBoundingSphere centre xyz + radius */
  const int64_t* point_offset;  /* [L+1] CSR offsets into points (may be NULL when no link has points) */
  const double* points;         /* [P][3] BodyDecomposition::getCollisionPoints (link frame) */
} b200df_model_desc;

int b200df_model_create(const b200df_model_desc* desc, b200df_model** out);
int b200df_model_destroy(b200df_model* model);

/* FK only: RobotState::updateLinkTransforms for n states.  q: host [n][n_vars] (q_dtype), out: host [n][L][12]. */
int b200df_model_fk(const b200df_model* model, const void* q, int q_dtype, int64_t n, double* link_transforms);

/* BodyDecomposition (collision_distance_field_types.cpp:290-334) for one link / world shape list, host side
 * set-up helper: spheres along the bounding cylinders (determineCollisionSpheres :46-79), lattice points
 * (find_internal_points.cpp:39-64, rasterised on the GPU) and the merged bounding sphere.
 * Two-pass: call with caps 0 to get the counts. */
int b200df_body_decomposition(int32_t n_shapes, const int32_t* shape_types, const double* dims /*[n][3]*/,
                              const double* poses /*[n][12]*/, double resolution, double padding,
                              double* spheres /*[cap][4]*/, int64_t sphere_cap, int64_t* n_spheres,
                              double* points /*[cap][3]*/, int64_t point_cap, int64_t* n_points,
                              double* bsphere /*[4]*/);

/* ---- collision group ---------------------------------------------------------------------------------------
 * Replaces DistanceFieldCacheEntry (collision_common_distance_field.h:56-167) as produced by
 * CollisionEnvDistanceField::generateDistanceFieldCacheEntry (:728-975): the group's link list, which links
 * check against the self field and which link pairs are tested (already resolved from the
 * AllowedCollisionMatrix by the host, getEntry()==ALWAYS semantics :785-805), plus the field parameters the
 * thresholds depend on. */
typedef struct b200df_group_desc
{
  int32_t n_group_links;
  const int32_t* group_links;   /* [n] model link indices in model order (getUpdatedLinkModelNames / all links) */
  const uint8_t* self_enabled;  /* [n] self_collision_enabled_ */
  const uint8_t* intra_enabled; /* [n][n] intra_group_collision_enabled_ (only j > i is read) */
  double resolution;                /* fields used with this group must have this resolution */
  double collision_tolerance;       /* collision_tolerance_ (default 0.0) */
  double max_propagation_distance;  /* max_propogation_distance_ (default 0.25) */
} b200df_group_desc;

int b200df_group_create(const b200df_model* model, const b200df_group_desc* desc, b200df_group** out);
int b200df_group_destroy(b200df_group* group);
int b200df_group_sphere_count(const b200df_group* group, int32_t* n_spheres); /* S of the gradient outputs */
/* The per-link distance fields of GroupStateRepresentation::link_distance_fields_
 * (getGroupStateRepresentation, collision_env_distance_field.cpp:1187-1210): for each group link WITH geometry, in
 * group order, a field in the link's own frame (cube of edge 2*bounding radius around the relative bounding-sphere
 * centre, filled with the link's BodyDecomposition points), or NULL.  enabled[i*G+j] != 0 when the ACM lets link j's
 * field act on link i's spheres (entry absent or NEVER, :388-404); pass NULL fields to switch the term off.
 * The fields must stay alive and unmodified while the group uses them.  G = links with geometry, at most 32. */
int b200df_group_link_count(const b200df_group* group, int32_t* n_links_with_geometry);
int b200df_group_set_link_fields(b200df_group* group, b200df_field* const* fields, const uint8_t* enabled);

/* ---- distance field ----------------------------------------------------------------------------------------
 * Replaces distance_field::PropagationDistanceField (propagation_distance_field.cpp) with an occupancy grid plus
 * an EXACT clamped squared EDT rebuilt on the GPU (separable passes).  origin is the grid CORNER exactly as in the
 * PropagationDistanceField constructor; CollisionEnvDistanceField passes centre - size/2 (:946-948, :1804-1806). */
typedef struct b200df_field_desc
{
  double size[3];
  double origin[3];
  double resolution;
  double max_distance;
  int32_t propagate_negative; /* use_signed_distance_field_ */
} b200df_field_desc;

int b200df_field_create(const b200df_field_desc* desc, b200df_field** out);
int b200df_field_destroy(b200df_field* field);
/* out4 = nx, ny, nz, max_distance_sq_ (propagation_distance_field.cpp:78, voxel_grid.h:387) */
int b200df_field_dims(const b200df_field* field, int32_t* out4);
/* the geometry the field was created with (getSizeX/Y/Z, getOriginX/Y/Z, getResolution, getUninitializedDistance) */
int b200df_field_get_desc(const b200df_field* field, b200df_field_desc* out);
int b200df_field_reset(b200df_field* field);                                        /* reset()                 :507-525 */
int b200df_field_add_points(b200df_field* field, const double* xyz, int64_t n);     /* addPointsToField       :176-195 */
int b200df_field_remove_points(b200df_field* field, const double* xyz, int64_t n);  /* removePointsFromField  :197-218 */
/* addPointsToField for a point cloud that already lives on the GPU (fp64 [n][3] device pointer, e.g. produced by a
 * perception pipeline); stream-ordered, no host synchronisation */
int b200df_field_add_points_device(b200df_field* field, const double* xyz_dev, int64_t n, void* stream);
int b200df_field_update_points(b200df_field* field, const double* old_xyz, int64_t n_old, const double* new_xyz,
                               int64_t n_new);                                      /* updatePointsInField    :120-174 */
int b200df_field_add_voxels(b200df_field* field, const int32_t* ijk, int64_t n);    /* addNewObstacleVoxels   :220-300 */
/* addShapeToField / removeShapeFromField (distance_field.cpp:221-226,320-330): findInternalPointsConvex
 * rasterised on the GPU.  pose = shape pose in the field frame; padding/scale as bodies::Body. */
int b200df_field_add_shape(b200df_field* field, int32_t shape_type, const double* dims3, const double* pose12,
                           double scale, double padding);
int b200df_field_remove_shape(b200df_field* field, int32_t shape_type, const double* dims3, const double* pose12,
                              double scale, double padding);
/* BodyDecomposition points of a shape rasterised in its own frame, then posed (the CollisionEnvDistanceField
 * world-object route: getBodyDecompositionCacheEntry + PosedBodyPointDecomposition, types.cpp:368-379, Q10) */
int b200df_field_add_posed_body(b200df_field* field, int32_t shape_type, const double* dims3,
                                const double* pose12, double padding);
/* the removePointsFromField half of notifyObjectChange (collision_env_distance_field.cpp:1731-1743) for the same
 * decomposition; voxels shared with another object are cleared as well, like the reference (no ref-count) */
int b200df_field_remove_posed_body(b200df_field* field, int32_t shape_type, const double* dims3,
                                   const double* pose12, double padding);
/* occupied octree leaves (centre xyz + edge) -> points per getOcTreePoints (distance_field.cpp:236-280): a leaf no
 * bigger than a voxel is its centre, a bigger one a lattice accumulated with += resolution (Q19).  Expanded and
 * voxelised on the device (one thread per lattice point; the loops are re-run literally for the rounding).  The
 * remove form clears the same voxels (removePointsFromField of the same decomposition). */
int b200df_field_add_octree_leaves(b200df_field* field, const double* leaves4, int64_t n);
int b200df_field_remove_octree_leaves(b200df_field* field, const double* leaves4, int64_t n);
/* Rebuilds the EDT if the occupancy changed (queries do this implicitly).  Returns the device time of the
 * rebuild in milliseconds through *ms when non-NULL. */
int b200df_field_rebuild(b200df_field* field, float* ms);
/* d^2 per voxel as int32 (negative != 0: negative_distance_square_) */
int b200df_field_download_d2(b200df_field* field, int32_t* out, int negative);
int b200df_field_download_occupancy(b200df_field* field, uint8_t* out);
/* DistanceField::getDistanceGradient (distance_field.cpp:62-86) for n host points:
 * dist[n], grad[n][3], in_bounds[n] */
int b200df_field_distance_gradient(b200df_field* field, const double* xyz, int64_t n, double* dist, double* grad,
                                   uint8_t* in_bounds);
/* writeToStream / readFromStream (propagation_distance_field.cpp:636-761): the bit-packed occupancy payload
 * (before / after zlib, which stays on the host) */
int b200df_field_pack_occupancy(b200df_field* field, uint8_t* out, int64_t cap, int64_t* n_bytes);
int b200df_field_unpack_occupancy(b200df_field* field, const uint8_t* bytes, int64_t n_bytes);
/* device view for replication across GPUs (NCCL broadcast / peer copy of the built arrays) */
int b200df_field_device_arrays(b200df_field* field, void** d2, void** neg_d2, int64_t* n_bytes, int32_t* elem_size);
/* after the arrays behind b200df_field_device_arrays were overwritten in place (e.g. by a broadcast): the distances
 * are served as they are.  The occupancy is NOT touched by this route: replicate it as well
 * (b200df_field_import_device / b200df_field_copy_from do) before editing the replica, or the next edit rebuilds
 * from the replica's own (stale) occupancy. */
int b200df_field_mark_built(b200df_field* field);
int b200df_field_copy_from(b200df_field* dst, b200df_field* src); /* same geometry; peer copy when on two GPUs */
/* copies of the built arrays to / from caller-owned device buffers (occupancy: n voxels bytes; d2 / neg_d2:
 * n_bytes each), e.g. the tensor an NCCL broadcast runs on.  NULL pointers are skipped. */
int b200df_field_export_device(b200df_field* field, void* occ_dev, void* d2_dev, void* neg_d2_dev, void* stream);
int b200df_field_import_device(b200df_field* field, const void* occ_dev, const void* d2_dev, const void* neg_d2_dev,
                               void* stream);
/* Reference-field mode.  The engine's own distances are the EXACT clamped squared EDT; the reference's
 * PropagationDistanceField::propagatePositive (propagation_distance_field.cpp:389-444) is a wavefront that
 * over-estimates d^2 on a few voxels in 10^4 and depends on the insertion order.  A caller that holds
 * the reference's own distance_square_ / negative_distance_square_ arrays (int per voxel, z fastest) uploads them
 * here; occupancy becomes (d2 == 0) and every query, verdict and gradient is then served from exactly these values
 * until the next occupancy edit (add / remove / reset / unpack), which rebuilds the exact EDT.  Values outside
 * [0, max_distance_sq_] are refused (B200DF_ERR_INVALID, field unchanged).  neg_d2 is required for signed fields. */
int b200df_field_upload_d2(b200df_field* field, const int32_t* d2, const int32_t* neg_d2);
int b200df_field_values_uploaded(const b200df_field* field, int32_t* uploaded); /* 1 while uploaded values are served */
/* PropagationDistanceField::getNearestCell (propagation_distance_field.h:411-431) for n cells ijk[n][3]:
 * dist[n] (negative inside obstacles of a signed field), pos[n][3], found[n].  The engine keeps no closest_point_
 * per voxel; pos is the first voxel (x-major scan of offsets) of the other kind at exactly the cell's squared
 * distance - one of the equidistant nearest cells, not necessarily the one the reference's wavefront recorded.
 * found = 0 (pos = the cell itself) for obstacle cells of unsigned fields, clamped cells and out-of-range cells. */
int b200df_field_nearest_cells(b200df_field* field, const int32_t* ijk, int64_t n, double* dist, int32_t* pos,
                               uint8_t* found);

/* ---- batched checks ----------------------------------------------------------------------------------------
 * N joint vectors -> N verdicts.  Replaces, per state, checkRobotCollision (:1500-1519), checkSelfCollision
 * (:254-261 -> :182-204) and checkCollision (:1441-1464) with req.contacts == false.
 * world / self may be NULL when the corresponding flag is not set.
 * Any batch size is served: up to 16 384 states run on a one-CTA-per-state kernel (a single state answers in ~20 us
 * through the host-buffer entry, which stages small inputs in a mapped page-locked buffer of a cached per-call
 * context), larger batches on the one-thread-per-state kernels, pipelined over PCIe in 2^20-state chunks.  The entry
 * points are re-entrant: concurrent calls on the same handles each take their own context. */
int b200df_check_batch(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                       int q_dtype, int64_t n, int flags, int precision, uint8_t* verdict);
int b200df_check_batch_device(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q_dev,
                              int q_dtype, int64_t n, int flags, int precision, uint8_t* verdict_dev, void* stream);

/* The kernels b200df_check_batch(_device) would run for these arguments; nothing is launched and the fields are not
 * rebuilt.  *route = B200DF_ROUTE_*; *why_not_fast = B200DF_NOFAST_* bits when precision is FAST but the route is not.
 * The fall-backs never change a verdict - they cost time (EXACT is ~3x slower than FAST) - and this call is how a
 * caller (or a test) sees them instead of having them happen silently. */
int b200df_check_route(const b200df_group* group, const b200df_field* world, const b200df_field* self, int q_dtype,
                       int64_t n, int flags, int precision, int32_t* route, int32_t* why_not_fast);

/* Batched motion validation: M straight joint-space motions qa[i] -> qb[i] (host, fp64 [m][n_vars]), each cut into
 * `segments` pieces; the states at t = j/segments, j = 1..segments, are interpolated on the device with the
 * reference's JointModel::interpolate arithmetic (revolute_joint_model.cpp:125-148, planar_joint_model.cpp:180-205;
 * var_is_angular[v] == 1 marks continuous-revolute / planar-theta variables that take the shortest arc, 2 the first and
 * 3 the other three quaternion variables (x, y, z, w) of a floating joint, which are slerped like
 * FloatingJointModel::interpolate, floating_joint_model.cpp:130-158; NULL = all linear)
 * and checked like b200df_check_batch.  first_invalid[i] = the first j whose state collides, 0 = motion valid — what
 * ompl::base::DiscreteMotionValidator::checkMotion + StateValidityChecker::isValid
 * (moveit_planners/ompl/ompl_interface/src/detail/state_validity_checker.cpp:76-121) decide one state at a time; the
 * last valid fraction of OMPL's lastValid is (first_invalid - 1) / segments.  Only the end states cross PCIe. */
int b200df_check_edges(const b200df_group* group, b200df_field* world, b200df_field* self, const double* qa,
                       const double* qb, int64_t m, int32_t segments, const uint8_t* var_is_angular, int flags,
                       int precision, int32_t* first_invalid);

/* getCollisionGradients (:1536-1557) for N states, flattened over the group's spheres (links with geometry, in
 * group order): dist[n][S], grad[n][S][3], type[n][S] (B200DF_TYPE_*), centers[n][S][3] (may be NULL).
 * flags: B200DF_CHECK_SELF = self field term of getSelfProximityGradients, B200DF_CHECK_INTRA =
 * getIntraGroupProximityGradients, B200DF_CHECK_ROBOT = getEnvironmentProximityGradients. */
int b200df_gradients_batch(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                           int q_dtype, int64_t n, int flags, double* dist, double* grad, uint8_t* type,
                           double* centers);

/* the same on device buffers (q, dist, grad, type, centers all device pointers; centers may be NULL); never
 * synchronises — for consumers that keep the gradients on the GPU */
int b200df_gradients_batch_device(const b200df_group* group, b200df_field* world, b200df_field* self,
                                  const void* q_dev, int q_dtype, int64_t n, int flags, double* dist_dev,
                                  double* grad_dev, uint8_t* type_dev, double* centers_dev, void* stream);

/* The two calls above with single-precision outputs: dist, grad and centers are the fp64 results rounded once while
 * they are written (type stays exact; "no candidate", DBL_MAX in the fp64 form, is FLT_MAX).  north_star's tolerance
 * for this path is 1e-5 m; rounding costs < 5e-7 for distances below 1 m and gradient components below 8.  29 bytes
 * per sphere instead of 57 cross PCIe - the host-buffer call is bound by exactly that copy (CHOMP's consumer,
 * chomp_optimizer.cpp:924-961, reads the values into its own arrays either way). */
int b200df_gradients_batch_f32(const b200df_group* group, b200df_field* world, b200df_field* self, const void* q,
                               int q_dtype, int64_t n, int flags, float* dist, float* grad, uint8_t* type,
                               float* centers);
int b200df_gradients_batch_device_f32(const b200df_group* group, b200df_field* world, b200df_field* self,
                                      const void* q_dev, int q_dtype, int64_t n, int flags, float* dist_dev,
                                      float* grad_dev, uint8_t* type_dev, float* centers_dev, void* stream);

/* distanceRobot / distanceSelf are stubs in the reference (collision_env_distance_field.h:108-123,178-198, F5).
 * Defined here as min over group spheres of (field distance - radius); out[n]. */
int b200df_distance_batch(const b200df_group* group, b200df_field* world, const void* q, int q_dtype, int64_t n,
                          double* min_clearance);

/* posed sphere centres (PosedBodySphereDecomposition::updatePose, types.cpp:390-397): out[n][S][3] */
int b200df_sphere_centers(const b200df_group* group, const void* q, int q_dtype, int64_t n, double* centers);

/* ---- single-process multi-GPU ---------------------------------------------------------------------------------
 * MoveIt is ONE process whose planner threads share a const PlanningScene
 * (moveit_ros/planning/planning_components_tools/src/evaluate_collision_checking_speed.cpp:44-57,121-124): a
 * b200df_multi replicates one (model, group) on several devices of the calling process.  Fields are replicated by
 * b200df_multi_set_field (rebuild once on the source's device, then one peer copy of occupancy + distances per replica
 * over NVLink; call it again after every change of the source - the only exchange of the path, no collective).
 * b200df_multi_check_batch cuts the batch into contiguous N/G shards, drives each device from its own host thread
 * through b200df_check_batch (H2D, kernels, D2H of its shard) and the verdicts land in the caller's single buffer:
 * the same bytes b200df_check_batch would have written. */
typedef struct b200df_multi b200df_multi;
int b200df_multi_create(const b200df_model_desc* model, const b200df_group_desc* group, const int32_t* devices,
                        int32_t n_devices, b200df_multi** out);
int b200df_multi_destroy(b200df_multi* multi);
int b200df_multi_device_count(const b200df_multi* multi, int32_t* n);
/* which: 0 = world field, 1 = self field; src may live on any device; NULL drops the replicas */
int b200df_multi_set_field(b200df_multi* multi, int which, b200df_field* src);
int b200df_multi_field(const b200df_multi* multi, int which, int32_t device_slot, b200df_field** out); /* a replica */
int b200df_multi_check_batch(const b200df_multi* multi, const void* q, int q_dtype, int64_t n, int flags, int precision,
                             uint8_t* verdict);
/* device-resident shards: q_dev[i] / verdict_dev[i] live on device slot i, n[i] states each, streams[i] (may be NULL)
 * is a stream of that device; queues the work on every device and returns without synchronising */
int b200df_multi_check_batch_device(const b200df_multi* multi, const void* const* q_dev, int q_dtype, const int64_t* n,
                                    int flags, int precision, uint8_t* const* verdict_dev, void* const* streams);

/* page-locked host memory for batch inputs / results: the host-buffer entry points copy at PCIe speed from / to such
 * buffers (pageable memory costs an extra staging copy in the driver; 342 MB of CHOMP gradients: 8 ms vs 78 ms) */
int b200df_host_alloc(size_t bytes, void** out);
int b200df_host_free(void* p);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t b200df_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200DF_H */
