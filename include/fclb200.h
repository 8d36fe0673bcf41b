/* fclb200.h -- C ABI of libfclb200.so, the B200 batched narrowphase engine.
 *
 * This is the drop-in boundary.  mind-fcl has no FFI/plugin
 * seam of its own: its public seam is the C++ template API
 *     fcl::collide(o1, tf1, o2, tf2, request, result)
 *         reference include/fcl/narrowphase/collision.h:55-64,
 *         impl      include/fcl/narrowphase/collision_interface-inl.h:13-44
 * plus the internal function-pointer table
 *     detail::CollisionFunctionMatrix<S>::collision_matrix[NODE][NODE]
 *         reference include/fcl/narrowphase/detail/collision_func_matrix.h:67-78.
 * Each entry point below replaces a LOOP of those calls over a batch of
 * independent (geometry pair, pose pair) queries.  The host-side C++ mirror of
 * the fcl API (include/fcl_b200/fcl.h) and the binding a mind-fcl maintainer
 * would add sit directly on top of this header.
 *
 * Conventions
 *   - plain C, POD structs, SoA outputs, caller-allocated buffers, int error
 *     codes (0 = ok), no exceptions, no callbacks.
 *   - scalar_type: FCLB_F32 or FCLB_F64 -- the scalar S the reference would be
 *     instantiated with.  All arithmetic is done in S on the device.
 *   - pose: 12 S = rotation 3x3 row-major, then translation xyz
 *     (the meaningful part of Eigen::Transform<S,3,Isometry>, common/types.h:91).
 *   - "_host" entry points take HOST pointers and do H2D + kernels + D2H on the
 *     engine's streams (pinned buffers from fclb_host_alloc make the copies
 *     asynchronous);  "_dev" entry points take DEVICE pointers (cudaMalloc'd or
 *     fclb_dev_alloc'd) and only launch kernels + synchronise.
 *   - handles are opaque, immutable after creation, usable from any thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns FCLB_ERR_NO_DEVICE.
 */
#ifndef FCLB200_H_
#define FCLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCLB_F32 0
#define FCLB_F64 1

/* error codes */
#define FCLB_OK 0
#define FCLB_ERR_NO_DEVICE 1
#define FCLB_ERR_CUDA 2
#define FCLB_ERR_BAD_ARG 3
#define FCLB_ERR_UNSUPPORTED 4
#define FCLB_ERR_CAPACITY 5

/* shape type codes; the first six equal cvx_collide::GJKShapeType
 * (reference include/fcl/cvx_collide/gjk_shape.h:12-30). */
#define FCLB_BOX 0       /* p = side x,y,z            (geometry/shape/box.h)       */
#define FCLB_SPHERE 1    /* p = radius                (geometry/shape/sphere.h)    */
#define FCLB_ELLIPSOID 2 /* p = radii x,y,z           (geometry/shape/ellipsoid.h) */
#define FCLB_CAPSULE 3   /* p = radius, lz            (geometry/shape/capsule.h)   */
#define FCLB_CONE 4      /* p = radius, lz            (geometry/shape/cone.h)      */
#define FCLB_CYLINDER 5  /* p = radius, lz            (geometry/shape/cylinder.h)  */
#define FCLB_CONVEX 6    /* geom = convex handle slot (geometry/shape/convex.h)    */

typedef struct {
  uint32_t type;  /* FCLB_* */
  uint32_t geom;  /* FCLB_CONVEX: index returned by fclb_convex_upload */
  double p[3];    /* parameters, converted to S on upload */
} fclb_shape;

typedef struct {
  uint32_t shape1, shape2; /* indices into the shape table */
} fclb_pair;

/* CollisionRequest<S> by value (reference narrowphase/collision_request.h:51-94)
 * plus the GJKSolver knobs fcl::collide fixes (gjk_solver-inl.h:1121-1130). */
#define FCLB_PEN_DISABLED 0
#define FCLB_PEN_DEFAULT_GJK_EPA 1
#define FCLB_PEN_DIRECTED 2
#define FCLB_PEN_INCREMENTAL_MIN 3
typedef struct {
  uint32_t max_contacts;     /* num_max_contacts_; 0 => every query returns 0 (collision-inl.h:79-84) */
  uint32_t penetration_mode; /* FCLB_PEN_* */
  double dir[3];             /* escape direction for the MPR penetration modes: UNIT vector, as useDirectedPenetration stores it */
  double binary_tol;         /* GJK/MPR tolerance; <=0 => 1e-6 (collision_request.h:66) */
  double distance_tol;       /* EPA tolerance;     <=0 => 1e-6 (collision_request.h:67) */
  uint32_t gjk_max_iter;     /* 0 => 128 */
  uint32_t epa_max_faces;    /* 0 => 256 */
  uint32_t epa_max_iter;     /* 0 => 255 */
  uint32_t flags;            /* reserved, 0 */
} fclb_request;

typedef uint64_t fclb_handle;

/* ---- engine ------------------------------------------------------------ */
int fclb_init(int device);             /* bind the calling process to one GPU (one process per GPU) */
int fclb_device_count(void);           /* 0 when no CUDA device is visible */
/* One process, several GPUs (SURVEY.md 8b "fclb_init(int n_devices)", 8e): one engine -- streams, scratch, a replica
 * of every geometry uploaded afterwards -- per device (n_devices <= 0: every visible device).  Every *_host batch entry
 * point then shards its batch by contiguous query range over the devices, one host thread per device, each device
 * writing its slice of the caller's (pinned) output arrays; there is no collective.  *_dev entry points, the
 * broadphase trees and fclb_scene_self_collide_* act on the calling thread's current device (fclb_set_device).
 * Call before the first upload. */
int fclb_init_devices(int n_devices);
int fclb_num_devices(void);            /* engines created so far */
int fclb_set_device(int slot);         /* thread-local: the engine later calls of this thread use */
const char* fclb_last_error(void);     /* thread-local message for the last non-zero return */
const char* fclb_version(void);

/* pinned host / device buffers for the batch arrays */
int fclb_host_alloc(void** p, size_t bytes);
int fclb_host_free(void* p);
int fclb_dev_alloc(void** p, size_t bytes);
int fclb_dev_free(void* p);
int fclb_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
int fclb_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
int fclb_synchronize(void);

/* ---- geometry upload (host pointers; device copies are immutable) -------- */
/* Convex<S>: vertices (n_verts x 3 doubles) + faces in the reference encoding
 * (count, v0, v1, ... per face; geometry/shape/convex.h:84-108).  The engine
 * derives the neighbour CSR, the 6 axis seeds and the interior point exactly as
 * Convex<S>'s constructor does (convex-inl.h:52-75,153-200,379-407).
 * *slot is the value to put in fclb_shape.geom. */
int fclb_convex_upload(const double* verts, int n_verts, const int* faces, int faces_len, int num_faces,
                       uint32_t* slot);
/* shape table (replaces the per-call ShapeBase<S> objects) */
int fclb_shapes_upload(const fclb_shape* shapes, uint32_t n_shapes, fclb_handle* table);
int fclb_release(fclb_handle h);

/* ---- batched queries ------------------------------------------------------ */
/* fcl::distance semantics == detail::GJKSolver<S>::shapeDistance
 * (gjk_solver-inl.h:801-808; closed forms :902-988; generic GJK :762-798):
 *   ok[q]=1, dist>0, p1/p2 = world-frame witness points   when separated
 *   ok[q]=0, dist=-1                                       otherwise.
 * gjk_tol<=0 => constants<S>::gjk_default_tolerance() (eps^(7/8));
 * gjk_max_iter==0 => 128.  Any of out_* may be NULL. */
int fclb_distance_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                             size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                             void* out_p1, void* out_p2, uint8_t* out_ok);
int fclb_distance_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                            void* out_p1, void* out_p2, uint8_t* out_ok);

/* Compact pose encoding for the host link.  FCLB_POSE_QT7: 7 S per pose = unit quaternion x, y, z, w then translation
 * x, y, z (28 B instead of 48 B in float).  The device expands it with the arithmetic of Eigen's
 * QuaternionBase::toRotationMatrix in S, i.e. to exactly the tf.linear() of a Transform3<S> built from that quaternion
 * (mind-fcl's tests build their poses that way: test/test_fcl_utility.h generateRandomTransform), then runs the same
 * kernels: results are those of the 12-S entry point on the expanded poses, bit for bit.  The PCIe-bound host path moves
 * 64 B instead of 104 B per query. */
#define FCLB_POSE_RT12 0
#define FCLB_POSE_QT7 1
int fclb_distance_batch_qt_host(fclb_handle shapes, const fclb_pair* pairs, const void* qt_poses1, const void* qt_poses2,
                                size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                                void* out_p1, void* out_p2, uint8_t* out_ok);
/* the expansion alone (DEVICE pointers): n poses of 7 S -> n poses of 12 S */
int fclb_expand_poses_dev(const void* qt_poses, size_t n, int scalar_type, void* out_poses12);

/* Signed distance == detail::GJKSolver<S>::shapeSignedDistance (gjk_solver-inl.h:810-880): always the generic
 * GJK (no closed forms), GJKSolver's default tolerances / limits (:1121-1130);
 *   separated:   ok = 1, dist > 0, witness points as above
 *   penetrating: ok = 1, dist = -(EPA depth), p1 / p2 = tf1 * EPA witness points
 *   otherwise:   ok = 0, dist = -1 (EPA failed, or GJK neither separated nor intersecting) */
int fclb_signed_distance_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                    size_t n, int scalar_type, void* out_dist, void* out_p1, void* out_p2,
                                    uint8_t* out_ok);
int fclb_signed_distance_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                   size_t n, int scalar_type, void* out_dist, void* out_p1, void* out_p2,
                                   uint8_t* out_ok);

/* fcl::collide semantics for shape-shape pairs
 * (collision_func_matrix-inl.h:340 ShapeShapeCollide -> shape_pair_intersect-inl.h:49).
 * out_contacts: max_keep records per query of 9 S = {b1, b2, normal[3], pos[3], depth}
 * (Contact<S>, narrowphase/contact.h:46-87; b1=b2=-1 for shape pairs);
 * out_counts[q] = result.numContacts().  out_contacts may be NULL. */
int fclb_collide_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep,
                            void* out_contacts, uint32_t* out_counts);
int fclb_collide_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                           size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep,
                           void* out_contacts, uint32_t* out_counts);

/* The cvx_collide path driven directly, as the reference's tests do
 * (test/cvx_collide/test_epa2_with_gjk2.cpp:76-162): GJK(max_iter, tol) and, on
 * Intersect, EPA(max_faces, max_iter, tol) on the GJK simplex.
 * out_gjk: GJK_Status (gjk.h:13-29); out_epa: EPA_Status (epa.h:14-22) or -1;
 * out_geom: 7 S per query {depth, p0[3], p1[3]} in shape-1's frame. */
int fclb_gjk_epa_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, const fclb_request* req, int32_t* out_gjk, int32_t* out_epa,
                            void* out_geom);
int fclb_gjk_epa_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                           size_t n, int scalar_type, const fclb_request* req, int32_t* out_gjk, int32_t* out_epa,
                           void* out_geom);

/* ---- translational continuous collision, shape vs shape -------------------------------------------
 * fcl::translational_ccd(o1, tf1, o1_displacement, o2, tf2, request, result) per query
 * (narrowphase/continuous_collision-inl.h:21-36 -> ShapePairTranslationalCollisionSolver::RunShapePair,
 * detail/ccd/shape_pair_ccd-inl.h:139-170): does shape 1, translated along displacement, ever touch shape 2?
 *   Box-Box                 the swept-box separating-axis test with its time-of-collision interval
 *                           (BoxPairTranslationalCCD::IsDisjoint, box_pair_ccd-inl.h), whatever the request type
 *   every other pair        MPR on the Minkowski difference of (shape 1 swept, shape 2) (gjk_ccd-inl.h:21-114);
 *                           kBoxApproximate first runs the swept-box test on the shapes' local AABBs (toc = its
 *                           interval), kOneTocSample derives one time sample from MPR's final portal (:116-186)
 * displacements: 4 S per query = TranslationalDisplacement{unit_axis_in_shape1 xyz, scalar_displacement}.
 *   out_hit[q] = 1 when a contact is reported;  out_toc[2q..] = ContinuousCollisionContact::toc (lower, upper),
 *   (-1, -1) when the request type computes none.  Bit-identical to the reference (tests/test_ccd_gpu.py). */
#define FCLB_CCD_NOT_REQUESTED 0
#define FCLB_CCD_BOX_APPROXIMATE 1
#define FCLB_CCD_ONE_TOC_SAMPLE 2
typedef struct {
  uint32_t request_type;           /* FCLB_CCD_* = TimeOfCollisionRequestType (detail/ccd/ccd_request.h:10-16) */
  uint32_t max_contacts;           /* num_max_contacts; a shape pair has at most one contact */
  double zero_movement_tolerance;  /* <= 0 => 1e-4 */
  double gjk_tolerance;            /* <= 0 => 1e-6 */
  int32_t max_gjk_iterations;      /* <= 0 => 128 */
  uint32_t flags;                  /* reserved, 0 */
} fclb_ccd_request;
int fclb_translational_ccd_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                      const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                      uint8_t* out_hit, void* out_toc);
int fclb_translational_ccd_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                     const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                     uint8_t* out_hit, void* out_toc);

/* ---- translational continuous collision, shape vs mesh ----------------------------------------------
 * fcl::translational_ccd(shape, tf_shape, displacement, BVHModel<OBB<S>>, tf_mesh, request, result) per query
 * (matrix entries ShapeBVH_ / BVH_ShapeTranslationalCollideImpl<Shape, OBB<S>>,
 * detail/ccd/translational_collision_func_matrix-inl.h:469-487 -> bvh_ccd_solver-inl.h:120-218 RunSweptBV): the mesh's
 * OBB tree is walked with the swept-box test (BoxPairTranslationalCCD::IsDisjoint, every node narrowing its parent's
 * time-of-collision interval) and every surviving triangle runs the swept-volume MPR against the shape
 * (RunShapeSimplex, shape_pair_ccd-inl.h:196-212).  `bvh` is a handle of fclb_bvh_upload / fclb_bvh_build: the
 * reference's BVHModel<OBB<S>> has the same hierarchy and internal boxes, its leaf boxes (3-point fit) are derived
 * from the triangles on the device.
 *   displacements  4 S per query: unit axis + scalar displacement of the MOVING object, in its own frame
 *   mesh_moves     0: the shape moves (fcl::translational_ccd(shape, ..., mesh, ...));
 *                  1: the mesh moves (fcl::translational_ccd(mesh, ..., shape, ...), RunMeshShape :572-584)
 *   out_counts[q]  ContinuousCollisionResult::num_contacts(), at most req->max_contacts
 *   out_prim[q * max_keep + k], out_toc[(q * max_keep + k) * 2 ..]   ContinuousCollisionContact::b2 (triangle id)
 *                  and ::toc of the k-th contact IN THE REFERENCE'S ORDER (its depth-first walk visits the right
 *                  child first and stops at max_contacts); -1 beyond the count.  toc is the leaf's swept-box
 *                  interval (kNotRequested, kBoxApproximate) or MPR's time sample (kOneTocSample).
 * Convex shapes with exactly 1, 2, 3 or 6 vertices return FCLB_ERR_UNSUPPORTED (computeBV<OBB, Convex> uses the
 * small-set fits for them).  Bit-identical to the reference (tests/test_ccd_mesh_gpu.py). */
int fclb_translational_ccd_mesh_batch_host(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids,
                                           const void* poses_shape, const void* poses_mesh, const void* displacements,
                                           size_t n, int scalar_type, const fclb_ccd_request* req, int mesh_moves,
                                           uint32_t max_keep, uint32_t* out_counts, int64_t* out_prim, void* out_toc);
int fclb_translational_ccd_mesh_batch_dev(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids,
                                          const void* poses_shape, const void* poses_mesh, const void* displacements,
                                          size_t n, int scalar_type, const fclb_ccd_request* req, int mesh_moves,
                                          uint32_t max_keep, uint32_t* out_counts, int64_t* out_prim, void* out_toc);

/* mesh vs mesh: fcl::translational_ccd(BVHModel<OBB<S>>, tf1, displacement, BVHModel<OBB<S>>, tf2, request, result)
 * (TranslationalDisplacementBVH_PairSolverImpl<S, OBB<S>>, bvh_ccd_solver-inl.h:425-551): the walk over node PAIRS, the
 * swept-volume MPR of (triangle 1 swept, triangle 2) per surviving leaf pair.  displacements: mesh 1's, in its frame.
 * out_prim[(q * max_keep + k) * 2 ..] = (b1, b2) triangle ids of the k-th contact in the reference's order. */
int fclb_translational_ccd_mesh_pair_batch_host(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                                const void* displacements, size_t n, int scalar_type,
                                                const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                int64_t* out_prim, void* out_toc);
int fclb_translational_ccd_mesh_pair_batch_dev(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                               const void* displacements, size_t n, int scalar_type,
                                               const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                               int64_t* out_prim, void* out_toc);

/* shape vs heightmap / octree: fcl::translational_ccd(shape, tf_shape, displacement, scene, tf_scene, request, result) and the
 * scene-first entry (TranslationalDisplacementHeightMapSolver::RunShapeHeightMap / RunHeightMapShape,
 * detail/ccd/heightmap_ccd_solver-inl.h:8-166; TranslationalDisplacementOctreeSolver::RunShapeOctree / RunOctreeShape,
 * detail/ccd/octree2_ccd_solver-inl.h:58-260): the scene's box hierarchy is walked with the fixed-orientation swept-box
 * test (box_pair_ccd_fixed_orientation-inl.h), every surviving pixel / voxel / fully occupied node runs
 * RunShapePair<Shape, Box> with the request's type.  scene_kind: FCLB_SCENE_HEIGHTMAP or FCLB_SCENE_OCTREE.
 *   scene_moves    0: the shape moves; 1: the scene moves (displacement in the scene's frame)
 *   out_code       ContinuousCollisionContact::b2 per contact: encodePixel / encodeOctree2Node, in the reference's order
 *   out_toc        2 S per contact (or (-1, -1) when the request type computes none)
 *   out_box        6 S per contact: ContinuousCollisionContact::o2_bv (min xyz, max xyz, scene frame); may be NULL */
int fclb_translational_ccd_scene_batch_host(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                            const void* poses_shape, const void* poses_scene, const void* displacements,
                                            size_t n, int scalar_type, const fclb_ccd_request* req, int scene_moves,
                                            uint32_t max_keep, uint32_t* out_counts, int64_t* out_code, void* out_toc,
                                            void* out_box);
int fclb_translational_ccd_scene_batch_dev(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                           const void* poses_shape, const void* poses_scene, const void* displacements,
                                           size_t n, int scalar_type, const fclb_ccd_request* req, int scene_moves,
                                           uint32_t max_keep, uint32_t* out_counts, int64_t* out_code, void* out_toc,
                                           void* out_box);

/* heightmap / octree vs mesh: fcl::translational_ccd(scene, tf_scene, displacement, BVHModel<OBB<S>>, tf_mesh, ...) and the
 * mesh-first entry (RunHeightMapObbBVH / RunObbBVH_HeightMap, heightmap_ccd_solver-inl.h:168-366; RunOctreeObbBVH /
 * RunObbBVH_Octree, octree2_ccd_solver-inl.h:225-470): the walk over (box node, mesh node) pairs with the OBB swept-box
 * test, RunShapeSimplex<Box>(pixel / voxel box swept, triangle) per surviving leaf pair.
 *   mesh_moves  0: the scene geometry moves (displacement in its frame); 1: the mesh moves (displacement in the mesh frame)
 *   out_ids[(q * max_keep + k) * 2 ..] = (pixel / node code, triangle id) of the k-th contact in the reference's order;
 *   out_toc 2 S, out_box 6 S (the scene box, ContinuousCollisionContact::o1_bv) per contact; out_toc / out_box may be NULL */
int fclb_translational_ccd_scene_mesh_batch_host(int scene_kind, fclb_handle scene, fclb_handle bvh, const void* poses_scene,
                                                 const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                                 const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep,
                                                 uint32_t* out_counts, int64_t* out_ids, void* out_toc, void* out_box);
int fclb_translational_ccd_scene_mesh_batch_dev(int scene_kind, fclb_handle scene, fclb_handle bvh, const void* poses_scene,
                                                const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                                const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep,
                                                uint32_t* out_counts, int64_t* out_ids, void* out_toc, void* out_box);

/* heightmap / octree vs heightmap / octree: fcl::translational_ccd(scene1, tf1, displacement, scene2, tf2, ...)
 * (RunHeightMapPair heightmap_ccd_solver-inl.h:636-779, RunHeightMapOctree / RunOctreeHeightMap :369-632, RunOctreePair
 * octree2_ccd_solver-inl.h:467-922).  Geometry 1 moves, the displacement is given in its frame.  A contact is a pair of
 * terminal boxes (bottom-layer pixel; fully occupied octree node; voxel of a partial leaf) whose swept-box test from
 * [0, 1] is not disjoint; its toc is that test's interval (for every request type: these routines never consult it).
 *   out_ids[(q * max_keep + k) * 2 ..] = (code on geometry 1, code on geometry 2) of the k-th contact in the reference's
 *     order; pixel = x << 16 | y; octree node = encodeOctree2Node(index, is_leaf, voxel) in an octree pair, the plain
 *     node_vector_index against a heightmap (as the reference writes b2 there).  For (octree, heightmap) the reference's
 *     own contact names the heightmap as o1; the arrays here stay in the caller's argument order.
 *   out_toc 2 S, out_box 12 S (box on geometry 1, box on geometry 2: o1_bv / o2_bv) per contact; either may be NULL */
int fclb_translational_ccd_scene_pair_batch_host(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                                 const void* poses2, const void* displacements, size_t n, int scalar_type,
                                                 const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                 int64_t* out_ids, void* out_toc, void* out_box);
int fclb_translational_ccd_scene_pair_batch_dev(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                                const void* poses2, const void* displacements, size_t n, int scalar_type,
                                                const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                int64_t* out_ids, void* out_toc, void* out_box);

/* ---- meshes: BVHModel<OBBRSS<S>> flattened by the caller ------------------------
 * (reference geometry/bvh/BVH_model.h:63-196, BV_node_base.h:50-82).  Only the
 * OBB half of OBBRSS is ever read by collide (math/bv/OBBRSS-inl.h:130-135).
 *   obb         15 S per node: axis 3x3 row-major (axis(i,j)), To xyz, extent xyz
 *   first_child BVNodeBase::first_child per node (>0: children at fc, fc+1;
 *               <0: leaf holding primitive -(fc+1)); node 0 is the root
 *   tri_verts   9 S per triangle, indexed by primitive id
 * The tree is built on the host by the caller (mind-fcl's own builder in an
 * integration); the device copy is immutable. */
int fclb_bvh_upload(const void* obb, const int32_t* first_child, int n_nodes, const void* tri_verts, int n_tris,
                    int scalar_type, fclb_handle* bvh);
/* Host-side mirror of BVHModel<OBBRSS<S>>::beginModel / addSubModel / endModel
 * (geometry/bvh/BVH_model-inl.h:402-570; OBBRSS fitter detail/BV_fitter-inl.h:324-345;
 * mean split detail/BV_splitter-inl.h:361-372): builds the tree on the host with the
 * reference's arithmetic (node-for-node identical OBBs) and uploads it.
 * verts: n_verts x 3 doubles (rounded once to S); tris: n_tris x 3 vertex indices. */
int fclb_bvh_build(const double* verts, int n_verts, const int32_t* tris, int n_tris, int scalar_type,
                   fclb_handle* bvh);
/* The same builder run ON THE DEVICE, level by level (fit sums in primitive order, the reference's in-place swap pass
 * replayed per node, node ids from the depth-first order): obb / first_child / triangle order identical to fclb_bvh_build
 * and to the reference's BVHModel (tests/test_bvh_build_gpu.py).  verts / tris are host arrays. */
int fclb_bvh_build_device(const double* verts, int n_verts, const int32_t* tris, int n_tris, int scalar_type,
                          fclb_handle* bvh);
/* same builder, host only (no GPU needed): obb / first_child / tri_verts must hold
 * 15*(2*n_tris-1) S / (2*n_tris-1) / 9*n_tris S entries; *n_nodes = nodes written. */
int fclb_bvh_build_host(const double* verts, int n_verts, const int32_t* tris, int n_tris, int scalar_type, void* obb,
                        int32_t* first_child, void* tri_verts, int* n_nodes);
/* Refit ON THE DEVICE: BVHModel::beginReplaceModel / replaceSubModel / endReplaceModel(refit = true, bottomup = false)
 * (geometry/bvh/BVH_model-inl.h:318-375 -> refitTreeTopDown :624-637): the vertices move, the topology stays, every node
 * is fitted again from the triangles it covers with the OBBRSS fitter of build time (detail/BV_fitter-inl.h:324-345).
 * tri_verts: the new 9 S per triangle, indexed by primitive id (HOST pointer for _host, DEVICE pointer for _dev).  One
 * warp per node; the node OBBs equal the reference's refit bit for bit.  The step before the path for a deforming /
 * re-perceived scene mesh: no host rebuild, no re-upload of the tree. */
int fclb_bvh_refit_host(fclb_handle bvh, const void* tri_verts, int n_tris);
int fclb_bvh_refit_dev(fclb_handle bvh, const void* tri_verts, int n_tris);
/* BVHModel::endReplaceModel(refit = true, bottomup = true) / endUpdateModel(true, true) -- the reference's DEFAULT arguments
 * (BVH_model.h:137): refitTreeBottomUp (BVH_model-inl.h:580-617): leaf box = 3-point fit of its triangle, inner box =
 * left + right (OBB<S>::operator+: merge_largedist / merge_smalldist, math/bv/OBB-inl.h:116-293).  Node-for-node identical
 * to the reference's OBBs (tests/test_bvh_refit_gpu.py). */
int fclb_bvh_refit_bottomup_host(fclb_handle bvh, const void* tri_verts, int n_tris);
int fclb_bvh_refit_bottomup_dev(fclb_handle bvh, const void* tri_verts, int n_tris);
int fclb_bvh_info(fclb_handle bvh, int* n_nodes, int* n_tris, int* scalar_type);
/* copies the tree back in the fclb_bvh_upload layout (any pointer may be NULL) */
int fclb_bvh_export(fclb_handle bvh, void* obb, int32_t* first_child, void* tri_verts);
int fclb_bvh_release(fclb_handle bvh);
/* fcl::collide(BVHModel<OBBRSS>, tf1, BVHModel<OBBRSS>, tf2, request, result) per query
 * (-> OrientedNodeBVHSolver::MeshIntersect, traversal/collision/bvh_solver-inl.h:75).
 * request.penetration_mode must be FCLB_PEN_DISABLED (boolean / counting collide):
 *   out_counts[q]   = result.numContacts() = min(#intersecting triangle pairs, max_contacts)
 *   out_first_pair  = (b1, b2) of ONE colliding triangle pair or (-1,-1); optional.
 *                     Not necessarily the pair the reference's DFS reports first. */
int fclb_bvh_collide_batch_host(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                                int scalar_type, const fclb_request* req, uint32_t* out_counts,
                                int32_t* out_first_pair);
int fclb_bvh_collide_batch_dev(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                               int scalar_type, const fclb_request* req, uint32_t* out_counts,
                               int32_t* out_first_pair);
/* fcl::collide(BVH, BVH) with every kept contact.  Boolean request: ids only (b1, b2 of every kept triangle pair, the
 * contact records are not written).  request.useDefaultPenetration(): contact generation of
 * Intersect::intersect_Triangle (traversal/collision/intersect-inl.h:795-888) through trianglePairIntersect
 * (shape_pair_intersect-inl.h:201-252): up to two contact points per intersecting triangle pair, sharing one
 * normal and depth.
 *   out_counts[q] = result.numContacts() = min(sum of contact points, max_contacts)
 *   out_ids[(q*max_keep + k)*2 ..]      = b1, b2 of the k-th stored contact (or -1, -1)
 *   out_contacts[(q*max_keep + k)*7 ..] = normal[3], pos[3], penetration_depth, as the reference reports them
 *     (it maps the mesh-1-frame contact with tf2 -- SimplexIntersect passes tf2 as the contact frame,
 *     shape_pair_intersect-inl.h:266 -- which is reproduced)
 * Which contacts are the first max_keep follows the device traversal order. */
int fclb_bvh_collide_contacts_batch_host(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                         size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep,
                                         uint32_t* out_counts, int32_t* out_ids, void* out_contacts);
int fclb_bvh_collide_contacts_batch_dev(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                        size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep,
                                        uint32_t* out_counts, int32_t* out_ids, void* out_contacts);
/* BV-pair and leaf-pair tests executed by the most recent mesh batch call */
int fclb_bvh_last_visit_counts(uint64_t* n_bv, uint64_t* n_leaf);

/* ---- mesh vs shape ----------------------------------------------------------------
 * fcl::collide(BVHModel<OBBRSS>, tf_mesh, Shape, tf_shape, request, result) per query
 * (collision_func_matrix-inl.h:390-408 -> OrientedNodeBVHSolver::MeshShapeIntersect,
 * traversal/collision/bvh_solver-inl.h:8-72; leaf = GJKSolver::shapeTriangleIntersect,
 * gjk_solver-inl.h:540-600).  shape_ids[q] indexes the shape table.
 * out_counts[q] = result.numContacts() for ANY request mode (with a penetration request the contacts are generated
 * as in fclb_scene_shape_contacts_batch and dropped); for a boolean request:
 *   out_counts[q]    = min(#triangles hit, max_contacts)
 *   out_first_tri[q] = b1 of ONE hit triangle or -1 (optional; not necessarily the DFS-first) */
int fclb_bvh_shape_collide_batch_host(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids,
                                      const void* poses_mesh, const void* poses_shape, size_t n, int scalar_type,
                                      const fclb_request* req, uint32_t* out_counts, int32_t* out_first_tri);
int fclb_bvh_shape_collide_batch_dev(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids,
                                     const void* poses_mesh, const void* poses_shape, size_t n, int scalar_type,
                                     const fclb_request* req, uint32_t* out_counts, int32_t* out_first_tri);
/* ---- heightmap vs shape ----------------------------------------------------------
 * LayeredHeightMap<S> (geometry/heightmap/layered_heightmap.h, flat_heightmap.h): the caller hands
 * over the BOTTOM layer verbatim -- heights in millimetres, index = y * full_x + x
 * (flat_heightmap-inl.h:121-124), full = 2 * half shape, half shapes powers of two -- plus the
 * bottom resolution; the coarser layers (2x2 max, layered_heightmap-inl.h:77-101) are rebuilt here.
 * upper_bound_mm: FlatHeightMap::height_upper_bound_in_mm (0 => the maximum height). */
int fclb_heightmap_upload(const uint16_t* heights_mm, uint32_t full_x, uint32_t full_y, double resolution_x,
                          double resolution_y, uint32_t upper_bound_mm, fclb_handle* hm);
int fclb_heightmap_release(fclb_handle hm);
/* LayeredHeightMap<S> built ON THE DEVICE from a point cloud (n_points x 3 S, device / host pointer): the
 * rasteriser of FlatHeightMap<S>::updateHeightsByPointGenerationFunctor (flat_heightmap-inl.h:249-272: per pixel the
 * maximum of uint16(z * 1000) over its points, negative z and out-of-range points dropped) as one atomicMax per
 * point, then the layer pyramid (layered_heightmap-inl.h:77-101).  Heights equal the reference's bit for bit
 * (a maximum is order-free).  This is the step before the path: the scene changes every perception cycle. */
int fclb_heightmap_build_dev(const void* points, size_t n_points, double resolution_x, double resolution_y,
                             uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, fclb_handle* hm);
int fclb_heightmap_build_points_host(const void* points, size_t n_points, double resolution_x, double resolution_y,
                                     uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, fclb_handle* hm);
/* shape of an uploaded / built map and a copy of one layer (0 = bottom, k = k levels above it) */
int fclb_heightmap_info(fclb_handle hm, uint32_t* n_layers, uint32_t* full_x, uint32_t* full_y, uint32_t* upper_bound_mm);
int fclb_heightmap_export(fclb_handle hm, uint32_t layer, uint16_t* heights_mm);
/* Host-side mirror of FlatHeightMap<S>::updateHeightsByPointGenerationFunctor (flat_heightmap-inl.h:249-272):
 * rasterises n_points (x, y, z doubles, rounded once to S) into heights_mm (2*half_x * 2*half_y, caller-zeroed
 * or holding earlier heights).  Host only, no GPU needed. */
int fclb_heightmap_build_host(const double* points, size_t n_points, double resolution_x, double resolution_y,
                              uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, uint16_t* heights_mm);
/* fcl::collide(HeightMapCollisionGeometry, tf_hm, Shape, tf_shape, request, result) per query
 * (collision_func_matrix-inl.h:99-118 -> heightmap_solver_traverse-inl.h:23-118; leaf = pixel Box vs Shape,
 * heightmap_solver_leaf-inl.h:10-31).  Any request mode (see fclb_bvh_shape_collide_batch); boolean request:
 *   out_counts[q]      = result.numContacts() = min(#pixel boxes hit, max_contacts)
 *   out_first_pixel[q] = b1 = encodePixel (x << 16 | y, heightmap_types.h:53-58) of ONE hit pixel or -1 */
int fclb_heightmap_shape_collide_batch_host(fclb_handle hm, fclb_handle shapes, const uint32_t* shape_ids,
                                            const void* poses_hm, const void* poses_shape, size_t n, int scalar_type,
                                            const fclb_request* req, uint32_t* out_counts, int32_t* out_first_pixel);
int fclb_heightmap_shape_collide_batch_dev(fclb_handle hm, fclb_handle shapes, const uint32_t* shape_ids,
                                           const void* poses_hm, const void* poses_shape, size_t n, int scalar_type,
                                           const fclb_request* req, uint32_t* out_counts, int32_t* out_first_pixel);
/* ---- octree vs shape ---------------------------------------------------------------
 * octree2::Octree<S> (geometry/octree2/octree.h, octree_node.h:21-48) handed over as its flat arrays:
 *   inner_children   8 x u32 per inner node, 0xffffffff = absent child (OctreeInnerNode::children); node 0 = root
 *   inner_full       inner_nodes_fully_occupied()[i] as bytes
 *   leaf_bits        OctreeLeafNode::child_occupied of every leaf-layer node (2x2x2 bitmask)
 *   pruned_or_null   OctreePruneInfo::prune_internal_nodes as bytes, or NULL
 *   root_aabb        root_bv(): min xyz, max xyz;  num_layers = n_layers()
 * fcl::collide(Octree2CollisionGeometry, tf_octree, Shape, tf_shape, request, result) per query
 * (collision_func_matrix-inl.h:253-273 -> octree2_solver_traverse-inl.h:12-136); any request mode; boolean request:
 *   out_counts[q]     = result.numContacts() = min(#voxel boxes hit, max_contacts)
 *   out_first_node[q] = b1 = encodeOctree2Node (octree2_solver_leaf-inl.h:10-20) of ONE hit box or -1 */
int fclb_octree_upload(const uint32_t* inner_children, const uint8_t* inner_full, uint32_t n_inner,
                       const uint8_t* leaf_bits, uint32_t n_leaf, const uint8_t* pruned_or_null,
                       const double* root_aabb, int num_layers, fclb_handle* octree);
int fclb_octree_release(fclb_handle octree);
/* Host-side mirror of octree2::Octree<S>(resolution, bottom_half_shape) + rebuildTree
 * (geometry/octree2/octree-inl.h:15-142, octree_construction-inl.h:10-74,111-205): inserts n_points
 * (x, y, z doubles, rounded once to S; out-of-range points dropped) in stream order, so inner / leaf nodes get the
 * reference's numbering (the contact ids of the octree kernels), then sets the fully-occupied flags.
 * bottom_half_shape: power of two >= 2.  Host only, no GPU needed.  Always writes *n_inner / *n_leaf (and
 * root_aabb[6], *num_layers when non-NULL); returns FCLB_ERR_CAPACITY without writing the arrays when they are
 * NULL or smaller than that -- call once for the sizes, once for the data (the size query keeps its tree for a data
 * call that follows on the same thread with the same arguments, so the points are inserted once). */
int fclb_octree_build_host(const double* points, size_t n_points, double resolution, uint32_t bottom_half_shape,
                           int scalar_type, uint32_t* inner_children, uint8_t* inner_full, uint32_t inner_capacity,
                           uint32_t* n_inner, uint8_t* leaf_bits, uint32_t leaf_capacity, uint32_t* n_leaf,
                           double* root_aabb, int* num_layers);
/* Host-side mirror of octree2::pruneOctreeByOBB (geometry/octree2/octree_prune-inl.h:10-103) = what
 * Octree2CollisionGeometry::pruneBy(obb, rebuild_octree = false) computes: inner nodes inside the OBB are marked
 * pruned, voxels whose centre is inside it are cleared from the leaf masks, and the fully-occupied flags are
 * re-derived.  obb: axis[9] row-major (columns = box directions), To[3], extent[3].  pruned / inner_full / leaf_bits
 * are IN-OUT (the OctreePruneInfo being extended): pass zeros and the tree's own flags / masks for a first prune, the
 * previous outputs for a further one; upload the three with fclb_octree_upload.  Host only, no GPU needed. */
int fclb_octree_prune_host(const uint32_t* inner_children, uint32_t n_inner, uint32_t n_leaf, const double* root_aabb,
                           int num_layers, const double* obb, int scalar_type, uint8_t* pruned, uint8_t* inner_full,
                           uint8_t* leaf_bits);
/* Host-side mirror of Octree<S>::rebuildAccordingToPruneInfo (octree_construction-inl.h:247-369) = the
 * rebuild_octree = true half of pruneBy: drops the pruned inner nodes and renumbers inner and leaf nodes exactly as the
 * reference does, then re-derives the fully-occupied flags.  pruned / leaf_bits: the outputs of
 * fclb_octree_prune_host.  The out arrays must hold n_inner x 8 / n_inner / n_leaf entries (the tree never grows);
 * root box and layer count are unchanged.  Host only, no GPU needed. */
int fclb_octree_consolidate_host(const uint32_t* inner_children, uint32_t n_inner, const uint8_t* pruned,
                                 const uint8_t* leaf_bits, uint32_t n_leaf, int num_layers, uint32_t* out_children,
                                 uint8_t* out_full, uint32_t* out_n_inner, uint8_t* out_leaf_bits, uint32_t* out_n_leaf);
/* octree2::Octree<S>::rebuildTree ON THE DEVICE (points: n x 3 S, DEVICE pointer for _dev, HOST pointer for
 * _points_host), with the reference's node numbering: a node's creation time in the reference is (index of the first
 * point of the stream that reaches it, its depth), so the device sorts the voxel path keys, folds them level by level
 * into unique prefixes with their first point, and ranks the nodes by that creation time -- no sequential insert.
 * inner_children / inner_full / leaf_bits equal the reference's arrays (and the host mirror's) exactly, so contact ids
 * are unchanged.  fclb_octree_export / _info hand the arrays back. */
int fclb_octree_build_dev(const void* points, size_t n_points, double resolution, uint32_t bottom_half_shape, int scalar_type,
                          fclb_handle* octree);
int fclb_octree_build_points_host(const void* points, size_t n_points, double resolution, uint32_t bottom_half_shape,
                                  int scalar_type, fclb_handle* octree);
int fclb_octree_info(fclb_handle octree, uint32_t* n_inner, uint32_t* n_leaf, int* num_layers, double* root_aabb);
int fclb_octree_export(fclb_handle octree, uint32_t* inner_children, uint8_t* inner_full, uint8_t* leaf_bits);
/* the same builder followed by fclb_octree_upload (no prune mask) */
int fclb_octree_build(const double* points, size_t n_points, double resolution, uint32_t bottom_half_shape, int scalar_type,
                      fclb_handle* octree);
int fclb_octree_shape_collide_batch_host(fclb_handle octree, fclb_handle shapes, const uint32_t* shape_ids,
                                         const void* poses_octree, const void* poses_shape, size_t n, int scalar_type,
                                         const fclb_request* req, uint32_t* out_counts, int64_t* out_first_node);
int fclb_octree_shape_collide_batch_dev(fclb_handle octree, fclb_handle shapes, const uint32_t* shape_ids,
                                        const void* poses_octree, const void* poses_shape, size_t n, int scalar_type,
                                        const fclb_request* req, uint32_t* out_counts, int64_t* out_first_node);
/* ---- every contact of a scene-vs-shape query ---------------------------------------------------
 * fcl::collide(scene geometry, tf1, Shape, tf2, request, result) with ALL its contacts, for every request mode:
 *   FCLB_PEN_DISABLED            the boolean traversal; ids only (the contact records are zero)
 *   FCLB_PEN_DEFAULT_GJK_EPA     request.useDefaultPenetration(): each leaf runs the shape-pair leaf stage with contacts --
 *                                ShapeSimplexIntersect -> shapeTriangleIntersect (GJK + EPA; Sphere: closed form) for a
 *                                mesh triangle (bvh_solver-inl.h:52-66, gjk_solver-inl.h:479-581), ShapeIntersect<Box, Shape>
 *                                (boxBox2 / closed forms / GJK + EPA) for a pixel or voxel box
 *                                (heightmap_solver_leaf-inl.h:10-31, octree2_solver_leaf-inl.h:22-44)
 *   FCLB_PEN_DIRECTED / _INCREMENTAL_MIN  collisionPenetrationMPR (narrowphase/collision_penetration-inl.h:189-252):
 *                                the boolean traversal, then computePenetrationMPR between the contact's leaf geometry
 *                                (the mesh triangle / the pixel or voxel Box, :34-95) and the shape
 *   out_counts[q]              = result.numContacts() (<= request.max_contacts)
 *   out_b1[q*max_keep + k]     = Contact::b1 of the k-th stored contact (triangle id, encodePixel,
 *                                encodeOctree2Node) or -1; k < min(count, max_keep)
 *   out_contacts[(q*max_keep + k)*7 ..] = normal[3], pos[3], penetration_depth (depth -1: MPR reported failure)
 * Which contacts are the first max_keep follows the device (MPR modes: traversal order; DefaultGJK_EPA: ascending b1),
 * not the reference's DFS order. */
#define FCLB_SCENE_BVH 0
#define FCLB_SCENE_HEIGHTMAP 1
#define FCLB_SCENE_OCTREE 2
int fclb_scene_shape_contacts_batch_host(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                         const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                         const fclb_request* req, uint32_t max_keep, uint32_t* out_counts,
                                         int64_t* out_b1, void* out_contacts);
int fclb_scene_shape_contacts_batch_dev(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                        const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                        const fclb_request* req, uint32_t max_keep, uint32_t* out_counts,
                                        int64_t* out_b1, void* out_contacts);
/* ---- scene vs scene: heightmap / octree against heightmap / octree / mesh ------------------------
 * fcl::collide(g1, tf1, g2, tf2, request, result) per query for the non-convex pairs of the collision matrix
 * (collision_func_matrix-inl.h:774-857), boolean request (penetration_mode FCLB_PEN_DISABLED):
 *   (HEIGHTMAP, HEIGHTMAP)  heightMapPairIntersect    traversal/heightmap/heightmap_solver_traverse-inl.h:189-296
 *   (HEIGHTMAP, BVH)        heightMapBVHIntersect     :298-404
 *   (HEIGHTMAP, OCTREE)     heightMapOctreeIntersect  :406-570
 *   (OCTREE, BVH)           octreeBVHIntersect        traversal/octree2/octree2_solver_traverse-inl.h:138-288
 *   (OCTREE, OCTREE)        octreePairIntersect       :290-447
 * A contact is a leaf pair that passes the reference's leaf test: two boxes (pixel box, voxel box, fully
 * occupied octree node) that the strict 15-axis FixedRotationBoxDisjoint does not separate
 * (heightmap_solver_leaf-inl.h:33-68, octree2_solver_leaf-inl.h:85-404), or a box and a triangle with
 * boxTriangleIntersect (heightmap_solver_leaf-inl.h:70-88, octree2_solver_leaf-inl.h:46-66).
 *   out_counts[q]          = result.numContacts() = min(#leaf pairs hit, max_contacts)
 *   out_b1/out_b2[q*max_keep + k] = Contact::b1 / b2 of the k-th stored contact (encodePixel,
 *                            encodeOctree2Node, triangle id; the octree side of a (HEIGHTMAP, OCTREE)
 *                            contact is the bare node_vector_index, as in the reference) or -1;
 *                            both NULL = counts only
 * Which contacts are the first max_keep follows the device traversal order, not the reference's. */
int fclb_scene_pair_collide_batch_host(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                       const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                       uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2);
int fclb_scene_pair_collide_batch_dev(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                      const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                      uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2);
/* The same pairs with a penetration request.  useDirectedPenetration(dir) / useIncrementalMinimumDistancePenetration(dir)
 * (collisionPenetrationMPR, narrowphase/collision_penetration-inl.h:189-252): the boolean traversal, then
 * computePenetrationMPR between the two leaf geometries of each contact (:34-95: a Box of Contact::o1_bv / o2_bv in
 * the pose of its heightmap / octree, the mesh triangle b2 in the mesh pose).  useDefaultPenetration(): every leaf pair
 * runs ShapeIntersect<Box, Box> (boxBox2, up to four contacts) or ShapeSimplexIntersect<Box> (GJK + EPA) on the two leaf
 * geometries (heightmap_solver_leaf-inl.h:56-88, octree2_solver_leaf-inl.h:46-404, incl. the reverse_tree12 cases).
 *   out_contacts[(q*max_keep + k)*7 ..] = normal[3], pos[3], penetration_depth (depth -1: MPR reported failure) */
int fclb_scene_pair_contacts_batch_host(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                        const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                        uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2,
                                        void* out_contacts);
int fclb_scene_pair_contacts_batch_dev(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                       const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                       uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2,
                                       void* out_contacts);
/* node tests and leaf (shape-triangle / shape-pixel) tests of the most recent scene batch call */
int fclb_scene_last_visit_counts(uint64_t* n_node, uint64_t* n_leaf);

/* ---- broadphase: device-side AABB tree ---------------------------------------------
 * detail::BinaryAABB_Tree<S, Alloc> (broadphase/binary_AABB_tree.h:44-58, -inl.h:70-573) over
 * BroadphaseObjectInfo{AABB, user_id} (broadphase/broadphase_common.h:13-22).
 * aabbs: 6 S per object = min xyz, max xyz.  The reported pair SET equals the reference's
 * ({(a,b): AABB_a overlaps AABB_b}, math/bv/AABB-inl.h:82-88); the order of the list, and for
 * self pairs which id comes first, follow the device tree (Morton-order linear BVH), not the
 * reference's median-split tree.  out_pairs: 2 ids per pair; *n_pairs = pairs found; if that
 * exceeds cap the call returns FCLB_ERR_CAPACITY after writing the first cap pairs' worth of
 * nothing useful -- call again with a larger buffer (out_pairs == NULL just counts). */
int fclb_broadphase_build_host(const void* aabbs, const uint64_t* user_ids, size_t n, int scalar_type,
                               fclb_handle* tree);                         /* Rebuild / BuildTreeExternal */
int fclb_broadphase_build_dev(const void* aabbs, const uint64_t* user_ids, size_t n, int scalar_type,
                              fclb_handle* tree);                          /* same, DEVICE input arrays  */
int fclb_broadphase_release(fclb_handle tree);
int fclb_broadphase_self_pairs_host(fclb_handle tree, uint64_t* out_pairs, size_t cap, size_t* n_pairs);  /* SelfCollision */
int fclb_broadphase_self_pairs_dev(fclb_handle tree, uint64_t* out_pairs, size_t cap, size_t* n_pairs);   /* DEVICE out buffer */
/* TreeCollision: pairs (id in tree_a, id in tree_b) */
int fclb_broadphase_tree_pairs_host(fclb_handle tree_a, fclb_handle tree_b, uint64_t* out_pairs, size_t cap,
                                    size_t* n_pairs);
/* SingleObjectCollision for n query boxes: pairs (leaf id, object id) */
int fclb_broadphase_query_pairs_host(fclb_handle tree, const void* aabbs, const uint64_t* object_ids, size_t n,
                                     uint64_t* out_pairs, size_t cap, size_t* n_pairs);
/* UpdateObjectAABB (binary_AABB_tree-inl.h:332-355) for n objects of a tree built with
 * fclb_broadphase_build_host.  As in the reference the leaf box GROWS to old U new (node.bv += new_AABB);
 * ancestors are refitted.  An unknown id fails the call (the reference returns false). */
int fclb_broadphase_update_host(fclb_handle tree, const uint64_t* user_ids, const void* new_aabbs, size_t n);
uint64_t fclb_broadphase_last_visits(void); /* node / leaf boxes tested by the most recent pair search */
/* CollisionObject<S>::computeAABB (narrowphase/collision_object-inl.h:141-154) for n objects:
 * out_aabbs = 6 S per object.  Tight translate when the rotation is the identity, else the
 * bounding-sphere box centre +- aabb_radius. */
int fclb_compute_aabb_batch_host(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n,
                                 int scalar_type, void* out_aabbs);
int fclb_compute_aabb_batch_dev(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n,
                                int scalar_type, void* out_aabbs);
/* candidate (object id, object id) pairs -> the per-query arrays fclb_collide_batch_dev consumes
 * (pairs of shape-table indices, pose of the first and of the second object).  DEVICE pointers. */
int fclb_gather_pairs_dev(const uint64_t* id_pairs, size_t n_pairs, const uint32_t* shape_ids, const void* poses,
                          int scalar_type, fclb_pair* out_pairs, void* out_poses1, void* out_poses2);

/* One scene end to end on the device -- the loop a planner runs per perception cycle:
 * computeAABB for every object (object i has user id i), tree build, SelfCollision, and boolean
 * fcl::collide (request as in fclb_collide_batch) on every candidate pair.
 *   *n_candidates = candidate pairs of the broadphase, *n_colliding = pairs with numContacts > 0;
 *   out_id_pairs / out_counts (optional, out_cap entries): the candidates and their numContacts.
 * _dev: shape_ids / poses are DEVICE arrays; _host: HOST arrays (H2D inside the call). */
int fclb_scene_self_collide_host(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n,
                                 int scalar_type, const fclb_request* req, size_t* n_candidates, size_t* n_colliding,
                                 uint64_t* out_id_pairs, uint32_t* out_counts, size_t out_cap);
int fclb_scene_self_collide_dev(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n,
                                int scalar_type, const fclb_request* req, size_t* n_candidates, size_t* n_colliding,
                                uint64_t* out_id_pairs, uint32_t* out_counts, size_t out_cap);

/* measured FP32 / FP64 FMA throughput of the bound device (FMA-chain microbenchmark, TFLOP/s):
 * the denominator for the compute-bound GJK / MPR / EPA kernels */
int fclb_measure_fp_peak(int scalar_type, double* tflops);
/* measured L2 read bandwidth (GB/s; 32 MB L2-resident buffer streamed by every SM with 128-bit loads): the denominator
 * for the traversal kernels, whose node / triangle / pixel arrays are L2-resident */
int fclb_measure_l2_bandwidth(double* gbs);

/* kernel launches issued by this process so far (bench.py's gpu_launches) */
uint64_t fclb_launch_count(void);
/* device time (ms, CUDA events on the engine's stream) of the most recent batch
 * call: query kernels only / whole device side including the bucketing pass */
double fclb_last_kernel_ms(void);
double fclb_last_call_ms(void);
/* per-launch records of the most recent batch call: kinds[i] = type1*8+type2 of
 * the bucket, counts[i] = queries in it, ms[i] = CUDA-event duration.
 * Returns the number of records (may exceed cap). */
int fclb_last_launches(int* kinds, uint64_t* counts, double* ms, int cap);
/* the engine's compute stream (cudaStream_t) so callers can order their own
 * work / events against it */
void* fclb_stream(void);

#ifdef __cplusplus
}
#endif
#endif /* FCLB200_H_ */
