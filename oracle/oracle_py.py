"""ctypes access to the two CPU oracles.  TEST INFRASTRUCTURE ONLY.

  * RefOracle  -> oracle/_ref/libfclref.so : the unmodified reference headers,
                  compiled against oracle/eigen_shim (kind "reference").
  * PortOracle -> oracle/liboracle.so      : our CPU restatement (kind "port").

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module;
nothing under mind-fcl_b200/ or include/ does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libfclref.so")
PORT_PATH = os.path.join(_HERE, "liboracle.so")


class Shape(C.Structure):
    _fields_ = [("type", C.c_uint32), ("geom", C.c_uint32), ("p", C.c_double * 3)]


class Request(C.Structure):
    _fields_ = [
        ("max_contacts", C.c_uint32), ("penetration_mode", C.c_uint32), ("dir", C.c_double * 3),
        ("binary_tol", C.c_double), ("distance_tol", C.c_double),
        ("gjk_max_iter", C.c_uint32), ("epa_max_faces", C.c_uint32), ("epa_max_iter", C.c_uint32),
        ("flags", C.c_uint32),
    ]


def _shape_array(shapes):
    arr = (Shape * len(shapes))()
    for i, (t, g, p) in enumerate(shapes):
        arr[i].type = t
        arr[i].geom = g
        p = list(p) + [0.0] * (3 - len(p))
        arr[i].p[:] = p
    return arr


def _request(**kw):
    r = Request()
    r.max_contacts = kw.get("max_contacts", 1)
    r.penetration_mode = kw.get("penetration_mode", 0)
    r.dir[:] = kw.get("direction", (0.0, 0.0, 0.0))
    r.binary_tol = kw.get("binary_tol", 0.0)
    r.distance_tol = kw.get("distance_tol", 0.0)
    r.gjk_max_iter = kw.get("gjk_max_iter", 0)
    r.epa_max_faces = kw.get("epa_max_faces", 0)
    r.epa_max_iter = kw.get("epa_max_iter", 0)
    r.flags = 0
    return r


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _st(dtype):
    return 0 if np.dtype(dtype) == np.float32 else 1


class _Oracle:
    prefix = ""
    path = ""
    kind = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(f"{self.path} missing: run `make -C oracle`")
        self.lib = C.CDLL(self.path)
        self._convex_slots = 0

    def available(self):
        return True

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def hardware_threads(self):
        return int(self.fn("hardware_threads")())

    def register_convex(self, verts, faces, num_faces):
        verts = np.ascontiguousarray(verts, np.float64)
        faces = np.ascontiguousarray(faces, np.int32)
        f = self.fn("register_convex")
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        return int(f(_p(verts), verts.shape[0], _p(faces), faces.size, num_faces))

    def distance_batch(self, shapes, pairs, poses1, poses2, gjk_tol=0.0, gjk_max_iter=0, threads=1):
        n = len(pairs)
        dt = poses1.dtype
        dist = np.zeros(n, dt)
        p1 = np.zeros((n, 3), dt)
        p2 = np.zeros((n, 3), dt)
        ok = np.zeros(n, np.uint8)
        arr = _shape_array(shapes)
        f = self.fn("distance_batch")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double,
                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        rc = f(_st(dt), C.cast(arr, C.c_void_p), len(shapes), _p(pairs), _p(poses1), _p(poses2), n, gjk_tol,
               gjk_max_iter, _p(dist), _p(p1), _p(p2), _p(ok), threads)
        assert rc == 0
        return dist, p1, p2, ok

    def collide_batch(self, shapes, pairs, poses1, poses2, max_keep=1, threads=1, want_contacts=True, **req):
        n = len(pairs)
        dt = poses1.dtype
        contacts = np.zeros((n, max_keep, 9), dt) if want_contacts else None
        counts = np.zeros(n, np.uint32)
        arr = _shape_array(shapes)
        r = _request(**req)
        f = self.fn("collide_batch")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        rc = f(_st(dt), C.cast(arr, C.c_void_p), len(shapes), _p(pairs), _p(poses1), _p(poses2), n,
               C.cast(C.pointer(r), C.c_void_p), max_keep, _p(contacts), _p(counts), threads)
        assert rc == 0
        return counts, contacts

    def gjk_epa_batch(self, shapes, pairs, poses1, poses2, mode=0, threads=1, **req):
        n = len(pairs)
        dt = poses1.dtype
        status = np.zeros(n, np.int32)
        epa = np.zeros(n, np.int32)
        mpr = np.zeros(n, np.int32)
        geom = np.zeros((n, 7), dt)
        iters = np.zeros((n, 2), np.uint32)
        arr = _shape_array(shapes)
        r = _request(**req)
        f = self.fn("gjk_epa_batch")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        rc = f(_st(dt), C.cast(arr, C.c_void_p), len(shapes), _p(pairs), _p(poses1), _p(poses2), n,
               C.cast(C.pointer(r), C.c_void_p), mode, _p(status), _p(epa), _p(mpr), _p(geom), _p(iters), threads)
        assert rc == 0
        return status, epa, mpr, geom, iters


class RefOracle(_Oracle):
    prefix = "fclref_"
    path = REF_PATH
    kind = "reference"

    def signed_distance_batch(self, shapes, pairs, poses1, poses2, threads=1):
        n = len(pairs)
        dt = poses1.dtype
        dist = np.zeros(n, dt)
        p1 = np.zeros((n, 3), dt)
        p2 = np.zeros((n, 3), dt)
        ok = np.zeros(n, np.uint8)
        arr = _shape_array(shapes)
        f = self.fn("signed_distance_batch")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), C.cast(arr, C.c_void_p), len(shapes), _p(pairs), _p(poses1), _p(poses2), n, _p(dist), _p(p1), _p(p2),
          _p(ok), threads)
        return dist, p1, p2, ok

    def translational_ccd_batch(self, shapes, pairs, poses1, poses2, disp, request_type=0, zero_tol=0.0, gjk_tol=0.0, max_iter=0,
                                threads=1):
        """fcl::translational_ccd per query: (hit u8 [n], toc [n, 2])"""
        n = len(pairs)
        dt = poses1.dtype
        hit = np.zeros(n, np.uint8)
        toc = np.zeros((n, 2), dt)
        arr = _shape_array(shapes)
        f = self.fn("translational_ccd_batch")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double,
                      C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), C.cast(arr, C.c_void_p), len(shapes), _p(pairs), _p(poses1), _p(poses2), _p(disp), n, request_type, zero_tol,
          gjk_tol, max_iter, _p(hit), _p(toc), threads)
        return hit, toc


    # ---- translational continuous collision, shape vs mesh (reference BVHModel<OBB<S>>) ----
    def bvh_obb_create(self, verts, tris):
        verts = np.ascontiguousarray(verts, np.float64)
        tris = np.ascontiguousarray(tris, np.int32)
        f = self.fn("bvh_obb_create")
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        return int(f(_p(verts), len(verts), _p(tris), len(tris)))

    def bvh_obb_export(self, mesh_id, n_nodes, dtype):
        obb = np.zeros((n_nodes, 15), dtype)
        child = np.zeros(n_nodes, np.int32)
        f = self.fn("bvh_obb_export")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        f(mesh_id, _st(dtype), _p(obb), _p(child))
        return obb, child

    def translational_ccd_mesh_batch(self, mesh_id, shapes, shape_ids, poses_shape, poses_mesh, disp, request_type=0,
                                     max_contacts=1, zero_tol=0.0, mesh_moves=False, keep=8, threads=1):
        """fcl::translational_ccd(shape, mesh) per query: (counts u32 [n], primitive ids i64 [n, keep], toc [n, keep, 2])"""
        n = len(shape_ids)
        dt = poses_shape.dtype
        counts = np.zeros(n, np.uint32)
        prim = np.full((n, keep), -1, np.int64)
        toc = np.full((n, keep, 2), -1, dt)
        arr = _shape_array(shapes)
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        f = self.fn("translational_ccd_mesh_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                      C.c_uint32, C.c_double, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), mesh_id, C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses_shape), _p(poses_mesh), _p(disp), n,
          request_type, max_contacts, zero_tol, 1 if mesh_moves else 0, keep, _p(counts), _p(prim), _p(toc), threads)
        return counts, prim, toc

    def translational_ccd_mesh_pair_batch(self, mesh1, mesh2, poses1, poses2, disp, request_type=0, max_contacts=1, zero_tol=0.0,
                                          keep=8, threads=1):
        """fcl::translational_ccd(mesh, mesh) per query: (counts [n], (b1, b2) i64 [n, keep, 2], toc [n, keep, 2])"""
        n = len(poses1)
        dt = poses1.dtype
        counts = np.zeros(n, np.uint32)
        prim = np.full((n, keep, 2), -1, np.int64)
        toc = np.full((n, keep, 2), -1, dt)
        f = self.fn("translational_ccd_mesh_pair_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_double,
                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), mesh1, mesh2, _p(poses1), _p(poses2), _p(disp), n, request_type, max_contacts, zero_tol, keep, _p(counts),
          _p(prim), _p(toc), threads)
        return counts, prim, toc

    def translational_ccd_scene_batch(self, kind, scene_id, shapes, shape_ids, poses_shape, poses_scene, disp, request_type=0,
                                      max_contacts=1, scene_moves=False, keep=8, threads=1):
        """fcl::translational_ccd(shape, heightmap | octree): (counts, codes i64 [n, keep], toc [n, keep, 2], boxes [n, keep, 6])"""
        n = len(shape_ids)
        dt = poses_shape.dtype
        counts = np.zeros(n, np.uint32)
        code = np.full((n, keep), -1, np.int64)
        toc = np.full((n, keep, 2), -1, dt)
        box = np.zeros((n, keep, 6), dt)
        arr = _shape_array(shapes)
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        f = self.fn("translational_ccd_scene_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                      C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), kind, scene_id, C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses_shape), _p(poses_scene), _p(disp), n,
          request_type, max_contacts, 1 if scene_moves else 0, keep, _p(counts), _p(code), _p(toc), _p(box), threads)
        return counts, code, toc, box

    def translational_ccd_scene_mesh_batch(self, kind, scene_id, obb_mesh_id, poses_scene, poses_mesh, disp, request_type=0,
                                           max_contacts=1, mesh_moves=False, keep=8, threads=1):
        """fcl::translational_ccd(heightmap | octree, mesh): (counts, (code, triangle) i64 [n, keep, 2], toc, boxes)"""
        n = len(poses_scene)
        dt = poses_scene.dtype
        counts = np.zeros(n, np.uint32)
        ids = np.full((n, keep, 2), -1, np.int64)
        toc = np.full((n, keep, 2), -1, dt)
        box = np.zeros((n, keep, 6), dt)
        f = self.fn("translational_ccd_scene_mesh_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_int,
                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), kind, scene_id, obb_mesh_id, _p(poses_scene), _p(poses_mesh), _p(disp), n, request_type, max_contacts,
          1 if mesh_moves else 0, keep, _p(counts), _p(ids), _p(toc), _p(box), threads)
        return counts, ids, toc, box

    def translational_ccd_scene_pair_batch(self, kind1, id1, kind2, id2, poses1, poses2, disp, request_type=0, max_contacts=1,
                                           keep=8, threads=1):
        """fcl::translational_ccd(heightmap | octree, heightmap | octree): (counts, (code1, code2) i64 [n, keep, 2], toc, boxes
        [n, keep, 12]) in the caller's argument order"""
        n = len(poses1)
        dt = poses1.dtype
        counts = np.zeros(n, np.uint32)
        ids = np.full((n, keep, 2), -1, np.int64)
        toc = np.full((n, keep, 2), -1, dt)
        box = np.zeros((n, keep, 12), dt)
        f = self.fn("translational_ccd_scene_pair_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32,
                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(dt), kind1, id1, kind2, id2, _p(poses1), _p(poses2), _p(disp), n, request_type, max_contacts, keep, _p(counts), _p(ids),
          _p(toc), _p(box), threads)
        return counts, ids, toc, box

    # ---- meshes (reference BVHModel<OBBRSS<S>>) ----
    def bvh_create(self, verts, tris):
        verts = np.ascontiguousarray(verts, np.float64)
        tris = np.ascontiguousarray(tris, np.int32)
        f = self.fn("bvh_create")
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        return int(f(_p(verts), len(verts), _p(tris), len(tris)))

    def bvh_refit(self, mesh_id, verts, bottomup=False):
        """beginReplaceModel / replaceSubModel / endReplaceModel(refit=True, bottomup) on the stored model"""
        verts = np.ascontiguousarray(verts, np.float64)
        f = self.fn("bvh_refit")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
        rc = int(f(mesh_id, _p(verts), len(verts), 1 if bottomup else 0))
        assert rc == 0, rc

    def bvh_export(self, mesh_id, dtype):
        st = _st(dtype)
        fn_nodes = self.fn("bvh_num_nodes")
        fn_nodes.argtypes = [C.c_int, C.c_int]
        n_nodes = int(fn_nodes(mesh_id, st))
        fn_tris = self.fn("bvh_num_tris")
        fn_tris.argtypes = [C.c_int]
        n_tris = int(fn_tris(mesh_id))
        obb = np.zeros((n_nodes, 15), dtype)
        child = np.zeros(n_nodes, np.int32)
        tri = np.zeros((n_tris, 9), dtype)
        f = self.fn("bvh_export")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        f(mesh_id, st, _p(obb), _p(child), _p(tri))
        return obb, child, tri

    def bvh_collide_batch(self, id1, id2, poses1, poses2, threads=1, want_pair=True, **req):
        n = len(poses1)
        counts = np.zeros(n, np.uint32)
        pair = np.zeros((n, 2), np.int32) if want_pair else None
        r = _request(**req)
        f = self.fn("bvh_collide_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_int]
        f(_st(poses1.dtype), id1, id2, _p(poses1), _p(poses2), n, C.cast(C.pointer(r), C.c_void_p), _p(counts),
          _p(pair), threads)
        return counts, pair

    def bvh_collide_contacts_batch(self, id1, id2, poses1, poses2, max_keep, threads=1, **req):
        n = len(poses1)
        counts = np.zeros(n, np.uint32)
        ids = np.zeros((n, max_keep, 2), np.int32)
        contacts = np.zeros((n, max_keep, 7), poses1.dtype)
        r = _request(**req)
        f = self.fn("bvh_collide_contacts_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p,
                      C.c_void_p, C.c_void_p, C.c_int]
        f(_st(poses1.dtype), id1, id2, _p(poses1), _p(poses2), n, C.cast(C.pointer(r), C.c_void_p), max_keep, _p(counts),
          _p(ids), _p(contacts), threads)
        return counts, ids, contacts

    def bvh_visit_counts(self, id1, id2, poses1, poses2, threads=1):
        n = len(poses1)
        n_bv = np.zeros(n, np.uint64)
        n_leaf = np.zeros(n, np.uint64)
        n_hit = np.zeros(n, np.uint32)
        f = self.fn("bvh_visit_counts")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_int]
        f(_st(poses1.dtype), id1, id2, _p(poses1), _p(poses2), n, _p(n_bv), _p(n_leaf), _p(n_hit), threads)
        return n_bv, n_leaf, n_hit

    # ---- mesh vs shape ----
    def mesh_shape_collide_batch(self, mesh_id, shapes, shape_ids, poses_mesh, poses_shape, threads=1, want_tri=True,
                                 **req):
        n = len(poses_mesh)
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        counts = np.zeros(n, np.uint32)
        tri = np.zeros(n, np.int32) if want_tri else None
        r = _request(**req)
        arr = _shape_array(shapes)
        f = self.fn("mesh_shape_collide_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(poses_mesh.dtype), mesh_id, C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses_mesh),
          _p(poses_shape), n, C.cast(C.pointer(r), C.c_void_p), _p(counts), _p(tri), threads)
        return counts, tri

    # ---- heightmaps ----
    def heightmap_create(self, points, resolution, half_shape):
        pts = np.ascontiguousarray(points, np.float64)
        f = self.fn("heightmap_create")
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int]
        return int(f(_p(pts), len(pts), resolution, half_shape))

    def heightmap_export(self, hm_id, dtype, half_shape):
        h = np.zeros((2 * half_shape, 2 * half_shape), np.uint16)
        f = self.fn("heightmap_export")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p]
        upper = int(f(hm_id, _st(dtype), _p(h)))
        return h, upper

    def heightmap_shape_collide_batch(self, hm_id, shapes, shape_ids, poses_hm, poses_shape, threads=1, want_pixel=True,
                                      **req):
        n = len(poses_hm)
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        counts = np.zeros(n, np.uint32)
        pix = np.zeros(n, np.int32) if want_pixel else None
        r = _request(**req)
        arr = _shape_array(shapes)
        f = self.fn("heightmap_shape_collide_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(poses_hm.dtype), hm_id, C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses_hm), _p(poses_shape), n,
          C.cast(C.pointer(r), C.c_void_p), _p(counts), _p(pix), threads)
        return counts, pix

    # ---- octrees ----
    def octree_create(self, points, resolution, half_shape):
        pts = np.ascontiguousarray(points, np.float64)
        f = self.fn("octree_create")
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int]
        return int(f(_p(pts), len(pts), resolution, half_shape))

    def octree_prune(self, oct_id, axis, center, extent):
        """Octree2CollisionGeometry::pruneBy(OBB(axis, center, extent), rebuild=False): id of the pruned geometry."""
        obb = np.concatenate([np.asarray(axis, np.float64).reshape(9), np.asarray(center, np.float64),
                              np.asarray(extent, np.float64)])
        f = self.fn("octree_prune")
        f.argtypes = [C.c_int, C.c_void_p]
        return int(f(oct_id, _p(np.ascontiguousarray(obb))))

    def octree_prune_rebuild(self, oct_id, axis, center, extent):
        """Octree2CollisionGeometry::pruneBy(OBB, rebuild=True): id of the consolidated, renumbered geometry."""
        obb = np.concatenate([np.asarray(axis, np.float64).reshape(9), np.asarray(center, np.float64),
                              np.asarray(extent, np.float64)])
        f = self.fn("octree_prune_rebuild")
        f.argtypes = [C.c_int, C.c_void_p]
        return int(f(oct_id, _p(np.ascontiguousarray(obb))))

    def octree_export_pruned(self, oct_id, dtype, n_inner):
        """prune_internal_nodes as bytes, or None when the geometry has no prune info"""
        out = np.zeros(n_inner, np.uint8)
        f = self.fn("octree_export_pruned")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p]
        return out if f(oct_id, _st(dtype), _p(out)) else None

    def octree_export(self, oct_id, dtype):
        """(inner_children [n,8] u32, inner_full [n] u8, leaf_bits [m] u8, root_aabb [6] f64, n_layers)"""
        sizes = np.zeros(3, np.uint32)
        f = self.fn("octree_sizes")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p]
        f(oct_id, _st(dtype), _p(sizes))
        ch = np.zeros((int(sizes[0]), 8), np.uint32)
        full = np.zeros(int(sizes[0]), np.uint8)
        leaf = np.zeros(int(sizes[1]), np.uint8)
        root = np.zeros(6, np.float64)
        g = self.fn("octree_export")
        g.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        g(oct_id, _st(dtype), _p(ch), _p(full), _p(leaf), _p(root))
        return ch, full, leaf, root, int(sizes[2])

    def octree_shape_collide_batch(self, oct_id, shapes, shape_ids, poses_oct, poses_shape, threads=1, **req):
        n = len(poses_oct)
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        counts = np.zeros(n, np.uint32)
        node = np.zeros(n, np.int64)
        r = _request(**req)
        arr = _shape_array(shapes)
        f = self.fn("octree_shape_collide_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(poses_oct.dtype), oct_id, C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses_oct), _p(poses_shape), n,
          C.cast(C.pointer(r), C.c_void_p), _p(counts), _p(node), threads)
        return counts, node

    # ---- every contact of a scene-vs-shape query (any request mode) ----
    def scene_shape_contacts_batch(self, kind, scene_id, shapes, shape_ids, poses_scene, poses_shape, max_keep, threads=1,
                                   **req):
        n = len(poses_scene)
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        counts = np.zeros(n, np.uint32)
        b1 = np.zeros((n, max_keep), np.int64)
        contacts = np.zeros((n, max_keep, 7), poses_scene.dtype)
        r = _request(**req)
        arr = _shape_array(shapes)
        f = self.fn("scene_shape_contacts_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                      C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(poses_scene.dtype), kind, scene_id, C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses_scene),
          _p(poses_shape), n, C.cast(C.pointer(r), C.c_void_p), max_keep, _p(counts), _p(b1), _p(contacts), threads)
        return counts, b1, contacts

    # ---- scene vs scene (heightmap / octree / mesh pairs) ----
    def scene_pair_collide_batch(self, kind1, id1, kind2, id2, poses1, poses2, max_keep, threads=1, want_contacts=False,
                                 **req):
        n = len(poses1)
        counts = np.zeros(n, np.uint32)
        b1 = np.zeros((n, max_keep), np.int64)
        b2 = np.zeros((n, max_keep), np.int64)
        contacts = np.zeros((n, max_keep, 7), poses1.dtype) if want_contacts else None
        r = _request(**req)
        f = self.fn("scene_pair_collide_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        f(_st(poses1.dtype), kind1, id1, kind2, id2, _p(poses1), _p(poses2), n, C.cast(C.pointer(r), C.c_void_p), max_keep,
          _p(counts), _p(b1), _p(b2), _p(contacts) if want_contacts else None, threads)
        if want_contacts:
            return counts, b1, b2, contacts
        return counts, b1, b2

    # ---- broadphase ----
    def compute_aabb_batch(self, shapes, shape_ids, poses):
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        out = np.zeros((len(ids), 6), poses.dtype)
        arr = _shape_array(shapes)
        f = self.fn("compute_aabb_batch")
        f.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        f(_st(poses.dtype), C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses), len(ids), _p(out))
        return out

    def broadphase_create(self, aabbs, user_ids):
        ids = np.ascontiguousarray(user_ids, np.uint64)
        f = self.fn("broadphase_create")
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        return int(f(_st(aabbs.dtype), _p(aabbs), _p(ids), len(ids))), _st(aabbs.dtype)

    def _pairs(self, name, argtypes, *args):
        f = self.fn(name)
        f.argtypes = argtypes + [C.c_void_p, C.c_size_t]
        f.restype = C.c_size_t
        n = int(f(*args, None, 0))
        out = np.zeros((n, 2), np.uint64)
        if n:
            f(*args, _p(out), n)
        return out

    def broadphase_self_pairs(self, tree):
        tid, st = tree
        return self._pairs("broadphase_self_pairs", [C.c_int, C.c_int], st, tid)

    def broadphase_tree_pairs(self, tree_a, tree_b):
        return self._pairs("broadphase_tree_pairs", [C.c_int, C.c_int, C.c_int], tree_a[1], tree_a[0], tree_b[0])

    def broadphase_query_pairs(self, tree, aabbs, object_ids):
        ids = np.ascontiguousarray(object_ids, np.uint64)
        return self._pairs("broadphase_query_pairs", [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t], tree[1],
                           tree[0], _p(aabbs), _p(ids), len(ids))

    def broadphase_update(self, tree, user_ids, aabbs):
        ids = np.ascontiguousarray(user_ids, np.uint64)
        f = self.fn("broadphase_update")
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        return int(f(tree[1], tree[0], _p(ids), _p(aabbs), len(ids)))

    def scene_self_collide(self, shapes, shape_ids, poses):
        """computeAABB + Rebuild + SelfCollision with boolean fcl::collide per candidate: (hits, candidates)."""
        ids = np.ascontiguousarray(shape_ids, np.uint32)
        arr = _shape_array(shapes)
        cand = C.c_size_t()
        f = self.fn("scene_self_collide")
        f.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        f.restype = C.c_size_t
        hits = int(f(_st(poses.dtype), C.cast(arr, C.c_void_p), len(shapes), _p(ids), _p(poses), len(ids), C.byref(cand)))
        return hits, int(cand.value)


class PortOracle(_Oracle):
    prefix = "fclport_"
    path = PORT_PATH
    kind = "port"


def have_ref():
    return os.path.exists(REF_PATH)


def have_port():
    return os.path.exists(PORT_PATH)
