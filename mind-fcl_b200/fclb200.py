"""ctypes binding of libfclb200.so (include/fclb200.h) used by tests/ and bench.py.

This is host plumbing only: it loads the in-tree shared library, declares the
C ABI, and moves numpy / torch buffers across it.  There is no Python compute
path and no CPU fallback: if the library is missing or no GPU is visible the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FCLB_LIB") or os.path.join(_HERE, "libfclb200.so")  # FCLB_LIB: an experimental build (profiles/scripts)

F32, F64 = 0, 1
BOX, SPHERE, ELLIPSOID, CAPSULE, CONE, CYLINDER, CONVEX = range(7)
PEN_DISABLED, PEN_DEFAULT_GJK_EPA, PEN_DIRECTED, PEN_INCREMENTAL_MIN = range(4)

# every symbol include/fclb200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "fclb_init", "fclb_device_count", "fclb_last_error", "fclb_version",
    "fclb_host_alloc", "fclb_host_free", "fclb_dev_alloc", "fclb_dev_free",
    "fclb_memcpy_h2d", "fclb_memcpy_d2h", "fclb_synchronize",
    "fclb_convex_upload", "fclb_shapes_upload", "fclb_release",
    "fclb_distance_batch_host", "fclb_distance_batch_dev",
    "fclb_signed_distance_batch_host", "fclb_signed_distance_batch_dev",
    "fclb_collide_batch_host", "fclb_collide_batch_dev",
    "fclb_gjk_epa_batch_host", "fclb_gjk_epa_batch_dev",
    "fclb_bvh_upload", "fclb_bvh_release", "fclb_bvh_collide_batch_host", "fclb_bvh_collide_batch_dev",
    "fclb_bvh_collide_contacts_batch_host", "fclb_bvh_collide_contacts_batch_dev",
    "fclb_bvh_last_visit_counts", "fclb_bvh_build", "fclb_bvh_build_host", "fclb_bvh_info", "fclb_bvh_export",
    "fclb_bvh_shape_collide_batch_host", "fclb_bvh_shape_collide_batch_dev", "fclb_scene_last_visit_counts",
    "fclb_heightmap_upload", "fclb_heightmap_release", "fclb_heightmap_build_host",
    "fclb_heightmap_build_dev", "fclb_heightmap_build_points_host", "fclb_heightmap_info", "fclb_heightmap_export",
    "fclb_heightmap_shape_collide_batch_host", "fclb_heightmap_shape_collide_batch_dev",
    "fclb_octree_upload", "fclb_octree_release", "fclb_octree_build_host", "fclb_octree_build", "fclb_octree_prune_host", "fclb_octree_consolidate_host",
    "fclb_octree_shape_collide_batch_host",
    "fclb_octree_shape_collide_batch_dev",
    "fclb_scene_shape_contacts_batch_host", "fclb_scene_shape_contacts_batch_dev",
    "fclb_scene_pair_collide_batch_host", "fclb_scene_pair_collide_batch_dev",
    "fclb_scene_pair_contacts_batch_host", "fclb_scene_pair_contacts_batch_dev",
    "fclb_broadphase_build_host", "fclb_broadphase_build_dev", "fclb_broadphase_release",
    "fclb_broadphase_self_pairs_host", "fclb_broadphase_self_pairs_dev", "fclb_broadphase_tree_pairs_host",
    "fclb_broadphase_query_pairs_host", "fclb_broadphase_update_host", "fclb_broadphase_last_visits",
    "fclb_compute_aabb_batch_host", "fclb_compute_aabb_batch_dev", "fclb_gather_pairs_dev",
    "fclb_scene_self_collide_host", "fclb_scene_self_collide_dev",
    "fclb_bvh_refit_host", "fclb_bvh_refit_dev", "fclb_bvh_refit_bottomup_host", "fclb_bvh_refit_bottomup_dev", "fclb_bvh_build_device", "fclb_octree_build_dev", "fclb_octree_build_points_host", "fclb_octree_info", "fclb_octree_export", "fclb_translational_ccd_batch_host", "fclb_translational_ccd_batch_dev", "fclb_translational_ccd_mesh_batch_host", "fclb_translational_ccd_mesh_batch_dev", "fclb_translational_ccd_mesh_pair_batch_host", "fclb_translational_ccd_mesh_pair_batch_dev", "fclb_translational_ccd_scene_batch_host", "fclb_translational_ccd_scene_batch_dev", "fclb_translational_ccd_scene_mesh_batch_host", "fclb_translational_ccd_scene_mesh_batch_dev", "fclb_translational_ccd_scene_pair_batch_host", "fclb_translational_ccd_scene_pair_batch_dev", "fclb_init_devices", "fclb_num_devices", "fclb_set_device", "fclb_distance_batch_qt_host", "fclb_expand_poses_dev",
    "fclb_measure_fp_peak", "fclb_measure_l2_bandwidth", "fclb_launch_count", "fclb_last_kernel_ms", "fclb_last_call_ms", "fclb_last_launches", "fclb_stream",
]


class Shape(C.Structure):
    _fields_ = [("type", C.c_uint32), ("geom", C.c_uint32), ("p", C.c_double * 3)]


class Request(C.Structure):
    _fields_ = [
        ("max_contacts", C.c_uint32), ("penetration_mode", C.c_uint32), ("dir", C.c_double * 3),
        ("binary_tol", C.c_double), ("distance_tol", C.c_double),
        ("gjk_max_iter", C.c_uint32), ("epa_max_faces", C.c_uint32), ("epa_max_iter", C.c_uint32),
        ("flags", C.c_uint32),
    ]


def make_request(max_contacts=1, penetration_mode=PEN_DISABLED, direction=(0.0, 0.0, 0.0), binary_tol=0.0,
                 distance_tol=0.0, gjk_max_iter=0, epa_max_faces=0, epa_max_iter=0) -> Request:
    r = Request()
    r.max_contacts = max_contacts
    r.penetration_mode = penetration_mode
    r.dir[:] = direction
    r.binary_tol = binary_tol
    r.distance_tol = distance_tol
    r.gjk_max_iter = gjk_max_iter
    r.epa_max_faces = epa_max_faces
    r.epa_max_iter = epa_max_iter
    r.flags = 0
    return r


def shape_array(shapes) -> C.Array:
    """shapes: iterable of (type, geom, (p0,p1,p2))."""
    arr = (Shape * len(shapes))()
    for i, (t, g, p) in enumerate(shapes):
        arr[i].type = t
        arr[i].geom = g
        p = list(p) + [0.0] * (3 - len(p))
        arr[i].p[:] = p
    return arr


class FclbError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FclbError(f"{LIB_PATH} is missing: run `make -C mind-fcl_b200/csrc` (or __graft_entry__.build())")
    lib = C.CDLL(LIB_PATH)
    lib.fclb_last_error.restype = C.c_char_p
    lib.fclb_version.restype = C.c_char_p
    lib.fclb_launch_count.restype = C.c_uint64
    lib.fclb_last_kernel_ms.restype = C.c_double
    lib.fclb_last_call_ms.restype = C.c_double
    lib.fclb_stream.restype = C.c_void_p
    lib.fclb_last_launches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
    lib.fclb_init.argtypes = [C.c_int]
    lib.fclb_host_alloc.argtypes = [C.POINTER(vp), sz]
    lib.fclb_host_free.argtypes = [vp]
    lib.fclb_dev_alloc.argtypes = [C.POINTER(vp), sz]
    lib.fclb_dev_free.argtypes = [vp]
    lib.fclb_memcpy_h2d.argtypes = [vp, vp, sz]
    lib.fclb_memcpy_d2h.argtypes = [vp, vp, sz]
    lib.fclb_convex_upload.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.POINTER(u32)]
    lib.fclb_shapes_upload.argtypes = [vp, u32, C.POINTER(C.c_uint64)]
    lib.fclb_release.argtypes = [C.c_uint64]
    dist_args = [C.c_uint64, vp, vp, vp, sz, C.c_int, C.c_double, u32, vp, vp, vp, vp]
    lib.fclb_distance_batch_host.argtypes = dist_args
    if hasattr(lib, "fclb_distance_batch_qt_host"):
        lib.fclb_distance_batch_qt_host.argtypes = dist_args
        lib.fclb_expand_poses_dev.argtypes = [vp, sz, C.c_int, vp]
    lib.fclb_distance_batch_dev.argtypes = dist_args
    col_args = [C.c_uint64, vp, vp, vp, sz, C.c_int, vp, u32, vp, vp]
    ge_args = [C.c_uint64, vp, vp, vp, sz, C.c_int, vp, vp, vp, vp]
    for name, args in (("fclb_collide_batch_host", col_args), ("fclb_collide_batch_dev", col_args),
                       ("fclb_gjk_epa_batch_host", ge_args), ("fclb_gjk_epa_batch_dev", ge_args)):
        if hasattr(lib, name):
            getattr(lib, name).argtypes = args
    if hasattr(lib, "fclb_bvh_upload"):
        lib.fclb_bvh_upload.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_bvh_release.argtypes = [C.c_uint64]
        bvh_args = [C.c_uint64, C.c_uint64, vp, vp, sz, C.c_int, vp, vp, vp]
        lib.fclb_bvh_collide_batch_host.argtypes = bvh_args
        lib.fclb_bvh_collide_batch_dev.argtypes = bvh_args
        lib.fclb_bvh_last_visit_counts.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    if hasattr(lib, "fclb_bvh_build"):
        lib.fclb_bvh_build.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_bvh_build_host.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.POINTER(C.c_int)]
        lib.fclb_bvh_info.argtypes = [C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.fclb_bvh_export.argtypes = [C.c_uint64, vp, vp, vp]
    if hasattr(lib, "fclb_bvh_shape_collide_batch_host"):
        bs_args = [C.c_uint64, C.c_uint64, vp, vp, vp, sz, C.c_int, vp, vp, vp]
        lib.fclb_bvh_shape_collide_batch_host.argtypes = bs_args
        lib.fclb_bvh_shape_collide_batch_dev.argtypes = bs_args
        lib.fclb_scene_last_visit_counts.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    if hasattr(lib, "fclb_heightmap_upload"):
        lib.fclb_heightmap_upload.argtypes = [vp, u32, u32, C.c_double, C.c_double, u32, C.POINTER(C.c_uint64)]
        lib.fclb_heightmap_release.argtypes = [C.c_uint64]
        lib.fclb_heightmap_build_host.argtypes = [vp, sz, C.c_double, C.c_double, u32, u32, C.c_int, vp]
    if hasattr(lib, "fclb_heightmap_build_dev"):
        hb_args = [vp, sz, C.c_double, C.c_double, u32, u32, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_heightmap_build_dev.argtypes = hb_args
        lib.fclb_heightmap_build_points_host.argtypes = hb_args
        lib.fclb_heightmap_info.argtypes = [C.c_uint64, vp, vp, vp, vp]
        lib.fclb_heightmap_export.argtypes = [C.c_uint64, u32, vp]
        hs_args = [C.c_uint64, C.c_uint64, vp, vp, vp, sz, C.c_int, vp, vp, vp]
        lib.fclb_heightmap_shape_collide_batch_host.argtypes = hs_args
        lib.fclb_heightmap_shape_collide_batch_dev.argtypes = hs_args
    if hasattr(lib, "fclb_broadphase_build_host"):
        szp = C.POINTER(C.c_size_t)
        lib.fclb_broadphase_build_host.argtypes = [vp, vp, sz, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_broadphase_build_dev.argtypes = [vp, vp, sz, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_broadphase_release.argtypes = [C.c_uint64]
        lib.fclb_broadphase_self_pairs_host.argtypes = [C.c_uint64, vp, sz, szp]
        lib.fclb_broadphase_self_pairs_dev.argtypes = [C.c_uint64, vp, sz, szp]
        lib.fclb_broadphase_tree_pairs_host.argtypes = [C.c_uint64, C.c_uint64, vp, sz, szp]
        lib.fclb_broadphase_query_pairs_host.argtypes = [C.c_uint64, vp, vp, sz, vp, sz, szp]
        lib.fclb_broadphase_update_host.argtypes = [C.c_uint64, vp, vp, sz]
        lib.fclb_broadphase_last_visits.restype = C.c_uint64
        lib.fclb_compute_aabb_batch_host.argtypes = [C.c_uint64, vp, vp, sz, C.c_int, vp]
        lib.fclb_compute_aabb_batch_dev.argtypes = [C.c_uint64, vp, vp, sz, C.c_int, vp]
        lib.fclb_gather_pairs_dev.argtypes = [vp, sz, vp, vp, C.c_int, vp, vp, vp]
        sc_args = [C.c_uint64, vp, vp, sz, C.c_int, vp, szp, szp, vp, vp, sz]
        lib.fclb_scene_self_collide_host.argtypes = sc_args
        lib.fclb_scene_self_collide_dev.argtypes = sc_args
    if hasattr(lib, "fclb_octree_upload"):
        lib.fclb_octree_upload.argtypes = [vp, vp, u32, vp, u32, vp, vp, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_octree_release.argtypes = [C.c_uint64]
    if hasattr(lib, "fclb_octree_build_host"):
        lib.fclb_octree_build_host.argtypes = [vp, sz, C.c_double, u32, C.c_int, vp, vp, u32, C.POINTER(u32), vp, u32,
                                               C.POINTER(u32), vp, C.POINTER(C.c_int)]
        lib.fclb_octree_build.argtypes = [vp, sz, C.c_double, u32, C.c_int, C.POINTER(C.c_uint64)]
        lib.fclb_octree_prune_host.argtypes = [vp, u32, u32, vp, C.c_int, vp, C.c_int, vp, vp, vp]
        lib.fclb_octree_consolidate_host.argtypes = [vp, u32, vp, vp, u32, C.c_int, vp, vp, C.POINTER(u32), vp, C.POINTER(u32)]
        os_args = [C.c_uint64, C.c_uint64, vp, vp, vp, sz, C.c_int, vp, vp, vp]
        lib.fclb_octree_shape_collide_batch_host.argtypes = os_args
        lib.fclb_octree_shape_collide_batch_dev.argtypes = os_args
    if hasattr(lib, "fclb_scene_shape_contacts_batch_host"):
        sc2 = [C.c_int, C.c_uint64, C.c_uint64, vp, vp, vp, sz, C.c_int, vp, u32, vp, vp, vp]
        lib.fclb_scene_shape_contacts_batch_host.argtypes = sc2
        lib.fclb_scene_shape_contacts_batch_dev.argtypes = sc2
    if hasattr(lib, "fclb_scene_pair_collide_batch_host"):
        sp_args = [C.c_int, C.c_uint64, C.c_int, C.c_uint64, vp, vp, sz, C.c_int, vp, u32, vp, vp, vp]
        lib.fclb_scene_pair_collide_batch_host.argtypes = sp_args
        lib.fclb_scene_pair_collide_batch_dev.argtypes = sp_args
        lib.fclb_scene_pair_contacts_batch_host.argtypes = sp_args + [vp]
        lib.fclb_scene_pair_contacts_batch_dev.argtypes = sp_args + [vp]
    if hasattr(lib, "fclb_bvh_collide_contacts_batch_host"):
        bc_args = [C.c_uint64, C.c_uint64, vp, vp, sz, C.c_int, vp, u32, vp, vp, vp]
        lib.fclb_bvh_collide_contacts_batch_host.argtypes = bc_args
        lib.fclb_bvh_collide_contacts_batch_dev.argtypes = bc_args
    if hasattr(lib, "fclb_signed_distance_batch_host"):
        sd_args = [C.c_uint64, vp, vp, vp, sz, C.c_int, vp, vp, vp, vp]
        lib.fclb_signed_distance_batch_host.argtypes = sd_args
        lib.fclb_signed_distance_batch_dev.argtypes = sd_args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise FclbError(f"fclb error {rc}: {load().fclb_last_error().decode()}")


def _ptr(a):
    """numpy array / torch tensor / int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    return C.cast(a, C.c_void_p)


def np_dtype(scalar_type):
    return np.float32 if scalar_type == F32 else np.float64


def init(device: int = 0) -> None:
    check(load().fclb_init(device))


def init_devices(n: int = 0) -> int:
    """one process, n GPUs (0: all visible): *_host batch calls shard by query index; returns the engine count"""
    check(load().fclb_init_devices(n))
    return int(load().fclb_num_devices())


def convex_upload(verts: np.ndarray, faces: np.ndarray, num_faces: int) -> int:
    verts = np.ascontiguousarray(verts, dtype=np.float64)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    slot = C.c_uint32()
    check(load().fclb_convex_upload(_ptr(verts), verts.shape[0], _ptr(faces), faces.size, num_faces, C.byref(slot)))
    return slot.value


def shapes_upload(shapes) -> int:
    arr = shape_array(shapes)
    h = C.c_uint64()
    check(load().fclb_shapes_upload(C.cast(arr, C.c_void_p), len(shapes), C.byref(h)))
    return h.value


def release(h: int) -> None:
    check(load().fclb_release(h))


@dataclass
class DistanceResult:
    dist: np.ndarray
    p1: np.ndarray
    p2: np.ndarray
    ok: np.ndarray


def distance_batch_host(table, pairs, poses1, poses2, scalar_type, gjk_tol=0.0, gjk_max_iter=0, out=None):
    """Host-buffer call (numpy or pinned torch tensors)."""
    n = len(pairs)
    dt = np_dtype(scalar_type)
    if out is None:
        out = DistanceResult(np.empty(n, dt), np.empty((n, 3), dt), np.empty((n, 3), dt), np.empty(n, np.uint8))
    check(load().fclb_distance_batch_host(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type, gjk_tol,
                                          gjk_max_iter, _ptr(out.dist), _ptr(out.p1), _ptr(out.p2), _ptr(out.ok)))
    return out


def distance_batch_qt_host(table, pairs, qt1, qt2, scalar_type, gjk_tol=0.0, gjk_max_iter=0):
    """distance_batch_host with FCLB_POSE_QT7 poses (n x 7: quaternion x, y, z, w, translation)"""
    n = len(pairs)
    dt = np_dtype(scalar_type)
    out = DistanceResult(np.zeros(n, dt), np.zeros((n, 3), dt), np.zeros((n, 3), dt), np.zeros(n, np.uint8))
    check(load().fclb_distance_batch_qt_host(table, _ptr(pairs), _ptr(qt1), _ptr(qt2), n, scalar_type, gjk_tol, gjk_max_iter,
                                             _ptr(out.dist), _ptr(out.p1), _ptr(out.p2), _ptr(out.ok)))
    return out


def distance_batch_dev(table, pairs, poses1, poses2, n, scalar_type, dist, p1, p2, ok, gjk_tol=0.0, gjk_max_iter=0):
    """Device-buffer call: every array argument is a CUDA tensor or raw device pointer."""
    check(load().fclb_distance_batch_dev(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type, gjk_tol,
                                         gjk_max_iter, _ptr(dist), _ptr(p1), _ptr(p2), _ptr(ok)))


def collide_batch_host(table, pairs, poses1, poses2, scalar_type, request: Request, max_keep=1, want_contacts=True):
    n = len(pairs)
    dt = np_dtype(scalar_type)
    contacts = np.zeros((n, max_keep, 9), dt) if want_contacts else None
    counts = np.zeros(n, np.uint32)
    check(load().fclb_collide_batch_host(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type,
                                         C.cast(C.pointer(request), C.c_void_p), max_keep, _ptr(contacts),
                                         _ptr(counts)))
    return counts, contacts


def collide_batch_dev(table, pairs, poses1, poses2, n, scalar_type, request: Request, max_keep, contacts, counts):
    check(load().fclb_collide_batch_dev(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type,
                                        C.cast(C.pointer(request), C.c_void_p), max_keep, _ptr(contacts),
                                        _ptr(counts)))


def gjk_epa_batch_host(table, pairs, poses1, poses2, scalar_type, request: Request):
    n = len(pairs)
    dt = np_dtype(scalar_type)
    gjk = np.zeros(n, np.int32)
    epa = np.zeros(n, np.int32)
    geom = np.zeros((n, 7), dt)
    check(load().fclb_gjk_epa_batch_host(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type,
                                         C.cast(C.pointer(request), C.c_void_p), _ptr(gjk), _ptr(epa), _ptr(geom)))
    return gjk, epa, geom


def gjk_epa_batch_dev(table, pairs, poses1, poses2, n, scalar_type, request: Request, gjk, epa, geom):
    check(load().fclb_gjk_epa_batch_dev(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type,
                                        C.cast(C.pointer(request), C.c_void_p), _ptr(gjk), _ptr(epa), _ptr(geom)))


def launch_count() -> int:
    return int(load().fclb_launch_count())


def last_kernel_ms() -> float:
    return float(load().fclb_last_kernel_ms())


def last_call_ms() -> float:
    return float(load().fclb_last_call_ms())


def last_launches():
    """[(type1, type2, n_queries, ms)] of the most recent batch call."""
    cap = 128
    kinds = (C.c_int * cap)()
    counts = (C.c_uint64 * cap)()
    ms = (C.c_double * cap)()
    n = load().fclb_last_launches(kinds, counts, ms, cap)
    return [(kinds[i] // 8, kinds[i] % 8, int(counts[i]), float(ms[i])) for i in range(min(n, cap))]


def stream_ptr() -> int:
    return int(load().fclb_stream())


def bvh_upload(obb: np.ndarray, first_child: np.ndarray, tri_verts: np.ndarray, scalar_type) -> int:
    dt = np_dtype(scalar_type)
    obb = np.ascontiguousarray(obb, dt)
    fc = np.ascontiguousarray(first_child, np.int32)
    tv = np.ascontiguousarray(tri_verts, dt)
    h = C.c_uint64()
    check(load().fclb_bvh_upload(_ptr(obb), _ptr(fc), len(fc), _ptr(tv), tv.size // 9, scalar_type, C.byref(h)))
    return h.value


def bvh_build(verts: np.ndarray, tris: np.ndarray, scalar_type) -> int:
    """BVHModel<OBBRSS<S>> beginModel/addSubModel/endModel on the host + upload."""
    v = np.ascontiguousarray(verts, np.float64)
    t = np.ascontiguousarray(tris, np.int32)
    h = C.c_uint64()
    check(load().fclb_bvh_build(_ptr(v), len(v), _ptr(t), len(t), scalar_type, C.byref(h)))
    return h.value


def bvh_build_device(verts: np.ndarray, tris: np.ndarray, scalar_type) -> int:
    """the same tree built on the device, level by level (fclb_bvh_build_device); last_kernel_ms() = the build launches"""
    v = np.ascontiguousarray(verts, np.float64)
    t = np.ascontiguousarray(tris, np.int32)
    h = C.c_uint64()
    fn = load().fclb_bvh_build_device
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    check(fn(_ptr(v), len(v), _ptr(t), len(t), scalar_type, C.byref(h)))
    return h.value


def bvh_build_host(verts: np.ndarray, tris: np.ndarray, scalar_type):
    """The host builder alone (no GPU): returns (obb [n,15], first_child [n], tri_verts [t,9])."""
    v = np.ascontiguousarray(verts, np.float64)
    t = np.ascontiguousarray(tris, np.int32)
    dt = np_dtype(scalar_type)
    cap = 2 * len(t) - 1
    obb = np.zeros((cap, 15), dt)
    fc = np.zeros(cap, np.int32)
    tv = np.zeros((len(t), 9), dt)
    n = C.c_int()
    check(load().fclb_bvh_build_host(_ptr(v), len(v), _ptr(t), len(t), scalar_type, _ptr(obb), _ptr(fc), _ptr(tv),
                                     C.byref(n)))
    return obb[:n.value], fc[:n.value], tv


def bvh_refit_host(h: int, tri_verts: np.ndarray, bottomup: bool = False) -> None:
    """refit on the device from the new triangle corners (n_tris x 9, the tree's scalar type); bottomup = the reference's
    default endReplaceModel(): leaf boxes from their triangles, inner boxes merged from their children"""
    t = np.ascontiguousarray(tri_verts)
    fn = load().fclb_bvh_refit_bottomup_host if bottomup else load().fclb_bvh_refit_host
    fn.argtypes = [C.c_uint64, C.c_void_p, C.c_int]
    check(fn(h, _ptr(t), len(t)))


def bvh_export(h: int):
    n, t, st = C.c_int(), C.c_int(), C.c_int()
    check(load().fclb_bvh_info(h, C.byref(n), C.byref(t), C.byref(st)))
    dt = np_dtype(st.value)
    obb = np.zeros((n.value, 15), dt)
    fc = np.zeros(n.value, np.int32)
    tv = np.zeros((t.value, 9), dt)
    check(load().fclb_bvh_export(h, _ptr(obb), _ptr(fc), _ptr(tv)))
    return obb, fc, tv


def bvh_release(h: int) -> None:
    check(load().fclb_bvh_release(h))


def bvh_collide_batch_host(bvh1, bvh2, poses1, poses2, scalar_type, request: Request, want_pair=False):
    n = len(poses1)
    counts = np.zeros(n, np.uint32)
    pair = np.zeros((n, 2), np.int32) if want_pair else None
    check(load().fclb_bvh_collide_batch_host(bvh1, bvh2, _ptr(poses1), _ptr(poses2), n, scalar_type,
                                             C.cast(C.pointer(request), C.c_void_p), _ptr(counts), _ptr(pair)))
    return counts, pair


def bvh_collide_batch_dev(bvh1, bvh2, poses1, poses2, n, scalar_type, request: Request, counts, pair=None):
    check(load().fclb_bvh_collide_batch_dev(bvh1, bvh2, _ptr(poses1), _ptr(poses2), n, scalar_type,
                                            C.cast(C.pointer(request), C.c_void_p), _ptr(counts), _ptr(pair)))


def bvh_last_visit_counts():
    a, b = C.c_uint64(), C.c_uint64()
    check(load().fclb_bvh_last_visit_counts(C.byref(a), C.byref(b)))
    return a.value, b.value


def bvh_shape_collide_batch_host(bvh, table, shape_ids, poses_mesh, poses_shape, scalar_type, request: Request,
                                 want_tri=False):
    """fcl::collide(BVHModel<OBBRSS>, tf_mesh, Shape, tf_shape) per query; returns (counts, first_tri)."""
    n = len(poses_mesh)
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    counts = np.zeros(n, np.uint32)
    tri = np.zeros(n, np.int32) if want_tri else None
    check(load().fclb_bvh_shape_collide_batch_host(bvh, table, _ptr(ids), _ptr(poses_mesh), _ptr(poses_shape), n,
                                                   scalar_type, C.cast(C.pointer(request), C.c_void_p), _ptr(counts),
                                                   _ptr(tri)))
    return counts, tri


def bvh_shape_collide_batch_dev(bvh, table, shape_ids, poses_mesh, poses_shape, n, scalar_type, request: Request, counts,
                                tri=None):
    check(load().fclb_bvh_shape_collide_batch_dev(bvh, table, _ptr(shape_ids), _ptr(poses_mesh), _ptr(poses_shape), n,
                                                  scalar_type, C.cast(C.pointer(request), C.c_void_p), _ptr(counts),
                                                  _ptr(tri)))


def scene_last_visit_counts():
    a, b = C.c_uint64(), C.c_uint64()
    check(load().fclb_scene_last_visit_counts(C.byref(a), C.byref(b)))
    return a.value, b.value


def heightmap_build_host(points, resolution, half_shape, scalar_type, heights=None):
    """FlatHeightMap<S>::updateHeightsByPointGenerationFunctor on the host: bottom-layer heights in mm."""
    pts = np.ascontiguousarray(points, np.float64)
    if heights is None:
        heights = np.zeros((2 * half_shape, 2 * half_shape), np.uint16)
    check(load().fclb_heightmap_build_host(_ptr(pts), len(pts), resolution, resolution, half_shape, half_shape,
                                           scalar_type, _ptr(heights)))
    return heights


def heightmap_upload(heights, resolution, upper_bound_mm=0) -> int:
    h = np.ascontiguousarray(heights, np.uint16)
    out = C.c_uint64()
    check(load().fclb_heightmap_upload(_ptr(h), h.shape[1], h.shape[0], resolution, resolution, upper_bound_mm,
                                       C.byref(out)))
    return out.value


def heightmap_build_points_host(points, resolution, half_shape, scalar_type) -> int:
    """LayeredHeightMap built on the device from a host point cloud (n x 3, converted to the scalar type)."""
    pts = np.ascontiguousarray(points, np_dtype(scalar_type))
    out = C.c_uint64()
    check(load().fclb_heightmap_build_points_host(_ptr(pts), len(pts), resolution, resolution, half_shape, half_shape,
                                                  scalar_type, C.byref(out)))
    return out.value


def heightmap_build_dev(points_dev, n_points, resolution, half_shape, scalar_type) -> int:
    """The same from a DEVICE point array (torch tensor / raw pointer, n x 3 of the scalar type)."""
    out = C.c_uint64()
    check(load().fclb_heightmap_build_dev(_ptr(points_dev), n_points, resolution, resolution, half_shape, half_shape,
                                          scalar_type, C.byref(out)))
    return out.value


def heightmap_info(h: int):
    v = (C.c_uint32 * 4)()
    check(load().fclb_heightmap_info(h, C.byref(v, 0), C.byref(v, 4), C.byref(v, 8), C.byref(v, 12)))
    return {"n_layers": v[0], "full_x": v[1], "full_y": v[2], "upper_bound_mm": v[3]}


def heightmap_export(h: int, layer: int = 0) -> np.ndarray:
    info = heightmap_info(h)
    out = np.zeros((info["full_y"] >> layer, info["full_x"] >> layer), np.uint16)
    check(load().fclb_heightmap_export(h, layer, _ptr(out)))
    return out


def heightmap_release(h: int) -> None:
    check(load().fclb_heightmap_release(h))


def heightmap_shape_collide_batch_host(hm, table, shape_ids, poses_hm, poses_shape, scalar_type, request: Request,
                                       want_pixel=False):
    n = len(poses_hm)
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    counts = np.zeros(n, np.uint32)
    pix = np.zeros(n, np.int32) if want_pixel else None
    check(load().fclb_heightmap_shape_collide_batch_host(hm, table, _ptr(ids), _ptr(poses_hm), _ptr(poses_shape), n,
                                                         scalar_type, C.cast(C.pointer(request), C.c_void_p),
                                                         _ptr(counts), _ptr(pix)))
    return counts, pix


def heightmap_shape_collide_batch_dev(hm, table, shape_ids, poses_hm, poses_shape, n, scalar_type, request: Request,
                                      counts, pix=None):
    check(load().fclb_heightmap_shape_collide_batch_dev(hm, table, _ptr(shape_ids), _ptr(poses_hm), _ptr(poses_shape), n,
                                                        scalar_type, C.cast(C.pointer(request), C.c_void_p),
                                                        _ptr(counts), _ptr(pix)))


# ---- broadphase ---------------------------------------------------------------------
def compute_aabb_batch_host(table, shape_ids, poses, scalar_type):
    """CollisionObject<S>::computeAABB per object: [n, 6] = min xyz, max xyz."""
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    out = np.zeros((len(ids), 6), np_dtype(scalar_type))
    check(load().fclb_compute_aabb_batch_host(table, _ptr(ids), _ptr(poses), len(ids), scalar_type, _ptr(out)))
    return out


def broadphase_build_host(aabbs, user_ids, scalar_type) -> int:
    b = np.ascontiguousarray(aabbs, np_dtype(scalar_type))
    ids = np.ascontiguousarray(user_ids, np.uint64)
    h = C.c_uint64()
    check(load().fclb_broadphase_build_host(_ptr(b), _ptr(ids), len(ids), scalar_type, C.byref(h)))
    return h.value


def broadphase_release(h: int) -> None:
    check(load().fclb_broadphase_release(h))


def _pairs_call(fn, *args):
    """count, allocate, fetch: returns [m, 2] uint64"""
    n = C.c_size_t()
    check(fn(*args, None, 0, C.byref(n)))
    out = np.zeros((n.value, 2), np.uint64)
    if n.value:
        check(fn(*args, _ptr(out), n.value, C.byref(n)))
    return out


def broadphase_self_pairs_host(tree):
    return _pairs_call(load().fclb_broadphase_self_pairs_host, tree)


def broadphase_tree_pairs_host(tree_a, tree_b):
    return _pairs_call(load().fclb_broadphase_tree_pairs_host, tree_a, tree_b)


def broadphase_query_pairs_host(tree, aabbs, object_ids, scalar_type):
    b = np.ascontiguousarray(aabbs, np_dtype(scalar_type))
    ids = np.ascontiguousarray(object_ids, np.uint64)
    return _pairs_call(load().fclb_broadphase_query_pairs_host, tree, _ptr(b), _ptr(ids), len(ids))


def broadphase_update_host(tree, user_ids, new_aabbs, scalar_type) -> None:
    b = np.ascontiguousarray(new_aabbs, np_dtype(scalar_type))
    ids = np.ascontiguousarray(user_ids, np.uint64)
    check(load().fclb_broadphase_update_host(tree, _ptr(ids), _ptr(b), len(ids)))


def broadphase_last_visits() -> int:
    return int(load().fclb_broadphase_last_visits())


def scene_self_collide(table, shape_ids, poses, n, scalar_type, request: Request, host=True, want_pairs=False):
    """computeAABB + tree build + SelfCollision + boolean collide per candidate for one scene.
    Returns (n_candidates, n_colliding[, id_pairs, counts])."""
    fn = load().fclb_scene_self_collide_host if host else load().fclb_scene_self_collide_dev
    cand, hits = C.c_size_t(), C.c_size_t()
    rq = C.cast(C.pointer(request), C.c_void_p)
    check(fn(table, _ptr(shape_ids), _ptr(poses), n, scalar_type, rq, C.byref(cand), C.byref(hits), None, None, 0))
    if not want_pairs:
        return cand.value, hits.value
    pairs = np.zeros((cand.value, 2), np.uint64)
    counts = np.zeros(cand.value, np.uint32)
    check(fn(table, _ptr(shape_ids), _ptr(poses), n, scalar_type, rq, C.byref(cand), C.byref(hits), _ptr(pairs),
             _ptr(counts), len(counts)))
    return cand.value, hits.value, pairs, counts


# ---- octrees ---------------------------------------------------------------------------
def octree_upload(inner_children, inner_full, leaf_bits, root_aabb, n_layers, pruned=None) -> int:
    ch = np.ascontiguousarray(inner_children, np.uint32)
    full = np.ascontiguousarray(inner_full, np.uint8)
    leaf = np.ascontiguousarray(leaf_bits, np.uint8)
    root = np.ascontiguousarray(root_aabb, np.float64)
    pr = None if pruned is None else np.ascontiguousarray(pruned, np.uint8)
    h = C.c_uint64()
    check(load().fclb_octree_upload(_ptr(ch), _ptr(full), len(full), _ptr(leaf), len(leaf), _ptr(pr), _ptr(root), n_layers,
                                    C.byref(h)))
    return h.value


def octree_build_host(points, resolution, half_shape, scalar_type):
    """The host builder alone (no GPU) = octree2::Octree<S>(resolution, half_shape).rebuildTree(points):
    (inner_children [n,8] u32, inner_full [n] u8, leaf_bits [m] u8, root_aabb [6] f64, n_layers)."""
    pts = np.ascontiguousarray(points, np.float64)
    ni, nl, layers = C.c_uint32(), C.c_uint32(), C.c_int()
    root = np.zeros(6, np.float64)
    fn = load().fclb_octree_build_host
    rc = fn(_ptr(pts), len(pts), resolution, half_shape, scalar_type, None, None, 0, C.byref(ni), None, 0, C.byref(nl),
            _ptr(root), C.byref(layers))
    if rc != 5:  # FCLB_ERR_CAPACITY is the answer to the size query
        check(rc if rc else 3)
    ch = np.zeros((ni.value, 8), np.uint32)
    full = np.zeros(ni.value, np.uint8)
    leaf = np.zeros(max(nl.value, 1), np.uint8)
    check(fn(_ptr(pts), len(pts), resolution, half_shape, scalar_type, _ptr(ch), _ptr(full), ni.value, C.byref(ni), _ptr(leaf),
             len(leaf), C.byref(nl), _ptr(root), C.byref(layers)))
    return ch, full, leaf[:nl.value], root, layers.value


def octree_prune_host(inner_children, inner_full, leaf_bits, root_aabb, n_layers, axis, center, extent, scalar_type,
                      pruned=None):
    """pruneOctreeByOBB on the flat arrays (no GPU): returns new (pruned, inner_full, leaf_bits); the inputs are kept.
    Pass the previous outputs (and pruned) to prune further."""
    ch = np.ascontiguousarray(inner_children, np.uint32)
    full = np.array(inner_full, np.uint8)
    leaf = np.array(leaf_bits, np.uint8)
    pr = np.zeros(len(full), np.uint8) if pruned is None else np.array(pruned, np.uint8)
    root = np.ascontiguousarray(root_aabb, np.float64)
    obb = np.ascontiguousarray(np.concatenate([np.asarray(axis, np.float64).reshape(9), np.asarray(center, np.float64),
                                               np.asarray(extent, np.float64)]))
    check(load().fclb_octree_prune_host(_ptr(ch), len(full), len(leaf), _ptr(root), n_layers, _ptr(obb), scalar_type,
                                        _ptr(pr), _ptr(full), _ptr(leaf)))
    return pr, full, leaf


def octree_consolidate_host(inner_children, pruned, leaf_bits, n_layers):
    """Octree::rebuildAccordingToPruneInfo on the flat arrays (no GPU): (inner_children [n',8], inner_full [n'], leaf_bits [m'])."""
    ch = np.ascontiguousarray(inner_children, np.uint32)
    pr = np.ascontiguousarray(pruned, np.uint8)
    leaf = np.ascontiguousarray(leaf_bits, np.uint8)
    o_ch = np.zeros_like(ch)
    o_full = np.zeros(len(pr), np.uint8)
    o_leaf = np.zeros(max(len(leaf), 1), np.uint8)
    ni, nl = C.c_uint32(), C.c_uint32()
    check(load().fclb_octree_consolidate_host(_ptr(ch), len(pr), _ptr(pr), _ptr(leaf), len(leaf), n_layers, _ptr(o_ch), _ptr(o_full),
                                              C.byref(ni), _ptr(o_leaf), C.byref(nl)))
    return o_ch[:ni.value].copy(), o_full[:ni.value].copy(), o_leaf[:nl.value].copy()


def octree_build(points, resolution, half_shape, scalar_type) -> int:
    """octree_build_host + upload: a device octree straight from a point cloud."""
    pts = np.ascontiguousarray(points, np.float64)
    h = C.c_uint64()
    check(load().fclb_octree_build(_ptr(pts), len(pts), resolution, half_shape, scalar_type, C.byref(h)))
    return h.value


def octree_build_points_host(points, resolution, half_shape, scalar_type) -> int:
    """Octree<S>::rebuildTree on the device from host points (n x 3, rounded once to S)"""
    pts = np.ascontiguousarray(np.asarray(points, np.float64).astype(np_dtype(scalar_type)))
    h = C.c_uint64()
    fn = load().fclb_octree_build_points_host
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
    check(fn(_ptr(pts), len(pts), resolution, half_shape, scalar_type, C.byref(h)))
    return h.value


def octree_export(h: int):
    """(inner_children [n,8] u32, inner_full [n] u8, leaf_bits [m] u8, root_aabb [6], n_layers) of a device octree"""
    ni, nl, layers = C.c_uint32(), C.c_uint32(), C.c_int()
    root = np.zeros(6, np.float64)
    info = load().fclb_octree_info
    info.argtypes = [C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.c_void_p]
    check(info(h, C.byref(ni), C.byref(nl), C.byref(layers), _ptr(root)))
    ch = np.zeros((ni.value, 8), np.uint32)
    full = np.zeros(ni.value, np.uint8)
    leaf = np.zeros(nl.value, np.uint8)
    ex = load().fclb_octree_export
    ex.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    check(ex(h, _ptr(ch), _ptr(full), _ptr(leaf)))
    return ch, full, leaf, root, layers.value


def octree_release(h: int) -> None:
    check(load().fclb_octree_release(h))


def octree_shape_collide_batch_host(octree, table, shape_ids, poses_octree, poses_shape, scalar_type, request: Request,
                                    want_node=False):
    n = len(poses_octree)
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    counts = np.zeros(n, np.uint32)
    node = np.zeros(n, np.int64) if want_node else None
    check(load().fclb_octree_shape_collide_batch_host(octree, table, _ptr(ids), _ptr(poses_octree), _ptr(poses_shape), n,
                                                      scalar_type, C.cast(C.pointer(request), C.c_void_p), _ptr(counts),
                                                      _ptr(node)))
    return counts, node


SCENE_BVH, SCENE_HEIGHTMAP, SCENE_OCTREE = 0, 1, 2


def scene_shape_contacts_batch_host(kind, scene, table, shape_ids, poses_scene, poses_shape, scalar_type, request: Request,
                                    max_keep):
    """fcl::collide(scene, tf1, Shape, tf2) with an MPR penetration request: (counts, b1 [n,k], contacts [n,k,7])."""
    n = len(poses_scene)
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    counts = np.zeros(n, np.uint32)
    b1 = np.zeros((n, max_keep), np.int64)
    contacts = np.zeros((n, max_keep, 7), np_dtype(scalar_type))
    check(load().fclb_scene_shape_contacts_batch_host(kind, scene, table, _ptr(ids), _ptr(poses_scene), _ptr(poses_shape), n,
                                                      scalar_type, C.cast(C.pointer(request), C.c_void_p), max_keep,
                                                      _ptr(counts), _ptr(b1), _ptr(contacts)))
    return counts, b1, contacts


def scene_pair_collide_batch_host(kind1, scene1, kind2, scene2, poses1, poses2, scalar_type, request: Request, max_keep=0):
    """fcl::collide(scene1, tf1, scene2, tf2), boolean request: (counts, b1 [n,k], b2 [n,k]); max_keep 0 = counts only."""
    n = len(poses1)
    counts = np.zeros(n, np.uint32)
    b1 = np.zeros((n, max_keep), np.int64) if max_keep else None
    b2 = np.zeros((n, max_keep), np.int64) if max_keep else None
    check(load().fclb_scene_pair_collide_batch_host(kind1, scene1, kind2, scene2, _ptr(poses1), _ptr(poses2), n, scalar_type,
                                                    C.cast(C.pointer(request), C.c_void_p), max_keep, _ptr(counts),
                                                    _ptr(b1), _ptr(b2)))
    return counts, b1, b2


def scene_pair_contacts_batch_host(kind1, scene1, kind2, scene2, poses1, poses2, scalar_type, request: Request, max_keep):
    """fcl::collide(scene1, tf1, scene2, tf2) with an MPR penetration request: (counts, b1, b2 [n,k], contacts [n,k,7])."""
    n = len(poses1)
    counts = np.zeros(n, np.uint32)
    b1 = np.zeros((n, max_keep), np.int64)
    b2 = np.zeros((n, max_keep), np.int64)
    contacts = np.zeros((n, max_keep, 7), np_dtype(scalar_type))
    check(load().fclb_scene_pair_contacts_batch_host(kind1, scene1, kind2, scene2, _ptr(poses1), _ptr(poses2), n, scalar_type,
                                                     C.cast(C.pointer(request), C.c_void_p), max_keep, _ptr(counts),
                                                     _ptr(b1), _ptr(b2), _ptr(contacts)))
    return counts, b1, b2, contacts


def measure_fp_peak(scalar_type) -> float:
    """FMA-chain microbenchmark on the bound device, TFLOP/s."""
    v = C.c_double()
    check(load().fclb_measure_fp_peak(scalar_type, C.byref(v)))
    return v.value


class CcdRequest(C.Structure):
    _fields_ = [("request_type", C.c_uint32), ("max_contacts", C.c_uint32), ("zero_movement_tolerance", C.c_double),
                ("gjk_tolerance", C.c_double), ("max_gjk_iterations", C.c_int32), ("flags", C.c_uint32)]


def translational_ccd_batch_host(table, pairs, poses1, poses2, displacements, scalar_type, request_type=0, zero_tol=0.0,
                                 gjk_tol=0.0, max_iter=0):
    """fcl::translational_ccd per query (shape pairs): (hit u8 [n], toc [n, 2])"""
    n = len(pairs)
    dt = np_dtype(scalar_type)
    hit = np.zeros(n, np.uint8)
    toc = np.zeros((n, 2), dt)
    r = CcdRequest(request_type, 1, zero_tol, gjk_tol, max_iter, 0)
    fn = load().fclb_translational_ccd_batch_host
    fn.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    check(fn(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), _ptr(displacements), n, scalar_type, C.cast(C.pointer(r), C.c_void_p),
             _ptr(hit), _ptr(toc)))
    return hit, toc


def translational_ccd_mesh_batch_host(bvh, table, shape_ids, poses_shape, poses_mesh, displacements, scalar_type, request_type=0,
                                      max_contacts=1, zero_tol=0.0, mesh_moves=False, max_keep=8):
    """fcl::translational_ccd(shape, mesh) per query: (counts u32 [n], triangle ids i64 [n, keep], toc [n, keep, 2])"""
    n = len(shape_ids)
    dt = np_dtype(scalar_type)
    counts = np.zeros(n, np.uint32)
    prim = np.zeros((n, max_keep), np.int64)
    toc = np.zeros((n, max_keep, 2), dt)
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    r = CcdRequest(request_type, max_contacts, zero_tol, 0.0, 0, 0)
    fn = load().fclb_translational_ccd_mesh_batch_host
    fn.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int,
                   C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    check(fn(bvh, table, _ptr(ids), _ptr(poses_shape), _ptr(poses_mesh), _ptr(displacements), n, scalar_type,
             C.cast(C.pointer(r), C.c_void_p), 1 if mesh_moves else 0, max_keep, _ptr(counts), _ptr(prim), _ptr(toc)))
    return counts, prim, toc


def translational_ccd_mesh_pair_batch_host(bvh1, bvh2, poses1, poses2, displacements, scalar_type, request_type=0, max_contacts=1,
                                           zero_tol=0.0, max_keep=8):
    """fcl::translational_ccd(mesh, mesh) per query: (counts u32 [n], (b1, b2) i64 [n, keep, 2], toc [n, keep, 2])"""
    n = len(poses1)
    dt = np_dtype(scalar_type)
    counts = np.zeros(n, np.uint32)
    prim = np.zeros((n, max_keep, 2), np.int64)
    toc = np.zeros((n, max_keep, 2), dt)
    r = CcdRequest(request_type, max_contacts, zero_tol, 0.0, 0, 0)
    fn = load().fclb_translational_ccd_mesh_pair_batch_host
    fn.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_uint32,
                   C.c_void_p, C.c_void_p, C.c_void_p]
    check(fn(bvh1, bvh2, _ptr(poses1), _ptr(poses2), _ptr(displacements), n, scalar_type, C.cast(C.pointer(r), C.c_void_p), max_keep,
             _ptr(counts), _ptr(prim), _ptr(toc)))
    return counts, prim, toc


def translational_ccd_scene_batch_host(scene_kind, scene, table, shape_ids, poses_shape, poses_scene, displacements, scalar_type,
                                       request_type=0, max_contacts=1, scene_moves=False, max_keep=8):
    """fcl::translational_ccd(shape, heightmap | octree): (counts, codes i64 [n, keep], toc [n, keep, 2], boxes [n, keep, 6])"""
    n = len(shape_ids)
    dt = np_dtype(scalar_type)
    counts = np.zeros(n, np.uint32)
    code = np.zeros((n, max_keep), np.int64)
    toc = np.zeros((n, max_keep, 2), dt)
    box = np.zeros((n, max_keep, 6), dt)
    ids = np.ascontiguousarray(shape_ids, np.uint32)
    r = CcdRequest(request_type, max_contacts, 0.0, 0.0, 0, 0)
    fn = load().fclb_translational_ccd_scene_batch_host
    fn.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                   C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    check(fn(scene_kind, scene, table, _ptr(ids), _ptr(poses_shape), _ptr(poses_scene), _ptr(displacements), n, scalar_type,
             C.cast(C.pointer(r), C.c_void_p), 1 if scene_moves else 0, max_keep, _ptr(counts), _ptr(code), _ptr(toc), _ptr(box)))
    return counts, code, toc, box


def translational_ccd_scene_mesh_batch_host(scene_kind, scene, bvh, poses_scene, poses_mesh, displacements, scalar_type, request_type=0,
                                            max_contacts=1, mesh_moves=False, max_keep=8):
    """fcl::translational_ccd(heightmap | octree, mesh): (counts, (code, triangle) i64 [n, keep, 2], toc [n, keep, 2], boxes)"""
    n = len(poses_scene)
    dt = np_dtype(scalar_type)
    counts = np.zeros(n, np.uint32)
    ids = np.zeros((n, max_keep, 2), np.int64)
    toc = np.zeros((n, max_keep, 2), dt)
    box = np.zeros((n, max_keep, 6), dt)
    r = CcdRequest(request_type, max_contacts, 0.0, 0.0, 0, 0)
    fn = load().fclb_translational_ccd_scene_mesh_batch_host
    fn.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int,
                   C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    check(fn(scene_kind, scene, bvh, _ptr(poses_scene), _ptr(poses_mesh), _ptr(displacements), n, scalar_type,
             C.cast(C.pointer(r), C.c_void_p), 1 if mesh_moves else 0, max_keep, _ptr(counts), _ptr(ids), _ptr(toc), _ptr(box)))
    return counts, ids, toc, box


def translational_ccd_scene_pair_batch_host(kind1, scene1, kind2, scene2, poses1, poses2, displacements, scalar_type, request_type=0,
                                            max_contacts=1, max_keep=8):
    """fcl::translational_ccd(heightmap | octree, heightmap | octree): (counts, (code1, code2) i64 [n, keep, 2], toc [n, keep, 2],
    boxes [n, keep, 12])"""
    n = len(poses1)
    dt = np_dtype(scalar_type)
    counts = np.zeros(n, np.uint32)
    ids = np.zeros((n, max_keep, 2), np.int64)
    toc = np.zeros((n, max_keep, 2), dt)
    box = np.zeros((n, max_keep, 12), dt)
    r = CcdRequest(request_type, max_contacts, 0.0, 0.0, 0, 0)
    fn = load().fclb_translational_ccd_scene_pair_batch_host
    fn.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                   C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    check(fn(kind1, scene1, kind2, scene2, _ptr(poses1), _ptr(poses2), _ptr(displacements), n, scalar_type,
             C.cast(C.pointer(r), C.c_void_p), max_keep, _ptr(counts), _ptr(ids), _ptr(toc), _ptr(box)))
    return counts, ids, toc, box


def measure_l2_bandwidth() -> float:
    """measured L2 read bandwidth of the bound device, GB/s"""
    v = C.c_double()
    check(load().fclb_measure_l2_bandwidth(C.byref(v)))
    return v.value


def bvh_collide_contacts_batch_host(bvh1, bvh2, poses1, poses2, scalar_type, request: Request, max_keep):
    """fcl::collide(BVH, BVH) with request.useDefaultPenetration(): (counts, ids [n,k,2], contacts [n,k,7])."""
    n = len(poses1)
    counts = np.zeros(n, np.uint32)
    ids = np.zeros((n, max_keep, 2), np.int32)
    contacts = np.zeros((n, max_keep, 7), np_dtype(scalar_type))
    check(load().fclb_bvh_collide_contacts_batch_host(bvh1, bvh2, _ptr(poses1), _ptr(poses2), n, scalar_type,
                                                      C.cast(C.pointer(request), C.c_void_p), max_keep, _ptr(counts),
                                                      _ptr(ids), _ptr(contacts)))
    return counts, ids, contacts


def signed_distance_batch_host(table, pairs, poses1, poses2, scalar_type):
    """GJKSolver::shapeSignedDistance per query: DistanceResult(dist, p1, p2, ok)."""
    n = len(pairs)
    dt = np_dtype(scalar_type)
    out = DistanceResult(np.zeros(n, dt), np.zeros((n, 3), dt), np.zeros((n, 3), dt), np.zeros(n, np.uint8))
    check(load().fclb_signed_distance_batch_host(table, _ptr(pairs), _ptr(poses1), _ptr(poses2), n, scalar_type,
                                                 _ptr(out.dist), _ptr(out.p1), _ptr(out.p2), _ptr(out.ok)))
    return out
