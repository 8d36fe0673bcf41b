"""Seeded synthetic workloads for the configs of BASELINE.json / SURVEY.md 8(d).

Poses follow the reference's test helper semantics (test/test_fcl_utility.h:346-411:
uniform translation in a box + three uniform Euler angles through eulerToMatrix),
but with an explicit numpy PCG64 seed instead of an unseeded rand(), computed in
float64 and then rounded ONCE to the scalar type S.  The rounded arrays are the
shared input of the CUDA path, the oracle port and the reference oracle.
Layout: pose = 12 S, rotation row-major then translation.
"""
from __future__ import annotations

import numpy as np

BOX, SPHERE, ELLIPSOID, CAPSULE, CONE, CYLINDER, CONVEX = range(7)
PAIR_DTYPE = np.dtype([("shape1", np.uint32), ("shape2", np.uint32)])


def euler_to_matrix(a, b, c):
    """Vectorised eulerToMatrix (test/test_fcl_utility.h:346-357)."""
    c1, c2, c3 = np.cos(a), np.cos(b), np.cos(c)
    s1, s2, s3 = np.sin(a), np.sin(b), np.sin(c)
    R = np.empty(a.shape + (3, 3), np.float64)
    R[..., 0, 0] = c1 * c2
    R[..., 0, 1] = -c2 * s1
    R[..., 0, 2] = s2
    R[..., 1, 0] = c3 * s1 + c1 * s2 * s3
    R[..., 1, 1] = c1 * c3 - s1 * s2 * s3
    R[..., 1, 2] = -c2 * s3
    R[..., 2, 0] = s1 * s3 - c1 * c3 * s2
    R[..., 2, 1] = c3 * s1 * s2 + c1 * s3
    R[..., 2, 2] = c2 * c3
    return R


def random_poses(rng: np.random.Generator, n: int, extent: float, dtype) -> np.ndarray:
    """n poses, translation uniform in [-extent, extent]^3, Euler angles in [0, 2pi)."""
    t = rng.uniform(-extent, extent, size=(n, 3))
    ang = rng.uniform(0.0, 2.0 * np.pi, size=(n, 3))
    R = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    out = np.empty((n, 12), np.float64)
    out[:, :9] = R.reshape(n, 9)
    out[:, 9:] = t
    return np.ascontiguousarray(out.astype(dtype))


def make_pairs(s1, s2) -> np.ndarray:
    p = np.empty(len(s1), PAIR_DTYPE)
    p["shape1"] = s1
    p["shape2"] = s2
    return p


def config_c2(n: int, dtype, seed: int = 2001):
    """C2: sphere / capsule / cylinder (round-robin) vs box, distance queries.

    Sphere(0.05), Capsule(0.05, 0.2), Cylinder(0.05, 0.2) vs Box(0.2,0.2,0.2), poses
    uniform in [-0.5,0.5]^3."""
    shapes = [(SPHERE, 0, (0.05,)), (CAPSULE, 0, (0.05, 0.2)), (CYLINDER, 0, (0.05, 0.2)), (BOX, 0, (0.2, 0.2, 0.2))]
    rng = np.random.Generator(np.random.PCG64(seed))
    poses1 = random_poses(rng, n, 0.5, dtype)
    poses2 = random_poses(rng, n, 0.5, dtype)
    s1 = (np.arange(n) % 3).astype(np.uint32)
    s2 = np.full(n, 3, np.uint32)
    return shapes, make_pairs(s1, s2), poses1, poses2


def expand_qt7(qt: np.ndarray) -> np.ndarray:
    """FCLB_POSE_QT7 (n x 7: unit quaternion x, y, z, w, translation) -> 12-S poses, with the arithmetic of Eigen's
    QuaternionBase::toRotationMatrix evaluated in the array's own scalar type (what the device does)."""
    S = qt.dtype.type
    x, y, z, w = qt[:, 0], qt[:, 1], qt[:, 2], qt[:, 3]
    tx, ty, tz = S(2) * x, S(2) * y, S(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    out = np.empty((len(qt), 12), qt.dtype)
    out[:, 0] = S(1) - (tyy + tzz)
    out[:, 1] = txy - twz
    out[:, 2] = txz + twy
    out[:, 3] = txy + twz
    out[:, 4] = S(1) - (txx + tzz)
    out[:, 5] = tyz - twx
    out[:, 6] = txz - twy
    out[:, 7] = tyz + twx
    out[:, 8] = S(1) - (txx + tyy)
    out[:, 9:] = qt[:, 4:]
    return np.ascontiguousarray(out)


def random_poses_qt(rng: np.random.Generator, n: int, extent: float, dtype) -> np.ndarray:
    """n poses as unit quaternion + translation (uniform rotations, translation uniform in [-extent, extent]^3),
    rounded ONCE to the scalar type: the shared input of every implementation (expand_qt7 gives the 12-S form)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    out = np.empty((n, 7), np.float64)
    out[:, :4] = q
    out[:, 4:] = rng.uniform(-extent, extent, size=(n, 3))
    return np.ascontiguousarray(out.astype(dtype))


def config_c2_qt(n: int, dtype, seed: int = 2001):
    """C2 with poses in the compact FCLB_POSE_QT7 encoding: (shapes, pairs, qt1, qt2); the 12-S poses every other
    implementation reads are expand_qt7(qt)."""
    shapes = [(SPHERE, 0, (0.05,)), (CAPSULE, 0, (0.05, 0.2)), (CYLINDER, 0, (0.05, 0.2)), (BOX, 0, (0.2, 0.2, 0.2))]
    rng = np.random.Generator(np.random.PCG64(seed))
    qt1 = random_poses_qt(rng, n, 0.5, dtype)
    qt2 = random_poses_qt(rng, n, 0.5, dtype)
    s1 = (np.arange(n) % 3).astype(np.uint32)
    s2 = np.full(n, 3, np.uint32)
    return shapes, make_pairs(s1, s2), qt1, qt2


def config_c1_boxes(n: int, dtype, seed: int = 1001):
    """C1: Box(2,1,0.5) vs Box(1,1,1), both poses uniform in [-2,2]^3
    (matches test/cvx_collide/test_epa2_with_gjk2.cpp:76-80)."""
    shapes = [(BOX, 0, (2.0, 1.0, 0.5)), (BOX, 0, (1.0, 1.0, 1.0))]
    rng = np.random.Generator(np.random.PCG64(seed))
    poses1 = random_poses(rng, n, 2.0, dtype)
    poses2 = random_poses(rng, n, 2.0, dtype)
    return shapes, make_pairs(np.zeros(n, np.uint32), np.ones(n, np.uint32)), poses1, poses2


def ellipsoid_mesh(rx, ry, rz, n_lat=8, n_lon=8):
    """Closed triangle mesh of an ellipsoid: 2 poles + (n_lat-1) rings of n_lon
    vertices => 58 vertices / 112 triangles for 8x8, the size of the reference's
    test convex (test/create_primitive_mesh-inl.h:148-182).  Returns
    (verts[n,3], faces in the reference encoding, num_faces)."""
    verts = [(0.0, 0.0, rz)]
    for i in range(1, n_lat):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            verts.append((rx * np.sin(th) * np.cos(ph), ry * np.sin(th) * np.sin(ph), rz * np.cos(th)))
    verts.append((0.0, 0.0, -rz))
    faces = []

    def ring(i, j):
        return 1 + (i - 1) * n_lon + (j % n_lon)

    for j in range(n_lon):
        faces.append((0, ring(1, j), ring(1, j + 1)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i + 1, j), ring(i + 1, j + 1), ring(i, j + 1)
            faces.append((a, b, c))
            faces.append((a, c, d))
    south = len(verts) - 1
    for j in range(n_lon):
        faces.append((south, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)))
    enc = []
    for f in faces:
        enc.extend((3,) + f)
    return np.asarray(verts, np.float64), np.asarray(enc, np.int32), len(faces)


def random_hull16(seed: int = 7, scale=(0.25, 0.2, 0.3)):
    """A 16-vertex convex polytope: two twisted octagonal rings (all 16 points are
    hull vertices by construction), triangulated.  Exercises the <=32-vertex
    linear-scan support path (convex-inl.h:133-150)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = 8
    verts = []
    for ring_i, (z, r) in enumerate(((0.6, 0.8), (-0.6, 1.0))):
        for j in range(m):
            ph = 2 * np.pi * (j + 0.5 * ring_i) / m
            rr = r * (1.0 + 0.02 * rng.uniform(-1, 1))
            verts.append((scale[0] * rr * np.cos(ph), scale[1] * rr * np.sin(ph), scale[2] * z))
    verts = np.asarray(verts, np.float64)
    faces = []
    # top and bottom caps as fans
    for j in range(1, m - 1):
        faces.append((0, j, j + 1))
        faces.append((m, m + j + 1, m + j))
    # side band
    for j in range(m):
        a, b = j, (j + 1) % m
        c, d = m + j, m + (j + 1) % m
        faces.append((a, c, b))
        faces.append((b, c, d))
    enc = []
    for f in faces:
        enc.extend((3,) + f)
    return verts, np.asarray(enc, np.int32), len(faces)


def config_c1_convex(n: int, dtype, seed: int = 1003):
    """C1b(ii): 58-vertex ellipsoid hull (0.2,0.3,0.4) vs a 16-vertex hull, poses in
    [-0.3,0.3]^3.  Returns the convex meshes separately: the caller uploads them and
    builds the shape table [(CONVEX, slot0), (CONVEX, slot1)]."""
    m0 = ellipsoid_mesh(0.2, 0.3, 0.4)
    m1 = random_hull16()
    rng = np.random.Generator(np.random.PCG64(seed))
    poses1 = random_poses(rng, n, 0.3, dtype)
    poses2 = random_poses(rng, n, 0.3, dtype)
    return (m0, m1), make_pairs(np.zeros(n, np.uint32), np.ones(n, np.uint32)), poses1, poses2


# ---- procedural meshes (config C3) ---------------------------------------------
def noisy_uv_sphere(n_lat=51, n_lon=100, radius=1.0, noise=0.08, seed=3001):
    """Closed UV sphere with smooth radial noise: 2*n_lon*(n_lat-1) triangles
    (10,000 for the defaults).  Returns (verts[n,3] float64, tris[m,3] int32)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.uniform(1.0, 4.0, size=(6, 3))
    ph = rng.uniform(0, 2 * np.pi, size=6)

    def bump(p):
        return sum(np.sin(p @ k[i] + ph[i]) for i in range(6)) / 6.0

    verts = [(0.0, 0.0, 1.0)]
    for i in range(1, n_lat):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            a = 2 * np.pi * j / n_lon
            verts.append((np.sin(th) * np.cos(a), np.sin(th) * np.sin(a), np.cos(th)))
    verts.append((0.0, 0.0, -1.0))
    verts = np.asarray(verts, np.float64)
    verts = verts * (radius * (1.0 + noise * bump(verts)))[:, None]

    def ring(i, j):
        return 1 + (i - 1) * n_lon + (j % n_lon)

    tris = []
    for j in range(n_lon):
        tris.append((0, ring(1, j), ring(1, j + 1)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i + 1, j), ring(i + 1, j + 1), ring(i, j + 1)
            tris.append((a, b, c))
            tris.append((a, c, d))
    south = len(verts) - 1
    for j in range(n_lon):
        tris.append((south, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)))
    return verts, np.asarray(tris, np.int32)


def noisy_torus(n_major=100, n_minor=50, R=0.75, r=0.3, noise=0.06, seed=3002):
    """Torus with smooth radial noise: 2*n_major*n_minor triangles (10,000 for the defaults)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.uniform(1.0, 4.0, size=(6, 3))
    ph = rng.uniform(0, 2 * np.pi, size=6)
    verts = []
    for i in range(n_major):
        u = 2 * np.pi * i / n_major
        for j in range(n_minor):
            v = 2 * np.pi * j / n_minor
            base = np.array([np.cos(u), np.sin(u), 0.0])
            p = base * R + r * (np.cos(v) * base + np.array([0.0, 0.0, np.sin(v)]))
            verts.append(p)
    verts = np.asarray(verts, np.float64)
    b = sum(np.sin(verts @ k[i] + ph[i]) for i in range(6)) / 6.0
    verts = verts * (1.0 + noise * b)[:, None]

    def idx(i, j):
        return (i % n_major) * n_minor + (j % n_minor)

    tris = []
    for i in range(n_major):
        for j in range(n_minor):
            a, bb, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            tris.append((a, bb, c))
            tris.append((a, c, d))
    return verts, np.asarray(tris, np.int32)


def config_c3_poses(n: int, dtype, seed: int = 3003, extent: float = 2.5):
    """C3: mesh 2 at identity, mesh 1 translated uniformly in [-extent,extent]^3 with a
    random rotation."""
    rng = np.random.Generator(np.random.PCG64(seed))
    poses1 = random_poses(rng, n, extent, dtype)
    poses2 = np.zeros((n, 12), dtype)
    poses2[:, 0] = poses2[:, 4] = poses2[:, 8] = 1
    return poses1, np.ascontiguousarray(poses2)


# ---- heightmaps (config C4) ------------------------------------------------------
def terrain_points(n_points: int, half_range: float, z_max: float = 0.3, seed: int = 4200):
    """Random 3-D points over a smooth terrain (pattern of the reference's
    test/geometry/heightmap/test_heightmap_bvh_collision.cpp:15-24: uniform x/y, bounded z);
    a few points fall outside the map or below z = 0 on purpose (both are ignored by
    FlatHeightMap::updateHeightsByPointGenerationFunctor)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xy = rng.uniform(-1.05 * half_range, 1.05 * half_range, size=(n_points, 2))
    k = rng.uniform(1.0, 3.0, size=(4, 2)) / half_range
    ph = rng.uniform(0, 2 * np.pi, size=4)
    base = sum(np.sin(xy @ k[i] + ph[i]) for i in range(4)) / 4.0
    z = z_max * (0.45 + 0.45 * base) + rng.uniform(-0.03, 0.03, size=n_points) * z_max
    return np.ascontiguousarray(np.concatenate([xy, z[:, None]], axis=1))


def compose_poses(a: np.ndarray, b: np.ndarray, dtype) -> np.ndarray:
    """Row-wise a * b for [n,12] pose arrays (float64 arithmetic, rounded once to dtype)."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    Ra, ta = a[:, :9].reshape(-1, 3, 3), a[:, 9:]
    Rb, tb = b[:, :9].reshape(-1, 3, 3), b[:, 9:]
    out = np.empty((len(a), 12), np.float64)
    out[:, :9] = (Ra @ Rb).reshape(-1, 9)
    out[:, 9:] = np.einsum("nij,nj->ni", Ra, tb) + ta
    return np.ascontiguousarray(out.astype(dtype))


def heightmap_query_poses(n: int, dtype, half_range: float, z_lo: float, z_hi: float, seed: int = 4300):
    """Heightmap poses (random rigid) and shape poses expressed relative to the map:
    shape translation uniform over the map footprint x [z_lo, z_hi], random rotation."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tf_hm = random_poses(rng, n, 1.0, np.float64)
    local = random_poses(rng, n, 1.0, np.float64)
    local[:, 9:11] = rng.uniform(-1.1 * half_range, 1.1 * half_range, size=(n, 2))
    local[:, 11] = rng.uniform(z_lo, z_hi, size=n)
    return np.ascontiguousarray(tf_hm.astype(dtype)), compose_poses(tf_hm, local, dtype)


# ---- broadphase scenes (config C5) -------------------------------------------------
def config_c5_scene(n_objects: int, dtype, seed: int = 5000, neighbours: float = 8.0):
    """C5: 40 % Box(5,10,20), 30 % Sphere(30) x0.1, 30 % Cylinder(10,40) x0.1 -- the mix of the
    reference's broadphase tests (test/broadphase/test_binary_AABB_tree_collision.cpp:13-42) scaled by
    0.1 -- at random poses in a cube sized for about `neighbours` overlapping-AABB neighbours per
    object; 3 % of the objects keep an identity rotation (tight-AABB branch of computeAABB).
    Returns (shapes, shape_ids[n], poses[n,12])."""
    shapes = [(BOX, 0, (0.5, 1.0, 2.0)), (SPHERE, 0, (3.0,)), (CYLINDER, 0, (1.0, 4.0))]
    rng = np.random.Generator(np.random.PCG64(seed))
    shape_ids = rng.choice(3, size=n_objects, p=(0.4, 0.3, 0.3)).astype(np.uint32)
    # world AABBs are bounding-sphere boxes of half width r: box 1.146, sphere 3, cylinder 2.236
    # (mean overlap volume of two boxes of half widths a, b is (2(a+b))^3)
    r = np.array([np.sqrt(0.25 ** 2 + 0.5 ** 2 + 1.0), 3.0, np.sqrt(1.0 + 1.0 + 4.0)])
    p = np.array([0.4, 0.3, 0.3])
    mean_vol = sum(p[i] * p[j] * (2 * (r[i] + r[j])) ** 3 for i in range(3) for j in range(3))
    side = (n_objects * mean_vol / neighbours) ** (1.0 / 3.0)
    poses = random_poses(rng, n_objects, side / 2.0, np.float64)
    ident = rng.uniform(size=n_objects) < 0.03
    poses[ident, :9] = np.eye(3).reshape(9)
    return shapes, shape_ids, np.ascontiguousarray(poses.astype(dtype))


# ---- robot-arm scene (config C4) ---------------------------------------------------
def ellipsoid_point_hull(n_verts: int, semi_axes, seed: int):
    """Convex hull of n_verts random points ON an ellipsoid (all of them are hull vertices).
    Returns (verts[n,3], faces in the reference encoding, num_faces) for fclb_convex_upload."""
    from scipy.spatial import ConvexHull

    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.normal(size=(n_verts, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    verts = d * np.asarray(semi_axes)[None, :]
    hull = ConvexHull(verts)
    assert len(hull.vertices) == n_verts
    enc = []
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = (int(v) for v in simplex)
        nrm = np.cross(verts[b] - verts[a], verts[c] - verts[a])
        if np.dot(nrm, eq[:3]) < 0:  # orient outward
            b, c = c, b
        enc.extend((3, a, b, c))
    return np.ascontiguousarray(verts), np.asarray(enc, np.int32), len(hull.simplices)


def c4_links(seed: int = 4001):
    """7 convex link hulls, 32-128 vertices, semi-axes 0.05-0.25."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_verts = [32, 48, 64, 80, 96, 112, 128]
    return [ellipsoid_point_hull(n_verts[i], rng.uniform(0.05, 0.25, size=3), seed + 1 + i) for i in range(7)]


def terrain_height(x, y):
    return 0.35 + 0.2 * np.sin(1.7 * x + 0.3) * np.cos(1.3 * y - 0.5) + 0.1 * np.sin(3.1 * x * y * 0.25 + 1.0)


def c4_scene_mesh(grid: int = 300, n_boxes: int = 1666, seed: int = 4050):
    """Procedural scene in [-2,2]^2 x [0,1.5]: a grid terrain (2*grid^2 triangles) plus n_boxes small
    boxes (12 triangles each) standing on it: 180,000 + 19,992 = 199,992 triangles for the defaults."""
    xs = np.linspace(-2.0, 2.0, grid + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    Z = terrain_height(X, Y)
    verts = [np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)]
    idx = np.arange((grid + 1) * (grid + 1)).reshape(grid + 1, grid + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    tris = [np.stack([a, b, c], axis=1), np.stack([a, c, d], axis=1)]
    rng = np.random.Generator(np.random.PCG64(seed))
    base = (grid + 1) * (grid + 1)
    corner = np.array([[i, j, k] for i in (-1, 1) for j in (-1, 1) for k in (-1, 1)], np.float64)
    box_tris = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6],
                         [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], np.int64)
    for _ in range(n_boxes):
        cx, cy = rng.uniform(-1.9, 1.9, size=2)
        h = rng.uniform(0.02, 0.08, size=3)
        cz = terrain_height(cx, cy) + h[2] * rng.uniform(0.5, 4.0)
        verts.append(corner * h[None, :] + np.array([cx, cy, cz])[None, :])
        tris.append(box_tris + base)
        base += 8
    return np.ascontiguousarray(np.concatenate(verts)), np.ascontiguousarray(np.concatenate(tris).astype(np.int32))


def c4_heightmap_points(n_points: int = 1_000_000, seed: int = 4060):
    """Random points on the same terrain for LayeredHeightMap(0.004, 512) (1024^2 bottom layer)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xy = rng.uniform(-2.048, 2.048, size=(n_points, 2))
    z = terrain_height(xy[:, 0], xy[:, 1]) + rng.uniform(-0.01, 0.01, size=n_points)
    return np.ascontiguousarray(np.concatenate([xy, z[:, None]], axis=1))


def config_c4_poses(n_configs: int, dtype, seed: int = 4100):
    """Per configuration 7 link poses: a random-walk chain (link i+1 sits 0.15-0.3 m from link i)
    starting above a random point of the workspace, heights 0.05-0.6 m above the terrain.
    Returns (shape_ids[7n], link poses[7n,12], identity poses[7n,12]) -- query 7c+i is link i of
    configuration c; the scene mesh and the heightmap sit at the identity."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = n_configs
    pos = np.empty((n, 7, 3))
    p = np.empty((n, 3))
    p[:, :2] = rng.uniform(-1.6, 1.6, size=(n, 2))
    p[:, 2] = terrain_height(p[:, 0], p[:, 1]) + rng.uniform(0.05, 0.6, size=n)
    for i in range(7):
        pos[:, i] = p
        step = rng.normal(size=(n, 3))
        step[:, 2] *= 0.4
        step /= np.linalg.norm(step, axis=1)[:, None]
        p = p + step * rng.uniform(0.15, 0.3, size=(n, 1))
        p[:, :2] = np.clip(p[:, :2], -1.9, 1.9)
        p[:, 2] = np.maximum(p[:, 2], terrain_height(p[:, 0], p[:, 1]) - 0.05)
    ang = rng.uniform(0.0, 2.0 * np.pi, size=(n * 7, 3))
    R = euler_to_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    poses = np.empty((n * 7, 12), np.float64)
    poses[:, :9] = R.reshape(-1, 9)
    poses[:, 9:] = pos.reshape(-1, 3)
    ident = np.zeros((n * 7, 12), dtype)
    ident[:, 0] = ident[:, 4] = ident[:, 8] = 1
    shape_ids = np.tile(np.arange(7, dtype=np.uint32), n)
    return shape_ids, np.ascontiguousarray(poses.astype(dtype)), np.ascontiguousarray(ident)
