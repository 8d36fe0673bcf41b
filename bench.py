#!/usr/bin/env python
"""bench.py -- headline benchmark: narrowphase queries/s on B200 vs host-CPU fcl.

Workload (BASELINE.json configs[1]): mixed-primitive
DISTANCE queries -- Sphere / Capsule / Cylinder (round-robin) vs Box, random
poses, Q = 10M pose pairs per GPU, scalar type float (double with --dtype f64).
One "step" = one pass of the hot path over the whole Q-query batch.

  value  : whole-job queries/s with the batch resident in HBM
           (fclb_distance_batch_dev), timed with CUDA events on the engine's stream.
  e2e    : same metric through the host-buffer C-ABI call
           (fclb_distance_batch_host): pinned host inputs -> H2D -> kernels -> D2H
           of every result array, all inside the timed region.
  roofline / cpu_baseline: see DESIGN.md "Measurement".

Launch: python bench.py [--gpus N --steps K --warmup W]  (N>1 under torchrun).
--impl reference times the reference's own CPU implementation (oracle/_ref, the
unmodified mind-fcl headers; falls back to the oracle port) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mind-fcl_b200"))

import scenes  # noqa: E402

METRIC = "narrowphase_queries_per_sec"
UNIT = "queries/s"
WORKLOADS = {
    "c2": "C2 mixed-primitive distance: sphere/capsule/cylinder vs box",
    "c1a": "C1a box-box fcl::collide with contacts (boxBox2), max_contacts=4",
    "c1b": "C1b box-box through cvx_collide GJK(128,1e-6)+EPA(256,255,1e-6)",
    "c1b_convex": "C1b convex-convex (58-vertex vs 16-vertex hulls) GJK+EPA",
    "c3": "C3 mesh-mesh BVHModel<OBBRSS> boolean collide, two 10k-triangle meshes, random relative poses",
    "c4": "C4 7 convex links vs 200k-triangle scene mesh + 1024^2 heightmap, boolean collide per (link, configuration)",
    "c5": "C5 broadphase + narrowphase: 100k mixed objects per scene (computeAABB, tree build, SelfCollision, "
          "boolean collide per candidate pair), 8 scenes per GPU per step",
}
WORKLOAD = WORKLOADS["c2"]


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    if oracle_py.have_ref():
        return oracle_py.RefOracle()
    if oracle_py.have_port():
        return oracle_py.PortOracle()
    raise RuntimeError("no CPU oracle built: run `make -C oracle`")


class Workload:
    """One bench configuration: synthetic inputs, the C-ABI calls of a step, byte counts, CPU leg."""

    def __init__(self, name, n, dtype_name, seed):
        self.name = name
        self.n = n
        self.dtype_name = dtype_name
        self.np_dtype = np.float32 if dtype_name == "f32" else np.float64
        self.sb = 4 if dtype_name == "f32" else 8
        self.convex = None
        if name == "c2":
            self.shapes, self.pairs, self.poses1, self.poses2 = scenes.config_c2(n, self.np_dtype, seed=2001 + seed)
            self.kind = "distance"
        elif name in ("c1a", "c1b"):
            self.shapes, self.pairs, self.poses1, self.poses2 = scenes.config_c1_boxes(n, self.np_dtype, seed=1001 + seed)
            self.kind = "collide" if name == "c1a" else "gjk_epa"
        elif name == "c1b_convex":
            self.convex, self.pairs, self.poses1, self.poses2 = scenes.config_c1_convex(n, self.np_dtype, seed=1003 + seed)
            self.shapes = None
            self.kind = "gjk_epa"
        else:
            raise SystemExit(f"unknown workload {name}")
        self.max_keep = 4 if name == "c1a" else 0
        self.req_kw = dict(max_contacts=4 if name == "c1a" else 1, penetration_mode=1)

    # -- bytes ---------------------------------------------------------------
    def h2d_bytes(self):
        return self.n * (24 * self.sb + 8)

    def d2h_bytes(self):
        if self.kind == "distance":
            return self.n * (7 * self.sb + 1)
        if self.kind == "collide":
            return self.n * (4 + self.max_keep * 9 * self.sb)
        return self.n * (8 + 7 * self.sb)

    def algorithmic_bytes_per_query(self):
        """2 poses in (24 S) + the result record."""
        if self.kind == "distance":
            return 24 * self.sb + 7 * self.sb + 1  # dist + 2 witness points + flag
        if self.kind == "collide":
            return 24 * self.sb + 4 + 7 * self.sb  # count + one contact (c-bar ~ 1 of colliding pairs)
        return 24 * self.sb + 1 + 7 * self.sb

    # -- device side -----------------------------------------------------------
    def setup(self, fclb, torch, dev):
        self.fclb = fclb
        n = self.n
        td = torch.float32 if self.dtype_name == "f32" else torch.float64
        self.st = fclb.F32 if self.dtype_name == "f32" else fclb.F64
        if self.convex is not None:
            slots = [fclb.convex_upload(*m) for m in self.convex]
            self.shapes = [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())]
        self.table = fclb.shapes_upload(self.shapes)
        self.req = fclb.make_request(**self.req_kw)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        self.h_pairs = pin(self.pairs.view(np.uint32).reshape(n, 2).view(np.int32))
        self.h_p1, self.h_p2 = pin(self.poses1), pin(self.poses2)
        self.d_pairs, self.d_p1, self.d_p2 = self.h_pairs.to(dev), self.h_p1.to(dev), self.h_p2.to(dev)

        def bufs(device, pinned):
            mk = (lambda *sh, dtype: torch.empty(*sh, dtype=dtype).pin_memory()) if pinned else \
                (lambda *sh, dtype: torch.empty(*sh, dtype=dtype, device=device))
            if self.kind == "distance":
                return [mk(n, dtype=td), mk(n, 3, dtype=td), mk(n, 3, dtype=td), mk(n, dtype=torch.uint8)]
            if self.kind == "collide":
                return [mk(n, self.max_keep, 9, dtype=td), mk(n, dtype=torch.int32)]
            return [mk(n, dtype=torch.int32), mk(n, dtype=torch.int32), mk(n, 7, dtype=td)]

        self.d_out = bufs(dev, False)
        self.h_out = bufs(None, True)

    def _call(self, host):
        f = self.fclb
        pairs, p1, p2 = (self.h_pairs, self.h_p1, self.h_p2) if host else (self.d_pairs, self.d_p1, self.d_p2)
        out = self.h_out if host else self.d_out
        lib = f.load()
        P = f._ptr
        if self.kind == "distance":
            fn = lib.fclb_distance_batch_host if host else lib.fclb_distance_batch_dev
            f.check(fn(self.table, P(pairs), P(p1), P(p2), self.n, self.st, 0.0, 0, P(out[0]), P(out[1]), P(out[2]), P(out[3])))
        elif self.kind == "collide":
            fn = lib.fclb_collide_batch_host if host else lib.fclb_collide_batch_dev
            import ctypes as C
            f.check(fn(self.table, P(pairs), P(p1), P(p2), self.n, self.st, C.cast(C.pointer(self.req), C.c_void_p),
                       self.max_keep, P(out[0]), P(out[1])))
        else:
            fn = lib.fclb_gjk_epa_batch_host if host else lib.fclb_gjk_epa_batch_dev
            import ctypes as C
            f.check(fn(self.table, P(pairs), P(p1), P(p2), self.n, self.st, C.cast(C.pointer(self.req), C.c_void_p),
                       P(out[0]), P(out[1]), P(out[2])))

    def step_dev(self):
        self._call(False)

    def step_host(self):
        self._call(True)

    # -- CPU leg -------------------------------------------------------------------
    def cpu_run(self, oracle, threads):
        shapes = self.shapes
        if self.convex is not None:
            slots = [oracle.register_convex(*m) for m in self.convex]
            shapes = [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())]
        if self.kind == "distance":
            return lambda: oracle.distance_batch(shapes, self.pairs, self.poses1, self.poses2, threads=threads)
        if self.kind == "collide":
            return lambda: oracle.collide_batch(shapes, self.pairs, self.poses1, self.poses2, max_keep=self.max_keep,
                                                threads=threads, **self.req_kw)
        return lambda: oracle.gjk_epa_batch(shapes, self.pairs, self.poses1, self.poses2, threads=threads)


class MeshWorkload:
    """C3: fcl::collide(BVHModel<OBBRSS>, BVHModel<OBBRSS>) per relative pose, boolean (max_contacts=1)."""

    kind = "bvh_collide"

    def __init__(self, name, n, dtype_name, seed):
        self.name = name
        self.n = n
        self.dtype_name = dtype_name
        self.np_dtype = np.float32 if dtype_name == "f32" else np.float64
        self.sb = 4 if dtype_name == "f32" else 8
        self.meshes = [scenes.noisy_uv_sphere(), scenes.noisy_torus()]
        self.poses1, self.poses2 = scenes.config_c3_poses(n, self.np_dtype, seed=3003 + seed)
        self.visits = None  # (BV-pair tests, leaf-pair tests) of one device launch

    def h2d_bytes(self):
        return self.n * 24 * self.sb

    def d2h_bytes(self):
        return self.n * 4

    def algorithmic_bytes_per_query(self):
        """poses in + result out + N_bv * 2 * node_bytes + N_leaf * 2 * tri_bytes, with
        N_bv / N_leaf = the node-pair / triangle-pair tests this launch executed (fclb_bvh_last_visit_counts)."""
        node_b, tri_b = 16 * self.sb, 9 * self.sb
        n_bv, n_leaf = self.visits if self.visits else (0, 0)
        return 24 * self.sb + 4 + (n_bv * 2 * node_b + n_leaf * 2 * tri_b) / max(self.n, 1)

    def setup(self, fclb, torch, dev):
        self.fclb = fclb
        self.st = fclb.F32 if self.dtype_name == "f32" else fclb.F64
        self.handles = [fclb.bvh_build(v, t, self.st) for v, t in self.meshes]
        self.req = fclb.make_request(max_contacts=1)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        self.h_p1, self.h_p2 = pin(self.poses1), pin(self.poses2)
        self.d_p1, self.d_p2 = self.h_p1.to(dev), self.h_p2.to(dev)
        self.d_out = torch.empty(self.n, dtype=torch.int32, device=dev)
        self.h_out = torch.empty(self.n, dtype=torch.int32).pin_memory()

    def step_dev(self):
        f = self.fclb
        f.bvh_collide_batch_dev(self.handles[0], self.handles[1], self.d_p1, self.d_p2, self.n, self.st, self.req,
                                self.d_out)
        self.visits = f.bvh_last_visit_counts()

    def step_host(self):
        import ctypes as C
        f = self.fclb
        f.check(f.load().fclb_bvh_collide_batch_host(self.handles[0], self.handles[1], f._ptr(self.h_p1),
                                                     f._ptr(self.h_p2), self.n, self.st,
                                                     C.cast(C.pointer(self.req), C.c_void_p), f._ptr(self.h_out), None))

    def roof_note(self):
        return ("both trees (2.6 MB f32) and triangles are L2-resident by construction of the config: the algorithmic "
                "node/triangle bytes are served by L2, not HBM, so 'achieved' is node+triangle fetch bandwidth quoted "
                "against the HBM peak as SURVEY.md 8(d) asks; the kernel is bound by FP32 issue of the 15-axis OBB test "
                "and SIMT divergence (DESIGN.md 4.5)")

    def extra(self):
        n_bv, n_leaf = self.visits if self.visits else (0, 0)
        return {"colliding_fraction": float((self.d_out != 0).float().mean().item()),
                "bv_pair_tests_per_query": n_bv / self.n, "leaf_pair_tests_per_query": n_leaf / self.n}

    def cpu_sample(self):
        return min(self.n, 100_000)

    def cpu_run(self, oracle, threads):
        m = self.cpu_sample()
        ids = [oracle.bvh_create(v, t) for v, t in self.meshes]
        return lambda: oracle.bvh_collide_batch(ids[0], ids[1], self.poses1[:m], self.poses2[:m], threads=threads,
                                                want_pair=False, max_contacts=1)


class ArmSceneWorkload:
    """C4: n configurations x 7 convex links, each link tested against the scene mesh (mesh-shape
    traversal) and against the heightmap (heightmap-shape scan): 14 n narrowphase queries per step."""

    kind = "scene_collide"

    def __init__(self, name, n, dtype_name, seed):
        self.name = name
        self.n_configs = n
        self.n = 14 * n  # narrowphase queries per step
        self.dtype_name = dtype_name
        self.np_dtype = np.float32 if dtype_name == "f32" else np.float64
        self.sb = 4 if dtype_name == "f32" else 8
        self.links = scenes.c4_links()
        self.mesh = scenes.c4_scene_mesh()
        self.hm_points = scenes.c4_heightmap_points()
        self.shape_ids, self.poses, self.ident = scenes.config_c4_poses(n, self.np_dtype, seed=4100 + seed)
        self.visits = {}

    def h2d_bytes(self):
        return 7 * self.n_configs * (4 + 24 * self.sb) * 2  # ids + two poses, for the mesh call and the heightmap call

    def d2h_bytes(self):
        return self.n * 4

    def algorithmic_bytes_per_query(self):
        """SURVEY.md 8(d) row C4, averaged over the two query kinds of a step:
        mesh: 24 S + 4 + N_node*64(128) + N_leaf*36(72); heightmap: 24 S + 4 + 2*N_pixels."""
        node_b, tri_b = 16 * self.sb, 9 * self.sb
        nb, nl = self.visits.get("mesh", (0, 0))
        px, _ = self.visits.get("hm", (0, 0))
        per_step = self.n * (24 * self.sb + 8) + nb * node_b + nl * tri_b + 2 * px
        return per_step / max(self.n, 1)

    def setup(self, fclb, torch, dev):
        self.fclb = fclb
        self.st = fclb.F32 if self.dtype_name == "f32" else fclb.F64
        slots = [fclb.convex_upload(*m) for m in self.links]
        self.shapes = [(scenes.CONVEX, s, ()) for s in slots]
        self.table = fclb.shapes_upload(self.shapes)
        self.bvh = fclb.bvh_build(self.mesh[0], self.mesh[1], self.st)
        self.heights = fclb.heightmap_build_host(self.hm_points, 0.004, 512, self.st)
        self.hm = fclb.heightmap_upload(self.heights, 0.004)
        self.req = fclb.make_request(max_contacts=1)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        m = len(self.shape_ids)
        self.h_ids = pin(self.shape_ids.view(np.int32))
        self.h_pose, self.h_ident = pin(self.poses), pin(self.ident)
        self.d_ids, self.d_pose, self.d_ident = self.h_ids.to(dev), self.h_pose.to(dev), self.h_ident.to(dev)
        self.d_out = [torch.empty(m, dtype=torch.int32, device=dev) for _ in range(2)]
        self.h_out = [torch.empty(m, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.m = m

    def step_dev(self):
        f = self.fclb
        f.bvh_shape_collide_batch_dev(self.bvh, self.table, self.d_ids, self.d_ident, self.d_pose, self.m, self.st,
                                      self.req, self.d_out[0])
        self.visits["mesh"] = f.scene_last_visit_counts()
        self.launch_ms = {"mesh-shape": f.last_kernel_ms()}
        f.heightmap_shape_collide_batch_dev(self.hm, self.table, self.d_ids, self.d_ident, self.d_pose, self.m, self.st,
                                            self.req, self.d_out[1])
        self.visits["hm"] = f.scene_last_visit_counts()
        self.launch_ms["heightmap-shape"] = f.last_kernel_ms()

    def step_host(self):
        import ctypes as C
        f = self.fclb
        lib, P = f.load(), f._ptr
        rq = C.cast(C.pointer(self.req), C.c_void_p)
        f.check(lib.fclb_bvh_shape_collide_batch_host(self.bvh, self.table, P(self.h_ids), P(self.h_ident), P(self.h_pose),
                                                      self.m, self.st, rq, P(self.h_out[0]), None))
        f.check(lib.fclb_heightmap_shape_collide_batch_host(self.hm, self.table, P(self.h_ids), P(self.h_ident),
                                                            P(self.h_pose), self.m, self.st, rq, P(self.h_out[1]), None))

    def kernel_records(self):
        node_b, tri_b = 16 * self.sb, 9 * self.sb
        nb, nl = self.visits.get("mesh", (0, 0))
        px, _ = self.visits.get("hm", (0, 0))
        io = 4 + 24 * self.sb + 4
        ms = getattr(self, "launch_ms", {})
        return [
            {"kernel": "mesh-shape[convex]", "queries": self.m, "avg_ms": ms.get("mesh-shape", 0.0),
             "bytes_per_query": io + (nb * node_b + nl * tri_b) / self.m},
            {"kernel": "heightmap-shape[convex]", "queries": self.m, "avg_ms": ms.get("heightmap-shape", 0.0),
             "bytes_per_query": io + 2 * px / self.m},
        ]

    def roof_note(self):
        return ("two kernels per step (mesh-shape traversal, heightmap-shape scan); 'achieved' counts the node / triangle / "
                "pixel bytes both fetched against the HBM peak; the 25.6 MB tree and the 2 MB grid are L2-resident, the "
                "kernels are bound by FP32 issue of the OBB SAT and the per-leaf MPR (DESIGN.md 4.6-4.7)")

    def extra(self):
        nb, nl = self.visits.get("mesh", (0, 0))
        px, bx = self.visits.get("hm", (0, 0))
        return {"configurations_per_step": self.n_configs,
                "mesh_colliding_fraction": float((self.d_out[0] != 0).float().mean().item()),
                "heightmap_colliding_fraction": float((self.d_out[1] != 0).float().mean().item()),
                "mesh_node_tests_per_query": nb / self.m, "mesh_leaf_tests_per_query": nl / self.m,
                "heightmap_pixels_per_query": px / self.m, "heightmap_boxes_per_query": bx / self.m,
                "kernel_ms": getattr(self, "launch_ms", {})}

    def cpu_sample(self):
        return 14 * min(self.n_configs, 3000)

    def cpu_run(self, oracle, threads):
        k = self.cpu_sample() // 2
        slots = [oracle.register_convex(*m) for m in self.links]
        shapes = [(scenes.CONVEX, s, ()) for s in slots]
        mid = oracle.bvh_create(*self.mesh)
        hid = oracle.heightmap_create(self.hm_points, 0.004, 512)

        def run():
            oracle.mesh_shape_collide_batch(mid, shapes, self.shape_ids[:k], self.ident[:k], self.poses[:k], threads=threads,
                                            want_tri=False, max_contacts=1)
            oracle.heightmap_shape_collide_batch(hid, shapes, self.shape_ids[:k], self.ident[:k], self.poses[:k],
                                                 threads=threads, want_pixel=False, max_contacts=1)
        return run


class BroadphaseWorkload:
    """C5: per scene computeAABB + tree build + SelfCollision + boolean collide on every candidate.
    A step processes `scenes_per_step` scenes; the metric counts candidate pairs (narrowphase queries)."""

    kind = "scene_self_collide"
    SCENES = 8
    cpu_threads = 1  # BinaryAABB_Tree::SelfCollision + its callback is one sequential loop in the reference

    def __init__(self, name, n, dtype_name, seed):
        self.name = name
        self.n_objects = n
        self.dtype_name = dtype_name
        self.np_dtype = np.float32 if dtype_name == "f32" else np.float64
        self.sb = 4 if dtype_name == "f32" else 8
        self.scenes = [scenes.config_c5_scene(n, self.np_dtype, seed=5000 + 8 * seed + k) for k in range(self.SCENES)]
        self.n = 0  # candidate pairs per step, known after the first step
        self.cand = [0] * self.SCENES
        self.hits = [0] * self.SCENES

    def h2d_bytes(self):
        return self.SCENES * self.n_objects * (4 + 12 * self.sb)

    def d2h_bytes(self):
        return self.SCENES * 16

    def algorithmic_bytes_per_query(self):
        """per candidate pair: 16 B of ids out of the broadphase + 2 gathered poses (24 S) + 8 B pair record
        written and read again by the narrowphase + 4 B count (N_pairs * 8 B ids + C1/C2 rows)"""
        return 16 + 2 * (24 * self.sb + 8) + 4

    def setup(self, fclb, torch, dev):
        self.fclb = fclb
        self.torch = torch
        self.st = fclb.F32 if self.dtype_name == "f32" else fclb.F64
        self.table = fclb.shapes_upload(self.scenes[0][0])
        self.req = fclb.make_request(max_contacts=1)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        self.h_ids = [pin(sc[1].view(np.int32)) for sc in self.scenes]
        self.h_pose = [pin(sc[2]) for sc in self.scenes]
        self.d_ids = [t.to(dev) for t in self.h_ids]
        self.d_pose = [t.to(dev) for t in self.h_pose]

    def _run(self, host):
        f = self.fclb
        ids, pose = (self.h_ids, self.h_pose) if host else (self.d_ids, self.d_pose)
        for k in range(self.SCENES):
            self.cand[k], self.hits[k] = f.scene_self_collide(self.table, ids[k], pose[k], self.n_objects, self.st, self.req,
                                                              host=host)
        self.n = sum(self.cand)

    def step_dev(self):
        self._run(False)

    def step_host(self):
        self._run(True)

    def roof_note(self):
        return ("a step is 8 scenes x (computeAABB, Morton sort, hierarchy, refit, pair search, gather, bucketed boolean "
                "collide); 'achieved' is candidate pairs x the bytes the narrowphase stage moves per pair over the whole "
                "step time, against the HBM peak; per-stage times are in profiles/")

    def extra(self):
        return {"objects_per_scene": self.n_objects, "scenes_per_step": self.SCENES,
                "candidate_pairs_per_scene": float(np.mean(self.cand)), "colliding_pairs_per_scene": float(np.mean(self.hits)),
                "objects_per_sec_broadphase_plus_narrowphase": None}

    def cpu_sample(self):
        return self.cand[0] if self.cand[0] else self.n_objects * 4

    def cpu_run(self, oracle, threads):
        shapes, ids, poses = self.scenes[0]

        def run():
            hits, cand = oracle.scene_self_collide(shapes, ids, poses)
            self.cand[0] = self.cand[0] or cand
        return run


def make_workload(name, n, dtype_name, seed):
    if name == "c3":
        return MeshWorkload(name, n, dtype_name, seed)
    if name == "c4":
        return ArmSceneWorkload(name, n, dtype_name, seed)
    if name == "c5":
        return BroadphaseWorkload(name, n, dtype_name, seed)
    return Workload(name, n, dtype_name, seed)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation on the host cores."""
    if rank != 0:
        return
    wl = make_workload(args.workload, args.queries, args.dtype, seed=0)
    oracle = load_oracle()
    threads = getattr(wl, "cpu_threads", os.cpu_count() or 1)
    fn = wl.cpu_run(oracle, threads)
    for _ in range(args.warmup):
        fn()
    m = wl.cpu_sample() if hasattr(wl, "cpu_sample") else wl.n  # (C5 learns its candidate count from a run)
    t = time.perf_counter()
    for _ in range(args.steps):
        fn()
    el = time.perf_counter() - t
    v = m * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "queries_per_step": wl.n, "note": "host CPU only; GPUs idle"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": oracle.kind,
                         "sample": f"{m} of the step's {wl.n} queries per step, {args.steps} timed steps"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


def bind_to_gpu_numa_node(torch, local_rank):
    """One process per GPU: run on (and therefore first-touch the pinned staging buffers of) the CPU socket the
    GPU hangs off, so that the host side of the e2e copies does not cross the socket interconnect.  Best effort."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


_OUT = sys.stdout


def claim_stdout():
    """Only the JSON line may reach stdout: anything else written to fd 1 (NCCL prints its version banner there when
    the box sets NCCL_DEBUG) is sent to stderr for the rest of the run."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step (default: the config's size)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3)
    if args.queries <= 0:
        args.queries = {"c2": 10_000_000, "c4": 100_000, "c5": 100_000}.get(args.workload, 1_000_000)

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import fclb200 as fclb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    fclb.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # each rank owns its own shard of the job: independent queries, replicated geometry
    wl = make_workload(args.workload, args.queries, args.dtype, seed=rank)
    wl.setup(fclb, torch, dev)
    n = wl.n
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(fclb.stream_ptr(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        launches0 = fclb.launch_count()
        per_launch = {}
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
            for (t1_, t2_, cnt, ms) in fclb.last_launches():
                per_launch.setdefault((t1_, t2_, cnt), []).append(ms)
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.time()
        barrier()
        ms_total = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return ms_total, fclb.launch_count() - launches0, per_launch, (t0, t1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches, per_launch, win = timed(wl.step_dev, args.steps, args.warmup)
    clocks = sampler.stop(*win) if rank == 0 else None
    e2e_steps = max(3, args.steps // 2)
    ms_e2e, _, _, _ = timed(wl.step_host, e2e_steps, 3)

    n = wl.n  # (C5 learns its candidate-pair count from the run itself)
    total_n = n * world
    if world > 1:
        tn = torch.tensor([float(n)], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        total_n = float(tn.item())
    value = total_n * args.steps / (ms_dev * 1e-3)
    e2e_value = total_n * e2e_steps / (ms_e2e * 1e-3)

    # dominant kernel of the step and its roofline
    names = {0: "box", 1: "sphere", 2: "ellipsoid", 3: "capsule", 4: "cone", 5: "cylinder", 6: "convex", 7: "triangle"}
    kern = []
    for (t1_, t2_, cnt), v in per_launch.items():
        label = f"{wl.kind}[{names.get(t1_, '?')}-{names.get(t2_, '?')}]" if t1_ >= 0 else wl.kind
        kern.append({"kernel": label, "queries": cnt, "avg_ms": float(np.mean(v))})
    if hasattr(wl, "kernel_records"):
        kern = wl.kernel_records()
    kern.sort(key=lambda k: -k["avg_ms"])
    peak, peak_src = measured_peaks()
    bpq = wl.algorithmic_bytes_per_query()
    roof = None
    if kern:
        top = kern[0]
        bpq = top.get("bytes_per_query", bpq)
        achieved = top["queries"] * bpq / (top["avg_ms"] * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                rec = json.load(f).get(top["kernel"] + ":" + args.dtype)
            if rec:  # one ncu capture, scaled linearly to this launch's query count
                traffic = rec["bytes"] * top["queries"] / rec["queries"]
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": top["kernel"], "peak_source": peak_src,
                "algorithmic_bytes_per_query": bpq, "queries_per_launch": top["queries"],
                "avg_launch_ms": top["avg_ms"],
                "note": wl.roof_note() if hasattr(wl, "roof_note") else
                "iterative GJK/EPA buckets are FP32/FP64-issue and latency bound, not HBM bound (DESIGN.md 4.3); "
                "closed-form buckets are the HBM-bound kernels; per-bucket figures under 'kernels'"}
        for k in kern:
            k["hbm_gbs"] = k["queries"] * k.get("bytes_per_query", bpq) / (k["avg_ms"] * 1e-3) / 1e9
            k["hbm_frac"] = k["hbm_gbs"] / peak

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "queries_per_gpu_per_step": n,
                   "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no flush needed" % (wl.h2d_bytes() / 1e6)
                   if wl.h2d_bytes() > 200e6 else "inputs %.0f MB/step: smaller than L2 on purpose of the config; "
                   "each step re-reads them after %.0f MB of result writes" % (wl.h2d_bytes() / 1e6, wl.d2h_bytes() / 1e6),
                   "sharding": "queries sharded by rank, geometry replicated, no collective on the data path"
                   + ("" if numa is None else "; rank processes bound to their GPU's NUMA node")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes(),
                "d2h_bytes_per_step": wl.d2h_bytes(), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                # the host link is what bounds this number once copies and kernels overlap: bytes moved per
                # second in each direction (PCIe is full duplex; Gen5 x16 peaks near 55-57 GB/s one way)
                "h2d_gbs": wl.h2d_bytes() / (ms_e2e / e2e_steps) / 1e6, "d2h_gbs": wl.d2h_bytes() / (ms_e2e / e2e_steps) / 1e6},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels": kern,
    }

    if hasattr(wl, "extra"):
        line["config"].update(wl.extra())
    if rank == 0:
        try:  # measured CUDA-core FMA peaks: the denominators for the compute-bound GJK / MPR / EPA kernels
            line["fp_peak_measured"] = {"fp32_tflops": fclb.measure_fp_peak(fclb.F32), "fp64_tflops": fclb.measure_fp_peak(fclb.F64),
                                        "how": "FMA chain, 16 accumulators per thread, 8 CTAs x 256 threads per SM, best of 5"}
        except Exception as ex:
            line["fp_peak_measured"] = {"error": str(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            oracle = load_oracle()
            threads = getattr(wl, "cpu_threads", os.cpu_count() or 1)
            fn = wl.cpu_run(oracle, threads)
            best = None
            for _ in range(3):
                t = time.perf_counter()
                fn()
                dt = time.perf_counter() - t
                best = dt if best is None else min(best, dt)
            m = wl.cpu_sample() if hasattr(wl, "cpu_sample") else n
            line["cpu_baseline"] = {"value": m / best, "unit": UNIT, "cores": threads, "kind": oracle.kind,
                                    "sample": f"{m} of the step's {n} queries, best of 3 ({best:.2f} s each)"}
        except Exception as ex:  # the CPU leg is a reported baseline, never the product path
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}

    if rank == 0:
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
