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
        self.qt1 = self.qt2 = None
        if name == "c2":
            # poses are generated as unit quaternion + translation (FCLB_POSE_QT7, 28 B in float) and expanded with Eigen's
            # toRotationMatrix arithmetic: the 12-S arrays are what the device-resident path, the 12-S host path and the
            # CPU reference read; the QT7 arrays are what the headline host path sends over PCIe
            self.shapes, self.pairs, self.qt1, self.qt2 = scenes.config_c2_qt(n, self.np_dtype, seed=2001 + seed)
            self.poses1, self.poses2 = scenes.expand_qt7(self.qt1), scenes.expand_qt7(self.qt2)
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
        if self.qt1 is not None:
            return self.n * (14 * self.sb + 8)
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
        if self.qt1 is not None:
            self.h_q1, self.h_q2 = pin(self.qt1), pin(self.qt2)

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
            if host and self.qt1 is not None and not getattr(self, "force_pose12", False):
                fn, p1, p2 = lib.fclb_distance_batch_qt_host, self.h_q1, self.h_q2
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

    # -- roofline ------------------------------------------------------------------
    def shape_types(self):
        shapes = self.shapes if self.shapes is not None else [(scenes.CONVEX, 0, ()), (scenes.CONVEX, 1, ())]
        return np.asarray([s[0] for s in shapes], np.int64)

    def bucket_mask(self, t1, t2):
        ty = self.shape_types()
        return (ty[self.pairs["shape1"]] == t1) & (ty[self.pairs["shape2"]] == t2)

    def bound_of(self, top):
        """closed-form buckets move bytes; the iterative GJK / EPA buckets issue flops"""
        t1, t2 = top.get("t1", -1), top.get("t2", -1)
        closed_distance = {(1, 0), (0, 1), (1, 3), (3, 1), (1, 5), (5, 1), (1, 1), (3, 3)}
        closed_collide = closed_distance - {(3, 3)} | {(0, 0)}
        if self.kind == "distance" and (t1, t2) in closed_distance:
            return "hbm"
        if self.kind == "collide" and (t1, t2) in closed_collide:
            return "hbm"
        return "fp32" if self.dtype_name == "f32" else "fp64"

    def teardown(self):
        self.fclb.release(self.table)
        self.d_out = self.h_out = self.d_pairs = self.d_p1 = self.d_p2 = self.h_pairs = self.h_p1 = self.h_p2 = None
        self.h_q1 = self.h_q2 = None


class MeshWorkload:
    """C3: fcl::collide(BVHModel<OBBRSS>, BVHModel<OBBRSS>) per relative pose, boolean (max_contacts=1)."""

    kind = "bvh_collide"
    bound = "l2"

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
        self.kernel_ms = f.last_kernel_ms()  # (read here: the host-buffer call that follows in the bench launches per chunk)

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

    def kernel_records(self):
        return [{"kernel": "bvh_collide", "queries": self.n, "avg_ms": self.kernel_ms,
                 "bytes_per_query": self.algorithmic_bytes_per_query()}]

    def teardown(self):
        for h in self.handles:
            self.fclb.bvh_release(h)


class ArmSceneWorkload:
    """C4: n configurations x 7 convex links, each link tested against the scene mesh (mesh-shape
    traversal) and against the heightmap (heightmap-shape scan): 14 n narrowphase queries per step."""

    kind = "scene_collide"
    bound = "l2"

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

    def teardown(self):
        self.fclb.bvh_release(self.bvh)
        self.fclb.heightmap_release(self.hm)
        self.fclb.release(self.table)


class BroadphaseWorkload:
    """C5: per scene computeAABB + tree build + SelfCollision + boolean collide on every candidate.
    A step processes `scenes_per_step` scenes; the metric counts candidate pairs (narrowphase queries)."""

    kind = "scene_self_collide"
    bound = "hbm"
    SCENES = 8

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
        return getattr(self, "cpu_cand", 0) or self.n_objects * 4 * self.SCENES

    def cpu_run(self, oracle, threads):
        """BinaryAABB_Tree::SelfCollision + its fcl::collide callback is one sequential loop per scene in the reference;
        the CPU arm runs the step's scenes in parallel, one host thread per scene."""
        self.cpu_threads_used = min(threads, self.SCENES)

        def run():
            out = [0] * self.SCENES

            def one(k):
                shapes, ids, poses = self.scenes[k]
                out[k] = oracle.scene_self_collide(shapes, ids, poses)[1]
            ts = [threading.Thread(target=one, args=(k,)) for k in range(self.SCENES)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            self.cpu_cand = sum(out)
        return run

    def kernel_records(self):
        ms = getattr(self, "measured_step_ms", 0.0)
        return [{"kernel": "scene_self_collide (8 scenes: aabb, sort, hierarchy, pairs, gather, narrowphase)", "queries": self.n,
                 "avg_ms": ms, "bytes_per_query": self.algorithmic_bytes_per_query()}]

    def teardown(self):
        self.fclb.release(self.table)


def make_workload(name, n, dtype_name, seed):
    if name == "c3":
        return MeshWorkload(name, n, dtype_name, seed)
    if name == "c4":
        return ArmSceneWorkload(name, n, dtype_name, seed)
    if name == "c5":
        return BroadphaseWorkload(name, n, dtype_name, seed)
    return Workload(name, n, dtype_name, seed)


DEFAULT_QUERIES = {"c2": 10_000_000, "c4": 100_000, "c5": 100_000}


def default_queries(name):
    return DEFAULT_QUERIES.get(name, 1_000_000)


# ---- SURVEY.md 8(d) flop / query model for the GJK / EPA buckets -----------------------------------------------
# flop/query = n_sup * F_sup + n_gjk * F_proj + n_epa * (F_epa0 + F_scan * (V + E + F)), all counts taken from the CPU
# oracles on the bench seed (deterministic properties of the input), constants as SURVEY.md fixes them.
F_PROJ = {1: 0, 2: 40, 3: 120, 4: 270}
F_EPA0, F_SCAN, F_DOT = 400, 2, 5
SHAPE_NAMES = {0: "box", 1: "sphere", 2: "ellipsoid", 3: "capsule", 4: "cone", 5: "cylinder", 6: "convex", 7: "triangle"}


def f_sup(t1, t2):
    if t1 == 6 or t2 == 6:
        return 33  # the two transforms; the Convex dot products are counted separately (5 flop each)
    if t1 == 0 and t2 == 0:
        return 48
    return 60


def flop_model(wl, t1, t2, sel, want_distance, with_epa, sample=200_000):
    """flop / query of one (type1, type2) bucket of a shape-pair workload, from the oracles' counters on the first
    `sample` queries of the bucket.  Returns None when no oracle is built."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ctypes as C
        import oracle_py
        port = oracle_py.PortOracle() if oracle_py.have_port() else None
        ref = oracle_py.RefOracle() if oracle_py.have_ref() else None
    except Exception:
        return None
    if port is None:
        return None
    idx = np.nonzero(sel)[0][:sample]
    if idx.size == 0:
        return None
    shapes = wl.shapes
    if wl.convex is not None:
        slots = [port.register_convex(*m) for m in wl.convex]
        shapes = [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())]
    pairs = np.ascontiguousarray(wl.pairs[idx])
    p1, p2 = np.ascontiguousarray(wl.poses1[idx]), np.ascontiguousarray(wl.poses2[idx])
    out = np.zeros(16, np.uint64)
    arr = oracle_py._shape_array(shapes)
    f = port.fn("gjk_work_counters")
    f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_uint32,
                  C.c_int, C.c_void_p, C.c_int]
    tol = 0.0 if want_distance else 1e-6  # shapeDistance: eps^(7/8); the collide / gjk_epa path: 1e-6
    f(oracle_py._st(p1.dtype), C.cast(arr, C.c_void_p), len(shapes), oracle_py._p(pairs), oracle_py._p(p1), oracle_py._p(p2),
      len(idx), tol, 0, 1 if want_distance else 0, oracle_py._p(out), os.cpu_count() or 1)
    m = float(out[0])
    sv, ex, dots = float(out[1]), float(out[2]), float(out[3])
    proj = {r: float(out[4 + r]) for r in range(5)}
    upd = {r: float(out[9 + r]) for r in range(5)}
    fs = f_sup(t1, t2)
    flops = sv * fs + ex * fs / 2 + dots * F_DOT
    flops += sum(proj[r] * F_PROJ.get(r, 0) for r in proj) + sum(upd[r] * F_PROJ.get(r, 0) for r in upd)
    model = {"sample_queries": int(m), "n_sup_per_query": (sv + ex / 2) / m, "n_gjk_per_query": sum(proj.values()) / m,
             "n_dist_per_query": sum(upd.values()) / m, "convex_dots_per_query": dots / m, "F_sup": fs,
             "F_proj_rank2/3/4": [40, 120, 270], "gjk_flop_per_query": flops / m}
    if with_epa and ref is not None:
        rshapes = wl.shapes
        if wl.convex is not None:
            slots = [ref.register_convex(*mm) for mm in wl.convex]
            rshapes = [(scenes.CONVEX, slots[0], ()), (scenes.CONVEX, slots[1], ())]
        _, epa, _, _, iters = ref.gjk_epa_batch(rshapes, pairs, p1, p2, threads=os.cpu_count() or 1)
        k = iters[:, 1].astype(np.float64) / 2.0  # EPA supportVertex evaluations = expansions (+ the initial polytope's)
        dots_per_sup = dots / max(sv, 1.0)
        # a closed triangulated polytope after i expansions: V = 4 + i, F = 4 + 2 i, E = 6 + 3 i  (Euler), so the
        # nearest-feature scans of k iterations touch sum_i (14 + 6 i) = 14 k + 3 k (k - 1) elements
        scan = 14.0 * k + 3.0 * k * (k - 1.0)
        epa_flops = k * (F_EPA0 + fs + dots_per_sup * F_DOT) + F_SCAN * scan
        model.update({"n_epa_per_query": float(k.mean()), "epa_flop_per_query": float(epa_flops.mean()),
                      "mean_V+E+F_scanned_per_iteration": float(scan.sum() / max(k.sum(), 1.0)),
                      "F_epa0": F_EPA0, "F_scan": F_SCAN, "intersecting_fraction": float((epa >= 0).mean())})
        flops += float(epa_flops.sum())
    model["flop_per_query"] = flops / m
    return model


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
    m = wl.cpu_sample() if hasattr(wl, "cpu_sample") else wl.n
    v = m * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "queries_per_step": wl.n, "queries_per_gpu_per_step": wl.n,
                   "note": "host CPU only; GPUs idle"},
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


class Ctx:
    """Everything a measurement needs: torch, the process group, the engine's stream, measured peaks."""

    def __init__(self, torch, dist, fclb, dev, rank, world, local_rank):
        self.torch, self.dist, self.fclb, self.dev = torch, dist, fclb, dev
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.stream = torch.cuda.ExternalStream(fclb.stream_ptr(), device=dev)
        self.hbm_peak, self.hbm_src = measured_peaks()
        self.fp_peak = {}
        self.l2_peak = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def peaks(self):
        """FP32 / FP64 FMA-chain and L2 read-bandwidth microbenchmarks, run once on this device."""
        if not self.fp_peak:
            f = self.fclb
            try:
                self.fp_peak = {"f32": f.measure_fp_peak(f.F32), "f64": f.measure_fp_peak(f.F64)}
                self.l2_peak = f.measure_l2_bandwidth()
            except Exception as ex:
                self.fp_peak = {"error": str(ex)}
        return self.fp_peak, self.l2_peak

    def timed(self, fn, steps, warmup):
        torch, fclb = self.torch, self.fclb
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        launches0 = fclb.launch_count()
        per_launch = {}
        t0 = time.time()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
            for (t1_, t2_, cnt, ms) in fclb.last_launches():
                per_launch.setdefault((t1_, t2_, cnt), []).append(ms)
        e1.record(self.stream)
        torch.cuda.synchronize()
        t1 = time.time()
        self.barrier()
        ms_total = self.max_over_ranks(e0.elapsed_time(e1))
        return ms_total, fclb.launch_count() - launches0, per_launch, (t0, t1)


def build_roofline(ctx, wl, args_dtype, kern, with_model):
    """Roofline of the dominant kernel against the ceiling that binds it (hbm / l2 / fp32 / fp64), with the per-bucket
    list inside.  `kern`: [{kernel, queries, avg_ms[, bytes_per_query][, t1, t2]}], sorted by time."""
    if not kern:
        return None
    fp_peak, l2_peak = ctx.peaks()
    top = kern[0]
    bpq = top.get("bytes_per_query", wl.algorithmic_bytes_per_query())
    sec = top["avg_ms"] * 1e-3
    hbm_gbs = top["queries"] * bpq / sec / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            rec = json.load(f).get(top["kernel"] + ":" + args_dtype)
        if rec:  # one ncu --set full capture of this kernel, scaled linearly to this launch's query count
            traffic = rec["bytes"] * top["queries"] / rec["queries"]
    bound = getattr(wl, "bound", None) or wl.bound_of(top)
    roof = {"kernel": top["kernel"], "queries_per_launch": top["queries"], "avg_launch_ms": top["avg_ms"], "traffic": traffic}
    hbm_view = {"achieved": hbm_gbs, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": hbm_gbs / ctx.hbm_peak,
                "algorithmic_bytes_per_query": bpq, "peak_source": ctx.hbm_src}
    if bound in ("fp32", "fp64") and with_model and "t1" in top:
        model = flop_model(wl, top["t1"], top["t2"], wl.bucket_mask(top["t1"], top["t2"]), wl.kind == "distance",
                           wl.kind == "gjk_epa")
        peak = fp_peak.get("f32" if bound == "fp32" else "f64") if isinstance(fp_peak, dict) else None
        if model and peak:
            ach = top["queries"] * model["flop_per_query"] / sec / 1e12
            roof.update({"bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "peak_source": "measured here: FMA chain, 16 accumulators per thread (an FMA counted as 2 flop; the "
                                        "product is compiled --fmad=false for parity with the reference, which halves what "
                                        "a mul+add stream can reach)",
                         "flop_model": model, "hbm_view": hbm_view})
    if "bound" not in roof and bound == "l2" and l2_peak:
        roof.update({"bound": "l2", "achieved": hbm_gbs, "peak": l2_peak, "unit": "GB/s", "frac": hbm_gbs / l2_peak,
                     "peak_source": "measured here: 32 MB L2-resident buffer streamed by every SM with 128-bit loads",
                     "algorithmic_bytes_per_query": bpq, "hbm_view": hbm_view})
    if "bound" not in roof:
        roof.update({"bound": "hbm", **hbm_view})
    roof["note"] = wl.roof_note() if hasattr(wl, "roof_note") else (
        "closed-form buckets are HBM-bound (bytes = 2 poses + result record); the iterative GJK / EPA buckets are bound by "
        "FP32 / FP64 issue and SIMT divergence: their fraction is the SURVEY.md 8(d) flop model over the measured FMA peak")
    for k in kern:
        b = k.get("bytes_per_query", bpq)
        k["hbm_gbs"] = k["queries"] * b / (k["avg_ms"] * 1e-3) / 1e9
        k["hbm_frac"] = k["hbm_gbs"] / ctx.hbm_peak
        k.pop("t1", None)
        k.pop("t2", None)
    roof["kernels"] = kern
    return roof


def measure(ctx, name, queries, dtype, steps, warmup, seed, with_cpu, with_model, clock_sampler=None):
    """One workload, the full record: device-resident value, e2e through the host-buffer C ABI, roofline, CPU baseline."""
    torch, fclb = ctx.torch, ctx.fclb
    wl = make_workload(name, queries, dtype, seed=seed)
    wl.setup(fclb, torch, ctx.dev)
    torch.cuda.synchronize()
    if clock_sampler:
        clock_sampler.start()
    ms_dev, launches, per_launch, win = ctx.timed(wl.step_dev, steps, warmup)
    clocks = clock_sampler.stop(*win) if clock_sampler else None
    e2e_steps = max(3, steps // 2)
    ms_e2e, _, _, _ = ctx.timed(wl.step_host, e2e_steps, 3)
    ms_e2e12 = None
    if getattr(wl, "qt1", None) is not None:  # the same host call with 12-S poses, for comparison
        wl.force_pose12 = True
        ms_e2e12, _, _, _ = ctx.timed(wl.step_host, e2e_steps, 2)
        wl.force_pose12 = False
    n = wl.n  # (C5 learns its candidate-pair count from the run itself)
    total_n = ctx.sum_over_ranks(n)
    value = total_n * steps / (ms_dev * 1e-3)
    e2e_value = total_n * e2e_steps / (ms_e2e * 1e-3)

    kern = []
    for (t1_, t2_, cnt), v in per_launch.items():
        label = f"{wl.kind}[{SHAPE_NAMES.get(t1_, '?')}-{SHAPE_NAMES.get(t2_, '?')}]" if t1_ >= 0 else wl.kind
        kern.append({"kernel": label, "queries": cnt, "avg_ms": float(np.mean(v)), "t1": t1_, "t2": t2_})
    wl.measured_step_ms = ms_dev / steps
    if hasattr(wl, "kernel_records"):
        kern = wl.kernel_records()
    kern.sort(key=lambda k: -k["avg_ms"])
    roof = build_roofline(ctx, wl, dtype, kern, with_model and ctx.rank == 0)
    rec = {
        "workload": name, "value": value, "unit": UNIT, "ms_per_step": ms_dev / steps, "steps": steps, "dtype": dtype,
        "config": {"workload": WORKLOADS[name], "queries_per_gpu_per_step": n, "queries_per_step": n,
                   "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no flush needed" % (wl.h2d_bytes() / 1e6)
                   if wl.h2d_bytes() > 200e6 else "inputs %.0f MB/step: smaller than L2 on purpose of the config; "
                   "each step re-reads them after %.0f MB of result writes" % (wl.h2d_bytes() / 1e6, wl.d2h_bytes() / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes(),
                "d2h_bytes_per_step": wl.d2h_bytes(), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                # the host link is what bounds this number once copies and kernels overlap: bytes moved per
                # second in each direction (PCIe is full duplex; Gen5 x16 peaks near 55-57 GB/s one way)
                "h2d_gbs": wl.h2d_bytes() / (ms_e2e / e2e_steps) / 1e6, "d2h_gbs": wl.d2h_bytes() / (ms_e2e / e2e_steps) / 1e6},
        "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
    }
    if ms_e2e12 is not None:
        rec["e2e"]["pose_encoding"] = ("FCLB_POSE_QT7: unit quaternion + translation, 7 S per pose, expanded on the device with "
                                       "Eigen's toRotationMatrix arithmetic (results bit-identical to the 12-S call)")
        rec["e2e"]["with_12S_poses"] = {"value": total_n * e2e_steps / (ms_e2e12 * 1e-3), "ms_per_step": ms_e2e12 / e2e_steps,
                                        "h2d_bytes_per_step": n * (24 * wl.sb + 8)}
    if hasattr(wl, "extra"):
        rec["config"].update(wl.extra())
    if with_cpu and ctx.rank == 0:
        try:
            oracle = load_oracle()
            threads = getattr(wl, "cpu_threads", os.cpu_count() or 1)
            fn = wl.cpu_run(oracle, threads)
            best = None
            for _ in range(3 if name == "c2" else 2):
                t = time.perf_counter()
                fn()
                dt = time.perf_counter() - t
                best = dt if best is None else min(best, dt)
            m = wl.cpu_sample() if hasattr(wl, "cpu_sample") else n
            rec["cpu_baseline"] = {"value": m / best, "unit": UNIT, "cores": threads, "kind": oracle.kind,
                                   "sample": f"{m} of the step's {n} queries, best of {3 if name == 'c2' else 2} ({best:.2f} s each)"}
        except Exception as ex:  # the CPU leg is a reported baseline, never the product path
            rec["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
    wl.teardown()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step (default: the config's size)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU gets the config's batch; strong: ONE batch of the config's size sharded by query index")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the per-config records of the other workloads")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3)
    explicit_queries = args.queries > 0
    if args.queries <= 0:
        args.queries = default_queries(args.workload)

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
    ctx = Ctx(torch, dist, fclb, dev, rank, world, local_rank)

    # each rank owns its own shard of the job: independent queries, replicated geometry, no collective on the data path
    per_gpu = args.queries
    if args.scaling == "strong":
        per_gpu = shard_size(args.queries, rank, world)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    head = measure(ctx, args.workload, per_gpu, args.dtype, args.steps, args.warmup, seed=rank,
                   with_cpu=(world == 1 and not args.no_cpu_baseline), with_model=True, clock_sampler=sampler)
    fp_peak, l2_peak = ctx.peaks()
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": head["config"], "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"],
        "kernels": (head["roofline"] or {}).get("kernels"),
    }
    line["config"]["sharding"] = ("queries sharded by rank, geometry replicated, no collective on the data path"
                                  + ("" if numa is None else "; rank processes bound to their GPU's NUMA node"))
    if "cpu_baseline" in head:
        line["cpu_baseline"] = head["cpu_baseline"]
    if rank == 0:
        line["fp_peak_measured"] = {"fp32_tflops": (fp_peak or {}).get("f32"), "fp64_tflops": (fp_peak or {}).get("f64"),
                                    "l2_read_gbs": l2_peak,
                                    "how": "FMA chain, 16 accumulators per thread, 8 CTAs x 256 threads per SM, best of 5; "
                                           "L2: 32 MB resident buffer, 128-bit loads, all SMs"}
    # the same job as ONE batch of the config's size sharded by query index over the N GPUs (strong scaling): the
    # north star's "query batch sharded across the GPUs of one box"
    if world > 1 and args.scaling == "weak" and not explicit_queries:
        st = measure(ctx, args.workload, shard_size(args.queries, rank, world), args.dtype, max(3, args.steps // 2), 3,
                     seed=rank, with_cpu=False, with_model=False)
        line["strong_scaling"] = {"total_queries_per_step": args.queries, "value": st["value"], "ms_per_step": st["ms_per_step"],
                                  "e2e": st["e2e"], "note": "one batch of the config's size, sharded by contiguous query range"}
    # every other config of BASELINE.json at its full size, one record each (N = 1 only: keeps the scaling runs short)
    if world == 1 and not args.no_workloads and args.workload == "c2" and not explicit_queries:
        others = []
        for name in ("c1a", "c1b", "c1b_convex", "c3", "c4", "c5"):
            try:
                r = measure(ctx, name, default_queries(name), args.dtype, min(args.steps, 5), 3, seed=0,
                            with_cpu=not args.no_cpu_baseline, with_model=True)
                r.pop("clocks", None)
                others.append(r)
            except Exception as ex:
                others.append({"workload": name, "error": f"{type(ex).__name__}: {ex}"})
        line["workloads"] = others

    if rank == 0:
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def shard_size(total, rank, world):
    """contiguous query-index ranges, remainder to the first ranks (mind-fcl_b200/sharding.py)"""
    base, rem = divmod(total, world)
    return base + (1 if rank < rem else 0)


if __name__ == "__main__":
    main()
