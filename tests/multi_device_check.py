"""Single-process multi-GPU check (run as a script by tests/test_multi_device_gpu.py, or by hand on a multi-GPU box):
fclb_init_devices(N) then the *_host entry points, whose batches are sharded by query index over the N engines, against
the reference oracle -- the results must not depend on how many devices served the batch."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("mind-fcl_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import fclb200 as fclb  # noqa: E402
import oracle_py  # noqa: E402
import scenes  # noqa: E402


def main():
    want = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n_dev = fclb.init_devices(want)
    print(f"engines: {n_dev}")
    ref = oracle_py.RefOracle()
    for dtype, st in ((np.float32, fclb.F32), (np.float64, fclb.F64)):
        # C2 distance, a batch that does not divide evenly
        n = 1_000_003
        shapes, pairs, p1, p2 = scenes.config_c2(n, dtype)
        table = fclb.shapes_upload(shapes)
        t = time.perf_counter()
        r = fclb.distance_batch_host(table, pairs, p1, p2, st)
        dt = time.perf_counter() - t
        e = ref.distance_batch(shapes, pairs, p1, p2, threads=16)
        assert np.array_equal(r.ok != 0, e[3] != 0)
        sep = e[3] != 0
        assert np.array_equal(r.dist[sep], e[0][sep]) and np.array_equal(r.p1[sep], e[1][sep])
        print(f"distance {np.dtype(dtype).name}: {n} queries over {n_dev} devices in {dt * 1e3:.1f} ms, bit-identical")
        # collide with contacts
        shapes_b, pairs_b, q1, q2 = scenes.config_c1_boxes(200_001, dtype)
        tb = fclb.shapes_upload(shapes_b)
        kw = dict(max_contacts=4, penetration_mode=1)
        counts, contacts = fclb.collide_batch_host(tb, pairs_b, q1, q2, st, fclb.make_request(**kw), max_keep=4)
        e_counts, e_contacts = ref.collide_batch(shapes_b, pairs_b, q1, q2, max_keep=4, threads=16, **kw)
        assert np.array_equal(counts, e_counts)
        m = np.arange(4)[None, :] < e_counts[:, None]
        assert np.array_equal(contacts[m][:, 2:], e_contacts[m][:, 2:])
        print(f"collide {np.dtype(dtype).name}: counts and contact records identical")
        # mesh-shape + heightmap-shape (replicated scene geometry), and DefaultGJK_EPA contacts of a mesh
        v, tri = scenes.noisy_uv_sphere(n_lat=13, n_lon=24)
        bvh = fclb.bvh_build(v, tri, st)
        mid = ref.bvh_create(v, tri)
        sh = [(scenes.BOX, 0, (0.3, 0.2, 0.25)), (scenes.CAPSULE, 0, (0.08, 0.3))]
        ts = fclb.shapes_upload(sh)
        rng = np.random.Generator(np.random.PCG64(3))
        nq = 20_001
        pm, ps = scenes.random_poses(rng, nq, 0.4, dtype), scenes.random_poses(rng, nq, 0.9, dtype)
        ids = (np.arange(nq) % 2).astype(np.uint32)
        c, _ = fclb.bvh_shape_collide_batch_host(bvh, ts, ids, pm, ps, st, fclb.make_request(max_contacts=2**31 - 1))
        ec, _ = ref.mesh_shape_collide_batch(mid, sh, ids, pm, ps, threads=16, max_contacts=2**31 - 1)
        assert np.array_equal(c, ec)
        req = fclb.make_request(max_contacts=2**31 - 1, penetration_mode=1)
        c2, b1, ct = fclb.scene_shape_contacts_batch_host(fclb.SCENE_BVH, bvh, ts, ids[:2001], pm[:2001], ps[:2001], st, req, 64)
        e2, eb1, ect = ref.scene_shape_contacts_batch(0, mid, sh, ids[:2001], pm[:2001], ps[:2001], 512, threads=16,
                                                      max_contacts=2**31 - 1, penetration_mode=1)
        assert np.array_equal(c2, e2)
        print(f"mesh-shape {np.dtype(dtype).name}: counts identical ({int(ec.sum())} contacts); DefaultGJK_EPA counts identical")
        fclb.bvh_release(bvh)
        for h in (table, tb, ts):
            fclb.release(h)
    print("MULTI_DEVICE_OK")


if __name__ == "__main__":
    main()
