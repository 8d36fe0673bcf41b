"""Builds tests/cpp/test_host_api.cpp against include/fcl_b200/fcl.h + libfclb200.so with g++
and runs it: the C++ mirror of the fcl API works end to end through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_reference_file(path, ref_oracle):
    """a mesh, a box, poses and the contacts the reference reports with useDefaultPenetration(), for the C++ program"""
    import numpy as np

    import scenes

    v, t = scenes.noisy_uv_sphere(n_lat=13, n_lon=24, radius=0.5, noise=0.05)
    side = (0.3, 0.2, 0.25)
    shapes = [(scenes.BOX, 0, side)]
    nq, keep = 400, 256
    rng = np.random.Generator(np.random.PCG64(17))
    pm = scenes.random_poses(rng, nq, 0.2, np.float64)
    ps = scenes.random_poses(rng, nq, 0.6, np.float64)
    ids = np.zeros(nq, np.uint32)
    mid = ref_oracle.bvh_create(v, t)
    counts, b1, contacts = ref_oracle.scene_shape_contacts_batch(0, mid, shapes, ids, pm, ps, keep, threads=8,
                                                                 max_contacts=keep, penetration_mode=1)
    assert int(counts.max()) < keep and int(counts.sum()) > 100
    with open(path, "wb") as f:
        f.write(np.asarray([len(v), len(t), nq, keep], np.int32).tobytes())
        f.write(np.ascontiguousarray(v, np.float64).tobytes())
        f.write(np.ascontiguousarray(t, np.int32).tobytes())
        f.write(np.asarray(side, np.float64).tobytes())
        for a in (pm, ps, counts.astype(np.uint32), b1.astype(np.int64), contacts.astype(np.float64)):
            f.write(np.ascontiguousarray(a).tobytes())


def test_cpp_host_api(tmp_path, ref_oracle):
    exe = str(tmp_path / "test_host_api")
    ref_file = str(tmp_path / "mesh_box_contacts.bin")
    write_reference_file(ref_file, ref_oracle)
    lib_dir = os.path.join(ROOT, "mind-fcl_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
           "-L", lib_dir, "-lfclb200", f"-Wl,-rpath,{lib_dir}", "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe, ref_file], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "ALL OK" in r.stdout
