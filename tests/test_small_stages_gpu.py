"""The chunked host pipelines (*_host entry points) with very short stages: a 1k-query first stage doubling up to 4k / 16k
stages, a tail tapering to 4k, and the scene-vs-shape entry points staged as well.  The stage sizes are read when the engine
starts, so the parity tests of those entry points are re-run in a child process with the switches set: same inputs, same
oracle comparisons, every batch cut into tens of ragged stages."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

FILES = ["test_distance_gpu.py", "test_collide_gpu.py", "test_bvh_gpu.py", "test_mesh_shape_gpu.py", "test_heightmap_gpu.py",
         "test_octree_gpu.py"]


def test_host_paths_with_short_stages():
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, FCLB_HOST_HEAD="1024", FCLB_HOST_CHUNK="16384", FCLB_HOST_TAPER="4096", FCLB_SCENE_HOST_STAGED="1")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "not stage_schedule", "-p", "no:cacheprovider"]
    p = subprocess.run(cmd + [os.path.join(here, f) for f in FILES], env=env, capture_output=True, text=True, timeout=1200,
                       cwd=os.path.dirname(here))
    tail = "\n".join(p.stdout.splitlines()[-15:])
    assert p.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
