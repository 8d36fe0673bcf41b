"""One process, several GPUs: fclb_init_devices + sharded *_host calls.  Needs >= 2 GPUs; the check
runs in its own process because the engine set must be chosen before the first upload."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_process_multi_gpu(fclb):
    n = fclb.load().fclb_device_count()
    script = os.path.join(ROOT, "tests", "multi_device_check.py")
    # with one GPU the same script still exercises the fclb_init_devices(1) path
    r = subprocess.run([sys.executable, script, str(min(n, 8))], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "MULTI_DEVICE_OK" in r.stdout
