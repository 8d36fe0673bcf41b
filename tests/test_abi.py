"""The C-ABI library loads, exports every symbol include/fclb200.h declares, and
fails loudly (no CPU fallback) when there is no GPU.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fclb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fclb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import fclb200

    lib = fclb200.load()
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/fclb200.h but not exported: {missing}"
    assert set(fclb200.EXPORTS) <= set(names)


def test_struct_layouts_match_header():
    import fclb200

    assert C.sizeof(fclb200.Shape) == 32          # u32 type, u32 geom, double p[3]
    assert C.sizeof(fclb200.Request) == 64        # see fclb_request
    import scenes

    assert scenes.PAIR_DTYPE.itemsize == 8


def test_no_cpu_fallback():
    import fclb200

    lib = fclb200.load()
    if lib.fclb_device_count() > 0:
        return  # on a GPU box this property is covered by the gpu tests loading the kernels
    rc = lib.fclb_init(0)
    assert rc == 1, "fclb_init must return FCLB_ERR_NO_DEVICE without a GPU"
    assert b"no CPU fallback" in lib.fclb_last_error()
    # compute entry points refuse too
    shapes = fclb200.shape_array([(0, 0, (1.0, 1.0, 1.0))])
    h = C.c_uint64()
    assert lib.fclb_shapes_upload(C.cast(shapes, C.c_void_p), 1, C.byref(h)) == 1
    pairs = np.zeros((1, 2), np.uint32)
    poses = np.zeros((1, 12), np.float32)
    out = np.zeros(1, np.float32)
    rc = lib.fclb_distance_batch_host(0, fclb200._ptr(pairs), fclb200._ptr(poses), fclb200._ptr(poses), 1, 0, 0.0, 0,
                                      fclb200._ptr(out), None, None, None)
    assert rc == 1


def test_version_string():
    import fclb200

    assert b"sm_100a" in fclb200.load().fclb_version()
