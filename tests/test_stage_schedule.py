"""Host logic of the chunked *_host pipelines (mind-fcl_b200/csrc/fclb_stages.h): the stage schedule tiles the batch exactly,
starts short for compute-bound calls (head), ends short for copy-bound calls (taper), and degenerates to equal stages /
one stage.  CPU only: the header is plain C++ and is compiled here with g++."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mirror(n, chunk, head, taper):
    """Python restatement of fclb::stageSizes."""
    chunk = max(chunk, 1)
    ramp = min(head, chunk) if head else chunk
    out, b = [], 0
    while b < n:
        rem, m = n - b, ramp
        if ramp < chunk:
            ramp = min(chunk, 2 * ramp)
            if rem < m + m // 2:
                m = rem
        elif taper and rem < 2 * chunk:
            m = max(taper, (rem // 2 + 4095) // 4096 * 4096)
            if rem < m + taper:
                m = rem
        m = min(m, rem)
        out.append(m)
        b += m
    return out


CASES = [
    (10_000_000, 1 << 21, 0, 1 << 19),   # C2 through fclb_distance_batch_qt_host (the defaults)
    (10_000_000, 1 << 21, 0, 0),         # equal stages
    (1_000_000, 1 << 19, 1 << 16, 0),    # C3 through fclb_bvh_collide_batch_host (the defaults)
    (1_000_000, 1 << 18, 1 << 16, 0),    # C1a through fclb_collide_batch_host
    (700_000, 700_000, 0, 0),            # one stage (scene-vs-shape calls, GJK + EPA collide batches)
    (200_001, 32768, 0, 4096), (50_001, 16384, 0, 0), (50_001, 4096, 0, 65536), (50_001, 16384, 1024, 0),
    (1, 1 << 21, 1 << 16, 1 << 19), (0, 1 << 21, 1 << 16, 1 << 19), (4097, 4096, 0, 4096), (12_345, 0, 0, 0),
    (3_000_000, 1 << 21, 1 << 16, 1 << 19),  # head and taper together
]


@pytest.fixture(scope="module")
def stages_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("stages") / "test_stages")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "mind-fcl_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "test_stages.cpp"), "-o", exe], check=True)
    return exe


def test_stage_schedule_matches_its_restatement(stages_exe):
    args = [str(v) for c in CASES if c[0] and c[1] for v in c]
    r = subprocess.run([stages_exe] + args, stdout=subprocess.PIPE, text=True, check=True)
    got = [[int(t) for t in line.split()] for line in r.stdout.splitlines()]
    want = [mirror(*c) for c in CASES if c[0] and c[1]]
    assert got == want


def test_stage_schedule_properties(stages_exe):
    for n, chunk, head, taper in CASES:
        r = subprocess.run([stages_exe, str(n), str(chunk), str(head), str(taper)], stdout=subprocess.PIPE, text=True, check=True)
        sizes = [int(t) for t in r.stdout.split()]
        assert sum(sizes) == n and all(s > 0 for s in sizes)
        if n == 0:
            assert sizes == []
            continue
        c = max(chunk, 1)
        if head and head < c and n > 2 * head:  # compute-bound shape: short first stage, doubling, never above the chunk
            assert sizes[0] == head
            k = 0
            while k + 1 < len(sizes) and sizes[k + 1] == 2 * sizes[k] and sizes[k + 1] <= c:
                k += 1
            assert k >= 1 or sizes[1] in (c, n - head)
        if taper and not head and taper < c and n >= 2 * c:  # copy-bound shape: full stages, then a shrinking tail
            assert sizes[0] == c and sizes[-1] < c + taper and min(sizes) >= min(taper, n)
            tail = [s for s in sizes if s != c]
            assert tail == sorted(tail, reverse=True) or tail[-1] >= taper
        if not head and not taper:  # equal stages, ragged last one
            assert all(s == c for s in sizes[:-1]) and sizes[-1] <= c
