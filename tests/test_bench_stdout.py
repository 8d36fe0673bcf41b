"""bench.py's contract is ONE JSON line on stdout.  NCCL writes its version banner to fd 1 when the box sets NCCL_DEBUG,
so bench.py re-points fd 1 at stderr and keeps a private handle for the line (claim_stdout): checked here with a raw
write to fd 1, the way a C library would."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_only_the_json_line_reaches_stdout():
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 0.0.0\\n'); print('a stray print'); "
            "print(json.dumps({'metric': 'x'}), file=bench._OUT, flush=True)") % ROOT
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x"}
    assert "NCCL version 0.0.0" in r.stderr and "a stray print" in r.stderr
