"""oracle/refio.py -- TEST INFRASTRUCTURE ONLY.

Reader for the binary container written by oracle/ref_driver.cpp
([name:32][dtype:1][count:int64][payload] records) and a helper that runs the
reference driver (oracle/_ref/ref_driver, the unmodified reference compiled by
oracle/Makefile).  Used by tests/, oracle/make_golden.py and bench.py's reference arm.
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")

_DT = {b"i": np.int32, b"d": np.float64, b"l": np.int64}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    while pos < len(buf):
        name = buf[pos:pos + 32].split(b"\0", 1)[0].decode()
        dt = _DT[buf[pos + 32:pos + 33]]
        n = int(np.frombuffer(buf, np.int64, 1, pos + 33)[0])
        pos += 41
        out[name] = np.frombuffer(buf, dt, n, pos).copy()
        pos += n * np.dtype(dt).itemsize
    return out


def have_ref_driver():
    return os.path.isfile(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)


def run_dump(**kw):
    """Run the reference on one case and return its dump as a dict of numpy arrays."""
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as t:
        path = t.name
    try:
        cmd = [REF_DRIVER, "dump", path]
        for k, v in kw.items():
            cmd += ["--" + k, str(v)]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        return read_dump(path)
    finally:
        os.unlink(path)


def run_bench(env=None, **kw):
    """Time the reference CPU path; returns the parsed JSON dict."""
    import json
    cmd = [REF_DRIVER, "bench"]
    for k, v in kw.items():
        cmd += ["--" + k, str(v)]
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, env=e)
    return json.loads(r.stdout.decode().strip().splitlines()[-1])
