#!/usr/bin/env python
"""SURVEY 8(f)-3 measurement: seconds to get a TahoeII .geom of n^3 Hex8 elements into memory --
the library's threaded reader (tb2_geom_open) against the reference executable's own input phase (its printed `Construction:`
time for a one-step explicit input on the same file: ModelManagerT/TahoeInputT parsing plus FEManagerT set-up).  Host only.
usage: python profiles/tools/geom_reader_bench.py [n=100]"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402
import tahoe_input as ti  # noqa: E402
from tahoe_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
work = tempfile.mkdtemp(prefix="tb2_geombench_")
X, conn, ns = ti.structured_cube(n, jitter=0.1)
path = os.path.join(work, "mesh.geom")
t0 = time.perf_counter()
with open(path, "w") as f:  # vectorised writer (ti.write_geom loops in Python)
    nn, ne = X.shape[0], conn.shape[0]
    f.write("*version\n1.0\n*title\nbench\n*dimensions\n%d\n3\n1\n1 %d 8\n%d\n" % (nn, ne, len(ns)))
    for sid in sorted(ns):
        f.write("%d %d\n" % (sid, len(ns[sid])))
    f.write("0\n*nodesets\n")
    for sid in sorted(ns):
        f.write("*set\n%d\n" % len(ns[sid]))
        np.savetxt(f, (ns[sid] + 1).reshape(-1, 1), fmt="%d")
    f.write("*sidesets\n*elements\n*set\n%d\n8\n" % ne)
    np.savetxt(f, np.column_stack([np.arange(1, ne + 1), conn + 1]), fmt="%d")
    f.write("*nodes\n%d\n3\n" % nn)
    np.savetxt(f, np.column_stack([np.arange(1, nn + 1), X]), fmt=["%d", "%.17e", "%.17e", "%.17e"])
size = os.path.getsize(path)
print("wrote %s: %.1f MB, %d elements, %d nodes (%.1f s)" % (path, size / 1e6, ne, nn, time.perf_counter() - t0))

lib = capi.lib()
best = 1e30
for _ in range(3):
    h = C.c_void_p()
    t0 = time.perf_counter()
    assert lib.tb2_geom_open(path.encode(), C.byref(h)) == 0
    best = min(best, time.perf_counter() - t0)
    lib.tb2_geom_close(h)
print("tb2_geom_open: %.3f s  (%.0f MB/s, %d host threads available)" % (best, size / 1e6 / best, os.cpu_count()))
Xr, blocks, nsr, _ = capi.read_geom(path)
assert np.array_equal(Xr, X) and np.array_equal(blocks[0], conn)

ref = os.path.join(REPO, "oracle", "_ref", "tahoe")
if os.path.exists(ref):
    desc = {"geometry_file": "mesh.geom", "time": {"num_steps": 1, "time_step": 1e-6, "schedules": [[(0.0, 1.0)]]},
            "integrator": "central_difference", "kbc": [{"nodeset": 1, "dof": d, "type": "fixed", "schedule": 0, "value": 0.0} for d in (1, 2, 3)],
            "fbc": [], "element": {"type": "total_lagrangian", "mass_type": "lumped_mass"},
            "material": {"type": "Simo_isotropic", "density": 1.0, "kappa": 1000.0, "mu": 5.0},
            "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}
    ti.write_xml(os.path.join(work, "run.xml"), desc)
    t0 = time.perf_counter()
    r = subprocess.run([ref, "-f", "run.xml"], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.perf_counter() - t0
    m = re.search(r"Construction:\s*([0-9.eE+-]+)\s*sec", r.stdout)
    print("reference executable: Construction %s s (wall of the whole one-step run %.1f s, rc %d)" % (m.group(1) if m else "?", wall, r.returncode))
import shutil
shutil.rmtree(work, ignore_errors=True)
