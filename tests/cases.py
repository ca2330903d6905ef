"""Shared helpers for the parity tests: load a golden fixture, derive BC arrays and the equation system."""
import glob
import json
import os

import numpy as np

import tahoe_input as ti

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
PCG = [n for n in ALL if n.endswith("_pcg")]  # PCGSolver_LS runs (a21): converged by nonlinear CG, iteration counts recorded
STATIC = [n for n in ALL if n not in PCG and ("static" in n or n.startswith("ref_mat") or n.startswith("ref_beam") or n.endswith("_traction")
                                            or n == "ref_traction_a")]
XS = [n for n in ALL if "_xs_" in n]  # <explicit_solid> runs; their `fint` dump comes from the classic element path, not used
EXPLICIT = [n for n in ALL if "explicit" in n and n not in XS]
TRACTION = [n for n in ALL if n.endswith("_traction") or n == "ref_traction_a"]  # natural_bc tractions
STRESS = [n for n in ALL if n.endswith("_stress")]  # nodal stress output
WITH_LHS = [n for n in ALL if n.startswith("syn_") and "static" in n]


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.z = z
        self.desc = json.loads(str(z["desc"]))
        self.nodesets = {int(k[3:]): z[k] for k in z.files if k.startswith("ns_")}
        self.sidesets = {int(k[3:]): z[k] for k in z.files if k.startswith("ss_")}
        # natural_bc tractions are formed by the implementation under test: traction_fn(conn, X, elem, facet, tract, coord_system,
        # scale, out) adds the nodal forces of one natural_bc list to `out` (oracle_lib.traction_force or the device binding)
        self.traction_fn = None
        self.X = np.ascontiguousarray(z["ref_coords"])
        self.conn = np.ascontiguousarray(z["ref_conn"])
        self.nn, self.ne = self.X.shape[0], self.conn.shape[0]
        t = self.desc["time"]
        self.nsteps, self.dt = t["num_steps"], t["time_step"]
        self.dump_steps = sorted(int(k[6:]) for k in z.files if k.startswith("ref_d_"))
        # profile_matrix (CCSMatrixT) renumbers equations for bandwidth; others keep node-major numbering
        self.renumbered = self.desc["solver"].get("matrix") == "profile_matrix"

    def ref(self, key):
        return self.z["ref_" + key]

    def bc(self, t):
        code, val, fext = ti.bc_arrays(self.desc, self.nodesets, self.nn, t)
        cards = ti.traction_cards(self.desc, self.sidesets)
        if cards and self.traction_fn is None:
            fext[:] = np.nan  # callers that only want the BC codes may leave traction_fn unset; the load itself is then unusable
        elif cards:
            for elem, facet, tract, system, sched in cards:
                scale = ti.schedule_value(self.desc["time"]["schedules"][sched], t)
                self.traction_fn(self.conn, self.X, elem, facet, tract, system, scale, fext)
        return code, val, fext


def relerr(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / s
