"""Shared helpers for the parity tests: load a golden fixture, derive BC arrays and the equation system."""
import glob
import json
import os

import numpy as np

import tahoe_input as ti

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
             if not os.path.basename(p).startswith("ref_contact_"))  # the contact fixtures have their own layout (tests/test_contact_golden.py)
PCG = [n for n in ALL if n.endswith("_pcg")]  # PCGSolver_LS runs (a21): converged by nonlinear CG, iteration counts recorded
STATIC = [n for n in ALL if n not in PCG and ("static" in n or n.startswith("ref_mat") or n.startswith("ref_beam") or n.endswith("_traction")
                                            or n == "ref_traction_a")]
XS = [n for n in ALL if "_xs_" in n]  # <explicit_solid> runs; their `fint` dump comes from the classic element path, not used
EXPLICIT = [n for n in ALL if "explicit" in n and n not in XS]
TRACTION = [n for n in ALL if n.endswith("_traction") or n == "ref_traction_a"]  # natural_bc tractions
IMPLICIT = [n for n in ALL if "implicit" in n]  # nonlinear_HHT runs: the inertia branches of a2 / a16
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


MASS_TYPE = {"consistent_mass": 1, "lumped_mass": 2}  # ContinuumElementT::MassTypeT


def implicit_dynamics(c, act, fint, inertia, solve, step_solve=None):
    """FEManagerT's step loop for the `nonlinear_HHT` integrator (IntegratorT_factory.cpp:45-47: NLHHTalpha(0.0), i.e. Newmark with
    beta = 1/4, gamma = 1/2) around NLSolver::Solve, restated for the tests.  The unknown of the Newton iteration is the acceleration
    increment: nNLHHTalpha::Predictor / ConsistentKBC / Corrector (nNLHHTalpha.cpp:24-160,219-229), element constants
    eLinearHHTalpha::eComputeParameters (constM = 1, constK = beta dt^2) and eNLHHTalpha (constMa = constKd = 1).
    FEManagerT::InitialCondition first solves the same system with dt = 0 for the initial acceleration (FEManagerT.cpp:2053-2080).
    fint(d) -> [nn,3]; inertia(a) -> M a [nn,3]; solve(d, constM, constK, R[act]) -> acceleration increment on the active dofs.
    step_solve(d, v, a, fext, dt) -> iteration number, when given, replaces the Newton loop (a resident driver under test).
    Yields (step, d, v, a, iteration_number); step 0 is the initial-condition solve."""
    beta, gamma = 0.25, 0.5
    s = c.desc["solver"]
    atol, rtol = float(s["abs_tolerance"]), float(s["rel_tolerance"])
    d, v, a = np.zeros_like(c.X), np.zeros_like(c.X), np.zeros_like(c.X)
    for k in range(0, c.nsteps + 1):
        dt = c.dt if k else 0.0
        code, val, fext = c.bc(k * c.dt)
        dcorr_a, vcorr_a = beta * dt * dt, gamma * dt
        d += dt * v + (1.0 - 2.0 * beta) * 0.5 * dt * dt * a  # Predictor
        v += (1.0 - gamma) * dt * a
        a[:] = 0.0
        fixed = code != 0  # ConsistentKBC for kFix / kDsp
        target = np.where(code == 2, val, 0.0)
        a[fixed] = (target[fixed] - d[fixed]) / dcorr_a if abs(dcorr_a) > 1e-12 else 0.0
        d[fixed] = target[fixed]
        v[fixed] += vcorr_a * a[fixed]
        if step_solve is not None:
            yield k, d, v, a, step_solve(d, v, a, fext, dt)
            continue
        it = -1
        R = (fext - fint(d) - inertia(a))[act]
        e0 = e = np.linalg.norm(R)
        while e0 >= atol and not (it >= 0 and (e / e0 < rtol or e < atol)):
            assert it < 25
            da = solve(d, 1.0, dcorr_a, R)
            d[act] += dcorr_a * da  # Corrector on the active equations
            v[act] += vcorr_a * da
            a[act] += da
            it += 1
            R = (fext - fint(d) - inertia(a))[act]
            e = np.linalg.norm(R)
        yield k, d, v, a, it
