#!/usr/bin/env python
"""bench.py -- explicit central-difference element-updates/s on a structured Hex8 cube (BASELINE.json configs[1] per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n EDGE] [--workload explicit|pcg] [--impl reference]

One "step" = one explicit central-difference step over the whole mesh: total-Lagrangian Neo-Hookean (SimoIso3D) internal
force over every element (K1), deterministic node gather, M^-1 R, corrector and the next predictor (K5).
  value   element-updates/s, device-resident (d, v, a stay in HBM), CUDA events on the library's stream, max over ranks
  e2e     the same step through the host-buffer entry point tb2_explicit_step_host: d, v, a come from pinned host memory
          and go back every step (what a drop-in does when Tahoe's FieldT stays authoritative)
  roofline / roofline_hbm    the dominant kernel (K1) against the measured FP64 peak (the bound) and the HBM peak, from per-launch CUDA events
          recorded inside the timed region
  cpu_baseline   the unmodified reference binary (oracle/_ref/tahoe, built from /root/reference by oracle/build_ref.mk)
          on a bounded sample of the same workload, one core; falls back to the C oracle port if the binary is absent.
`--impl reference` times that reference binary alone on all host cores (one serial Tahoe process per core; Tahoe's
classic element path is single-threaded and no MPI exists on this box).
N > 1 (torchrun): weak scaling, one EDGE^3 brick per rank, interface-node force sums over NCCL.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402

METRIC = "element-updates/s"
MATERIAL = {"type": "Simo_isotropic", "density": 1.0, "kappa": 1000.0, "mu": 5.0}  # legacy_hex_*.xml
REF_BIN = os.path.join(REPO, "oracle", "_ref", "tahoe")


EXCHANGE = "peer"  # --exchange: interface sums over NVLink peer memory (tb2_comm_peer_*), or "nccl" = the packed ncclAllReduce


def comm_init(m, dist, comm):
    """tb2_comm_init on mesh m (+ the peer-window import when EXCHANGE is "peer": the 64-byte handles travel by all_gather_object)"""
    def gather(b):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, b)
        return out
    m.comm_init(*comm, all_gather=gather if EXCHANGE == "peer" else None)
    return m.comm_peer_enabled()


def stable_dt(n):
    return 0.25 * (1.0 / n) / np.sqrt((MATERIAL["kappa"] + 4.0 * MATERIAL["mu"] / 3.0) / MATERIAL["density"])


def initial_displacement(X):
    """smooth field + seeded noise (u = 1e-2 smooth + 1e-4 noise).  The smooth part is a 0.3 % homogeneous strain
    tapered to zero at the clamped x = 0 face (no displacement jump at the boundary condition) and the noise scales with the
    element size, so the same field is a gentle start at every mesh size."""
    rng = np.random.default_rng(12345)
    A = rng.standard_normal((3, 3))
    h = 1.0 / max(round((X.shape[0]) ** (1.0 / 3.0)) - 1, 1)
    return 1e-2 * 0.3 * X[:, :1] * (X @ A.T) + 1.5e-4 * h * rng.standard_normal(X.shape)


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock + throttle-reason sampler for the timed region.  NVML in a thread (5 ms period: the timed region of a small
    --steps run is only tens of ms); falls back to the recipe's `nvidia-smi -lms` line when pynvml is unusable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu, uuid=None):
        self.gpu, self.uuid, self.proc, self.lines = gpu, uuid, None, []
        self.nvml, self.handle, self.samples, self.smax, self.bits, self.power = None, None, [], None, 0, []
        self.stop_flag = threading.Event()

    def _nvml_open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                for cand in (self.uuid, self.uuid.encode()):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, h
            return True
        except Exception:
            return False

    def _nvml_loop(self):
        p, h = self.nvml, self.handle
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(p.nvmlDeviceGetClockInfo(h, p.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(p.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.bits |= int(p.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.power.append(p.nvmlDeviceGetPowerUsage(h) * 1e-3)
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def start(self):
        if self._nvml_open():
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.t.join(timeout=2)
            reasons = sorted(name for bit, name in self.REASONS if self.bits & bit)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None, "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 20"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the unmodified reference binary on a bounded sample
# ---------------------------------------------------------------------------------------------------------------------
XS_MATERIAL = {"type": "explicit_neo_hookean", "density": 1.0, "kappa": 1000.0, "mu": 5.0}  # vectorized_hex_*.xml (level.5/explicit_benchmark)


def _write_reference_case(work, n, nsteps, element="total_lagrangian", material=None):
    import tahoe_input as ti
    X, conn, ns = ti.structured_cube(n, jitter=0.1)
    if not os.path.exists(os.path.join(work, "mesh.geom")):
        ti.write_geom(os.path.join(work, "mesh.geom"), X, conn, ns)
    desc = {"geometry_file": "mesh.geom",
            "time": {"num_steps": nsteps, "time_step": float(stable_dt(n)), "schedules": [[(0.0, 1.0)]]},
            "integrator": "central_difference",
            "kbc": [{"nodeset": 1, "dof": d, "type": "fixed", "schedule": 0, "value": 0.0} for d in (1, 2, 3)],
            "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02 / (n * n)}],
            "element": {"type": element, "mass_type": "lumped_mass"}, "material": material or MATERIAL,
            "solver": {"type": "linear_solver", "matrix": "diagonal_matrix"}}
    ti.write_xml(os.path.join(work, "run_%d.xml" % nsteps), desc)
    return conn.shape[0]


def _run_reference_batch(work, nsteps, procs, omp_threads=1):
    t0 = time.perf_counter()
    ps = [subprocess.Popen([REF_BIN, "-f", "run_%d.xml" % nsteps], cwd=work, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                           env=dict(os.environ, OMP_NUM_THREADS=str(omp_threads))) for _ in range(procs)]
    rc = [p.wait() for p in ps]
    if any(rc):
        raise RuntimeError("reference binary failed: %s" % rc)
    return time.perf_counter() - t0


def reference_rate(n, s_lo, s_hi, procs, min_seconds=2.0, element="total_lagrangian", material=None, omp_threads=1):
    """element-updates/s of `procs` concurrent serial reference processes on an n^3 cube, from the wall-clock difference
    between an s_hi-step and an s_lo-step run (cancels input parsing, set-up and FEManagerT::InitialCondition).  If the difference
    is shorter than min_seconds (a short --steps request) the long run is repeated with 4x the steps, so the rate never comes
    from timer noise.  Returns (rate, seconds, elements, timed steps)."""
    work = tempfile.mkdtemp(prefix="tb2_ref_")
    try:
        ne = _write_reference_case(work, n, s_lo, element, material)
        t_lo = _run_reference_batch(work, s_lo, procs, omp_threads)
        k = s_hi - s_lo
        for attempt in range(5):
            _write_reference_case(work, n, s_lo + k, element, material)
            dt = _run_reference_batch(work, s_lo + k, procs, omp_threads) - t_lo
            if dt >= min_seconds or attempt == 4:
                break
            k *= 4
    finally:
        shutil.rmtree(work, ignore_errors=True)
    dt = max(dt, 1e-3)
    return procs * ne * k / dt, dt, ne, k


PLUGIN_BIN = os.path.join(REPO, "tahoe_b200", "host", "_build", "tahoe_b200")


def plugin_leg(n=48, s_lo=10, s_hi=110, ref_s_hi=20):
    """The drop-in itself (tahoe_b200/host: the reference's libraries + the plugin classes, tests/test_plugin_binary.py): the same XML
    family through the plugin EXECUTABLE -- (a) cuda_total_lagrangian with Tahoe's own linear_solver / diagonal_matrix / nExplicitCD
    on the host (u up, force down every step), (b) the resident pair integrator="CUDA_central_difference" + <CUDA_explicit_solver>
    -- and through the unmodified reference, wall(s_hi steps) - wall(s_lo steps) each.  Everything outside the hot path (FieldT
    bookkeeping, FEManagerT's step loop) is Tahoe's host code in all three."""
    if not (os.path.exists(PLUGIN_BIN) and os.path.exists(REF_BIN)):
        return None
    import tahoe_input as ti
    work = tempfile.mkdtemp(prefix="tb2_plugin_")
    try:
        X, conn, ns = ti.structured_cube(n, jitter=0.1)
        ti.write_geom(os.path.join(work, "mesh.geom"), X, conn, ns)
        ne = conn.shape[0]

        def case(name, nsteps, element, integrator, solver):
            desc = {"geometry_file": "mesh.geom", "output_inc": nsteps,
                    "time": {"num_steps": nsteps, "time_step": float(stable_dt(n)), "schedules": [[(0.0, 1.0)]]},
                    "integrator": integrator,
                    "kbc": [{"nodeset": 1, "dof": d, "type": "fixed", "schedule": 0, "value": 0.0} for d in (1, 2, 3)],
                    "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02 / (n * n)}],
                    "element": {"type": "total_lagrangian", "tag": element, "mass_type": "lumped_mass", "nodal_output": True},
                    "material": MATERIAL, "solver": solver}
            ti.write_xml(os.path.join(work, "%s_%d.xml" % (name, nsteps)), desc)

        def wall(binary, name, nsteps):
            """seconds of FEManagerT::Solve as the executable reports them (`Solution:`, FEExecutionManagerT.cpp:513-525: clock(), i.e. CPU
            time of the single host thread -- it spins in the CUDA synchronisations, so for these runs it is the wall time of the phase
            without the process start-up, CUDA context creation and input parsing that a wall clock around the process would add)"""
            import re
            r = subprocess.run([binary, "-f", "%s_%d.xml" % (name, nsteps)], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            m_ = re.search(r"Solution:\s*([0-9.eE+-]+)\s*sec", r.stdout)
            if r.returncode or "End Execution" not in r.stdout or not m_:
                raise RuntimeError("%s failed: %s" % (name, r.stdout[-800:]))
            return float(m_.group(1))

        host = {"type": "linear_solver", "matrix": "diagonal_matrix"}
        variants = {"reference": (REF_BIN, "total_lagrangian", "central_difference", host),
                    "plugin_host_integrator": (PLUGIN_BIN, "cuda_total_lagrangian", "central_difference", host),
                    "plugin_resident": (PLUGIN_BIN, "cuda_total_lagrangian", "CUDA_central_difference",
                                        {"type": "CUDA_explicit_solver", "matrix": "diagonal_matrix"})}
        out = {"workload": "%d^3=%d-element cube, total_lagrangian + Simo_isotropic, explicit central difference, one output step at the end; "
                           "the executable's own `Solution:` seconds, (%d steps) - (%d steps), best of 2 each (%d steps for the reference)" % (n, ne, s_hi, s_lo, ref_s_hi)}
        for name, (binary, element, integrator, solver) in variants.items():
            hi = ref_s_hi if name == "reference" else s_hi  # the classic element loop takes ~0.2 s per step at this size
            for k in (s_lo, hi):
                case(name, k, element, integrator, solver)
            wall(binary, name, s_lo)  # warm file cache; CUDA context creation is inside both timed runs and cancels
            t_lo = min(wall(binary, name, s_lo) for _ in range(2))
            t_hi = min(wall(binary, name, hi) for _ in range(2))
            dt = max(t_hi - t_lo, 1e-4)
            out[name] = {"ms_per_step": 1e3 * dt / (hi - s_lo), "element_updates_per_s": ne * (hi - s_lo) / dt, "timed_steps": hi - s_lo}
        out["resident_vs_reference"] = out["plugin_resident"]["element_updates_per_s"] / out["reference"]["element_updates_per_s"]
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def oracle_port_rate(n, steps):
    """fallback CPU baseline: the plain-C oracle port (kind = "port"), one core"""
    import oracle_lib as orc
    import tahoe_input as ti
    X, conn, ns = ti.structured_cube(n, jitter=0.1)
    mat = orc.material(MATERIAL)
    u = initial_displacement(X)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.internal_force(orc.TOTAL_LAGRANGIAN, mat, conn, X, u)
    return conn.shape[0] * steps / (time.perf_counter() - t0), conn.shape[0]


def cpu_baseline(sample_n=40, s_lo=2, s_hi=42):
    if os.path.exists(REF_BIN):
        rate, dt, ne, k = reference_rate(sample_n, s_lo, s_hi, 1)
        return {"value": rate, "unit": METRIC, "cores": 1, "kind": "reference",
                "sample": "%d^3=%d-element jittered cube, %d explicit steps of oracle/_ref/tahoe (classic total_lagrangian + Simo_isotropic, "
                          "lumped mass, central_difference), wall(%d steps) - wall(%d steps) = %.2f s" % (sample_n, ne, k, s_lo + k, s_lo, dt)}
    rate, ne = oracle_port_rate(24, 4)
    return {"value": rate, "unit": METRIC, "cores": 1, "kind": "port",
            "sample": "24^3=%d elements x 4 internal-force sweeps of oracle/tahoe_oracle.c (reference binary absent)" % ne}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 32
    w, k = max(args.warmup, 1), max(args.steps, 1)
    k = min(k, 60)
    if os.path.exists(REF_BIN):
        rate, dt, ne, k = reference_rate(n, w, w + k, cores)
        kind, sample = "reference", ("%d concurrent serial oracle/_ref/tahoe processes (one per host core; no MPI/METIS on this box), each a %d^3=%d-element "
                                     "jittered cube, %d timed explicit steps (wall(%d) - wall(%d) steps)" % (cores, n, ne, k, w + k, w))
    else:
        rate, ne = oracle_port_rate(24, k)
        dt, cores, kind, sample = ne * k / rate, 1, "port", "oracle/tahoe_oracle.c internal-force sweeps on 24^3 (reference binary absent)"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": METRIC, "n_gpus": args.gpus, "steps": k, "warmup": w,
            "ms_per_step": 1e3 * dt / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            # the workload family and metric of the GPU arm; what is actually timed is `reference_workload` (CPU rates are
            # taken per element on sizes the reference finishes in minutes -- a cache-resident size without communication favours the CPU)
            "config": dict(workload_config(args.n, args.gpus), same_config=False,
                           reference_workload="%d^3=%d-element cube x %d concurrent serial processes (one per host core), per-element rate "
                                              "extrapolated to the %d^3 workload" % (n, ne, cores, args.n)),
            "cpu_baseline": {"value": rate, "unit": METRIC, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def pcg_cpu_reference(n=16):
    """the reference's own iterative static solve -- <PCG_solver> with a <diagonal_matrix> preconditioner
    (PCGSolver_LS: nonlinear CG, >= 2 element residual sweeps per iteration) -- on a small_strain + SSKStV cube of the PCG leg's
    family, one host core.  Iterations are counted from the printed error history, seconds are the binary's own `Solution:` time."""
    import re
    import tahoe_input as ti
    if not os.path.exists(REF_BIN):
        return None
    work = tempfile.mkdtemp(prefix="tb2_refpcg_")
    try:
        X, conn, ns = ti.structured_cube(n, jitter=0.1)
        ti.write_geom(os.path.join(work, "mesh.geom"), X, conn, ns)
        desc = {"geometry_file": "mesh.geom", "time": {"num_steps": 1, "time_step": 1.0, "schedules": [[(0.0, 0.0), (1.0, 1.0)]]},
                "integrator": "static",
                "kbc": [{"nodeset": 1, "dof": d, "type": "fixed", "schedule": 0, "value": 0.0} for d in (1, 2, 3)],
                "fbc": [{"nodeset": 2, "dof": 1, "schedule": 1, "value": 0.02 / (n * n)}],
                "element": {"type": "small_strain"}, "material": {"type": "small_strain_StVenant", "density": 1.0, "E": 100.0, "nu": 0.25},
                "solver": {"type": "PCG_solver", "output_flag": "all_iterations", "abs_tolerance": "1.0e-14", "divergence_tolerance": "1.0e+06",
                           "line_search_iterations": "10", "line_search_tolerance": "0.1", "max_iterations": "5000", "max_step": "2.5",
                           "quick_solve_iter": "100", "rel_tolerance": "1.0e-08", "restart": "50", "matrix": "diagonal_matrix"}}
        ti.write_xml(os.path.join(work, "pcg.xml"), desc)
        r = subprocess.run([REF_BIN, "-f", "pcg.xml"], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=600)
        its = len([ln for ln in r.stdout.splitlines() if "Relative error =" in ln]) - 1
        sec = re.search(r"Solution:\s*([0-9.eE+-]+)\s*sec", r.stdout)
        if r.returncode != 0 or its < 1 or not sec:
            return {"unavailable": "reference PCG_solver run failed (rc %d)" % r.returncode}
        neq = 3 * (X.shape[0] - len(ns[1]))
        seconds = float(sec.group(1))
        return {"value": neq * its / max(seconds, 1e-6), "unit": "DOF-iters/s", "cores": 1, "kind": "reference", "iterations": its,
                "num_equations": neq, "seconds": seconds,
                "sample": "oracle/_ref/tahoe, <PCG_solver><diagonal_matrix/> (nonlinear CG with line search, rel. tolerance 1e-8) on a "
                          "%d^3=%d-element small_strain + SSKStV jittered cube, one static step" % (n, conn.shape[0])}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def workload_config(n, gpus):
    return {"workload": "BASELINE.json configs[1]: %d^3=%d-element structured Hex8 cube per GPU (jitter 0.1 h, seed 12345), total_lagrangian + "
                        "Simo_isotropic Neo-Hookean (kappa=1000, mu=5, rho=1), 8 integration points, lumped mass, explicit central difference"
                        % (n, n ** 3),
            "elements_per_gpu": n ** 3, "partition": "1 brick" if gpus == 1 else "%d bricks, NCCL interface-node force sum" % gpus,
            "l2": "inputs larger than L2 (per-step working set ~%d MB vs 126 MB L2)" % (n ** 3 * (192 + 32 + 9 * 24) // 10 ** 6)}


# ---------------------------------------------------------------------------------------------------------------------
# second half of BASELINE.json's metric: PCG DOF-iterations/s (configs[2] scaled to one GPU's share)
# ---------------------------------------------------------------------------------------------------------------------
def run_pcg(torch, dist, capi, tmesh, local, rank, world, n, iters, hbm_peak):
    """small-strain elastic cube, n^3 elements per GPU: device assembly (K3) of the CSR tangent, then `iters` Jacobi-PCG iterations
    (K6-K8) with device-resident vectors.  N > 1: every rank holds the sub-domain matrix of its brick; per iteration one packed
    interface sum of A_loc p and two scalar all-reduces over NCCL.  Returns the "pcg" object of the JSON line (rank 0)."""
    if world == 1:
        X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
        part = None
    else:
        gx, gy, gz = tmesh.brick_grid(world)
        part = tmesh.partition_cube(n * gx, n * gy, n * gz, world, rank, jitter=0.1)
        X, conn, ns = part["coords"], part["conn"], part["nodesets"]
    m = capi.Mesh(X, conn, device=local)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm_init(m, dist, (rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"]))
    g = capi.Group(m, capi.SMALL_STRAIN, capi.material({"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    t0 = time.perf_counter()
    eqs = capi.Equations(m, code)
    A = capi.Matrix(eqs)
    m.synchronize()
    t_struct = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    u0 = torch.zeros(X.shape, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    A.form_stiffness(g, u0)  # warm-up (includes the one-off colouring)
    m.synchronize()
    A.clear()
    ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.synchronize()
    ea0.record(stream)  # the assembly pipeline (element kernel || row gather on two streams) joins the mesh stream at its end
    A.form_stiffness(g, u0)
    ea1.record(stream)
    m.synchronize()
    torch.cuda.synchronize()
    t_assembly_ms = float(ea0.elapsed_time(ea1))
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 1e-3
    active = eqs.eqnos() > 0
    b = torch.from_numpy(fext[active]).to(dev)
    x = torch.zeros_like(b)
    neq_glob = torch.tensor([float((active & (part["owned"][:, None] > 0)).sum()) if part else float(A.neq), float(A.nnz), float(conn.shape[0])],
                            device=dev, dtype=torch.float64)
    torch.cuda.synchronize()
    A.pcg(b, x, rtol=0.0, atol=0.0, max_iter=iters)  # warm-up with the timed call's arguments (one-off CUDA-graph capture of 16 iterations)
    x.zero_()
    m.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.all_reduce(neq_glob, op=dist.ReduceOp.SUM)
        dist.barrier()
    # timed pass: as a caller runs it (single GPU: iterations replayed from the CUDA graph; no per-launch events in the way)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    it, rn = A.pcg(b, x, rtol=0.0, atol=0.0, max_iter=iters)
    e1.record(stream)
    m.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if it < iters or not np.isfinite(rn):
        raise SystemExit("bench.py: PCG ran %d of %d iterations, |r| = %g" % (it, iters, rn))
    iters = it  # whole batches of 16 are launched: the count the library reports is the one that ran
    # profiled pass: the same solve again with per-launch CUDA events, for the SpMV's own duration and its share
    x.zero_()
    m.synchronize()
    m.profile_begin()
    it2, rn = A.pcg(b, x, rtol=0.0, atol=0.0, max_iter=min(iters, 48))
    ms_cat, cnt_cat, launches = m.profile_end()
    ms_profiled_iter = (float(ms_cat[3]) + float(ms_cat[4]) + float(ms_cat[6])) / max(it2, 1)
    if world == 1:
        launches = 5 * iters + 4  # timed pass: 5 kernels per iteration + set-up (counted, not measured: the graph replays them)
    else:
        # distributed iteration over peer memory: update, publishing SpMV, pull, scalar step (4 of ours); with the
        # NCCL exchange: update, interface rows, pack, interior rows, unpack, partial sums, scalars (7 of ours + 2 NCCL kernels)
        launches = (4 if m.comm_peer_enabled() else 7) * iters + 12
    # configs[2] as one call: NLSolver::Solve on the device (tb2_newton_solve: K1 residual, K3 tangent, Jacobi-PCG to 1e-8, update)
    newton = None
    if world == 1:
        work = capi.NonlinearPCG(g, eqs, capi.nlpcg_params())
        prm = capi.newton_params(abs_tolerance=1e-30, rel_tolerance=1e-6, max_iterations=5, pcg_rel_tolerance=1e-8, pcg_max_iterations=4000)
        # one untimed call first: it is the first use of this path's kernels in the process (lazy module loading, graph instantiation,
        # staging allocations) -- r01h-j saw 0.53 / 0.76 / 1.11 s for the identical 1029 iterations when that was inside the timing
        warm = capi.newton_params(abs_tolerance=1e-30, rel_tolerance=1e-6, max_iterations=5, pcg_rel_tolerance=1e-8, pcg_max_iterations=32)
        capi.newton_solve_host(work, A, warm, np.zeros_like(X), fext, solve_max_iterations=1)
        m.synchronize()
        un = np.zeros_like(X)
        t0 = time.perf_counter()
        st, nit, nerr, nerr0, lin = capi.newton_solve_host(work, A, prm, un, fext)
        t_newton = time.perf_counter() - t0
        newton = {"status": {0: "continue", 1: "converged", 2: "failed"}[st], "newton_iterations": nit + 1, "pcg_iterations": int(lin),
                  "seconds": t_newton, "relative_residual": nerr / nerr0 if nerr0 > 0 else None,
                  "dof_iters_per_s": float(A.neq) * lin / t_newton,
                  "note": "host-buffer call: H2D of u and fext, residual + tangent assembly + PCG (rel 1e-8) + update per Newton iteration, D2H of u"}
        work.close()
    tt = torch.tensor([ms, float(ms_cat[3]) / max(int(cnt_cat[3]), 1), float(ms_cat[6]) / max(it2, 1), t_assembly_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, spmv_ms, comm_ms, t_assembly_ms = (float(v) for v in tt.tolist())
    neq_total, nnz_total, ne_total = (float(v) for v in neq_glob.tolist())
    # algorithmic bytes of this rank's SpMV launch.  SURVEY.md 8d's plain-CSR figure is 12 B/nnz + 12 B/row; the node-grouped
    # format this kernel reads stores the column indices once per 3-row node group, so its compulsory traffic is
    # 8 B/nnz (values) + 4/3 B/nnz (indices) + 20 B/row (rowptr, x, y) -- the smaller figure is the honest numerator.
    spmv_bytes = (8.0 + 4.0 / 3.0) * A.nnz + 20.0 * A.neq
    spmv_bytes_csr = 12.0 * A.nnz + 12.0 * A.neq
    iter_bytes = spmv_bytes + 128.0 * A.neq         # + 16 vector passes of the unfused PCG iteration
    out = {"metric": "PCG DOF-iters/s", "value": neq_total * iters / (ms * 1e-3), "unit": "DOF-iters/s", "n_gpus": world, "iterations": iters,
           "ms_per_iteration": ms / iters, "num_equations": int(neq_total), "nnz": int(nnz_total), "gpu_launches": int(launches),
           "workload": "BASELINE.json configs[2] at %d^3=%d small_strain + SSKStV elements per GPU (%d GPUs: %d elements, %d equations), CSR "
                       "assembled on the device (K3)%s, Jacobi-PCG with device-resident vectors"
                       % (n, n ** 3, world, int(ne_total), int(neq_total), " as sub-domain matrices" if world > 1 else ""),
           "assembly": {"ms": t_assembly_ms, "elements_per_s": ne_total / (t_assembly_ms * 1e-3), "structure_build_s": t_struct,
                        "kernels": "k_element_stiffness (symmetric-half element matrices -> packed [element][300] scratch) + k_assemble_gather "
                                   "(ordered row gather into the CSR); no colouring, no atomics",
                        "fp64_tflops": K3_FLOP_PER_ELEMENT * conn.shape[0] / (t_assembly_ms * 1e-3) * 1e-12,
                        "scratch_plus_matrix_gbs": (2 * 300 * 8.0 * conn.shape[0] + 2 * 8.0 * A.nnz) / (t_assembly_ms * 1e-3) * 1e-9},
           "newton": newton,
           "roofline": {"bound": "hbm", "kernel": "k_spmv (K6)", "achieved": spmv_bytes / (spmv_ms * 1e-3) * 1e-9, "peak": hbm_peak,
                        "unit": "GB/s", "frac": spmv_bytes / (spmv_ms * 1e-3) * 1e-9 / hbm_peak, "avg_launch_ms": spmv_ms,
                        "share_of_iteration": spmv_ms / ms_profiled_iter if ms_profiled_iter > 0 else None, "traffic": PCG_SPMV_TRAFFIC_BYTES_PER_NNZ * A.nnz,
                        "algorithmic_bytes_per_launch": spmv_bytes, "plain_csr_bytes_per_launch": spmv_bytes_csr,
                        "plain_csr_equivalent_gbs": spmv_bytes_csr / (spmv_ms * 1e-3) * 1e-9,
                        "note": "algorithmic bytes = node-grouped CSR: 8 B/nnz values + 4/3 B/nnz indices (read once per 3-row node group) + "
                                "20 B/row; SURVEY.md 8d's plain-CSR figure (12 B/nnz + 12 B/row) is given beside it"},
           "interface_exchange_ms_per_iteration": comm_ms if world > 1 else None,
           "iteration_hbm_frac": iter_bytes * iters / (ms * 1e-3) * 1e-9 / hbm_peak}
    A.close(); eqs.close(); g.close(); m.close()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own fast path, <explicit_solid> (ExplicitElementT, 8-point Hex8, ExplNeoHookeanT)
# ---------------------------------------------------------------------------------------------------------------------
def run_explicit_solid(torch, capi, tmesh, local, n, steps, warmup, hbm_peak, with_cpu):
    """device-resident explicit steps of the explicit_solid element on one n^3 cube, and (with_cpu) the reference executable on the
    same input family with its OpenMP batch loop on all host cores (ExplicitElementT.cpp:697-704) and with one thread"""
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    m = capi.Mesh(X, conn, device=local)
    g = capi.Group(m, capi.UPDATED_LAGRANGIAN, capi.material(XS_MATERIAL))
    dt_cfl = g.stable_time_step()
    ex = capi.Explicit(g)
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 0.02 / (n * n)
    ex.set_bc(code, np.zeros_like(X), fext)
    ex.set_state(initial_displacement(X), np.zeros_like(X), np.zeros_like(X))
    dt = float(stable_dt(n))
    stream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local))
    ex.run(dt, warmup)
    m.synchronize()
    m.profile_reserve(steps * 16 + 64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.profile_begin()
    e0.record(stream)
    ex.run(dt, steps)
    e1.record(stream)
    ms_cat, cnt_cat, launches = m.profile_end()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d, v, a = ex.get_state()
    if not np.isfinite(d).all():
        raise SystemExit("bench.py: non-finite state in the explicit_solid leg")
    ne = conn.shape[0]
    k1_ms = float(ms_cat[0]) / steps
    out = {"metric": METRIC, "value": ne * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "gpu_launches": int(launches),
           "workload": "%d^3=%d-element jittered cube, <explicit_solid>: 8-point Hex8, ExplNeoHookeanT kappa=1000 mu=5, lumped mass, central difference; "
                       "CFL estimate of the element (ExplicitElementT::ComputeStableTimeStep) %.3e, dt used %.3e" % (n, ne, dt_cfl, dt),
           "roofline": {"bound": "hbm", "kernel": "k_internal_force<UL,ExplNeoHookean> (K1)", "achieved": 104.0 * ne / (k1_ms * 1e-3) * 1e-9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": 104.0 * ne / (k1_ms * 1e-3) * 1e-9 / hbm_peak, "ms_per_step": k1_ms,
                        "share_of_step": k1_ms * steps / ms, "note": "FP64-pipe bound like the SimoIso3D sweep; no cube root in this law"},
           "cpu_reference": None}
    ex.close(); g.close(); m.close()
    if with_cpu and os.path.exists(REF_BIN):
        cores = os.cpu_count() or 1
        nref = 32  # 256 batches of 128 elements >= 4 x threads: the reference's OpenMP loop engages (ExplicitElementT.cpp:701)
        r_omp, t_omp, ne_ref, k_omp = reference_rate(nref, 3, 13, 1, 2.0, "explicit_solid", XS_MATERIAL, cores)
        r_one, t_one, _, k_one = reference_rate(nref, 3, 13, 1, 2.0, "explicit_solid", XS_MATERIAL, 1)
        out["cpu_reference"] = {"omp": {"value": r_omp, "threads": cores, "steps": k_omp, "seconds": t_omp},
                                "serial": {"value": r_one, "threads": 1, "steps": k_one, "seconds": t_one}, "unit": METRIC,
                                "sample": "oracle/_ref/tahoe on a %d^3=%d-element <explicit_solid> input of the same family" % (nref, ne_ref)}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: UpdatedLagrangianT + J2Simo3D, implicit static step with the (matrix-free) nonlinear PCG
# ---------------------------------------------------------------------------------------------------------------------
def run_nlpcg_j2(torch, capi, tmesh, local, n, iters):
    """one load step of the resident PCGSolver_LS twin (tb2_nlpcg_solve) on an n^3 J2 cube pulled past yield: `iters` nonlinear-CG
    iterations (each 2-4 residual sweeps of the return-mapping kernel + preconditioner refreshes).  J2Simo3D's tangent is
    non-symmetric (J2Simo3D.cpp:18-21), so the reference solves this config with LU or with PCG_solver; the latter is this path."""
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    m = capi.Mesh(X, conn, device=local)
    mat = {"type": "Simo_J2", "E": 100.0, "nu": 0.25, "density": 1.0, "hardening": {"type": "linear_function", "a": 0.05, "b": 0.25}}  # mat.09.a.xml
    g = capi.Group(m, capi.UPDATED_LAGRANGIAN, capi.material(mat))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    code[ns[2], 0] = 1
    eqs = capi.Equations(m, code)
    solver = capi.NonlinearPCG(g, eqs, capi.nlpcg_params(restart=30, line_search_iterations=10, line_search_tolerance=0.1, rel_tolerance=1e-30,
                                                         abs_tolerance=1e-30, max_iterations=10 ** 6))
    dev = torch.device("cuda", local)
    u = np.zeros_like(X)
    u[:, 0] = 0.02 * X[:, 0]   # 2 % stretch (well past yield: sigma_Y / E ~ 0.3 %), homogeneous start; the solver finds the lateral contraction
    d_u = torch.from_numpy(u).to(dev)
    d_ul = torch.zeros_like(d_u)
    d_f = torch.zeros_like(d_u)
    torch.cuda.synchronize()
    st, it, err, err0 = solver.solve(d_u, d_f, d_ul, solve_max_iterations=2)  # warm-up (allocation of the yielded elements' history)
    sw0, pc0 = solver.counters()
    m.synchronize()
    t0 = time.perf_counter()
    st, it, err, err0 = solver.solve(d_u, d_f, d_ul, solve_max_iterations=iters)
    m.synchronize()
    dt = time.perf_counter() - t0
    sw1, pc1 = solver.counters()
    data, flags, alloc = (None, None, None)
    n_alloc = int(g.get_history()[2].sum()) if n <= 64 else None
    out = {"workload": "BASELINE.json configs[3] at %d^3=%d updated_lagrangian + Simo_J2 elements (E=100, nu=0.25, linear hardening 0.05 a + 0.25), "
                       "2 %% stretch, %d equations; nonlinear PCG (restart 30, 10 line-search evaluations), device resident" % (n, conn.shape[0], eqs.neq),
           "iterations": iters, "seconds": dt, "residual_sweeps": sw1 - sw0, "preconditioner_sweeps": pc1 - pc0,
           "dof_iters_per_s": eqs.neq * iters / dt, "element_sweeps_per_s": conn.shape[0] * (sw1 - sw0) / dt,
           "relative_residual": err / err0 if err0 > 0 else None, "yielded_elements": n_alloc}
    solver.close(); eqs.close(); g.close(); m.close()
    return out


def run_newton_j2(capi, tmesh, local, n):
    """BASELINE.json configs[3] as stated -- updated_lagrangian + Simo_J2, implicit Newton with a Krylov solve on the device CSR: one
    load step (1 % stretch: every point yields) through tb2_newton_solve_host.  The consistent tangent is non-symmetric
    (J2Simo3D.cpp:18-21), so the linear solve is the Jacobi-preconditioned BiCGStab (tb2_matrix_bicgstab), 2 SpMVs per iteration."""
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    m = capi.Mesh(X, conn, device=local)
    mat = {"type": "Simo_J2", "E": 100.0, "nu": 0.25, "density": 1.0, "hardening": {"type": "linear_function", "a": 0.05, "b": 0.25}}  # mat.09.a.xml
    g = capi.Group(m, capi.UPDATED_LAGRANGIAN, capi.material(mat))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    code[ns[2], 0] = 2
    val = np.zeros_like(X)
    val[ns[2], 0] = 0.01
    eqs = capi.Equations(m, code)
    A = capi.Matrix(eqs)
    work = capi.NonlinearPCG(g, eqs, capi.nlpcg_params())
    prm = capi.newton_params(abs_tolerance=1e-30, rel_tolerance=1e-8, max_iterations=25, pcg_rel_tolerance=1e-9, pcg_max_iterations=40000)
    nsteps, stretch = 8, 0.01  # 0.125 % per load step: Newton from the last converged state stays inside its basin (0.25 % does not)
    u = np.zeros_like(X)
    u_last = np.zeros_like(X)
    m.synchronize()
    t0 = time.perf_counter()
    newton_its, lin, st, err, err0 = 0, 0, 1, 0.0, 0.0
    for k in range(1, nsteps + 1):  # FEManagerT's load steps: prescribed values of the step, Newton, CloseStep (history update)
        u[code == 2] = (stretch * k / nsteps) * val[code == 2] / 0.01
        st, nit, err, err0, li = capi.newton_solve_host(work, A, prm, u, np.zeros_like(X), u_last=u_last)
        newton_its += nit + 1
        lin += int(li)
        if st != 1:
            break
        g.close_step()
        u_last = u.copy()
    m.synchronize()
    dt = time.perf_counter() - t0
    n_alloc = int(g.get_history()[2].sum()) if n <= 64 else None
    out = {"workload": "BASELINE.json configs[3] at %d^3=%d updated_lagrangian + Simo_J2 elements, %d equations, %d non-zeros: %d load steps to "
                       "%g %% stretch, Newton to 1e-8 with Jacobi-BiCGStab to 1e-9 on the device-assembled non-symmetric CSR"
                       % (n, conn.shape[0], eqs.neq, A.nnz, nsteps, 100 * stretch),
           "status": {0: "continue", 1: "converged", 2: "failed"}[st], "newton_iterations": newton_its, "bicgstab_iterations": int(lin),
           "seconds": dt, "relative_residual": err / err0 if err0 > 0 else None, "dof_iters_per_s": float(eqs.neq) * lin / dt,
           "spmv_per_s": 2.0 * lin / dt, "yielded_elements": n_alloc,
           "note": "host-buffer calls: include residual sweeps, tangent assemblies (K3, full 24x24 element matrices) and the H2D/D2H of u"}
    work.close(); A.close(); eqs.close(); g.close(); m.close()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def parity_check(torch, dist, capi, tmesh, rank, world, local):
    """Before anything is timed: the partitioned run must reproduce the single-GPU run of the same mesh.  N > 1: 16^3 elements per
    rank, 8 explicit steps (TL Neo-Hookean) and one small-strain Jacobi-PCG solve on the element-partitioned cube against rank 0's
    own single-GPU run of the whole cube -- max relative difference over d, v, a and over the PCG solution, and whether all sharers
    of an interface node hold bitwise identical values.  N = 1: the same steps twice, bit for bit.  Returns the "parity" object."""
    n = 16
    dev = local

    def explicit(X, conn, ns, comm, nsteps=8):
        m = capi.Mesh(X, conn, device=dev)
        if comm:
            comm_init(m, dist, comm)
        g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material(MATERIAL))
        ex = capi.Explicit(g)
        code = np.zeros(X.shape, np.uint8)
        code[ns[1]] = 1
        fext = np.zeros_like(X)
        fext[ns[2], 0] = 1e-3
        ex.set_bc(code, np.zeros_like(X), fext)
        ex.set_state(0.01 * X @ np.array([[0.3, -0.2, 0.1], [0.05, 0.4, -0.3], [0.2, 0.1, -0.25]]) + 1e-3 * np.sin(7.0 * X[:, ::-1]),
                     np.zeros_like(X), np.zeros_like(X))
        ex.run(2e-4, nsteps)
        out = ex.get_state()
        # the implicit side: K x = f with this rank's sub-domain matrix
        g2 = capi.Group(m, capi.SMALL_STRAIN, capi.material({"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}))
        eqs = capi.Equations(m, code)
        A = capi.Matrix(eqs)
        A.form_stiffness_host(g2, np.zeros_like(X))
        act = eqs.eqnos() > 0
        x, it, rn = A.pcg_host(fext[act], rtol=1e-12, max_iter=20000)
        xs = np.zeros_like(X)
        xs[act] = x
        A.close(); eqs.close(); g2.close(); ex.close(); g.close(); m.close()
        return out + (xs, it)

    if world == 1:
        X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
        a, b = explicit(X, conn, ns, None), explicit(X, conn, ns, None)
        same = all(np.array_equal(p, q) for p, q in zip(a[:4], b[:4]))
        if not same:
            raise SystemExit("bench.py: two runs of the same steps differ bitwise")
        return {"kind": "rerun", "bitwise_rerun": True, "note": "N = 1: 8 explicit steps + one PCG solve of a 16^3 cube run twice, identical bits; "
                "parity against the oracle is tests/ (-m gpu), incl. configs[1] at full size"}
    gx, gy, gz = tmesh.brick_grid(world)
    dims = (n * gx, n * gy, n * gz)
    part = tmesh.partition_cube(*dims, world, rank, jitter=0.1)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    d, v, a, xs, it = explicit(part["coords"], part["conn"], part["nodesets"],
                               (rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"]))
    out = [None] * world
    dist.all_gather_object(out, {"gid": part["node_gid"], "d": d, "v": v, "a": a, "xs": xs, "it": it})
    res = None
    if rank == 0:
        X, conn, ns = tmesh.structured_cube(*dims, jitter=0.1)
        d1, v1, a1, xs1, it1 = explicit(X, conn, ns, None)
        worst, worst_pcg, bitwise, seen = 0.0, 0.0, True, {}
        for o in out:
            for got, ref in ((o["d"], d1), (o["v"], v1), (o["a"], a1)):
                worst = max(worst, float(np.abs(got - ref[o["gid"]]).max() / max(np.abs(ref).max(), 1e-300)))
            worst_pcg = max(worst_pcg, float(np.abs(o["xs"] - xs1[o["gid"]]).max() / np.abs(xs1).max()))
            rows = np.hstack([o["d"], o["v"], o["a"]])
            for gid, row in zip(o["gid"], rows):
                prev = seen.get(gid)
                if prev is not None and not np.array_equal(prev, row):
                    bitwise = False
                seen[gid] = row
        res = {"kind": "partitioned vs single GPU", "max_rel": worst, "bitwise_sharers": bitwise, "pcg_max_rel": worst_pcg,
               "pcg_iterations": {"single": int(it1), "partitioned": [int(o["it"]) for o in out]},
               "workload": "%dx%dx%d elements in %d bricks, 8 explicit steps (d, v, a) and one Jacobi-PCG solve to 1e-12" % (dims + (world,))}
    flag = torch.tensor([0.0], device="cuda", dtype=torch.float64)
    if rank == 0 and (not res["max_rel"] < 1e-10 or not res["bitwise_sharers"] or not res["pcg_max_rel"] < 1e-8):
        flag += 1.0
    dist.all_reduce(flag)
    if float(flag.item()) > 0:
        if rank == 0:
            print("bench.py: multi-GPU parity check FAILED: %s" % json.dumps(res), file=sys.stderr)
        raise SystemExit(3)
    return res


def run_shuffled(torch, capi, tmesh, local, n, steps, dt):
    """the headline workload on a node- and element-shuffled copy of the mesh (no gather locality to rely on): element-updates/s"""
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    rng = np.random.default_rng(7)
    nperm = rng.permutation(X.shape[0])
    Xs = np.empty_like(X)
    Xs[nperm] = X
    conn_s = np.ascontiguousarray(nperm[conn][rng.permutation(conn.shape[0])].astype(np.int32))
    m = capi.Mesh(Xs, conn_s, device=local)
    g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material(MATERIAL))
    ex = capi.Explicit(g)
    code = np.zeros(Xs.shape, np.uint8)
    code[nperm[ns[1]]] = 1
    ex.set_bc(code, np.zeros_like(Xs), np.zeros_like(Xs))
    ex.set_state(initial_displacement(Xs), np.zeros_like(Xs), np.zeros_like(Xs))
    stream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local))
    ex.run(dt, 5)
    m.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ex.run(dt, steps)
    e1.record(stream)
    m.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d, v, a = ex.get_state()
    ok = bool(np.isfinite(d).all())
    ex.close(); g.close(); m.close()
    if not ok:
        raise SystemExit("bench.py: non-finite state on the shuffled mesh")
    # the same shuffled mesh after the host-side locality renumbering a caller can apply before tb2_mesh_create (tahoe_b200.mesh.renumber)
    t0 = time.perf_counter()
    Xr, conn_r, ns_r, new_of_old = tmesh.renumber(Xs, conn_s, {1: nperm[ns[1]]})
    t_renumber = time.perf_counter() - t0
    m = capi.Mesh(Xr, conn_r, device=local)
    g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material(MATERIAL))
    ex = capi.Explicit(g)
    code = np.zeros(Xr.shape, np.uint8)
    code[ns_r[1]] = 1
    ex.set_bc(code, np.zeros_like(Xr), np.zeros_like(Xr))
    u_r = np.empty_like(Xr)
    u_r[new_of_old] = initial_displacement(Xs)  # the field of the shuffled run, node by node (its noise term goes by node number)
    ex.set_state(u_r, np.zeros_like(Xr), np.zeros_like(Xr))
    stream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local))
    ex.run(dt, 5)
    m.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ex.run(dt, steps)
    e1.record(stream)
    m.synchronize()
    torch.cuda.synchronize()
    ms_r = e0.elapsed_time(e1)
    dr = ex.get_state()[0]
    ex.close(); g.close(); m.close()
    # both runs took 5 + steps steps from the same physical state: the renumbered result, mapped back, is the shuffled one
    same = float(np.abs(dr[new_of_old] - d).max() / max(np.abs(d).max(), 1e-300))
    if not same < 1e-10:
        raise SystemExit("bench.py: the renumbered mesh gives a different answer (%.3e)" % same)
    return {"value": conn.shape[0] * steps / (ms * 1e-3), "unit": METRIC, "ms_per_step": ms / steps, "steps": steps,
            "workload": "the same %d^3 cube with node and element numbers permuted at random (seed 7)" % n,
            "after_locality_renumbering": {"value": conn.shape[0] * steps / (ms_r * 1e-3), "ms_per_step": ms_r / steps,
                                           "host_seconds": t_renumber, "max_rel_difference_mapped_back": same,
                                           "note": "tahoe_b200.mesh.renumber on the host before tb2_mesh_create: nodes by (z cell, y cell, x), "
                                                   "elements by the same key of their centroid"}}


def run_contact(torch, capi, tmesh, local, n, steps):
    """two stacked n^3 cubes, the upper one coming down on the clamped lower one (the level.5 impact benchmark's shape),
    contact_3D_penalty with friction and viscous damping inside the resident step: element-updates/s with the contact force re-formed
    every step, the same run without it, and the two contact kernels' own time"""
    Xl, cl, nsl = tmesh.structured_cube(n, jitter=0.0)
    X = np.vstack([Xl, Xl + np.array([0.0, 0.0, 1.0])])
    conn = np.vstack([cl, cl + Xl.shape[0]]).astype(np.int32)
    nn_l, px = Xl.shape[0], n + 1
    j, i = [a.ravel() for a in np.meshgrid(np.arange(n), np.arange(n), indexing="ij")]
    top = lambda ii, jj: n * px * px + jj * px + ii
    bot = lambda ii, jj: nn_l + jj * px + ii
    pairs = np.concatenate([np.stack([top(i, j), top(i + 1, j), top(i + 1, j + 1), bot(i + 1, j)], axis=1),
                            np.stack([top(i, j), top(i + 1, j + 1), top(i, j + 1), bot(i, j + 1)], axis=1)]).astype(np.int32)
    area = np.full(pairs.shape[0], 1.0 / n ** 2)
    mat = {"type": "Simo_isotropic", "density": 1.0, "kappa": 1000.0, "mu": 400.0}
    dt = 0.25 / n / np.sqrt((1000.0 + 4.0 * 400.0 / 3.0) / 1.0)
    code = np.zeros(X.shape, np.uint8)
    code[nsl[5]] = 1
    v0 = np.zeros_like(X)
    v0[nn_l:] = [0.4, 0.0, -5.0]
    m = capi.Mesh(X, conn, device=local)
    g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material(mat))
    ex = capi.Explicit(g)
    contact = capi.Contact(m, 2000.0, 0.3, 1e-3, 20.0)
    contact.set_pairs(pairs, area)
    ex.set_bc(code, np.zeros_like(X), np.zeros_like(X))
    stream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local))
    out = {}
    for tag, att in (("without_contact", None), ("with_contact", contact)):
        ex.attach_contact(att)
        ex.set_state(np.zeros_like(X), v0, np.zeros_like(X))
        ex.run(dt, 5)
        m.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ex.run(dt, steps)
        e1.record(stream)
        m.synchronize()
        torch.cuda.synchronize()
        out[tag] = e0.elapsed_time(e1) / steps
    ncontact, hmax = contact.tracking()
    m.profile_begin()
    ex.run(dt, 20)
    ms_cat, cnt_cat, _ = m.profile_end()
    d = ex.get_state()[0]
    ok = bool(np.isfinite(d).all())
    # the same impact with the pair list maintained by the search on the device (both faces strike each other, as in the reference's
    # level.5 benchmark): tb2_explicit_run searches before the first step and after every step
    px = n + 1
    jj, ii = [a.ravel() for a in np.meshgrid(np.arange(px), np.arange(px), indexing="ij")]
    lower = np.concatenate([np.stack([top(i, j), top(i + 1, j), top(i + 1, j + 1)], axis=1), np.stack([top(i, j), top(i + 1, j + 1), top(i, j + 1)], axis=1)])
    upper = np.concatenate([np.stack([bot(i, j), bot(i + 1, j + 1), bot(i + 1, j)], axis=1), np.stack([bot(i, j), bot(i, j + 1), bot(i + 1, j + 1)], axis=1)])
    facets = np.concatenate([lower, upper]).astype(np.int32)
    surf = np.concatenate([np.zeros(len(lower), np.int32), np.ones(len(upper), np.int32)])
    strikers = np.concatenate([top(ii, jj), bot(ii, jj)]).astype(np.int32)
    w = np.ones(px)
    w[[0, -1]] = 0.5
    sarea = np.tile(np.outer(w, w).ravel() / n ** 2, 2)
    searching = capi.Contact(m, 2000.0, 0.3, 1e-3, 20.0)
    searching.set_surfaces(facets, surf, strikers, sarea)
    ex.attach_contact(searching)
    v0 = v0 * 0.1  # a gentler approach: the search keeps a striker only while it is within half a facet size of its facet (Contact3DT::Intersect)
    ex.set_state(np.zeros_like(X), v0, np.zeros_like(X))
    ex.run(dt, 5)
    m.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ex.run(dt, steps)
    e1.record(stream)
    m.synchronize()
    torch.cuda.synchronize()
    ms_search = e0.elapsed_time(e1) / steps
    npairs_search = int(searching.pairs()[0].shape[0])
    ok = ok and bool(np.isfinite(ex.get_state()[0]).all())
    ex.close(); contact.close(); searching.close(); g.close(); m.close()
    if not ok or ncontact == 0:
        raise RuntimeError("contact leg: non-finite state or no pair in contact")
    return {"value": conn.shape[0] / (out["with_contact"] * 1e-3), "unit": METRIC, "ms_per_step": out["with_contact"],
            "ms_per_step_without_contact": out["without_contact"], "contact_kernels_ms_per_step": float(ms_cat[7]) / 20, "pairs": int(pairs.shape[0]),
            "pairs_in_contact": int(ncontact), "deepest_penetration": float(hmax), "steps": steps,
            "with_device_search": {"ms_per_step": ms_search, "value": conn.shape[0] / (ms_search * 1e-3), "strikers": int(strikers.shape[0]),
                                   "facets": int(facets.shape[0]), "pairs_after_last_step": npairs_search,
                                   "note": "two-sided contact, pair list rebuilt by tb2_contact_search after every step (steps run one at a time)"},
            "workload": "two stacked %d^3 cubes (%d elements), contact_3D_penalty K=2000 mu=0.3 c=20 on %d striker-facet pairs, force re-formed on the "
                        "device before every sweep (tb2_explicit_attach_contact)" % (n, conn.shape[0], pairs.shape[0])}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from tahoe_b200 import capi, mesh as tmesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (the env wins over /etc/nccl.conf)
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d" % (args.gpus, world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    # run (and first-touch the pinned host buffers) on the CPUs NVML reports as local to this rank's GPU: the e2e leg's D2H copies
    # then stay on the GPU's own PCIe root / NUMA node instead of all ranks sharing one socket's path (round 1: 23 % e2e efficiency at N = 8)
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local).uuid)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.lib()
    n = args.n
    parity = None if args.no_parity else parity_check(torch, dist, capi, tmesh, rank, world, local)
    # ---- mesh: one n^3 brick per rank (weak scaling)
    if world == 1:
        X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
        part = None
    else:
        gx, gy, gz = tmesh.brick_grid(world)
        part = tmesh.partition_cube(n * gx, n * gy, n * gz, world, rank, jitter=0.1)
        X, conn, ns = part["coords"], part["conn"], part["nodesets"]
    ne_local, nn_local = conn.shape[0], X.shape[0]
    m = capi.Mesh(X, conn, device=local)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm_init(m, dist, (rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"]))
    g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material(MATERIAL))
    ex = capi.Explicit(g)
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1  # x = 0 face clamped
    fext = np.zeros_like(X)
    if world == 1:
        fext[ns[2], 0] = 0.02 / (n * n)  # nodal share of a 0.02 traction (<< mu) on the x = 1 face: the same load at every mesh size
    u0 = initial_displacement(X)
    ex.set_bc(code, np.zeros_like(X), fext)
    ex.set_state(u0, np.zeros_like(X), np.zeros_like(X))
    dt = float(stable_dt(n * (1 if world == 1 else max(tmesh.brick_grid(world)))))
    stream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local))

    def barrier():
        m.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timed region
    ex.run(dt, args.warmup)
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    if not args.no_profile:
        m.profile_reserve(args.steps * 16 + 64)
    clocks = Clocks(local, uuid)
    if rank == 0:
        clocks.start()  # before the barrier: the sampler's start-up must not skew rank 0 against the others
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not args.no_profile:
        m.profile_begin()
    ev0.record(stream)
    ex.run(dt, args.steps)
    ev1.record(stream)
    if args.no_profile:  # experiment mode: no per-launch events inside the timed region (roofline objects are then meaningless)
        m.synchronize()
        ms_cat, cnt_cat, launches = np.ones(8), np.ones(8, np.int64), 0
    else:
        ms_cat, cnt_cat, launches = m.profile_end()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    d, v, a = ex.get_state()
    if not (np.isfinite(d).all() and np.isfinite(v).all()):
        raise SystemExit("bench.py: non-finite state after the timed region")
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(ne_local)], device="cuda", dtype=torch.float64)
    # per-STEP device time of each kernel = sum of its launches in the step (the slab pipeline launches K1 / K5 in chunks)
    kt = torch.tensor([ms_cat[0] / args.steps, ms_cat[1] / args.steps, ms_cat[6] / args.steps], device="cuda",
                      dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ne_total = float(tot.item())
    value = ne_total * args.steps / (ms * 1e-3)

    # ---- e2e: the resident drop-in's call (tb2_explicit_run_async): every step takes its inputs (the two schedule scalars) from the
    # host and delivers the step's displacement field to pinned host memory; v, a never leave the device.  The D2H copy of step s
    # runs beside the kernels of step s+1 (two host buffers); the host reads each delivered field before its buffer is reused.
    e2e_steps = max(4, min(args.steps, 60))
    hbuf = [torch.zeros(X.shape, dtype=torch.float64).pin_memory() for _ in range(2)]
    one = np.ones(1)

    def e2e_loop(nst):
        tickets, acc = [], 0.0
        for s in range(nst):
            if s >= 2:
                ex.wait(tickets[s - 2])
                acc += float(hbuf[s & 1][-1, 0])  # the host consumes the result of step s-2
            tickets.append(ex.run_async(dt, 1, hbuf[s & 1].data_ptr(), one, one))
        for t in tickets[-2:]:
            ex.wait(t)
        return acc

    e2e_loop(4)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    te = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = ne_total * e2e_steps / (float(te.item()) * 1e-3)
    if not all(np.isfinite(h.numpy()).all() for h in hbuf):
        raise SystemExit("bench.py: non-finite state after the e2e region")
    # the same through tb2_explicit_step_host (Tahoe's FieldT authoritative: d, v, a cross the bus both ways every step), for reference
    hd, hv, ha = (torch.zeros(X.shape, dtype=torch.float64).pin_memory() for _ in range(3))
    d, v, a = ex.get_state()
    hd.copy_(torch.from_numpy(d)); hv.copy_(torch.from_numpy(v)); ha.copy_(torch.from_numpy(a))
    ex.step_host_ptr(dt, hd.data_ptr(), hv.data_ptr(), ha.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        ex.step_host_ptr(dt, hd.data_ptr(), hv.data_ptr(), ha.data_ptr())
    m.synchronize()
    fieldt_ms = 1e3 * (time.perf_counter() - t0) / 5
    barrier()

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        fp64_peak = capi.measure_fp64_peak(local)
        k1_ms, k5_ms, comm_ms = (float(x) for x in kt.tolist())
        # algorithmic bytes (SURVEY.md 8d / DESIGN.md): K1 = 104 B/element (conn 32 + X 24 + u 24 + f 24), K5 = 192 B/node.
        # Per LAUNCH: one launch of the slab pipeline sweeps ne/launches_per_step elements.
        k1_lps = max(int(cnt_cat[0]) // args.steps, 1)
        k5_lps = max(int(cnt_cat[1]) // args.steps, 1)
        k1_launch_ms = k1_ms / k1_lps
        k1_bytes = 104.0 * ne_local / k1_lps
        k1_flops = args.k1_flop_per_element * ne_local / k1_lps
        k1_insts = args.k1_fp64_inst_per_element * ne_local / k1_lps
        roof = {"bound": "hbm", "kernel": "k_internal_force_neo<SimoIso> (K1)", "achieved": k1_bytes / (k1_launch_ms * 1e-3) * 1e-9, "peak": hbm_peak,
                "unit": "GB/s", "frac": k1_bytes / (k1_launch_ms * 1e-3) * 1e-9 / hbm_peak,
                "traffic": args.k1_traffic_bytes_per_element * ne_local / k1_lps, "peak_source": hbm_src,
                "avg_launch_ms": k1_launch_ms, "launches_per_step": k1_lps, "elements_per_launch": ne_local / k1_lps,
                "algorithmic_bytes_per_launch": k1_bytes, "share_of_step": k1_ms * args.steps / ms,
                "note": "K1 is FP64-pipe bound (arithmetic intensity ~%.0f flop/B >> machine balance): see roofline; traffic = dram read+write "
                        "per launch from the ncu --set full capture in profiles/ (per-element figure x elements per launch)" % (args.k1_flop_per_element / 104.0)}
        roof64 = {"bound": "fp64", "kernel": roof["kernel"], "achieved": k1_flops / (k1_launch_ms * 1e-3) * 1e-12, "peak": fp64_peak, "unit": "TFLOP/s",
                  "frac": k1_flops / (k1_launch_ms * 1e-3) * 1e-12 / fp64_peak, "flop_per_element": args.k1_flop_per_element,
                  "fp64_inst_per_element": args.k1_fp64_inst_per_element,
                  "pipe_frac": k1_insts / (k1_launch_ms * 1e-3) * 1e-12 / (0.5 * fp64_peak),
                  "peak_source": "measured in this run (tb2_measure_fp64_peak: dependent DFMA chains; 1 DFMA = 2 flop)",
                  "note": "frac counts real flops (DFMA 2, DMUL/DADD 1) against the all-FMA peak; pipe_frac counts FP64 instructions against the pipe's "
                          "issue rate (peak/2 instructions/s), i.e. what ncu reports as sm__pipe_fp64_cycles_active"}
        k5_launch_ms = k5_ms / k5_lps
        k5_bytes = (123.0 + (24.0 if fext.any() else 0.0)) * nn_local / k5_lps  # d, v in and out, 1/m, codes (+ fext) in: see tb2_node_update.cuh
        roof_k5 = {"bound": "hbm", "kernel": "k_cd_node_update (gather + K5)", "achieved": k5_bytes / (k5_launch_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": k5_bytes / (k5_launch_ms * 1e-3) * 1e-9 / hbm_peak, "avg_launch_ms": k5_launch_ms, "launches_per_step": k5_lps,
                   "share_of_step": k5_ms * args.steps / ms,
                   "note": "algorithmic 123-147 B/node (a stays 0 and fint is written by the last step only); the kernel also reads the "
                           "192 B/element force scratch and the 32 B/node incidence table"}
        step_bytes = 104.0 * ne_local + 192.0 * nn_local
        line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(n, world), "clocks": clk,
                "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": int(16 * world),
                        "d2h_bytes_per_step": int(24 * nn_local * world), "steps": e2e_steps, "ms_per_step": float(te.item()) / e2e_steps,
                        "api": "tb2_explicit_run_async + tb2_explicit_wait: one step per call, the step's schedule scalars in, the step's "
                               "displacement field out to pinned host memory and read by the host; v, a stay resident (the CUDA integrator of "
                               "the plugin); timed by the host clock around the loop incl. the final waits",
                        "fieldt_authoritative_ms_per_step": fieldt_ms,
                        "fieldt_note": "tb2_explicit_step_host: d, v, a in and out every step (%d B each way)" % (72 * nn_local)},
                "gpu_launches": int(launches),
                # the roofline that bounds the dominant kernel is the FP64 pipe (not one of the contract's "hbm" | "tensor": K1 is FP64
                # vector arithmetic at ~34 flop/B); its HBM view, with the ncu traffic figure, is kept beside it
                "roofline": dict(roof64, traffic=roof["traffic"], avg_launch_ms=roof["avg_launch_ms"], launches_per_step=roof["launches_per_step"],
                                 elements_per_launch=roof["elements_per_launch"], share_of_step=roof["share_of_step"],
                                 algorithmic_flop_per_launch=k1_flops, hbm_frac=roof["frac"]),
                "roofline_hbm": roof, "roofline_k5": roof_k5,
                "schedule": ("K1 and K5 back to back on one stream (2 launches per step)" if world == 1 else
                             "two lanes: boundary elements -> partial interface forces %s -> interface nodes on the comm stream beside "
                             "interior elements -> private nodes on the main stream"
                             % ("published to the rank's NVLink-mapped window, pulled and summed by the interface-node kernel of every sharer"
                                if m.comm_peer_enabled() else "packed, ncclAllReduce")),
                "exchange": ("peer" if m.comm_peer_enabled() else "nccl") if world > 1 else None,
                "step_hbm_frac": step_bytes * args.steps / (ms * 1e-3) * 1e-9 / hbm_peak,
                "interface_exchange_ms": comm_ms if world > 1 else None,
                "parity": parity,
                "cpu_baseline": cpu_baseline() if (world == 1 and not args.no_cpu_baseline) else None}
    ex.close(); g.close(); m.close()
    hbm_peak_all = 6650.0
    try:
        hbm_peak_all = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        pass
    pcg = None if args.no_pcg else run_pcg(torch, dist, capi, tmesh, local, rank, world, args.pcg_n, args.pcg_iters, hbm_peak_all)
    xs = None
    if world == 1 and not args.no_explicit_solid:
        xs = run_explicit_solid(torch, capi, tmesh, local, n, min(args.steps, 200), args.warmup, hbm_peak_all, not args.no_cpu_baseline)
    shuffled = run_shuffled(torch, capi, tmesh, local, n, min(args.steps, 100), dt) if (world == 1 and not args.no_shuffled) else None
    contact = None
    if world == 1 and not args.no_contact:
        try:
            contact = run_contact(torch, capi, tmesh, local, 79, min(args.steps, 200))
        except Exception as e:  # a side leg must not cost the line
            contact = {"error": str(e)[-300:]}
    j2 = None
    if world == 1 and args.nlpcg_n > 0:
        j2 = run_nlpcg_j2(torch, capi, tmesh, local, args.nlpcg_n, args.nlpcg_iters)
    nj2 = None
    if world == 1 and args.newton_j2_n > 0:
        try:
            nj2 = run_newton_j2(capi, tmesh, local, args.newton_j2_n)
        except capi.Tb2Error as e:  # a failed load step must not cost the line
            nj2 = {"error": str(e)[-300:]}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and not args.no_plugin:
            try:
                line["plugin"] = plugin_leg()
            except Exception as e:  # the plugin leg must never cost the line
                line["plugin"] = {"error": str(e)[-400:]}
        if shuffled:
            line["shuffled_numbering"] = shuffled
        if contact:
            line["contact"] = contact
        if nj2:
            line["newton_j2"] = nj2
        if j2:
            line["nlpcg_j2"] = j2
        if xs:
            line["explicit_solid"] = xs
        if pcg:
            if world == 1 and not args.no_cpu_baseline:
                pcg["cpu_reference"] = pcg_cpu_reference()
            line["pcg"] = pcg
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--n", "--edge", dest="n", type=int, default=100, help="cube edge in elements per GPU (100 -> 1M elements, configs[1])")
    ap.add_argument("--impl", default="tahoe_b200", choices=["tahoe_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pcg", action="store_true", help="skip the PCG DOF-iters/s leg")
    ap.add_argument("--no-explicit-solid", action="store_true", help="skip the <explicit_solid> leg (SURVEY.md 8f-1)")
    ap.add_argument("--nlpcg-n", type=int, default=48, help="cube edge of the J2 nonlinear-PCG leg (configs[3]; 159 -> 4M elements; 0 = skip)")
    ap.add_argument("--nlpcg-iters", type=int, default=20)
    ap.add_argument("--newton-j2-n", type=int, default=40, help="cube edge of the J2 Newton + BiCGStab leg (configs[3] as stated; 159 -> 4M elements; 0 = skip)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: interface sums pulled over NVLink peer memory inside the consuming kernels (default), or the packed ncclAllReduce")
    ap.add_argument("--no-parity", action="store_true", help="skip the partitioned-vs-single-GPU check that precedes the timing")
    ap.add_argument("--no-shuffled", action="store_true", help="skip the shuffled-numbering leg")
    ap.add_argument("--no-contact", action="store_true", help="skip the contact_3D_penalty leg (SURVEY.md 8f-4)")
    ap.add_argument("--no-plugin", action="store_true", help="skip the plugin-executable leg")
    ap.add_argument("--no-profile", action="store_true", help="experiments only: no per-launch CUDA events in the timed region")
    ap.add_argument("--pcg-n", type=int, default=100, help="cube edge of the implicit small-strain case (100 -> 3.06M equations, 245M nnz)")
    ap.add_argument("--pcg-iters", type=int, default=100)
    # per-element figures of K1 taken from the committed ncu capture (profiles/), see DESIGN.md
    ap.add_argument("--k1-flop-per-element", type=float, default=K1_FLOP_PER_ELEMENT)
    ap.add_argument("--k1-fp64-inst-per-element", type=float, default=K1_FP64_INST_PER_ELEMENT)
    ap.add_argument("--k1-traffic-bytes-per-element", type=float, default=K1_TRAFFIC_BYTES_PER_ELEMENT)
    args = ap.parse_args()
    global EXCHANGE
    EXCHANGE = args.exchange
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: whatever a library prints to fd 1 meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
    sys.stdout.flush()


# Per-element figures of K1 <TL, SimoIso> (k_internal_force_neo) from the ncu --set full capture profiles/r02_k1_details.csv (one launch over
# 10^6 elements): thread-level DFMA 905 + DMUL 539 + DADD 426 = 1870 FP64 instructions = 2775 flop (round 1: 2174 / 3490: the pair-wise
# integration-point loop executes fewer); dram read 81.5 MB + write 135.5 MB = 217 B/element (the 192 B/element force scratch is the bulk).
K1_FLOP_PER_ELEMENT = 2775.0
K1_FP64_INST_PER_ELEMENT = 1870.0
K1_TRAFFIC_BYTES_PER_ELEMENT = 217.0
# K3 element kernel <small strain, SSKStV, symmetric half> (profiles/r01e_k3_*): DFMA 7961 + DMUL 4683 + DADD 2979 per element
K3_FLOP_PER_ELEMENT = 2 * 7961.0 + 4683.0 + 2979.0
# k_spmv DRAM read+write per stored non-zero (ncu --set full, profiles/r01c_spmv_full: 2.465 GB for 242,991,882 nnz)
PCG_SPMV_TRAFFIC_BYTES_PER_NNZ = 10.15

if __name__ == "__main__":
    main()
