"""Multi-GPU parity script (run under torchrun, one rank per GPU; launched by tests/test_gpu_multi.py):
explicit central difference on an element-partitioned cube with interface-node force sums over the sharers must reproduce the
single-GPU run of the whole cube (rank 0 computes it on its own device) to 1e-12, and all sharers of an interface node
must hold bitwise identical values.  argv[1] = "peer" (default: the exchange over NVLink peer memory, tb2_comm_peer_*) or
"nccl" (the packed ncclAllReduce)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from tahoe_b200 import capi, mesh as tmesh  # noqa: E402

MAT = {"type": "Simo_isotropic", "kappa": 1000.0, "mu": 5.0, "density": 1.0}


def field(X):
    return 0.01 * X @ np.array([[0.3, -0.2, 0.1], [0.05, 0.4, -0.3], [0.2, 0.1, -0.25]]) + 1e-3 * np.sin(7.0 * X[:, ::-1])


EXCHANGE = sys.argv[1] if len(sys.argv) > 1 else "peer"


def gather_bytes(b):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, b)
    return out


def setup(X, conn, ns, device, comm=None):
    m = capi.Mesh(X, conn, device=device)
    if comm:
        m.comm_init(*comm, all_gather=gather_bytes if EXCHANGE == "peer" else None)
        assert m.comm_peer_enabled() == (EXCHANGE == "peer"), "peer-memory exchange could not be set up on this box"
    g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material(MAT))
    ex = capi.Explicit(g)
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 1e-3  # on interface nodes every sharer holds the full nodal load (it is not summed)
    ex.set_bc(code, np.zeros_like(X), fext)
    ex.set_state(field(X), np.zeros_like(X), np.zeros_like(X))
    return m, g, ex


def static_pcg(m, X, ns):
    """small-strain elastic solve K x = f with the sub-domain matrix of this rank (tb2_matrix_pcg picks the distributed path
    when the mesh has a communicator): x on the local nodes, iteration count"""
    g = capi.Group(m, capi.SMALL_STRAIN, capi.material({"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    eqs = capi.Equations(m, code)
    A = capi.Matrix(eqs)
    A.form_stiffness_host(g, np.zeros_like(X))
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 1e-3   # consistent on all sharers
    fext[ns[2], 2] = -4e-4
    act = eqs.eqnos() > 0
    x, it, rn = A.pcg_host(fext[act], rtol=1e-13, max_iter=20000)
    d = np.zeros_like(X)
    d[act] = x
    A.close(); eqs.close(); g.close()
    return d, it


def dynamic_pcg(m, X, ns):
    """effective system of an implicit step, (M + beta dt^2 K) x = f, from the rank's unassembled sub-domain matrices
    (tb2_form_stiffness + tb2_matrix_scale + tb2_form_mass with a consistent mass): x on the local nodes, iteration count"""
    g = capi.Group(m, capi.SMALL_STRAIN, capi.material({"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    eqs = capi.Equations(m, code)
    A = capi.Matrix(eqs)
    A.form_stiffness_host(g, np.zeros_like(X))
    A.scale(0.25 * 0.05 ** 2)
    A.form_mass(g, 1, 1.0)
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 1e-3
    fext[ns[4], 1] = -2e-4
    act = eqs.eqnos() > 0
    x, it, rn = A.pcg_host(fext[act], rtol=1e-13, max_iter=20000)
    d = np.zeros_like(X)
    d[act] = x
    A.close(); eqs.close(); g.close()
    return d, it


def nonlinear_solves(m, X, ns):
    """the two resident nonlinear drivers on this rank's sub-domain (the library takes the distributed path when the mesh has a
    communicator): PCGSolver_LS twin and Newton + Jacobi-PCG on a total-Lagrangian Neo-Hookean block.  Returns displacements and
    iteration counts."""
    g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material({"type": "Simo_isotropic", "E": 100.0, "nu": 0.25, "density": 1.0}))
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    eqs = capi.Equations(m, code)
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 2e-3  # full nodal load on every sharer
    fext[ns[2], 1] = 5e-4
    solver = capi.NonlinearPCG(g, eqs, capi.nlpcg_params(restart=30, line_search_iterations=10, line_search_tolerance=0.1,
                                                         rel_tolerance=1e-10, abs_tolerance=1e-14, max_iterations=2000))
    d_cg = np.zeros_like(X)
    st, it_cg, err, err0 = solver.solve_host(d_cg, fext)
    assert st == solver.CONVERGED, "nonlinear PCG status %d" % st
    A = capi.Matrix(eqs)
    d_nw = np.zeros_like(X)
    st, it_nw, err, err0, lin = capi.newton_solve_host(solver, A, capi.newton_params(abs_tolerance=1e-14, rel_tolerance=1e-10,
                                                                                      pcg_rel_tolerance=1e-13), d_nw, fext)
    assert st == 1, "Newton status %d" % st
    A.close(); solver.close(); eqs.close(); g.close()
    return d_cg, it_cg, d_nw, it_nw


def general_mesh_phase(rank, world, local):
    """a mesh with shuffled node and element numbering, partitioned by recursive coordinate bisection (tmesh.partition_mesh):
    explicit steps on the parts must reproduce the single-GPU run of the whole mesh"""
    X, conn, ns = tmesh.structured_cube(14, 11, 9, jitter=0.15)
    rng = np.random.default_rng(5)
    nperm = rng.permutation(X.shape[0])
    Xs = np.empty_like(X)
    Xs[nperm] = X
    conn_s = np.ascontiguousarray(nperm[conn][rng.permutation(conn.shape[0])].astype(np.int32))
    ns_s = {k: np.sort(nperm[v]).astype(np.int32) for k, v in ns.items()}
    # element owners from the library's own partitioner (tb2_partition_rcb), the rank's description from the harness
    part = tmesh.partition_mesh(Xs, conn_s, world, rank, ns_s, owner=capi.partition_rcb(Xs, conn_s, world))
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = (rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"])
    m, g, ex = setup(part["coords"], part["conn"], part["nodesets"], local, comm)
    dt, nsteps = 2e-4, 12
    ex.initial_condition()
    ex.run(dt, nsteps)
    d, v, a = ex.get_state()
    out = [None] * world
    dist.all_gather_object(out, {"gid": part["node_gid"], "d": d, "v": v, "a": a})  # also the barrier before the teardown
    ex.close(); g.close(); m.close()
    ok = True
    if rank == 0:
        m1, g1, ex1 = setup(Xs, conn_s, ns_s, local)
        ex1.initial_condition()
        ex1.run(dt, nsteps)
        ref = dict(zip("dva", ex1.get_state()))
        for r, o in enumerate(out):
            for nm in "dva":
                err = np.abs(o[nm] - ref[nm][o["gid"]]).max() / max(np.abs(ref[nm]).max(), 1e-300)
                if not err < 1e-12:
                    print("general mesh phase: rank %d field %s differs from the single-GPU run: %.3e" % (r, nm, err))
                    ok = False
        ex1.close(); g1.close(); m1.close()
        print("multi_gpu_check: general (RCB-partitioned, shuffled) mesh world=%d %s" % (world, "OK" if ok else "FAILED"))
    return ok


def pipelined_phase(rank, world, local):
    """64^3 elements per rank -> 3+ slab chunks: the overlapped schedule (boundary elements first, all-reduce beside the slab
    pipeline, interface nodes updated last) must give bitwise the fields of the serial schedule (single-step calls), and both
    must match the single-GPU run of the whole mesh"""
    n, nsteps = 64, 6
    gx, gy, gz = tmesh.brick_grid(world)
    dims = (n * gx, n * gy, n * gz)
    dt = 0.25 / max(dims) / np.sqrt(1000.0 + 20.0 / 3.0)
    part = tmesh.partition_cube(*dims, world, rank, jitter=0.1)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = (rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"])
    m, g, ex = setup(part["coords"], part["conn"], part["nodesets"], local, comm)
    ex.run(dt, nsteps)
    piped = ex.get_state()
    ex.set_state(field(part["coords"]), np.zeros_like(part["coords"]), np.zeros_like(part["coords"]))
    for _ in range(nsteps):
        ex.run(dt, 1)
    serial = ex.get_state()
    ok = all(np.array_equal(a, b) for a, b in zip(piped, serial))
    if not ok:
        print("rank %d: overlapped schedule differs bitwise from the serial one: %s"
              % (rank, [float(np.abs(a - b).max()) for a, b in zip(piped, serial)]))
    out = [None] * world
    dist.all_gather_object(out, {"gid": part["node_gid"], "d": piped[0], "v": piped[1], "a": piped[2]})
    ex.close(); g.close(); m.close()
    if rank == 0:
        X, conn, ns = tmesh.structured_cube(*dims, jitter=0.1)
        m1, g1, ex1 = setup(X, conn, ns, local)
        ex1.run(dt, nsteps)
        ref = dict(zip("dva", ex1.get_state()))
        for r, o in enumerate(out):
            for nm in "dva":
                err = np.abs(o[nm] - ref[nm][o["gid"]]).max() / max(np.abs(ref[nm]).max(), 1e-300)
                if not err < 1e-12:
                    print("pipelined phase: rank %d field %s differs from the single-GPU run: %.3e" % (r, nm, err))
                    ok = False
        ex1.close(); g1.close(); m1.close()
        print("multi_gpu_check: pipelined phase world=%d %s" % (world, "OK" if ok else "FAILED"))
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = (12, 10, 8)
    dt, nsteps = 2e-4, 25
    part = tmesh.partition_cube(*dims, world, rank, jitter=0.15)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    m, g, ex = setup(part["coords"], part["conn"], part["nodesets"], local,
                     (rank, world, uid[0], part["if_nodes"], part["if_slots"], part["n_global_interface"], part["owned"]))
    ex.initial_condition()
    ex.run(dt, nsteps)
    d, v, a = ex.get_state()
    mass = ex.mass_host()
    xs, its = static_pcg(m, part["coords"], part["nodesets"])
    xd, itd = dynamic_pcg(m, part["coords"], part["nodesets"])
    d_cg, it_cg, d_nw, it_nw = nonlinear_solves(m, part["coords"], part["nodesets"])
    nn_glob = (dims[0] + 1) * (dims[1] + 1) * (dims[2] + 1)
    # gather every rank's fields on rank 0 keyed by global node id
    out = [None] * world
    dist.all_gather_object(out, {"gid": part["node_gid"], "d": d, "v": v, "a": a, "mass": mass, "xs": xs, "its": its, "xd": xd, "itd": itd,
                                 "d_cg": d_cg, "it_cg": it_cg, "d_nw": d_nw, "it_nw": it_nw})
    ok = True
    if rank == 0:
        X, conn, ns = tmesh.structured_cube(*dims, jitter=0.15)
        m1, g1, ex1 = setup(X, conn, ns, local)
        ex1.initial_condition()
        ex1.run(dt, nsteps)
        d1, v1, a1 = ex1.get_state()
        mass1 = ex1.mass_host()
        xs1, its1 = static_pcg(m1, X, ns)
        xd1, itd1 = dynamic_pcg(m1, X, ns)
        d_cg1, it_cg1, d_nw1, it_nw1 = nonlinear_solves(m1, X, ns)
        print("nonlinear PCG iterations: single GPU %d, partitioned %s; Newton: %d, %s"
              % (it_cg1, [o["it_cg"] for o in out], it_nw1, [o["it_nw"] for o in out]))
        print("PCG iterations: single GPU %d, partitioned %s" % (its1, [o["its"] for o in out]))
        seen = {}
        for r, o in enumerate(out):
            err = np.abs(o["xs"] - xs1[o["gid"]]).max() / np.abs(xs1).max()
            if not err < 1e-9 or abs(o["its"] - its1) > 3:
                print("rank %d static PCG solution differs from the single-GPU solve: %.3e (its %d vs %d)" % (r, err, o["its"], its1))
                ok = False
            err = np.abs(o["xd"] - xd1[o["gid"]]).max() / np.abs(xd1).max()
            if not err < 1e-9 or abs(o["itd"] - itd1) > 3:
                print("rank %d implicit-step system (M + beta dt^2 K) differs from the single-GPU solve: %.3e (its %d vs %d)" % (r, err, o["itd"], itd1))
                ok = False
            for nm, ref, cnt, cnt1, tol in (("d_cg", d_cg1, o["it_cg"], it_cg1, 1e-7), ("d_nw", d_nw1, o["it_nw"], it_nw1, 1e-9)):
                err = np.abs(o[nm] - ref[o["gid"]]).max() / np.abs(ref).max()
                if not err < tol or abs(cnt - cnt1) > max(2, 0.1 * cnt1):
                    print("rank %d %s differs from the single-GPU solve: %.3e (iterations %d vs %d)" % (r, nm, err, cnt, cnt1))
                    ok = False
            for nm, ref in (("d", d1), ("v", v1), ("a", a1), ("mass", mass1)):
                err = np.abs(o[nm] - ref[o["gid"]]).max() / max(np.abs(ref).max(), 1e-300)
                if not err < 1e-12:
                    print("rank %d field %s differs from the single-GPU run: %.3e" % (r, nm, err))
                    ok = False
            for gid, row in zip(o["gid"], np.hstack([o["d"], o["v"], o["a"]])):
                if gid in seen and not np.array_equal(seen[gid], row):
                    print("node %d differs bitwise between sharers" % gid)
                    ok = False
                    break
                seen[gid] = row
        assert len(seen) == nn_glob
        print("multi_gpu_check: world=%d exchange=%s %s" % (world, EXCHANGE, "OK" if ok else "FAILED"))
    dist.barrier()  # no rank tears its exchange window down while a peer may still pull from it
    ex.close(); g.close(); m.close()
    ok = general_mesh_phase(rank, world, local) and ok
    ok = pipelined_phase(rank, world, local) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
