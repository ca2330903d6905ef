"""Parity at BASELINE.json's full single-GPU sizes through size-independent properties (the oracle cannot run these sizes in
seconds): configs[1] = 100^3 total-Lagrangian Neo-Hookean elements, explicit; configs[2] = 200^3 small-strain elements with the
device-assembled CSR (24 M equations, 1.94 G non-zeros).  Objectivity, equilibrium of internal forces, the closed-form sparsity
count, rigid-body null space, symmetry, K u = fint(u) for the linear element, CG residual reduction, bit-reproducibility."""
import numpy as np
import pytest

from tahoe_b200 import mesh as tmesh

pytestmark = pytest.mark.gpu
TOL = 1.0e-10  # BASELINE.json: nodal forces and fields to a relative 1e-10 (max-norm over the field / max-norm of the reference)


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def oracle():
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


def _shuffled(X, conn, ns, seed):
    """the same mesh with node and element numbers permuted (worst-case gather locality; the reference accepts any numbering)"""
    rng = np.random.default_rng(seed)
    nperm = rng.permutation(X.shape[0])
    Xs = np.empty_like(X)
    Xs[nperm] = X
    conn_s = np.ascontiguousarray(nperm[conn][rng.permutation(conn.shape[0])].astype(np.int32))
    return Xs, conn_s, {k: np.sort(nperm[v]).astype(np.int32) for k, v in ns.items()}, nperm


@pytest.fixture(scope="module")
def tb2():
    from tahoe_b200 import capi
    capi.lib()
    assert capi.device_count() >= 1
    return capi


def _free_gb():
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    return free / 2 ** 30


def test_c2_explicit_full_size_properties(tb2):
    n = 100
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    assert conn.shape[0] == 10 ** 6
    mesh = tb2.Mesh(X, conn)
    mat = tb2.material({"type": "Simo_isotropic", "kappa": 1000.0, "mu": 5.0, "density": 1.0})
    grp = tb2.Group(mesh, tb2.TOTAL_LAGRANGIAN, mat)
    # objectivity: rigid rotation + translation -> no internal force
    th = 0.4
    Q = np.array([[np.cos(th), 0, np.sin(th)], [0, 1.0, 0], [-np.sin(th), 0, np.cos(th)]])
    f = grp.internal_force_host(X @ (Q.T - np.eye(3)) + np.array([0.3, 0.1, -0.2]))
    assert np.abs(f).max() < 1e-9 * mat.kappa / n ** 2
    # self-equilibrium (zero resultant force and moment) and the patch test for a homogeneous deformation
    u = 0.01 * X @ np.random.default_rng(0).standard_normal((3, 3))
    f = grp.internal_force_host(u)
    scale = np.abs(f).sum()
    assert np.abs(f.sum(axis=0)).max() < 1e-12 * scale
    assert np.abs(np.cross(X + u, f).sum(axis=0)).max() < 1e-11 * scale
    k, j, i = np.meshgrid(np.arange(n + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    interior = ((i > 0) & (i < n) & (j > 0) & (j < n) & (k > 0) & (k < n)).ravel()
    assert np.abs(f[interior]).max() < 1e-9 * np.abs(f).max()
    assert np.array_equal(f, grp.internal_force_host(u))  # no float atomics: identical bits
    # lumped mass: total mass of the unit cube, positivity
    m = grp.lumped_mass_host()
    assert abs(m[:, 0].sum() - 1.0) < 1e-10 and m.min() > 0
    # explicit run: free vibration from a smooth start conserves linear momentum (no external load, no constraints) and total
    # energy to the accuracy of central differences; two runs give the same bits
    ex = tb2.Explicit(grp)
    code = np.zeros(X.shape, np.uint8)
    ex.set_bc(code, np.zeros_like(X), np.zeros_like(X))
    v0 = 1e-2 * np.sin(2 * np.pi * X[:, ::-1])
    dt = 0.25 / n / np.sqrt(1000.0 + 20.0 / 3.0)
    out = []
    for _ in range(2):
        ex.set_state(np.zeros_like(X), v0, np.zeros_like(X))
        ex.run(dt, 50)
        out.append(ex.get_state())
    d, v, a = out[0]
    assert all(np.array_equal(p, q) for p, q in zip(out[0], out[1]))
    p0, p1 = (m * v0).sum(axis=0), (m * v).sum(axis=0)
    assert np.abs(p1 - p0).max() < 1e-10 * np.abs(m * v0).sum()
    assert np.isfinite(d).all() and np.abs(d).max() < 1e-2


@pytest.mark.parametrize("numbering", ["structured", "shuffled"])
def test_c2_full_size_against_the_oracle(tb2, oracle, numbering):
    """configs[1] at its full size, entry by entry against the C oracle (it sweeps 10^6 elements in seconds): internal force of
    the 100^3 total-Lagrangian Neo-Hookean cube and d, v, a after 3 explicit steps, on the generator's numbering and on a node-
    and element-shuffled copy of the mesh"""
    n = 100
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    if numbering == "shuffled":
        X, conn, ns, _ = _shuffled(X, conn, ns, 7)
    desc = {"type": "Simo_isotropic", "kappa": 1000.0, "mu": 5.0, "density": 1.0}
    omat = oracle.material(desc)
    rng = np.random.default_rng(3)
    # a 1 % strain field tapered to zero at the clamped x = 0 face (no displacement jump at the boundary condition) + noise
    u = 0.01 * X[:, :1] * (X @ rng.standard_normal((3, 3))) + 1e-4 / n * rng.standard_normal(X.shape)
    mesh = tb2.Mesh(X, conn)
    grp = tb2.Group(mesh, tb2.TOTAL_LAGRANGIAN, tb2.material(desc))
    err, f_ref = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, omat, conn, X, u)
    assert err == 0 and relerr(grp.internal_force_host(u), f_ref) < TOL
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    fext = np.zeros_like(X)
    fext[ns[2], 0] = 0.02 / n ** 2
    ex = tb2.Explicit(grp)
    ex.set_bc(code, np.zeros_like(X), fext)
    ex.set_state(u, np.zeros_like(X), np.zeros_like(X))
    dt = 0.5 / n / np.sqrt(1000.0 + 20.0 / 3.0)
    ex.run(dt, 3)
    d, v, a = ex.get_state()
    mass = oracle.lumped_mass(1.0, conn, X)
    d0, v0, a0 = u.copy(), np.zeros_like(X), np.zeros_like(X)
    for _ in range(3):
        oracle.cd_predictor(dt, d0, v0, a0, code, np.zeros_like(X))
        _, fi = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, omat, conn, X, d0)
        oracle.cd_corrector(dt, v0, a0, fext - fi, mass, code)
    assert relerr(d, d0) < TOL and relerr(v, v0) < TOL and relerr(a, a0) < TOL


@pytest.mark.parametrize("numbering", ["structured", "shuffled"])
def test_c3_tangent_50_cubed_against_the_oracle(tb2, oracle, numbering):
    """configs[2] family at 50^3 = 125 k elements (384 k equations, 30 M non-zeros): sparsity bit-exact, assembled values to 1e-10"""
    n = 50
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    if numbering == "shuffled":
        X, conn, ns, _ = _shuffled(X, conn, ns, 11)
    desc = {"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}
    code = np.zeros(X.shape, np.uint8)
    code[ns[1]] = 1
    eq, neq = oracle.equation_numbers(code)
    rp, ci = oracle.csr_structure(conn, eq, neq)
    u = np.zeros_like(X)
    err, kv = oracle.assemble_stiffness(oracle.SMALL_STRAIN, oracle.material(desc), conn, X, u, eq, neq, rp, ci)
    mesh = tb2.Mesh(X, conn)
    grp = tb2.Group(mesh, tb2.SMALL_STRAIN, tb2.material(desc))
    A = tb2.Matrix(tb2.Equations(mesh, code))
    A.form_stiffness_host(grp, u)
    rowptr, colind, val = A.csr()
    assert err == 0 and np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
    assert relerr(val, kv) < TOL


def test_c3_static_full_size_properties(tb2):
    if _free_gb() < 60:
        pytest.skip("needs ~45 GB of device memory")
    n = 200
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    mesh = tb2.Mesh(X, conn)
    grp = tb2.Group(mesh, tb2.SMALL_STRAIN, tb2.material({"type": "small_strain_StVenant", "E": 100.0, "nu": 0.25, "density": 1.0}))
    code = np.zeros(X.shape, np.uint8)
    eqs = tb2.Equations(mesh, code)  # no constraints: K keeps the 6 rigid-body modes in its null space
    assert eqs.neq == 3 * (n + 1) ** 3
    A = tb2.Matrix(eqs)
    assert A.nnz == 9 * (3 * (n + 1) - 2) ** 3  # 1.944 G non-zeros, within 8 % of the int32 limit of the reference's MSR
    A.form_stiffness_host(grp, np.zeros_like(X))
    kmax = 100.0 / n
    y = A.multx_host(np.tile([1.0, 2.0, -0.5], X.shape[0]))
    assert np.abs(y).max() < 1e-10 * kmax * 10
    assert np.abs(A.multx_host(np.cross(np.array([0.2, -0.1, 0.4]), X).ravel())).max() < 1e-10 * kmax * 10
    rng = np.random.default_rng(5)
    u = 1e-3 * rng.standard_normal(X.shape)
    Ku = A.multx_host(u.ravel())
    f = grp.internal_force_host(u)
    assert np.abs(Ku.reshape(-1, 3) - f).max() < 1e-11 * np.abs(f).max()  # K u = fint(u): assembly and sweep agree entry by entry
    w = 1e-3 * rng.standard_normal(X.shape).ravel()
    Kw = A.multx_host(w)
    assert abs(w @ Ku - u.ravel() @ Kw) < 1e-12 * abs(w @ Ku)  # symmetry
    assert np.array_equal(Ku, A.multx_host(u.ravel()))  # deterministic SpMV
    # constrained system
    code[ns[1]] = 1
    eqs2 = tb2.Equations(mesh, code)
    A2 = tb2.Matrix(eqs2)
    A2.form_stiffness_host(grp, np.zeros_like(X))
    b = np.zeros_like(X)
    b[ns[2], 0] = 1e-3
    rhs = b[eqs2.eqnos() > 0]
    # CG minimises the energy functional Phi(x) = x.Ax/2 - b.x over growing Krylov spaces: Phi decreases monotonically (the 2-norm
    # of the residual need not), and the recurrence residual the solver reports is the true residual
    phi = []
    for iters in (16, 48):
        x, it, rn = A2.pcg_host(rhs, rtol=0.0, atol=0.0, max_iter=iters)
        assert it == iters and np.isfinite(rn)
        Ax = A2.multx_host(x)
        phi.append(0.5 * x @ Ax - rhs @ x)
        assert abs(np.linalg.norm(rhs - Ax) - rn) < 1e-8 * np.linalg.norm(rhs)
    assert phi[1] < phi[0] < 0.0
