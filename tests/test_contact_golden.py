"""SURVEY 8(f)-4, contact_3D_penalty: the oracle's restatement of PenaltyContact3DT::RHSDriver against the reference's own runs
(tests/golden/ref_contact_*.npz, written by tests/golden/make_golden.py with oracle/_ref/tahoe_dump --contact): the contact group's
FormRHS on the active equations at every dumped state, for the reference's static two-cube case (penalty force only) and its explicit
sliding case with velocity-based Coulomb friction and normal viscous damping."""
import os

import numpy as np
import pytest

import oracle_lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ref_contact_cubes_1", "ref_contact_sliding_friction"]
# further inputs of the reference (a second static case, quasi-static sliding, the damped impact, the Hertz sphere on a mixed mesh):
# they pin the oracle's force and search on the CPU; the device tests run on CASES
ORACLE_CASES = CASES + ["ref_contact_cubes_2", "ref_contact_sliding_3d", "ref_contact_impact_damped", "ref_contact_hertz_explicit"]


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def reference_states(g):
    """(step, pairs, area, d, v or None, contact rhs on the active equations) of every dumped step"""
    for k in g["steps"]:
        v = g.get("ref_v_%d" % k)
        yield int(k), g["ref_cpairs_%d" % k], g["ref_carea_%d" % k], g["ref_d_%d" % k], v, g["ref_crhs_%d" % k]


def to_equations(f, eqnos):
    out = np.zeros(int(eqnos.max()))
    act = eqnos > 0
    out[eqnos[act] - 1] = f[act]
    return out


@pytest.fixture(scope="module")
def oracle():
    oracle_lib.build()
    return oracle_lib


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_contact_force_matches_the_reference(oracle, name):
    g = load(name)
    K, mu, eps, visc = g["ref_cparams"]
    X, eqnos = g["ref_coords"], g["ref_eqnos"]
    seen_contact = 0
    for k, pairs, area, d, v, want in reference_states(g):
        f, n, hmax = oracle.contact_force(pairs, area, X, d, v, K=K, mu=mu, eps=eps, visc=visc)
        got = to_equations(f, eqnos)
        scale = max(np.abs(want).max(), 1e-300)
        assert np.abs(got - want).max() < 1e-12 * scale, (name, k)
        assert hmax <= 0.0 and n <= pairs.shape[0]
        seen_contact += n
    assert seen_contact > 0
    if name == "ref_contact_sliding_friction":
        assert mu > 0 and visc > 0  # the fixture exercises all three force terms
        # ... and they matter: without friction / damping the force differs visibly
        k, pairs, area, d, v, want = list(reference_states(g))[-1]
        f0, _, _ = oracle.contact_force(pairs, area, X, d, None, K=K)
        assert np.abs(to_equations(f0, eqnos) - want).max() > 1e-3 * np.abs(want).max()


def test_contact_force_is_self_equilibrated(oracle):
    """every pair's 12-vector sums to zero force (Newton's third law in all three terms)"""
    g = load("ref_contact_sliding_friction")
    K, mu, eps, visc = g["ref_cparams"]
    k, pairs, area, d, v, _ = list(reference_states(g))[-1]
    f, n, _ = oracle.contact_force(pairs, area, g["ref_coords"], d, v, K=K, mu=mu, eps=eps, visc=visc)
    assert n > 0 and np.abs(f.sum(axis=0)).max() < 1e-12 * np.abs(f).sum()


# ---- the device path (tb2_contact_*) against the same fixtures and against the oracle ---------------------------------------------

@pytest.fixture(scope="module")
def tb2():
    from tahoe_b200 import capi
    capi.lib()
    assert capi.device_count() >= 1
    return capi


def _dummy_mesh(tb2, X):
    """the contact force needs the coordinates only: a one-element connectivity over the first 8 nodes carries them to the device"""
    return tb2.Mesh(X, np.arange(8, dtype=np.int32).reshape(1, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_contact_force_matches_the_reference(tb2, oracle, name):
    g = load(name)
    K, mu, eps, visc = g["ref_cparams"]
    X, eqnos = g["ref_coords"], g["ref_eqnos"]
    mesh = _dummy_mesh(tb2, X)
    contact = tb2.Contact(mesh, K, mu, eps, visc)
    for k, pairs, area, d, v, want in reference_states(g):
        contact.set_pairs(pairs, area)
        f = contact.form_host(d, v)
        got = to_equations(f, eqnos)
        assert np.abs(got - want).max() < 1e-10 * np.abs(want).max(), (name, k)  # BASELINE.json: forces to a relative 1e-10
        f_o, n_o, h_o = oracle.contact_force(pairs, area, X, d, v, K=K, mu=mu, eps=eps, visc=visc)
        assert np.abs(f - f_o).max() < 1e-12 * np.abs(f_o).max()
        assert contact.tracking() == (n_o, h_o)  # pairs in contact and deepest penetration: exact
        assert np.array_equal(f, contact.form_host(d, v))  # ordered sums, no float atomics
        base = np.full_like(X, 0.25)
        assert np.abs(contact.form_host(d, v, out=base.copy()) - (base + f)).max() < 1e-15 * max(np.abs(f).max(), 1.0)  # accumulate


@pytest.mark.gpu
def test_device_contact_force_on_a_large_random_pair_list(tb2, oracle):
    """2 x 10^5 pairs over 10^5 nodes with many pairs per node (a striker in several pairs, facets shared): open and closed gaps,
    friction and damping -- against the oracle entry by entry, tracking data exact, an empty list gives zero force"""
    rng = np.random.default_rng(11)
    nn, npairs = 100_000, 200_000
    X = rng.random((nn, 3)) * 20.0
    u = 0.01 * rng.standard_normal((nn, 3))
    v = rng.standard_normal((nn, 3))
    base = rng.integers(0, nn - 4, npairs)
    pairs = np.stack([base, base + 1, base + 2, base + 3], axis=1).astype(np.int32)
    area = rng.random(npairs) + 0.5
    mesh = _dummy_mesh(tb2, X)
    contact = tb2.Contact(mesh, 500.0, 0.3, 1e-3, 20.0)
    contact.set_pairs(pairs, area)
    f = contact.form_host(u, v, constKd=1.0)
    f_o, n_o, h_o = oracle.contact_force(pairs, area, X, u, v, K=500.0, mu=0.3, eps=1e-3, visc=20.0)
    assert 0.2 * npairs < n_o < 0.8 * npairs  # both branches are exercised
    assert np.abs(f - f_o).max() < 1e-11 * np.abs(f_o).max()
    assert contact.tracking() == (n_o, h_o)
    contact.set_pairs(np.zeros((0, 4), np.int32), np.zeros(0))
    assert not contact.form_host(u, v).any() and contact.tracking() == (0, 0.0)
    with pytest.raises(tb2.Tb2Error):
        contact.set_pairs(np.array([[0, 1, 2, nn]], np.int32), np.ones(1))  # node out of range
    with pytest.raises(tb2.Tb2Error):
        contact.form_host(u)  # friction without velocities


@pytest.mark.gpu
def test_resident_explicit_step_with_contact_matches_the_oracle_loop(tb2, oracle):
    """Impact of two stacked cubes in the resident explicit step (tb2_explicit_attach_contact): the upper cube's bottom nodes strike
    the triangulated top face of the clamped lower cube; 60 steps of predictor / element sweep / contact force on the predicted d, v /
    corrector against the same loop written with the oracle's functions, friction and damping on."""
    from tahoe_b200 import mesh as tmesh
    n = 4
    Xl, cl, nsl = tmesh.structured_cube(n, jitter=0.0)
    Xu = Xl + np.array([0.0, 0.0, 1.0])  # zero initial gap
    X = np.vstack([Xl, Xu])
    conn = np.vstack([cl, cl + Xl.shape[0]]).astype(np.int32)
    nn_l = Xl.shape[0]
    px = n + 1
    top = lambda i, j: n * px * px + j * px + i               # lower cube, z = 1 face
    bot = lambda i, j: nn_l + j * px + i                      # upper cube, z = 1 face (its bottom)
    pairs = []
    for j in range(n):
        for i in range(n):
            pairs.append([top(i, j), top(i + 1, j), top(i + 1, j + 1), bot(i + 1, j)])      # normal (x2-x1) x (x3-x1) = +z
            pairs.append([top(i, j), top(i + 1, j + 1), top(i, j + 1), bot(i, j + 1)])
    pairs = np.asarray(pairs, np.int32)
    area = np.full(len(pairs), 1.0 / n ** 2)
    desc = {"type": "Simo_isotropic", "kappa": 1000.0, "mu": 400.0, "density": 1.0}
    K, mu, eps, visc = 2000.0, 0.3, 1e-3, 20.0
    code = np.zeros(X.shape, np.uint8)
    code[nsl[5]] = 1  # bottom face of the lower cube clamped
    v0 = np.zeros_like(X)
    v0[nn_l:] = [0.4, 0.0, -5.0]  # the upper cube comes down sliding
    dt, nsteps = 2.0e-4, 60
    mesh = tb2.Mesh(X, conn)
    grp = tb2.Group(mesh, tb2.TOTAL_LAGRANGIAN, tb2.material(desc))
    ex = tb2.Explicit(grp)
    contact = tb2.Contact(mesh, K, mu, eps, visc)
    contact.set_pairs(pairs, area)
    ex.attach_contact(contact)
    ex.set_bc(code, np.zeros_like(X), np.zeros_like(X))
    ex.set_state(np.zeros_like(X), v0, np.zeros_like(X))
    ex.run(dt, nsteps)
    d, v, a = ex.get_state()
    ncontact, hmax = contact.tracking()
    # the same loop with the oracle
    omat = oracle.material(desc)
    mass = oracle.lumped_mass(1.0, conn, X)
    d0, w0, a0 = np.zeros_like(X), v0.copy(), np.zeros_like(X)
    seen = 0
    for _ in range(nsteps):
        oracle.cd_predictor(dt, d0, w0, a0, code, np.zeros_like(X))
        err, fi = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, omat, conn, X, d0)
        assert err == 0
        fc, nc, hm = oracle.contact_force(pairs, area, X, d0, w0, K=K, mu=mu, eps=eps, visc=visc)
        seen += nc
        oracle.cd_corrector(dt, w0, a0, fc - fi, mass, code)
    assert seen > 10 * nsteps and nc > 0 and hm < 0.0  # the cubes are in contact through the run
    assert (ncontact, hmax) == (nc, hm)  # the tracking data of the last step
    for got, want in ((d, d0), (v, w0), (a, a0)):
        assert np.abs(got - want).max() < TOL_FIELD * np.abs(want).max()
    # the contact matters: without it the result differs at first order
    ex.attach_contact(None)
    ex.set_state(np.zeros_like(X), v0, np.zeros_like(X))
    ex.run(dt, nsteps)
    assert np.abs(ex.get_state()[0] - d0).max() > 1e-2 * np.abs(d0).max()


TOL_FIELD = 1.0e-10  # BASELINE.json: fields to a relative 1e-10


# ---- the search (Contact3DT::SetActiveStrikers) -------------------------------------------------------------------------------------

def _pair_set(pairs):
    return sorted(map(tuple, np.asarray(pairs).tolist()))


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_contact_search_matches_the_reference_pair_lists(oracle, name):
    """the reference's own search (grid + Intersect) left the pair list of every dumped step on the configuration X + d of that step:
    the oracle's search finds the same striker-facet pairs (as a set: the reference's row order follows its grid traversal)"""
    g = load(name)
    X = g["ref_coords"]
    for k, pairs, area, d, v, _ in reference_states(g):
        hit, gap = oracle.contact_search(g["ref_cfacets"], g["ref_cfacet_surface"], g["ref_cstrikers"], X + d)
        got, got_area = oracle.pairs_from_hits(g["ref_cfacets"], g["ref_cstrikers"], g["ref_cstriker_area"], hit)
        assert _pair_set(got) == _pair_set(pairs), (name, k)
        # the areas ride along with their strikers
        want_area = dict(zip(pairs[:, 3].tolist(), area.tolist()))
        assert all(want_area[s] == a for s, a in zip(got[:, 3].tolist(), got_area.tolist()))


def _two_cube_surfaces(n):
    """two stacked n^3 cubes (zero gap): mesh, triangulated contact surfaces (top of the lower cube = surface 0, bottom of the upper
    cube = surface 1, both with outward normals), strikers = the nodes of both faces, tributary areas"""
    from tahoe_b200 import mesh as tmesh
    Xl, cl, nsl = tmesh.structured_cube(n, jitter=0.0)
    X = np.vstack([Xl, Xl + np.array([0.0, 0.0, 1.0])])
    conn = np.vstack([cl, cl + Xl.shape[0]]).astype(np.int32)
    nn_l, px = Xl.shape[0], n + 1
    top = lambda i, j: n * px * px + j * px + i
    bot = lambda i, j: nn_l + j * px + i
    facets, surf = [], []
    for j in range(n):
        for i in range(n):
            facets += [[top(i, j), top(i + 1, j), top(i + 1, j + 1)], [top(i, j), top(i + 1, j + 1), top(i, j + 1)]]      # normal +z
            surf += [0, 0]
    for j in range(n):
        for i in range(n):
            facets += [[bot(i, j), bot(i + 1, j + 1), bot(i + 1, j)], [bot(i, j), bot(i, j + 1), bot(i + 1, j + 1)]]      # normal -z
            surf += [1, 1]
    strikers = np.array([top(i, j) for j in range(px) for i in range(px)] + [bot(i, j) for j in range(px) for i in range(px)], np.int32)
    w = np.ones(px)
    w[[0, -1]] = 0.5
    area = np.concatenate([np.outer(w, w).ravel(), np.outer(w, w).ravel()]) / n ** 2
    code = np.zeros(X.shape, np.uint8)
    code[nsl[5]] = 1
    return X, conn, np.asarray(facets, np.int32), np.asarray(surf, np.int32), strikers, area, code, nn_l


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_contact_search_matches_the_reference_pair_lists(tb2, oracle, name):
    g = load(name)
    X = g["ref_coords"]
    mesh = _dummy_mesh(tb2, X)
    K, mu, eps, visc = g["ref_cparams"]
    contact = tb2.Contact(mesh, K, mu, eps, visc)
    contact.set_surfaces(g["ref_cfacets"], g["ref_cfacet_surface"], g["ref_cstrikers"], g["ref_cstriker_area"])
    for k, pairs, area, d, v, want in reference_states(g):
        n = contact.search_host(d)
        got, got_area = contact.pairs()
        assert n == pairs.shape[0] and _pair_set(got) == _pair_set(pairs), (name, k)
        hit, _ = oracle.contact_search(g["ref_cfacets"], g["ref_cfacet_surface"], g["ref_cstrikers"], X + d)
        o_pairs, o_area = oracle.pairs_from_hits(g["ref_cfacets"], g["ref_cstrikers"], g["ref_cstriker_area"], hit)
        assert np.array_equal(got, o_pairs) and np.array_equal(got_area, o_area)  # striker order: identical rows
        # and the force on the searched list is the reference's
        f = contact.form_host(d, v)
        assert np.abs(to_equations(f, g["ref_eqnos"]) - want).max() < 1e-10 * np.abs(want).max()


@pytest.mark.gpu
def test_resident_impact_with_the_search_on_the_device(tb2, oracle):
    """two cubes, both faces strikers of each other (two-sided contact as in the reference's impact benchmark): tb2_explicit_run searches
    before the first step and after every step; the same loop with the oracle's search and force gives the same fields"""
    n = 4
    X, conn, facets, surf, strikers, area, code, nn_l = _two_cube_surfaces(n)
    desc = {"type": "Simo_isotropic", "kappa": 1000.0, "mu": 400.0, "density": 1.0}
    K, mu, eps, visc = 2000.0, 0.3, 1e-3, 20.0
    v0 = np.zeros_like(X)
    v0[nn_l:] = [0.4, 0.0, -5.0]
    dt, nsteps = 2.0e-4, 50
    mesh = tb2.Mesh(X, conn)
    grp = tb2.Group(mesh, tb2.TOTAL_LAGRANGIAN, tb2.material(desc))
    ex = tb2.Explicit(grp)
    contact = tb2.Contact(mesh, K, mu, eps, visc)
    contact.set_surfaces(facets, surf, strikers, area)
    ex.attach_contact(contact)
    ex.set_bc(code, np.zeros_like(X), np.zeros_like(X))
    ex.set_state(np.zeros_like(X), v0, np.zeros_like(X))
    ex.run(dt, nsteps)
    d, v, a = ex.get_state()
    omat = oracle.material(desc)
    mass = oracle.lumped_mass(1.0, conn, X)
    d0, w0, a0 = np.zeros_like(X), v0.copy(), np.zeros_like(X)
    hit, _ = oracle.contact_search(facets, surf, strikers, X + d0)
    npairs_seen = []
    for _ in range(nsteps):
        pairs, parea = oracle.pairs_from_hits(facets, strikers, area, hit)
        oracle.cd_predictor(dt, d0, w0, a0, code, np.zeros_like(X))
        err, fi = oracle.internal_force(oracle.TOTAL_LAGRANGIAN, omat, conn, X, d0)
        assert err == 0
        fc, nc, hm = oracle.contact_force(pairs, parea, X, d0, w0, K=K, mu=mu, eps=eps, visc=visc)
        oracle.cd_corrector(dt, w0, a0, fc - fi, mass, code)
        hit, _ = oracle.contact_search(facets, surf, strikers, X + d0)
        npairs_seen.append(pairs.shape[0])
    assert max(npairs_seen) > 10 and nc > 0
    got_pairs, _ = contact.pairs()
    want_pairs, _ = oracle.pairs_from_hits(facets, strikers, area, hit)
    assert np.array_equal(got_pairs, want_pairs)  # the list after the last step
    for got, want in ((d, d0), (v, w0), (a, a0)):
        assert np.abs(got - want).max() < TOL_FIELD * np.abs(want).max()
