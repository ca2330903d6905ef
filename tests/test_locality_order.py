"""tahoe_b200.mesh.locality_order / renumber: a valid renumbering (permutations, same mesh), and on a shuffled structured cube it
restores consecutive node rows -- the property the gathers of K1 / K5 live on."""
import numpy as np

from tahoe_b200 import mesh as tmesh


def _shuffle(X, conn, ns, seed):
    rng = np.random.default_rng(seed)
    nperm = rng.permutation(X.shape[0])
    Xs = np.empty_like(X)
    Xs[nperm] = X
    conn_s = np.ascontiguousarray(nperm[conn][rng.permutation(conn.shape[0])].astype(np.int32))
    return Xs, conn_s, {k: np.sort(nperm[v]).astype(np.int32) for k, v in ns.items()}


def test_renumbering_is_a_permutation_of_the_same_mesh():
    X, conn, ns = tmesh.structured_cube(9, 7, 5, jitter=0.15)
    Xs, cs, nss = _shuffle(X, conn, ns, 3)
    Xr, cr, nsr, new_of_old = tmesh.renumber(Xs, cs, nss)
    assert sorted(new_of_old.tolist()) == list(range(X.shape[0]))
    assert np.array_equal(Xr[new_of_old], Xs)                      # every node keeps its coordinates
    # the same set of elements, each with its nodes in the same local order (HexahedronT node order matters)
    a = {tuple(map(tuple, Xr[e])) for e in cr}
    b = {tuple(map(tuple, X[e])) for e in conn}
    assert a == b
    for k in ns:
        assert np.array_equal(np.sort(Xr[nsr[k]], axis=0), np.sort(X[ns[k]], axis=0))


def test_renumbering_restores_rows_of_consecutive_nodes():
    n = 12
    X, conn, ns = tmesh.structured_cube(n, jitter=0.1)
    Xs, cs, _ = _shuffle(X, conn, ns, 5)

    def spread(c):  # mean distance in node number between an element's first two nodes (x neighbours) and its span
        return np.abs(c[:, 1].astype(np.int64) - c[:, 0]).mean(), (c.max(axis=1).astype(np.int64) - c.min(axis=1)).mean()

    Xr, cr, _, _ = tmesh.renumber(Xs, cs)
    assert spread(conn)[0] == 1.0
    assert spread(cs)[0] > 100.0
    assert spread(cr)[0] == 1.0                                    # x neighbours are consecutive again
    assert spread(cr)[1] == spread(conn)[1]                        # and the element's node span is the structured one
    # elements come in rows too: consecutive elements share a face far more often than not
    shared = np.array([len(set(cr[e]) & set(cr[e + 1])) for e in range(cr.shape[0] - 1)])
    assert (shared == 4).mean() > 0.85


import pytest


@pytest.mark.gpu
def test_renumbered_mesh_gives_the_same_explicit_run():
    """5 explicit steps on a shuffled cube and on its renumbered copy: the same fields node by node (sums run in another order: 1e-10)"""
    from tahoe_b200 import capi
    X, conn, ns = tmesh.structured_cube(10, jitter=0.1)
    Xs, cs, nss = _shuffle(X, conn, ns, 9)
    Xr, cr, nsr, new_of_old = tmesh.renumber(Xs, cs, nss)
    out = []
    for XX, cc, nn in ((Xs, cs, nss), (Xr, cr, nsr)):
        m = capi.Mesh(XX, cc)
        g = capi.Group(m, capi.TOTAL_LAGRANGIAN, capi.material({"type": "Simo_isotropic", "kappa": 1000.0, "mu": 5.0, "density": 1.0}))
        ex = capi.Explicit(g)
        code = np.zeros(XX.shape, np.uint8)
        code[nn[1]] = 1
        fext = np.zeros_like(XX)
        fext[nn[2], 0] = 0.02
        ex.set_bc(code, np.zeros_like(XX), fext)
        ex.set_state(0.01 * XX[:, :1] * XX, np.zeros_like(XX), np.zeros_like(XX))
        ex.run(2e-4, 5)
        out.append(ex.get_state())
        ex.close(); g.close(); m.close()
    for a, b in zip(out[0], out[1]):
        assert np.abs(b[new_of_old] - a).max() < 1e-10 * np.abs(a).max()
