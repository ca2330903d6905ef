"""The library's own partitioner (tb2_partition_rcb / tb2_partition_part: host code behind the C ABI, what a C++ host program calls)
against the Python harness's (tahoe_b200/mesh.py), on shuffled meshes and rank counts that are not powers of two: the same element
owners and the same per-rank description, array by array; plus the properties tb2_comm_init relies on."""
import numpy as np
import pytest

from tahoe_b200 import capi, mesh as tmesh


def _mesh(seed, dims=(9, 7, 6)):
    X, conn, ns = tmesh.structured_cube(*dims, jitter=0.15)
    rng = np.random.default_rng(seed)
    nperm = rng.permutation(X.shape[0])
    Xs = np.empty_like(X)
    Xs[nperm] = X
    return Xs, np.ascontiguousarray(nperm[conn][rng.permutation(conn.shape[0])].astype(np.int32))


@pytest.mark.parametrize("nparts", [1, 2, 3, 5, 8])
def test_native_partition_equals_the_harness_partition(nparts):
    capi.lib()
    X, conn = _mesh(nparts)
    owner = capi.partition_rcb(X, conn, nparts)
    assert np.array_equal(owner, tmesh.rcb_element_owner(X, conn, nparts))
    counts = np.bincount(owner, minlength=nparts)
    assert counts.max() - counts.min() <= nparts  # balanced to the rounding of the weighted medians
    slots_seen = {}
    owned_count = np.zeros(X.shape[0], np.int32)
    for rank in range(nparts):
        ref = tmesh.partition_mesh(X, conn, nparts, rank, owner=owner)
        node_gid, elem_gid, lconn, if_nodes, if_slots, nglob, owned = capi.partition_part(X.shape[0], conn, owner, nparts, rank)
        assert np.array_equal(node_gid, ref["node_gid"]) and np.array_equal(elem_gid, ref["elem_gid"])
        assert np.array_equal(lconn, ref["conn"])
        assert np.array_equal(if_nodes, ref["if_nodes"]) and np.array_equal(if_slots, ref["if_slots"])
        assert nglob == ref["n_global_interface"] and np.array_equal(owned, ref["owned"])
        # what the exchange relies on: a slot names the same global node on every sharer, every node has exactly one owner
        for ln, s in zip(if_nodes, if_slots):
            assert slots_seen.setdefault(int(s), int(node_gid[ln])) == int(node_gid[ln])
        owned_count[node_gid[owned > 0]] += 1
    assert (owned_count == 1).all()
    assert len(slots_seen) == (nglob if nparts > 1 else 0)


def test_bad_input_is_rejected():
    capi.lib()
    X, conn = _mesh(0)
    bad = conn.copy()
    bad[3, 2] = X.shape[0]
    with pytest.raises(capi.Tb2Error):
        capi.partition_rcb(X, bad, 2)
    owner = np.zeros(conn.shape[0], np.int32)
    owner[0] = 7
    with pytest.raises(capi.Tb2Error):
        capi.partition_part(X.shape[0], conn, owner, 2, 0)
