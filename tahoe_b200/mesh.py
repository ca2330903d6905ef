"""Synthetic structured Hex8 meshes and the element partitioner for the multi-GPU path.

Node id = k(ny+1)(nx+1) + j(nx+1) + i, element order k-major / i-minor, connectivity in HexahedronT order
(HexahedronT.cpp:23-25; same layout as benchmark_XML/level.5/explicit_benchmark/generate_3d_mesh.py:47-57).  Interior nodes are
jittered by +-jitter*h with a counter-based hash of (seed, global node id, dof), so any partition of the mesh sees
bit-identical coordinates.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def jitter_u01(node_ids, dof, seed=12345):
    with np.errstate(over="ignore"):
        key = np.uint64(seed) * np.uint64(0x100000001B3) + node_ids.astype(np.uint64) * np.uint64(3) + np.uint64(dof)
    return (_splitmix64(key) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _node_coords(i, j, k, dims, lengths, jitter, seed):
    nx, ny, nz = dims
    h = np.array([lengths[0] / nx, lengths[1] / ny, lengths[2] / nz])
    coords = np.stack([i * h[0], j * h[1], k * h[2]], axis=1).astype(np.float64)
    ids = (k.astype(np.int64) * (ny + 1) + j) * (nx + 1) + i
    if jitter:
        interior = (i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz)
        for d in range(3):
            coords[:, d] += np.where(interior, (jitter_u01(ids, d, seed) - 0.5) * 2.0 * jitter * h[d], 0.0)
    return coords, ids


def structured_cube(nx, ny=None, nz=None, jitter=0.1, seed=12345, lengths=(1.0, 1.0, 1.0)):
    """coords [nn,3], conn [ne,8] int32 0-based, nodesets {1: x=0, 2: x=L, 3: y=0, 4: y=L, 5: z=0, 6: z=L}"""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    part = brick(nx, ny, nz, (0, nx), (0, ny), (0, nz), jitter, seed, lengths)
    return part["coords"], part["conn"], part["nodesets"]


def brick(nx, ny, nz, xr, yr, zr, jitter=0.1, seed=12345, lengths=(1.0, 1.0, 1.0)):
    """the sub-brick of elements [xr) x [yr) x [zr) of the nx*ny*nz cube with local node numbering.
    Returns dict(coords, conn, nodesets (local ids), node_gid [nn_local] global node ids, elem_gid)"""
    (x0, x1), (y0, y1), (z0, z1) = xr, yr, zr
    lx, ly, lz = x1 - x0, y1 - y0, z1 - z0
    px, py = lx + 1, ly + 1
    k, j, i = np.meshgrid(np.arange(z0, z1 + 1), np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    coords, gid = _node_coords(i, j, k, (nx, ny, nz), lengths, jitter, seed)
    ek, ej, ei = np.meshgrid(np.arange(lz), np.arange(ly), np.arange(lx), indexing="ij")
    n0 = (ek * py * px + ej * px + ei).ravel()
    conn = np.stack([n0, n0 + 1, n0 + 1 + px, n0 + px, n0 + px * py, n0 + 1 + px * py, n0 + 1 + px + px * py, n0 + px + px * py],
                    axis=1).astype(np.int32)
    egid = (((ek + z0).astype(np.int64) * ny + (ej + y0)) * nx + (ei + x0)).ravel()
    loc = np.arange(len(i), dtype=np.int32)
    nodesets = {1: loc[i == 0], 2: loc[i == nx], 3: loc[j == 0], 4: loc[j == ny], 5: loc[k == 0], 6: loc[k == nz]}
    return {"coords": coords, "conn": conn, "nodesets": nodesets, "node_gid": gid, "elem_gid": egid, "ijk": (i, j, k)}


def brick_grid(nranks):
    """processor grid (px,py,pz) with px*py*pz = nranks, as cubic as possible (8 -> 2x2x2, 4 -> 1x2x2, 2 -> 1x1x2)"""
    best = None
    for a in range(1, nranks + 1):
        if nranks % a:
            continue
        for b in range(1, nranks // a + 1):
            if (nranks // a) % b:
                continue
            c = nranks // a // b
            key = (max(a, b, c) - min(a, b, c), a, b)
            if best is None or key < best[0]:
                best = (key, (a, b, c))
    return tuple(sorted(best[1]))


def _splits(n, p):
    return [(n * r) // p for r in range(p + 1)]


def partition_cube(nx, ny, nz, nranks, rank, jitter=0.1, seed=12345, lengths=(1.0, 1.0, 1.0)):
    """Element partition of the cube into nranks bricks (METIS-style k-way on a structured mesh degenerates to this).
    Returns the rank's brick plus the interface description tb2_comm_init wants:
      if_nodes  local ids of nodes shared with another rank,
      if_slots  their index in the packed global interface vector (global interface nodes sorted by global node id),
      n_global_interface, owned [nn_local] uint8 (owner = lowest rank touching the node)."""
    gx, gy, gz = brick_grid(nranks)
    sx, sy, sz = _splits(nx, gx), _splits(ny, gy), _splits(nz, gz)
    rx, ry, rz = rank % gx, (rank // gx) % gy, rank // (gx * gy)
    part = brick(nx, ny, nz, (sx[rx], sx[rx + 1]), (sy[ry], sy[ry + 1]), (sz[rz], sz[rz + 1]), jitter, seed, lengths)
    i, j, k = part["ijk"]

    def cut_info(v, cuts):
        """per node: is on an internal cut plane; lowest brick index touching it along this axis"""
        inner = np.asarray(cuts[1:-1])
        on = np.isin(v, inner)
        low = np.searchsorted(np.asarray(cuts), v, side="right") - 1  # brick containing v as lower face
        low = np.where(on, low - 1, np.minimum(low, len(cuts) - 2))
        return on, low

    onx, lowx = cut_info(i, sx)
    ony, lowy = cut_info(j, sy)
    onz, lowz = cut_info(k, sz)
    shared = onx | ony | onz
    owner = (lowz * gy + lowy) * gx + lowx
    part["owned"] = (owner == rank).astype(np.uint8)
    part["if_nodes"] = np.nonzero(shared)[0].astype(np.int32)
    # global interface numbering: closed form count of nodes on cut planes, ordered by global node id.  Computed by
    # enumerating the (small) set of interface nodes of the whole cube plane by plane.
    gi = _global_interface_ids(nx, ny, nz, sx, sy, sz)
    part["n_global_interface"] = len(gi)
    part["if_slots"] = np.searchsorted(gi, part["node_gid"][part["if_nodes"]]).astype(np.int32)
    part["grid"] = (gx, gy, gz)
    return part


def _global_interface_ids(nx, ny, nz, sx, sy, sz):
    px, py, pz = nx + 1, ny + 1, nz + 1
    ids = []
    for c in sx[1:-1]:
        k, j = np.meshgrid(np.arange(pz, dtype=np.int64), np.arange(py, dtype=np.int64), indexing="ij")
        ids.append(((k * py + j) * px + c).ravel())
    for c in sy[1:-1]:
        k, i = np.meshgrid(np.arange(pz, dtype=np.int64), np.arange(px, dtype=np.int64), indexing="ij")
        ids.append(((k * py + c) * px + i).ravel())
    for c in sz[1:-1]:
        j, i = np.meshgrid(np.arange(py, dtype=np.int64), np.arange(px, dtype=np.int64), indexing="ij")
        ids.append(((c * py + j) * px + i).ravel())
    if not ids:
        return np.zeros(0, np.int64)
    return np.unique(np.concatenate(ids))


# ---------------------------------------------------------------------------------------------------------------------
# general meshes: element partition by recursive coordinate bisection (METIS is absent in this image; the reference
# itself ships a non-METIS partitioner, GraphBaseT::Partition)
# ---------------------------------------------------------------------------------------------------------------------
def rcb_element_owner(coords, conn, nranks):
    """owner rank of every element: recursive coordinate bisection of the element centroids along the longest extent, split at
    the weighted median so that rank counts need not be powers of two.  Deterministic (stable sorts, ties by element id)."""
    cent = coords[conn].mean(axis=1)
    owner = np.zeros(conn.shape[0], np.int32)

    def split(ids, r0, nr):
        if nr == 1:
            owner[ids] = r0
            return
        nl = nr // 2
        c = cent[ids]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = ids[np.lexsort((ids, c[:, axis]))]
        cut = (len(ids) * nl) // nr
        split(order[:cut], r0, nl)
        split(order[cut:], r0 + nl, nr - nl)

    split(np.arange(conn.shape[0]), 0, nranks)
    return owner


def partition_mesh(coords, conn, nranks, rank, nodesets=None, owner=None):
    """This rank's part of an arbitrary Hex8 mesh, with the interface description tb2_comm_init wants (same contract as
    partition_cube).  Elements keep their global order inside a part (so summation orders follow the global numbering), local
    nodes are numbered by ascending global id; an interface node is owned by the lowest rank that touches it.
    Every rank computes the same global description from the whole connectivity (host work, once)."""
    owner = rcb_element_owner(coords, conn, nranks) if owner is None else np.asarray(owner, np.int32)
    nn = coords.shape[0]
    # ranks touching each node: bit mask (<= 64 ranks on one box)
    touch = np.zeros(nn, np.uint64)
    for r in range(nranks):
        nodes = np.unique(conn[owner == r])
        touch[nodes] |= np.uint64(1) << np.uint64(r)
    nshare = np.zeros(nn, np.int32)
    lowest = np.full(nn, -1, np.int32)
    for r in range(nranks - 1, -1, -1):
        has = (touch >> np.uint64(r)) & np.uint64(1) != 0
        nshare += has
        lowest[has] = r
    gi = np.nonzero(nshare > 1)[0].astype(np.int64)  # global interface nodes, ascending global id = slot order
    mine = owner == rank
    elem_gid = np.nonzero(mine)[0].astype(np.int64)
    node_gid = np.unique(conn[mine]).astype(np.int64)
    g2l = np.full(nn, -1, np.int32)
    g2l[node_gid] = np.arange(len(node_gid), dtype=np.int32)
    part = {"coords": np.ascontiguousarray(coords[node_gid]), "conn": np.ascontiguousarray(g2l[conn[mine]]).astype(np.int32),
            "node_gid": node_gid, "elem_gid": elem_gid, "owner": owner}
    part["nodesets"] = {k: g2l[v][g2l[v] >= 0].astype(np.int32) for k, v in (nodesets or {}).items()}
    shared_local = np.nonzero(nshare[node_gid] > 1)[0].astype(np.int32)
    part["if_nodes"] = shared_local
    part["if_slots"] = np.searchsorted(gi, node_gid[shared_local]).astype(np.int32)
    part["n_global_interface"] = len(gi)
    part["owned"] = (lowest[node_gid] == rank).astype(np.uint8)
    return part


def locality_order(coords, conn):
    """A renumbering for meshes whose node / element numbers carry no locality (mesh generators, merged parts): the kernels gather
    nodal data by element and element forces by node, so their memory traffic follows the numbering (bench.py `shuffled_numbering`:
    a randomly numbered cube runs at 0.43 of the structured one).  The reference accepts any numbering and offers bandwidth
    renumbering for its direct solvers; this is the analogue for the gathers: nodes sorted by (z cell, y cell, x) and elements by the
    same key of their centroid, cells of one mean element edge -- on a grid-like mesh that restores rows of consecutive nodes.

    Returns (new_of_old_node [nn], elem_order [ne]): new coordinates are coords[argsort(new_of_old_node)], new connectivity
    new_of_old_node[conn[elem_order]]; a nodal result r_new maps back as r_old = r_new[new_of_old_node]."""
    coords = np.asarray(coords, np.float64)
    conn = np.asarray(conn)
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    h = (np.prod(np.maximum(hi - lo, 1e-300)) / max(conn.shape[0], 1)) ** (1.0 / 3.0)

    def key(p, shift):
        cell = np.floor((p - lo) / h + shift).astype(np.int64)
        return np.lexsort((p[:, 0], cell[:, 1], cell[:, 2]))  # last key is the primary one

    node_sorted = key(coords, 0.5)  # old ids in their new order; nodes sit near cell corners, centroids near cell centres
    new_of_old = np.empty(coords.shape[0], np.int64)
    new_of_old[node_sorted] = np.arange(coords.shape[0])
    elem_order = key(coords[conn].mean(axis=1), 0.0)
    return new_of_old.astype(np.int32), elem_order.astype(np.int64)


def renumber(coords, conn, nodesets=None):
    """coords, conn (and node sets) in the numbering of locality_order, plus the map back: (X, conn, nodesets, new_of_old_node)"""
    new_of_old, elem_order = locality_order(coords, conn)
    X = np.empty_like(np.asarray(coords, np.float64))
    X[new_of_old] = coords
    c = np.ascontiguousarray(new_of_old[np.asarray(conn)[elem_order]].astype(np.int32))
    ns = {k: np.sort(new_of_old[v]).astype(np.int32) for k, v in (nodesets or {}).items()}
    return X, c, ns, new_of_old
