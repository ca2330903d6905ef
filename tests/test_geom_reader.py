"""the library's multi-threaded TahoeII .geom reader (tb2_geom_*; host code, runs without a GPU) against the
reference's own geometry files as the reference reads them (coordinates / connectivity of the golden fixtures come from
ModelManagerT through oracle/ref_dump.cpp) and against the test-side reader on generated meshes."""
import os
import shutil
import tempfile

import numpy as np
import pytest

import tahoe_input as ti
from cases import Case

REF_GEOM = "/root/reference/benchmark_XML/level.0/geometry"


@pytest.fixture(scope="module")
def capi():
    from tahoe_b200 import capi
    capi.lib()
    return capi


def _same(capi, path):
    X, blocks, ns, ss = capi.read_geom(path)
    X0, conn0, ns0 = ti.read_geom(path)
    assert np.array_equal(X, X0) and np.array_equal(np.concatenate(blocks), conn0)
    # a negative entry ("all model nodes", beam.1.geom) stays -1 in the library; the test-side reader shifts it like an id
    assert sorted(ns) == sorted(ns0) and all(np.array_equal(ns[k], np.where(ns0[k] < 0, -1, ns0[k])) for k in ns0)
    ss0 = ti.read_sidesets(path)
    assert sorted(ss) == sorted(ss0) and all(np.array_equal(ss[k], ss0[k]) for k in ss0)
    return X, blocks, ns, ss


@pytest.mark.parametrize("threads_worth", [3, 40])
def test_generated_mesh_round_trip(capi, threads_worth):
    """inline sections, jittered coordinates at full precision, side sets; the 40^3 file (6 MB) goes through the threaded path"""
    n = threads_worth
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        X, conn, ns = ti.structured_cube(n, jitter=0.1)
        X = ti.warp(X)
        path = os.path.join(work, "mesh.geom")
        ti.write_geom(path, X, conn, ns, sidesets=ti.cube_side_sets(n))
        Xr, blocks, nsr, ssr = _same(capi, path)
        assert np.array_equal(Xr, X) and np.array_equal(blocks[0], conn)  # %.17e round-trips every double
        assert all(np.array_equal(nsr[k], ns[k]) for k in ns)
        assert all(np.array_equal(ssr[k], ti.cube_side_sets(n)[k]) for k in ssr)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_several_element_blocks(capi):
    """a .geom with two element sets (one per material of a multi-material group): block ids, sizes and rows in file order"""
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        n = 5
        X, conn, ns = ti.structured_cube(n, jitter=0.1)
        path = os.path.join(work, "mesh.geom")
        ti.write_geom(path, X, conn, ns, block_sizes=[2 * n * n, 3 * n * n])
        Xr, blocks, nsr, _ = capi.read_geom(path)
        assert [b.shape for b in blocks] == [(2 * n * n, 8), (3 * n * n, 8)]
        assert np.array_equal(np.concatenate(blocks), conn) and np.array_equal(Xr, X)
        X0, conn0, _ = ti.read_geom(path)  # the test-side reader concatenates the blocks in file order too
        assert np.array_equal(conn0, conn)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_node_records_in_any_order_and_comments(capi):
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        X, conn, ns = ti.structured_cube(2, jitter=0.0)
        path = os.path.join(work, "mesh.geom")
        ti.write_geom(path, X, conn, ns)
        text = open(path).read()
        head, nodes = text.split("*nodes\n")
        lines = nodes.strip().splitlines()
        body = lines[2:]
        body = body[::-1]  # node records reversed: ids decide where a row goes
        body.insert(3, "# a comment line in the middle of the records")
        body[5] += "   # trailing comment"
        open(path, "w").write(head + "*nodes\n" + "\n".join(lines[:2] + body) + "\n")
        Xr, blocks, _, _ = capi.read_geom(path)
        assert np.array_equal(Xr, X) and np.array_equal(blocks[0], conn)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_element_records_in_any_order_go_to_their_numbered_row(capi):
    """ModelFileT::GetElementSet -> nArray2DT::ReadNumbered (nArray2DT.h:1071) places a record "<number> n1..n8" at row number - 1;
    side sets address elements by that number.  Shuffled element records (through the threaded path too: 30^3) must come back in
    numbered order, so the side-set entries still point at the elements they were written for"""
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        for n in (3, 30):
            X, conn, ns = ti.structured_cube(n, jitter=0.1)
            path = os.path.join(work, "mesh%d.geom" % n)
            ti.write_geom(path, X, conn, ns, sidesets=ti.cube_side_sets(n))
            text = open(path).read()
            head, rest = text.split("8  # number of element nodes\n")
            body, tail = rest.split("# end elements")
            lines = body.strip().splitlines()
            rng = np.random.default_rng(n)
            lines = [lines[k] for k in rng.permutation(len(lines))]
            open(path, "w").write(head + "8  # number of element nodes\n" + "\n".join(lines) + "\n# end elements" + tail)
            Xr, blocks, nsr, ssr = capi.read_geom(path)
            assert np.array_equal(blocks[0], conn) and np.array_equal(Xr, X)
            assert all(np.array_equal(ssr[k], ti.cube_side_sets(n)[k]) for k in ssr)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_repeated_and_out_of_range_record_numbers_are_rejected(capi):
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        X, conn, ns = ti.structured_cube(2, jitter=0.0)
        path = os.path.join(work, "mesh.geom")
        ti.write_geom(path, X, conn, ns)
        text = open(path).read()
        for bad in (text.replace("\n2 2 3 6 5 11 12 15 14\n", "\n1 2 3 6 5 11 12 15 14\n"),      # element number 1 twice
                    text.replace("\n2 2 3 6 5 11 12 15 14\n", "\n9 2 3 6 5 11 12 15 14\n"),      # element number past nel
                    text.replace("\n2 5.0", "\n1 5.0", 1)):                                       # node number 1 twice
            assert bad != text
            open(path, "w").write(bad)
            with pytest.raises(capi.Tb2Error):
                capi.read_geom(path)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_malformed_files_are_rejected(capi):
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        X, conn, ns = ti.structured_cube(2, jitter=0.0)
        path = os.path.join(work, "mesh.geom")
        ti.write_geom(path, X, conn, ns)
        text = open(path).read()
        for bad in (text.replace("*nodes", "*knots"), text[:len(text) // 2], text.replace("27  # number of nodes", "26  # number of nodes", 1),
                    text.replace("\n1 1 2 5 4 10 11 14 13\n", "\n1 1 2 5 4 10 11 14 99\n")):
            assert bad != text
            open(path, "w").write(bad)
            with pytest.raises(capi.Tb2Error):
                capi.read_geom(path)
        with pytest.raises(capi.Tb2Error):
            capi.read_geom(os.path.join(work, "missing.geom"))
    finally:
        shutil.rmtree(work, ignore_errors=True)


@pytest.mark.skipif(not os.path.isdir(REF_GEOM), reason="the reference tree exists in the authoring container only")
@pytest.mark.parametrize("geom,fixture", [("cube.1.geom", "ref_traction_a"), ("beam.1.geom", "ref_beam_newton")])
def test_reference_geometry_files_as_the_reference_reads_them(capi, geom, fixture):
    """external .elem / .node files (cube.1.geom style); coordinates and connectivity equal what ModelManagerT handed the
    reference run that wrote the golden fixture"""
    X, blocks, ns, ss = _same(capi, os.path.join(REF_GEOM, geom))
    c = Case(fixture)
    assert np.array_equal(X, c.X) and np.array_equal(np.concatenate(blocks), c.conn)
    assert all(np.array_equal(ns[k], np.where(c.nodesets[k] < 0, -1, c.nodesets[k])) for k in c.nodesets)
    assert all(np.array_equal(ss[k], c.sidesets[k]) for k in c.sidesets)


def test_node_sets_and_side_sets_in_external_files(capi):
    """the bodies of node sets and side sets may live in files the main file names (tri3.geom, quad4.*.geom of the reference): the
    same arrays as with the bodies inline, and side sets keep their element block"""
    import re
    work = tempfile.mkdtemp(prefix="tb2_geom_")
    try:
        n = 4
        X, conn, ns = ti.structured_cube(n, jitter=0.1)
        ss = ti.cube_side_sets(n)
        inline = os.path.join(work, "inline.geom")
        ti.write_geom(inline, X, conn, ns, sidesets={1: ss[6], 2: ss[5]}, block_sizes=[2 * n * n, 2 * n * n], sideset_blocks={1: 1, 2: 2})
        text = open(inline).read()
        head, rest = text.split("*nodesets\n", 1)
        nodesets, rest = rest.split("# end node sets\n*sidesets\n", 1)
        sidesets, rest = rest.split("*elements\n", 1)
        out = head + "*nodesets\n"
        for k, body in enumerate([b for b in nodesets.split("*set\n") if b.strip()]):
            name = "split.geom.ns%d" % k
            if k % 2 == 0:  # every other set goes to its own file
                open(os.path.join(work, name), "w").write(body)
                out += "*set\n%s\n" % name
            else:
                out += "*set\n" + body
        out += "# end node sets\n*sidesets\n"
        for k, body in enumerate([b for b in sidesets.split("*set\n") if b.strip()]):
            name = "split.geom.ss%d" % k
            open(os.path.join(work, name), "w").write(body)
            out += "*set\n%s\n" % name
        out += "*elements\n" + rest
        split = os.path.join(work, "split.geom")
        open(split, "w").write(out)
        a, b = capi.read_geom(inline), capi.read_geom(split)
        assert np.array_equal(a[0], b[0]) and all(np.array_equal(p, q) for p, q in zip(a[1], b[1]))
        assert sorted(a[2]) == sorted(b[2]) == sorted(ns) and all(np.array_equal(a[2][k], b[2][k]) for k in a[2])
        assert sorted(a[3]) == sorted(b[3]) == [1, 2] and all(np.array_equal(a[3][k], b[3][k]) for k in a[3])
        assert all(np.array_equal(b[2][k], np.asarray(ns[k])) for k in ns)
        assert np.array_equal(b[3][1], np.asarray(ss[6])) and np.array_equal(b[3][2], np.asarray(ss[5]))
        os.remove(os.path.join(work, "split.geom.ss1"))
        with pytest.raises(capi.Tb2Error):
            capi.read_geom(split)  # a named file that does not exist
    finally:
        shutil.rmtree(work, ignore_errors=True)
