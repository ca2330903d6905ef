"""ctypes binding of libtahoe_b200.so (include/tahoe_b200.h).

This is the harness-side binding used by tests/, bench.py and __graft_entry__.py; the product's host side is the C++
plugin layer in tahoe_b200/host/ which calls the same C ABI.  There is no fallback of any kind: if the shared library
is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtahoe_b200.so")

SMALL_STRAIN, TOTAL_LAGRANGIAN, UPDATED_LAGRANGIAN, SMALL_STRAIN_BBAR = 0, 1, 2, 3
SSKSTV, FDKSTV, SIMO_ISO, J2_SIMO = 0, 1, 2, 3
FORM_OF = {"small_strain": 0, "total_lagrangian": 1, "updated_lagrangian": 2, "small_strain_B-bar": 3,
           "explicit_solid": 2}  # ExplicitElementT derives from UpdatedLagrangianT


def form_of(element):
    """formulation code of a parsed element description: <small_strain strain_displacement="B-bar"> is its own code"""
    if element["type"] == "small_strain" and element.get("strain_displacement", "standard") == "B-bar":
        return FORM_OF["small_strain_B-bar"]
    return FORM_OF[element["type"]]
KIND_OF = {"small_strain_StVenant": 0, "large_strain_StVenant": 1, "Simo_isotropic": 2, "Simo_J2": 3,
           "explicit_neo_hookean": 4, "explicit_J2": 5}  # 4, 5: <explicit_solid> materials (ExplNeoHookeanT, ExplJ2PlasticityT)
STATUS = {0: "ok", 1: "bad_jacobian", 2: "j2_local", 3: "cuda", 4: "argument", 5: "size", 6: "pcg_breakdown", 7: "comm"}
J2_BLOCK = 5 * 48 + 64  # doubles per element in the reference's ElementCardT layout

# the exported symbols of include/tahoe_b200.h (checked by tests/test_capi_symbols.py against the header text)
SYMBOLS = """tb2_version tb2_last_error tb2_device_count tb2_malloc tb2_free tb2_memcpy_h2d tb2_memcpy_d2h tb2_host_register
tb2_host_unregister tb2_profile_reserve tb2_profile_begin tb2_profile_end tb2_mesh_synchronize tb2_measure_fp64_peak tb2_mesh_create tb2_mesh_destroy tb2_mesh_sizes tb2_mesh_device tb2_mesh_stream tb2_mesh_colouring
tb2_group_create tb2_group_destroy tb2_form_internal_force tb2_form_internal_force_host tb2_group_status tb2_form_lumped_mass
tb2_form_lumped_mass_host tb2_group_set_element_status tb2_group_stable_time_step tb2_group_set_mass_scaling tb2_group_get_explicit_history tb2_group_nodal_stress tb2_group_nodal_stress_host tb2_group_nodal_stress_at tb2_group_nodal_stress_at_host tb2_group_close_step tb2_group_reset_step tb2_group_get_history tb2_group_set_history
tb2_form_inertial_force tb2_form_inertial_force_host tb2_form_mass tb2_matrix_scale tb2_newton_solve_dynamic tb2_newton_solve_dynamic_host
tb2_geom_open tb2_geom_close tb2_geom_sizes tb2_geom_coords tb2_geom_block tb2_geom_nodeset tb2_geom_sideset
tb2_traction_create tb2_traction_destroy tb2_traction_form tb2_traction_form_host
tb2_explicit_create tb2_explicit_destroy tb2_explicit_set_state tb2_explicit_get_state tb2_explicit_set_bc
tb2_explicit_initial_condition tb2_explicit_run tb2_explicit_step_host tb2_explicit_run_async tb2_explicit_wait tb2_explicit_update_bc_values tb2_matrix_pcg_converged tb2_matrix_bicgstab tb2_matrix_bicgstab_host tb2_explicit_device_array tb2_equations_create
tb2_equations_destroy tb2_equations_count tb2_equations_get tb2_equations_device tb2_matrix_create tb2_matrix_create_csr tb2_matrix_set_values tb2_matrix_destroy
tb2_matrix_nnz tb2_matrix_get_csr tb2_matrix_get_msr tb2_matrix_clear tb2_form_stiffness tb2_form_stiffness_host
tb2_form_stiffness_diagonal tb2_form_stiffness_diagonal_host
tb2_nlpcg_create tb2_nlpcg_destroy tb2_nlpcg_solve tb2_nlpcg_solve_host tb2_nlpcg_counters tb2_newton_solve tb2_newton_solve_host
tb2_matrix_multx tb2_matrix_multx_host tb2_matrix_copy_diagonal tb2_matrix_copy_diagonal_host tb2_matrix_pcg tb2_matrix_pcg_host tb2_equations_gather
tb2_equations_scatter_add tb2_comm_unique_id tb2_comm_init tb2_comm_destroy tb2_comm_sum_interface
tb2_comm_peer_export tb2_comm_peer_import tb2_comm_peer_enabled tb2_comm_peer_disable tb2_partition_rcb tb2_partition_part tb2_secant_search_host
tb2_explicit_attach_contact tb2_contact_create tb2_contact_destroy tb2_contact_set_pairs tb2_contact_form tb2_contact_form_host tb2_contact_tracking
tb2_contact_set_surfaces tb2_contact_search tb2_contact_get_pairs tb2_contact_has_surfaces""".split()


class Tb2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tahoe_b200: %s (%d): %s" % (STATUS.get(code, "?"), code, msg))
        self.code = code


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("hard_kind", C.c_int32), ("mu", C.c_double), ("lam", C.c_double), ("kappa", C.c_double),
                ("density", C.c_double), ("hard", C.c_double * 4), ("num_knots", C.c_int32), ("spline_fixity", C.c_int32),
                ("knot_x", C.c_double * 16), ("knot_y", C.c_double * 16)]


_lib = None


def lib():
    """load the C-ABI library; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libtahoe_b200.so is not built (%s); run __graft_entry__.build() -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.tb2_version.restype = C.c_char_p
        L.tb2_last_error.restype = C.c_char_p
        L.tb2_mesh_stream.restype = C.c_void_p
        L.tb2_explicit_device_array.restype = C.c_void_p
        L.tb2_equations_device.restype = C.c_void_p
        _lib = L
    return _lib


def last_error():
    return lib().tb2_last_error().decode()


def _chk(code):
    if code != 0:
        raise Tb2Error(code, last_error())


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _dp(x):
    """device pointer from an int, a torch tensor or None"""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, np.float64)


def material(desc_mat):
    """tb2_material from the XML material description (IsotropicT::TakeParameterList, IsotropicT.cpp:32-45)"""
    m = Material()
    m.kind = KIND_OF[desc_mat["type"]]
    m.density = desc_mat.get("density", 1.0)
    if "E" in desc_mat:
        E, nu = desc_mat["E"], desc_mat["nu"]
        m.mu = 0.5 * E / (1.0 + nu)
        m.lam = 2.0 * m.mu * nu / (1.0 - 2.0 * nu)
        m.kappa = m.lam + 2.0 / 3.0 * m.mu
    else:
        m.mu, m.kappa = desc_mat["mu"], desc_mat["kappa"]
        m.lam = m.kappa - 2.0 * m.mu / 3.0
    if desc_mat["type"] == "explicit_J2":  # ExplJ2PlasticityT: hard[0] = sigma_Y, hard[1] = H
        m.hard[0], m.hard[1] = desc_mat["sigma_Y"], desc_mat["hardening_modulus"]
    h = desc_mat.get("hardening")
    if h:
        if h["type"] == "linear_function":
            m.hard_kind = 0
            m.hard[0], m.hard[1] = h["a"], h["b"]
        elif h["type"] == "power_law":  # PowerLawT: a (b + c x)^n
            m.hard_kind = 2
            for i, k in enumerate("abcn"):
                m.hard[i] = h[k]
        elif h["type"] == "cubic_spline":  # CubicSplineT: knots + fixity, coefficients are formed by the library
            m.hard_kind = 3
            m.num_knots = len(h["points"])
            m.spline_fixity = {"parabolic": 0, "free_run": 1}[h["fixity"]]
            for i, (x, y) in enumerate(h["points"]):
                m.knot_x[i], m.knot_y[i] = x, y
        else:
            m.hard_kind = 1
            for i, k in enumerate("abcd"):
                m.hard[i] = h[k]
    return m


def device_count():
    n = C.c_int(0)
    _chk(lib().tb2_device_count(C.byref(n)))
    return n.value


def host_register(arr):
    """pin a host array (tb2_host_register): async copies, and kernel-written results in tb2_explicit_step_host"""
    _chk(lib().tb2_host_register(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes)))


def host_unregister(arr):
    _chk(lib().tb2_host_unregister(C.c_void_p(arr.ctypes.data)))


def memcpy_d2h(device, out, dptr):
    """device pointer -> host array (tb2_memcpy_d2h)"""
    _chk(lib().tb2_memcpy_d2h(int(device), _p(out), C.c_void_p(dptr), C.c_size_t(out.nbytes)))


def measure_fp64_peak(device=0):
    t = C.c_double(0.0)
    _chk(lib().tb2_measure_fp64_peak(int(device), C.byref(t)))
    return t.value


class _Handle:
    """Owner of one C handle.  Children (group of a mesh, matrix of an equation set ...) are closed before their parent no
    matter in which order Python finalises the objects (the cyclic GC gives no order): parent and child reference each other
    strongly, and whichever is finalised first closes the sub-tree bottom-up."""
    _destroy = None

    def _init_handle(self, parent=None):
        self.h = C.c_void_p()
        self._children = []
        self._parent = parent
        if parent is not None:
            parent._children.append(self)

    def close(self):
        if not getattr(self, "h", None):
            return
        for c in list(self._children):
            c.close()
        self._children = []
        getattr(lib(), self._destroy)(self.h)
        self.h = C.c_void_p()
        if self._parent is not None and self in self._parent._children:
            self._parent._children.remove(self)
        self._parent = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Mesh(_Handle):
    _destroy = "tb2_mesh_destroy"

    def __init__(self, coords, conn, device=0):
        coords = _f64(coords)
        conn = np.ascontiguousarray(conn, np.int32)
        assert coords.ndim == 2 and coords.shape[1] == 3 and conn.ndim == 2 and conn.shape[1] == 8
        self.nn, self.ne, self.device = coords.shape[0], conn.shape[0], device
        self._init_handle()
        _chk(lib().tb2_mesh_create(device, C.c_int64(self.nn), C.c_int64(self.ne), _p(conn), _p(coords), C.byref(self.h)))

    @property
    def stream(self):
        return lib().tb2_mesh_stream(self.h)

    def synchronize(self):
        _chk(lib().tb2_mesh_synchronize(self.h))

    def profile_reserve(self, records):
        _chk(lib().tb2_profile_reserve(self.h, C.c_int64(int(records))))

    def profile_begin(self):
        _chk(lib().tb2_profile_begin(self.h))

    def profile_end(self):
        """-> (ms[8], count[8], kernel_launches) per category of include/tahoe_b200.h"""
        ms, cnt, n = np.zeros(8), np.zeros(8, np.int64), C.c_int64(0)
        _chk(lib().tb2_profile_end(self.h, _p(ms), _p(cnt), C.byref(n)))
        return ms, cnt, n.value

    def colouring(self):
        col = np.zeros(self.ne, np.int32)
        n = C.c_int32(0)
        _chk(lib().tb2_mesh_colouring(self.h, _p(col), C.byref(n)))
        return n.value, col

    # ---- multi-GPU
    def comm_init(self, rank, nranks, uid, if_nodes, if_slots, n_global_interface, owned, all_gather=None):
        """all_gather: callable bytes -> list of every rank's bytes in rank order (the host program's own all-gather); when
        given, the exchange windows are mapped across the ranks and the hot loops exchange over NVLink peer memory
        (tb2_comm_peer_export / _import) instead of the packed ncclAllReduce"""
        if_nodes = np.ascontiguousarray(if_nodes, np.int32)
        if_slots = np.ascontiguousarray(if_slots, np.int32)
        owned = np.ascontiguousarray(owned, np.uint8)
        _chk(lib().tb2_comm_init(self.h, rank, nranks, uid, C.c_int64(len(if_nodes)), _p(if_nodes), _p(if_slots),
                                 C.c_int64(n_global_interface), _p(owned)))
        if all_gather is not None and nranks > 1:
            self.comm_enable_peer(all_gather)

    def comm_enable_peer(self, all_gather):
        """export / all-gather / import; every step is collective, so a failure on one rank (no IPC between the processes, GPUs
        outside one NVLink domain) is agreed on by all of them and the mesh stays on the packed ncclAllReduce.  Returns whether the
        peer exchange is on."""
        import sys
        buf = C.create_string_buffer(64)
        ok = lib().tb2_comm_peer_export(self.h, buf) == 0
        err = "" if ok else last_error()
        handles = all_gather(buf.raw if ok else b"\0" * 64)
        if ok and all(h != b"\0" * 64 for h in handles):
            ok = lib().tb2_comm_peer_import(self.h, b"".join(handles)) == 0
            err = "" if ok else last_error()
        else:
            ok = False
        if not all(f == b"1" for f in all_gather(b"1" if ok else b"0")):
            _chk(lib().tb2_comm_peer_disable(self.h))
            print("tahoe_b200: peer-memory exchange unavailable (%s); using the packed ncclAllReduce" % (err or "another rank failed"), file=sys.stderr)
            return False
        return True

    def comm_peer_enabled(self):
        return bool(lib().tb2_comm_peer_enabled(self.h))

    def sum_interface(self, d_nodal):
        _chk(lib().tb2_comm_sum_interface(self.h, _dp(d_nodal)))


def partition_rcb(coords, conn, nparts):
    """tb2_partition_rcb: owner rank of every element"""
    coords, conn = _f64(coords), np.ascontiguousarray(conn, np.int32)
    owner = np.zeros(conn.shape[0], np.int32)
    _chk(lib().tb2_partition_rcb(C.c_int64(coords.shape[0]), C.c_int64(conn.shape[0]), _p(conn), _p(coords), int(nparts), _p(owner)))
    return owner


def partition_part(nn, conn, owner, nparts, rank):
    """tb2_partition_part: (node_gid, elem_gid, local_conn, if_nodes, if_slots, n_global_interface, owned) of one rank"""
    conn, owner = np.ascontiguousarray(conn, np.int32), np.ascontiguousarray(owner, np.int32)
    n = [C.c_int64(0) for _ in range(4)]
    args = (C.c_int64(nn), C.c_int64(conn.shape[0]), _p(conn), _p(owner), int(nparts), int(rank))
    _chk(lib().tb2_partition_part(*args, C.byref(n[0]), C.byref(n[1]), C.byref(n[2]), C.byref(n[3]), None, None, None, None, None, None))
    nl, nel, nif, nglob = (v.value for v in n)
    node_gid, elem_gid = np.zeros(nl, np.int64), np.zeros(nel, np.int64)
    lconn, if_nodes, if_slots, owned = np.zeros((nel, 8), np.int32), np.zeros(nif, np.int32), np.zeros(nif, np.int32), np.zeros(nl, np.uint8)
    _chk(lib().tb2_partition_part(*args, C.byref(n[0]), C.byref(n[1]), C.byref(n[2]), C.byref(n[3]), _p(node_gid), _p(elem_gid), _p(lconn),
                                  _p(if_nodes), _p(if_slots), _p(owned)))
    return node_gid, elem_gid, lconn, if_nodes, if_slots, nglob, owned


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _chk(lib().tb2_comm_unique_id(buf))
    return buf.raw


class Group(_Handle):
    _destroy = "tb2_group_destroy"

    def __init__(self, mesh, form, mat):
        self.mesh, self.form, self.mat = mesh, form, mat
        self._init_handle(mesh)
        _chk(lib().tb2_group_create(mesh.h, int(form), C.byref(mat), C.byref(self.h)))

    def internal_force_host(self, u, u_last=None, iteration=0):
        u, u_last = _f64(u), _f64(u_last)
        f = np.zeros((self.mesh.nn, 3))
        _chk(lib().tb2_form_internal_force_host(self.h, _p(u), _p(u_last), int(iteration), _p(f)))
        return f

    def stiffness_diagonal_host(self, u, u_last=None, iteration=0):
        """diag K(u) per nodal dof: DiagonalMatrixT kDiagOnly assembly of the element loop"""
        u, u_last = _f64(u), _f64(u_last)
        d = np.zeros((self.mesh.nn, 3))
        _chk(lib().tb2_form_stiffness_diagonal_host(self.h, _p(u), _p(u_last), int(iteration), _p(d)))
        return d

    def internal_force(self, d_u, d_u_last, iteration, d_f):
        _chk(lib().tb2_form_internal_force(self.h, _dp(d_u), _dp(d_u_last), int(iteration), _dp(d_f)))

    def status(self):
        bad = C.c_int64(-1)
        return lib().tb2_group_status(self.h, C.byref(bad)), bad.value

    def lumped_mass_host(self):
        m = np.zeros((self.mesh.nn, 3))
        _chk(lib().tb2_form_lumped_mass_host(self.h, _p(m)))
        return m

    def set_element_status(self, off):
        """ElementCardT::kOFF flags: off[ne] != 0 switches an element off (None: all on)"""
        _chk(lib().tb2_group_set_element_status(self.h, None if off is None else _p(np.ascontiguousarray(np.asarray(off) != 0, np.uint8))))

    def stable_time_step(self):
        """ExplicitElementT::ComputeStableTimeStep"""
        dt = C.c_double(0.0)
        _chk(lib().tb2_group_stable_time_step(self.h, C.byref(dt)))
        return dt.value

    def set_mass_scaling(self, target_dt, scale_factor=0.9):
        """ExplicitElementT::ApplyMassScaling (fixed): returns (number of scaled elements, largest factor, factors[ne])"""
        n, mx, sc = C.c_int64(0), C.c_double(0.0), np.zeros(self.mesh.ne)
        _chk(lib().tb2_group_set_mass_scaling(self.h, C.c_double(target_dt), C.c_double(scale_factor), C.byref(n), C.byref(mx), _p(sc)))
        return n.value, mx.value, sc

    def explicit_history(self):
        """ExplJ2PlasticityT history [ip][16][element]"""
        h = np.zeros((8, 16, self.mesh.ne))
        _chk(lib().tb2_group_get_explicit_history(self.h, _p(h)))
        return h

    def inertial_force_host(self, mass_type, acc, scale=1.0):
        """scale * M a [nn][3] (ContinuumElementT::FormMa; mass_type 1 consistent, 2 lumped)"""
        out = np.zeros((self.mesh.nn, 3))
        _chk(lib().tb2_form_inertial_force_host(self.h, int(mass_type), C.c_double(scale), _p(_f64(acc)), _p(out)))
        return out

    def nodal_stress_host(self, u, u_last=None, iteration=0):
        """extrapolated + averaged nodal Cauchy stress [nn][6] (SolidElementT::ComputeOutput); J2: pass the last converged displacement"""
        out = np.zeros((self.mesh.nn, 6))
        _chk(lib().tb2_group_nodal_stress_at_host(self.h, _p(_f64(u)), _p(_f64(u_last)), int(iteration), _p(out)))
        return out

    def close_step(self):
        _chk(lib().tb2_group_close_step(self.h))

    def reset_step(self):
        _chk(lib().tb2_group_reset_step(self.h))

    def get_history(self):
        ne = self.mesh.ne
        data, flags, alloc = np.zeros((ne, J2_BLOCK)), np.zeros((ne, 8), np.int32), np.zeros(ne, np.int32)
        _chk(lib().tb2_group_get_history(self.h, _p(data), _p(flags), _p(alloc)))
        return data, flags, alloc

    def set_history(self, data, flags, alloc):
        data, flags, alloc = _f64(data), np.ascontiguousarray(flags, np.int32), np.ascontiguousarray(alloc, np.int32)
        _chk(lib().tb2_group_set_history(self.h, _p(data), _p(flags), _p(alloc)))


def read_geom(path):
    """TahoeII .geom -> (coords [nn,3], [conn per block], {node set id: nodes}, {side set id: [[element, facet]]}), all 0-based,
    through the library's multi-threaded reader (tb2_geom_*)"""
    h = C.c_void_p()
    _chk(lib().tb2_geom_open(path.encode(), C.byref(h)))
    try:
        nn, nsd, nb, nns, nss = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _chk(lib().tb2_geom_sizes(h, C.byref(nn), C.byref(nsd), C.byref(nb), C.byref(nns), C.byref(nss)))
        X = np.zeros((nn.value, 3))
        _chk(lib().tb2_geom_coords(h, _p(X)))
        blocks, nodesets, sidesets = [], {}, {}
        for k in range(nb.value):
            bid, nel, nen = C.c_int32(), C.c_int64(), C.c_int32()
            _chk(lib().tb2_geom_block(h, k, C.byref(bid), C.byref(nel), C.byref(nen), None))
            conn = np.zeros((nel.value, nen.value), np.int32)
            _chk(lib().tb2_geom_block(h, k, None, None, None, _p(conn)))
            blocks.append(conn)
        for k in range(nns.value):
            sid, n = C.c_int32(), C.c_int64()
            _chk(lib().tb2_geom_nodeset(h, k, C.byref(sid), C.byref(n), None))
            nodes = np.zeros(n.value, np.int32)
            _chk(lib().tb2_geom_nodeset(h, k, None, None, _p(nodes)))
            nodesets[sid.value] = nodes
        for k in range(nss.value):
            sid, blk, n = C.c_int32(), C.c_int32(), C.c_int64()
            _chk(lib().tb2_geom_sideset(h, k, C.byref(sid), C.byref(blk), C.byref(n), None))
            sides = np.zeros((n.value, 2), np.int32)
            _chk(lib().tb2_geom_sideset(h, k, None, None, None, _p(sides)))
            sidesets[sid.value] = sides
        return X, blocks, nodesets, sidesets
    finally:
        lib().tb2_geom_close(h)


class Traction(_Handle):
    """natural_bc traction cards of a group (ContinuumElementT::ApplyTractionBC): elem / facet 0-based [ncards], tract [ncards][4][3]
    (or one vector [3] for all cards and nodes), coord_system "global" | "local" """
    _destroy = "tb2_traction_destroy"

    def __init__(self, mesh, elem, facet, tract, coord_system="global"):
        elem = np.ascontiguousarray(elem, np.int32)
        facet = np.ascontiguousarray(facet, np.int32)
        tract = _f64(np.broadcast_to(np.asarray(tract, np.float64), (elem.shape[0], 4, 3)))
        self.mesh = mesh
        self._init_handle(mesh)
        _chk(lib().tb2_traction_create(mesh.h, C.c_int64(elem.shape[0]), _p(elem), _p(facet), _p(tract),
                                       {"global": 0, "local": 1}[coord_system], C.byref(self.h)))

    def form_host(self, scale=1.0, out=None):
        """nodal forces [nn][3]; added to `out` when given"""
        f = np.zeros((self.mesh.nn, 3)) if out is None else _f64(out)
        _chk(lib().tb2_traction_form_host(self.h, C.c_double(scale), 0 if out is None else 1, _p(f)))
        return f


class Contact(_Handle):
    """contact_3D_penalty force of one PenaltyContact3DT group: parameters at construction, the active striker-facet pairs
    (three facet nodes + striker, 0-based) and the strikers' areas from the host search with set_pairs"""
    _destroy = "tb2_contact_destroy"

    def __init__(self, mesh, penalty_stiffness, friction_coefficient=0.0, friction_epsilon=1.0e-6, viscous_damping=0.0):
        self.mesh = mesh
        self._init_handle(mesh)
        _chk(lib().tb2_contact_create(mesh.h, C.c_double(penalty_stiffness), C.c_double(friction_coefficient), C.c_double(friction_epsilon),
                                      C.c_double(viscous_damping), C.byref(self.h)))

    def set_pairs(self, pairs, area):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 4)
        area = _f64(area)
        _chk(lib().tb2_contact_set_pairs(self.h, C.c_int64(pairs.shape[0]), _p(pairs), _p(area)))

    def form_host(self, u, v=None, constKd=1.0, out=None):
        """nodal forces [nn][3] (residual sign); added to `out` when given"""
        f = np.zeros((self.mesh.nn, 3)) if out is None else _f64(out)
        u = _f64(u)
        vv = None if v is None else _f64(v)
        _chk(lib().tb2_contact_form_host(self.h, C.c_double(constKd), _p(u), _p(vv) if vv is not None else None, 0 if out is None else 1, _p(f)))
        return f

    def set_surfaces(self, facets, facet_surface, strikers, striker_area):
        """what the device search works on: triangulated surfaces (node triples + surface id), striker nodes and their areas"""
        facets = np.ascontiguousarray(facets, np.int32).reshape(-1, 3)
        facet_surface = np.ascontiguousarray(facet_surface, np.int32)
        strikers = np.ascontiguousarray(strikers, np.int32)
        striker_area = _f64(striker_area)
        _chk(lib().tb2_contact_set_surfaces(self.h, C.c_int64(facets.shape[0]), _p(facets), _p(facet_surface), C.c_int64(strikers.shape[0]),
                                            _p(strikers), _p(striker_area)))

    def search_host(self, u):
        """search on X + u (host array): the number of active pairs; they become the group's pair list"""
        import torch
        d = torch.from_numpy(_f64(u)).to(torch.device("cuda", self.mesh.device))
        n = C.c_int64(0)
        _chk(lib().tb2_contact_search(self.h, C.c_void_p(d.data_ptr()), C.byref(n)))
        return n.value

    def pairs(self):
        n = C.c_int64(0)
        _chk(lib().tb2_contact_get_pairs(self.h, C.byref(n), None, None))
        pairs, area = np.zeros((n.value, 4), np.int32), np.zeros(n.value)
        if n.value:
            _chk(lib().tb2_contact_get_pairs(self.h, C.byref(n), _p(pairs), _p(area)))
        return pairs, area

    def tracking(self):
        n, h = C.c_int(0), C.c_double(0.0)
        _chk(lib().tb2_contact_tracking(self.h, C.byref(n), C.byref(h)))
        return n.value, h.value


class Explicit(_Handle):
    _destroy = "tb2_explicit_destroy"

    def __init__(self, group):
        self.group = group
        self.nn = group.mesh.nn
        self._init_handle(group)
        _chk(lib().tb2_explicit_create(group.h, C.byref(self.h)))

    def attach_contact(self, contact):
        """a capi.Contact whose force joins the residual of every step (None detaches)"""
        self._contact = contact
        _chk(lib().tb2_explicit_attach_contact(self.h, contact.h if contact is not None else None))

    def set_state(self, d=None, v=None, a=None):
        _chk(lib().tb2_explicit_set_state(self.h, _p(_f64(d)), _p(_f64(v)), _p(_f64(a))))

    def get_state(self):
        d, v, a = (np.zeros((self.nn, 3)) for _ in range(3))
        _chk(lib().tb2_explicit_get_state(self.h, _p(d), _p(v), _p(a)))
        return d, v, a

    def set_bc(self, code=None, value=None, fext=None):
        code = None if code is None else np.ascontiguousarray(code, np.uint8)
        _chk(lib().tb2_explicit_set_bc(self.h, _p(code), _p(_f64(value)), _p(_f64(fext))))

    def update_bc_values(self, dofs, values):
        """sparse refresh of prescribed values: dofs = nodal dof indices 3 n + i"""
        dofs = np.ascontiguousarray(dofs, np.int64)
        _chk(lib().tb2_explicit_update_bc_values(self.h, C.c_int64(len(dofs)), _p(dofs), _p(_f64(values))))

    def initial_condition(self):
        _chk(lib().tb2_explicit_initial_condition(self.h))

    def run(self, dt, nsteps, fext_scale=None, value_scale=None):
        fs, vs = _f64(fext_scale), _f64(value_scale)
        assert fs is None or len(fs) >= nsteps
        assert vs is None or len(vs) >= nsteps
        _chk(lib().tb2_explicit_run(self.h, C.c_double(dt), int(nsteps), _p(fs), _p(vs)))

    def step_host(self, dt, d, v, a):
        """d, v, a: C-contiguous float64 host arrays (ideally pinned), updated in place"""
        _chk(lib().tb2_explicit_step_host(self.h, C.c_double(dt), C.c_void_p(d.ctypes.data), C.c_void_p(v.ctypes.data),
                                          C.c_void_p(a.ctypes.data)))

    def run_async(self, dt, nsteps, out_d, fext_scale=None, value_scale=None):
        """nsteps resident steps, then d -> out_d (host array or raw pointer, ideally pinned) without waiting; returns the ticket"""
        fs, vs = _f64(fext_scale), _f64(value_scale)
        ptr = out_d if isinstance(out_d, int) else out_d.ctypes.data
        t = C.c_int(-1)
        _chk(lib().tb2_explicit_run_async(self.h, C.c_double(dt), int(nsteps), _p(fs), _p(vs), C.c_void_p(ptr), C.byref(t)))
        return t.value

    def wait(self, ticket):
        _chk(lib().tb2_explicit_wait(self.h, int(ticket)))

    def step_host_ptr(self, dt, pd, pv, pa):
        _chk(lib().tb2_explicit_step_host(self.h, C.c_double(dt), C.c_void_p(pd), C.c_void_p(pv), C.c_void_p(pa)))

    def device_array(self, which):
        return lib().tb2_explicit_device_array(self.h, int(which))

    def mass_host(self):
        m = np.zeros((self.nn, 3))
        _chk(lib().tb2_memcpy_d2h(self.group.mesh.device, _p(m), C.c_void_p(self.device_array(3)), C.c_size_t(m.nbytes)))
        return m


class Equations(_Handle):
    _destroy = "tb2_equations_destroy"

    def __init__(self, mesh, bc_code):
        self.mesh = mesh
        bc = np.ascontiguousarray(np.asarray(bc_code) != 0, np.uint8)
        assert bc.shape == (mesh.nn, 3)
        self._init_handle(mesh)
        _chk(lib().tb2_equations_create(mesh.h, _p(bc), C.byref(self.h)))
        n = C.c_int64(0)
        _chk(lib().tb2_equations_count(self.h, C.byref(n)))
        self.neq = n.value

    def eqnos(self):
        eq = np.zeros((self.mesh.nn, 3), np.int32)
        _chk(lib().tb2_equations_get(self.h, _p(eq)))
        return eq

    def gather(self, d_nodal, d_eqvec):
        _chk(lib().tb2_equations_gather(self.h, _dp(d_nodal), _dp(d_eqvec)))

    def scatter_add(self, scale, d_eqvec, d_nodal):
        _chk(lib().tb2_equations_scatter_add(self.h, C.c_double(scale), _dp(d_eqvec), _dp(d_nodal)))


class Matrix(_Handle):
    _destroy = "tb2_matrix_destroy"

    def __init__(self, eqs):
        self.eqs = eqs
        self.neq = eqs.neq
        self._init_handle(eqs)
        _chk(lib().tb2_matrix_create(eqs.h, C.byref(self.h)))
        n = C.c_int64(0)
        _chk(lib().tb2_matrix_nnz(self.h, C.byref(n)))
        self.nnz = n.value

    def csr(self, values=True):
        rowptr, colind = np.zeros(self.neq + 1, np.int64), np.zeros(self.nnz, np.int32)
        val = np.zeros(self.nnz) if values else None
        _chk(lib().tb2_matrix_get_csr(self.h, _p(rowptr), _p(colind), _p(val)))
        return rowptr, colind, val

    def msr(self, upper_only):
        n = C.c_int64(0)
        _chk(lib().tb2_matrix_get_msr(self.h, int(upper_only), None, C.byref(n)))
        bindx = np.zeros(n.value, np.int32)
        _chk(lib().tb2_matrix_get_msr(self.h, int(upper_only), _p(bindx), C.byref(n)))
        return bindx

    def clear(self):
        _chk(lib().tb2_matrix_clear(self.h))

    def form_stiffness_host(self, group, u, u_last=None, iteration=0):
        _chk(lib().tb2_form_stiffness_host(group.h, self.h, _p(_f64(u)), _p(_f64(u_last)), int(iteration)))

    def form_mass(self, group, mass_type, constM=1.0):
        """A += constM * M (ContinuumElementT::FormMass through the element LHS loop)"""
        _chk(lib().tb2_form_mass(group.h, self.h, int(mass_type), C.c_double(constM)))
        _chk(lib().tb2_group_status(group.h, None))

    def scale(self, s):
        _chk(lib().tb2_matrix_scale(self.h, C.c_double(s)))

    def form_stiffness(self, group, d_u, d_u_last=None, iteration=0):
        _chk(lib().tb2_form_stiffness(group.h, self.h, _dp(d_u), _dp(d_u_last), int(iteration)))

    def multx_host(self, x):
        x = _f64(x)
        y = np.zeros_like(x)
        _chk(lib().tb2_matrix_multx_host(self.h, _p(x), _p(y)))
        return y

    def multx(self, d_x, d_y):
        _chk(lib().tb2_matrix_multx(self.h, _dp(d_x), _dp(d_y)))

    def pcg_host(self, b, x0=None, rtol=1e-12, atol=0.0, max_iter=10000):
        b = _f64(b)
        x = np.zeros_like(b) if x0 is None else _f64(x0).copy()
        it, rn = C.c_int(0), C.c_double(0.0)
        _chk(lib().tb2_matrix_pcg_host(self.h, _p(b), _p(x), C.c_double(rtol), C.c_double(atol), int(max_iter), C.byref(it), C.byref(rn)))
        return x, it.value, rn.value

    def bicgstab_host(self, b, x0=None, rtol=1e-12, atol=0.0, max_iter=10000):
        """Jacobi-preconditioned BiCGStab (non-symmetric tangents): solution, iterations, |r|"""
        x = np.zeros(self.neq) if x0 is None else np.ascontiguousarray(x0, np.float64).copy()
        it, rn = C.c_int(0), C.c_double(0.0)
        _chk(lib().tb2_matrix_bicgstab_host(self.h, _p(_f64(b)), _p(x), C.c_double(rtol), C.c_double(atol), int(max_iter), C.byref(it), C.byref(rn)))
        return x, it.value, rn.value

    def pcg_converged(self):
        """(converged, |r|/|r0|) of the last pcg / pcg_host call"""
        c, r = C.c_int(1), C.c_double(0.0)
        _chk(lib().tb2_matrix_pcg_converged(self.h, C.byref(c), C.byref(r)))
        return bool(c.value), r.value

    def pcg(self, d_b, d_x, rtol=1e-12, atol=0.0, max_iter=10000):
        it, rn = C.c_int(0), C.c_double(0.0)
        _chk(lib().tb2_matrix_pcg(self.h, _dp(d_b), _dp(d_x), C.c_double(rtol), C.c_double(atol), int(max_iter), C.byref(it), C.byref(rn)))
        return it.value, rn.value


class NlpcgParams(C.Structure):
    """tb2_nlpcg_params: the <PCG_solver> attributes"""
    _fields_ = [("restart", C.c_int32), ("line_search_iterations", C.c_int32), ("line_search_tolerance", C.c_double),
                ("max_step", C.c_double), ("abs_tolerance", C.c_double), ("rel_tolerance", C.c_double),
                ("divergence_tolerance", C.c_double), ("max_iterations", C.c_int32), ("min_iterations", C.c_int32)]


def nlpcg_params(solver=None, **kw):
    """from a parsed <PCG_solver> description (attribute strings) and/or keywords; defaults of PCGSolver_LS / NLSolver"""
    d = dict(restart=50, line_search_iterations=3, line_search_tolerance=0.25, max_step=2.5, abs_tolerance=1e-10,
             rel_tolerance=1e-12, divergence_tolerance=10.0, max_iterations=300, min_iterations=0)
    for src in (solver or {}, kw):
        for k, v in src.items():
            if k in d:
                d[k] = type(d[k])(float(v))
    return NlpcgParams(**d)


class NonlinearPCG(_Handle):
    """device twin of Tahoe's PCG_solver + diagonal_matrix (PCGSolver_LS)"""
    _destroy = "tb2_nlpcg_destroy"
    CONTINUE, CONVERGED, FAILED = 0, 1, 2

    def __init__(self, group, eqs, params):
        self.group, self.eqs, self.params = group, eqs, params
        self._init_handle(eqs)
        _chk(lib().tb2_nlpcg_create(group.h, eqs.h, C.byref(params), C.byref(self.h)))

    def solve_host(self, u, fext, u_last=None, solve_max_iterations=-1):
        """u [nn][3] is updated in place; returns (status, iterations, error, error0)"""
        assert u.dtype == np.float64 and u.flags.c_contiguous
        st, it, e, e0 = C.c_int(0), C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        _chk(lib().tb2_nlpcg_solve_host(self.h, _p(u), _p(_f64(u_last)), _p(_f64(fext)), int(solve_max_iterations), C.byref(st),
                                        C.byref(it), C.byref(e), C.byref(e0)))
        return st.value, it.value, e.value, e0.value

    def solve(self, d_u, d_fext, d_u_last=None, solve_max_iterations=-1):
        st, it, e, e0 = C.c_int(0), C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        _chk(lib().tb2_nlpcg_solve(self.h, _dp(d_u), _dp(d_u_last), _dp(d_fext), int(solve_max_iterations), C.byref(st), C.byref(it),
                                   C.byref(e), C.byref(e0)))
        return st.value, it.value, e.value, e0.value

    def counters(self):
        a, b = C.c_int64(0), C.c_int64(0)
        _chk(lib().tb2_nlpcg_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


class NewtonParams(C.Structure):
    """tb2_newton_params: <nonlinear_solver> attributes + those of <CUDA_PCG_matrix/>"""
    _fields_ = [("abs_tolerance", C.c_double), ("rel_tolerance", C.c_double), ("divergence_tolerance", C.c_double),
                ("max_iterations", C.c_int32), ("min_iterations", C.c_int32), ("reform_tangent_iterations", C.c_int32),
                ("pcg_rel_tolerance", C.c_double), ("pcg_abs_tolerance", C.c_double), ("pcg_max_iterations", C.c_int32)]


def newton_params(solver=None, **kw):
    d = dict(abs_tolerance=1e-10, rel_tolerance=1e-12, divergence_tolerance=1e3, max_iterations=25, min_iterations=0,
             reform_tangent_iterations=1, pcg_rel_tolerance=1e-13, pcg_abs_tolerance=0.0, pcg_max_iterations=20000)
    for src in (solver or {}, kw):
        for k, v in src.items():
            if k in d:
                d[k] = type(d[k])(float(v))
    return NewtonParams(**d)


def newton_solve_host(work, A, params, u, fext, u_last=None, solve_max_iterations=-1):
    """NLSolver::Solve on the device (tb2_newton_solve_host): u [nn][3] updated in place; returns (status, iterations, error,
    error0, linear iterations)"""
    st, it, e, e0, lin = C.c_int(0), C.c_int(0), C.c_double(0.0), C.c_double(0.0), C.c_int64(0)
    _chk(lib().tb2_newton_solve_host(work.h, A.h, C.byref(params), _p(u), _p(_f64(u_last)), _p(_f64(fext)), int(solve_max_iterations),
                                     C.byref(st), C.byref(it), C.byref(e), C.byref(e0), C.byref(lin)))
    return st.value, it.value, e.value, e0.value, lin.value


class Dynamics(C.Structure):
    """tb2_dynamics: element / nodal constants of an implicit integrator step"""
    _fields_ = [("mass_type", C.c_int32), ("constM", C.c_double), ("constK", C.c_double), ("constMa", C.c_double), ("constKd", C.c_double),
                ("dcorr_a", C.c_double), ("vcorr_a", C.c_double)]


def hht_dynamics(mass_type, dt, alpha=0.0):
    """constants of NLHHTalpha(alpha) (HHTalpha::Set2ndOrder, eLinearHHTalpha / eNLHHTalpha::eComputeParameters,
    nNLHHTalpha::nComputeParameters); `nonlinear_HHT` is alpha = 0: beta = 1/4, gamma = 1/2"""
    gamma, beta = 0.5 * (1.0 - 2.0 * alpha), 0.25 * (1.0 - alpha) ** 2
    return Dynamics(int(mass_type), 1.0, (1.0 + alpha) * beta * dt * dt, 1.0, 1.0 + alpha, beta * dt * dt, gamma * dt)


def newton_solve_dynamic_host(work, A, params, dyn, u, v, a, fext, u_last=None, solve_max_iterations=-1):
    """one implicit step's Newton solve on the device (tb2_newton_solve_dynamic_host): u, v, a [nn][3] updated in place; returns
    (status, iterations, error, error0, linear iterations)"""
    st, it, e, e0, lin = C.c_int(0), C.c_int(0), C.c_double(0.0), C.c_double(0.0), C.c_int64(0)
    _chk(lib().tb2_newton_solve_dynamic_host(work.h, A.h, C.byref(params), C.byref(dyn), _p(u), _p(v), _p(a), _p(_f64(u_last)),
                                             _p(_f64(fext)), int(solve_max_iterations), C.byref(st), C.byref(it), C.byref(e), C.byref(e0),
                                             C.byref(lin)))
    return st.value, it.value, e.value, e0.value, lin.value
