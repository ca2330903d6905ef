"""ctypes binding of oracle/tahoe_oracle.c (test infrastructure -- the checker, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

SMALL_STRAIN, TOTAL_LAGRANGIAN, UPDATED_LAGRANGIAN = 0, 1, 2
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


class Material(C.Structure):
    _fields_ = [("kind", C.c_int), ("mu", C.c_double), ("lam", C.c_double), ("kappa", C.c_double),
                ("density", C.c_double), ("hard_kind", C.c_int), ("hard", C.c_double * 4),
                ("nknots", C.c_int), ("knot_x", C.c_double * 16), ("spline", C.c_double * (17 * 4))]


J2_DTYPE = np.dtype([("b_bar", "f8", 6), ("unit_norm", "f8", 6), ("beta_bar", "f8", 6), ("b_bar_trial", "f8", 6),
                     ("beta_bar_trial", "f8", 6), ("internal", "f8", 8), ("flag", "i4"), ("_pad", "i4")])


def build(force=False):
    src = os.path.join(ORACLE_DIR, "tahoe_oracle.c")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(ORACLE_DIR, "tahoe_oracle.h")))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_csr_structure.restype = C.c_int64
        _lib.orc_msr_structure.restype = C.c_int64
        _lib.orc_set_equation_numbers.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def material(desc_mat):
    """orc_material_t from the XML material description (IsotropicT::TakeParameterList)"""
    m = Material()
    kind = KIND_OF[desc_mat["type"]]
    if "E" in desc_mat:
        lib().orc_material_from_E_nu(C.byref(m), kind, C.c_double(desc_mat["E"]), C.c_double(desc_mat["nu"]),
                                     C.c_double(desc_mat["density"]))
    else:  # IsotropicT::Set_mu_kappa
        m.kind, m.mu, m.kappa, m.density = kind, desc_mat["mu"], desc_mat["kappa"], desc_mat["density"]
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
        elif h["type"] == "cubic_spline":
            m.hard_kind = 3
            pts = np.asarray(h["points"], np.float64)
            x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
            err = lib().orc_material_set_spline(C.byref(m), len(x), _p(x), _p(y), {"parabolic": 0, "free_run": 1}[h["fixity"]])
            assert err == 0
        else:
            m.hard_kind = 1
            for i, k in enumerate("abcd"):
                m.hard[i] = h[k]
    return m


def hardening(mat, alpha):
    """K(alpha), K'(alpha) of a J2 material"""
    K, dK = C.c_double(0), C.c_double(0)
    lib().orc_hardening(C.byref(mat), C.c_double(alpha), C.byref(K), C.byref(dK))
    return K.value, dK.value


def internal_force(form, mat, conn, X, u, u_last=None, j2=None, alloc=None, iteration=0):
    conn = np.ascontiguousarray(conn, np.int32)
    f = np.zeros_like(X)
    bad = C.c_int64(-1)
    err = lib().orc_internal_force(form, C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), _p(X), _p(u), _p(u_last),
                                   _p(j2), _p(alloc), iteration, _p(f), C.byref(bad))
    return err, f


def lumped_mass(density, conn, X):
    conn = np.ascontiguousarray(conn, np.int32)
    m = np.zeros_like(X)
    err = lib().orc_lumped_mass(C.c_double(density), C.c_int64(conn.shape[0]), _p(conn), _p(X), _p(m))
    assert err == 0
    return m


def equation_numbers(bc_code):
    bc = np.ascontiguousarray(bc_code != 0, np.uint8)
    eq = np.zeros(bc.shape, np.int32)
    neq = lib().orc_set_equation_numbers(C.c_int64(bc.shape[0]), _p(bc), _p(eq))
    return eq, int(neq)


def csr_structure(conn, eqnos, neq, upper_only=False):
    conn = np.ascontiguousarray(conn, np.int32)
    rowptr = np.zeros(neq + 1, np.int64)
    args = (C.c_int64(conn.shape[0]), _p(conn), C.c_int64(eqnos.shape[0]), _p(eqnos), C.c_int64(neq), int(upper_only))
    nnz = lib().orc_csr_structure(*args, _p(rowptr), None)
    colind = np.zeros(nnz, np.int32)
    lib().orc_csr_structure(*args, _p(rowptr), _p(colind))
    return rowptr, colind


def msr_structure(conn, eqnos, neq, upper_only):
    conn = np.ascontiguousarray(conn, np.int32)
    args = (C.c_int64(conn.shape[0]), _p(conn), C.c_int64(eqnos.shape[0]), _p(eqnos), C.c_int64(neq), int(upper_only))
    n = lib().orc_msr_structure(*args, None)
    bindx = np.zeros(n, np.int32)
    lib().orc_msr_structure(*args, _p(bindx))
    return bindx


def colouring(conn, nn):
    conn = np.ascontiguousarray(conn, np.int32)
    col = np.zeros(conn.shape[0], np.int32)
    ncol = lib().orc_greedy_colouring(C.c_int64(conn.shape[0]), _p(conn), C.c_int64(nn), _p(col))
    return ncol, col


def assemble_stiffness(form, mat, conn, X, u, eqnos, neq, rowptr, colind, u_last=None, j2=None, alloc=None, iteration=0):
    conn = np.ascontiguousarray(conn, np.int32)
    val = np.zeros(colind.shape[0])
    err = lib().orc_assemble_stiffness(form, C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), _p(X), _p(u), _p(u_last),
                                       _p(j2), _p(alloc), iteration, _p(eqnos), C.c_int64(neq), _p(rowptr), _p(colind),
                                       _p(val))
    return err, val


def element_stiffness(form, mat, Xe, ue, ul=None, j2=None, alloc=None, iteration=0):
    Ke = np.zeros(576)
    err = lib().orc_element_stiffness(form, C.byref(mat), _p(np.ascontiguousarray(Xe)), _p(np.ascontiguousarray(ue)),
                                      _p(ul), _p(j2), _p(alloc), iteration, _p(Ke))
    return err, Ke.reshape(24, 24).T  # column-major -> [r, c]


def element_force(form, mat, Xe, ue, ul=None, j2=None, alloc=None, iteration=0):
    fe = np.zeros(24)
    err = lib().orc_element_force(form, C.byref(mat), _p(np.ascontiguousarray(Xe)), _p(np.ascontiguousarray(ue)),
                                  _p(ul), _p(j2), _p(alloc), iteration, _p(fe))
    return err, fe


def spmv(rowptr, colind, val, x):
    y = np.zeros_like(x)
    lib().orc_csr_spmv(C.c_int64(x.shape[0]), _p(rowptr), _p(colind), _p(val), _p(x), _p(y))
    return y


def pcg_jacobi(rowptr, colind, val, b, x0=None, rtol=1e-12, atol=0.0, max_iter=10000):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    rn = C.c_double(0.0)
    it = lib().orc_pcg_jacobi(C.c_int64(b.shape[0]), _p(rowptr), _p(colind), _p(val), _p(b), _p(x), C.c_double(rtol),
                              C.c_double(atol), int(max_iter), C.byref(rn))
    return x, it, rn.value


def cd_predictor(dt, d, v, a, bc, bcval):
    lib().orc_cd_predictor(C.c_int64(d.size), C.c_double(dt), _p(d), _p(v), _p(a), _p(bc), _p(bcval))


def cd_corrector(dt, v, a, R, mass, bc):
    lib().orc_cd_corrector(C.c_int64(v.size), C.c_double(dt), _p(v), _p(a), _p(R), _p(mass), _p(bc))


def j2_update(mat, j2, alloc):
    for e in np.nonzero(alloc)[0]:
        lib().orc_j2_update(C.byref(mat), C.c_void_p(j2.ctypes.data + int(e) * 8 * J2_DTYPE.itemsize))


def j2_reset(j2, alloc):
    for e in np.nonzero(alloc)[0]:
        lib().orc_j2_reset(C.c_void_p(j2.ctypes.data + int(e) * 8 * J2_DTYPE.itemsize))


class NlpcgParams(C.Structure):
    """orc_nlpcg_params_t: the <PCG_solver> attributes (PCGSolver_LS.cpp:48-100, NLSolver::DefineParameters)"""
    _fields_ = [("restart", C.c_int), ("ls_iterations", C.c_int), ("ls_tolerance", C.c_double), ("max_step", C.c_double),
                ("abs_tol", C.c_double), ("rel_tol", C.c_double), ("div_tol", C.c_double), ("max_iterations", C.c_int),
                ("min_iterations", C.c_int), ("solve_max_iterations", C.c_int)]


def nlpcg_params(solver):
    """from the parsed <PCG_solver> description (strings as in the XML); defaults of PCGSolver_LS / NLSolver"""
    g = lambda k, d: type(d)(float(solver.get(k, d))) if not isinstance(d, int) else int(solver.get(k, d))
    return NlpcgParams(g("restart", 50), g("line_search_iterations", 3), g("line_search_tolerance", 0.25), g("max_step", 2.5),
                       g("abs_tolerance", 1e-10), g("rel_tolerance", 1e-12), g("divergence_tolerance", 10.0),
                       g("max_iterations", 300), g("min_iterations", 0), -1)


def stiffness_diagonal(form, mat, conn, X, u, u_last=None, j2=None, alloc=None, iteration=0):
    conn = np.ascontiguousarray(conn, np.int32)
    d = np.zeros_like(X)
    err = lib().orc_stiffness_diagonal(form, C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), _p(X), _p(u), _p(u_last), _p(j2),
                                       _p(alloc), iteration, _p(d))
    return err, d


def nlpcg_solve(form, mat, conn, X, u, eqnos, neq, fext, prm, u_last=None, j2=None, alloc=None):
    """PCGSolver_LS::Solve restated; u is updated in place.  returns (status, iterations, error, error0)"""
    conn = np.ascontiguousarray(conn, np.int32)
    it, err, err0 = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
    st = lib().orc_nlpcg_solve(form, C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), C.c_int64(X.shape[0]), _p(X), _p(u),
                               _p(u_last), _p(j2), _p(alloc), _p(np.ascontiguousarray(eqnos, np.int32)), C.c_int64(neq),
                               _p(np.ascontiguousarray(fext, np.float64)), C.byref(prm), C.byref(it), C.byref(err), C.byref(err0))
    return st, it.value, err.value, err0.value


def explicit_solid_force(mat, conn, X, u, hist=None):
    """+B^T sigma of <explicit_solid>; hist [ne][8][16] is updated in place (every evaluation, as the reference does)"""
    conn = np.ascontiguousarray(conn, np.int32)
    f = np.zeros_like(X)
    err = lib().orc_explicit_solid_force(C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), _p(X), _p(np.ascontiguousarray(u)), _p(hist), _p(f))
    return err, f


def explicit_solid_history(ne):
    h = np.zeros((ne, 8, 16))
    lib().orc_explicit_solid_init_history(C.c_int64(ne), _p(h))
    return h


def explicit_solid_stable_dt(mat, conn, X):
    conn = np.ascontiguousarray(conn, np.int32)
    lib().orc_explicit_solid_stable_dt.restype = C.c_double
    return lib().orc_explicit_solid_stable_dt(C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), _p(X))


def explicit_solid_mass_scale(mat, conn, X, target_dt, scale_factor):
    conn = np.ascontiguousarray(conn, np.int32)
    s = np.zeros(conn.shape[0])
    lib().orc_explicit_solid_mass_scale(C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), _p(X), C.c_double(target_dt), C.c_double(scale_factor), _p(s))
    return s


def lumped_mass_scaled(density, conn, X, scale):
    conn = np.ascontiguousarray(conn, np.int32)
    m = np.zeros_like(X)
    err = lib().orc_lumped_mass_scaled(C.c_double(density), C.c_int64(conn.shape[0]), _p(conn), _p(X), _p(scale), _p(m))
    assert err == 0
    return m


def nodal_stress(form, mat, conn, X, u, u_last=None, j2=None, alloc=None, iteration=0):
    conn = np.ascontiguousarray(conn, np.int32)
    out = np.zeros((X.shape[0], 6))
    if u_last is None:
        err = lib().orc_nodal_stress(form, C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), C.c_int64(X.shape[0]), _p(X),
                                     _p(np.ascontiguousarray(u)), _p(out))
    else:
        err = lib().orc_nodal_stress_history(form, C.byref(mat), C.c_int64(conn.shape[0]), _p(conn), C.c_int64(X.shape[0]), _p(X),
                                             _p(np.ascontiguousarray(u)), _p(np.ascontiguousarray(u_last)), _p(j2), _p(alloc),
                                             int(iteration), _p(out))
    return err, out


def explicit_material_stress(mat, F, h=None):
    """one evaluation of ExplNeoHookeanT / ExplJ2PlasticityT (F 3x3 row-major; h[16] updated in place for the J2 law)"""
    sig = np.zeros(6)
    lib().orc_explicit_material_stress(C.byref(mat), _p(np.ascontiguousarray(F, np.float64).ravel()), _p(h), _p(sig))
    return sig


def traction_force(conn, X, elem, facet, tract, coord_system="global", scale=1.0, out=None):
    """ContinuumElementT::ApplyTractionBC: nodal forces [nn][3] of the facet cards (tract [ncards][4][3] or one vector)"""
    conn = np.ascontiguousarray(conn, np.int32)
    elem = np.ascontiguousarray(elem, np.int32)
    facet = np.ascontiguousarray(facet, np.int32)
    tract = np.ascontiguousarray(np.broadcast_to(np.asarray(tract, np.float64), (elem.shape[0], 4, 3)))
    f = np.zeros_like(X) if out is None else out
    err = lib().orc_traction_force(C.c_int64(elem.shape[0]), _p(elem), _p(facet), _p(conn), _p(np.ascontiguousarray(X)), _p(tract),
                                   {"global": 0, "local": 1}[coord_system], C.c_double(scale), _p(f))
    assert err == 0, err
    return f


def contact_search(facets, facet_surface, strikers, x):
    """Contact3DT::SetActiveStrikers on the configuration x: (hit facet per striker or -1, gap of the hit)"""
    facets = np.ascontiguousarray(facets, np.int32).reshape(-1, 3)
    facet_surface = np.ascontiguousarray(facet_surface, np.int32)
    strikers = np.ascontiguousarray(strikers, np.int32)
    x = np.ascontiguousarray(x, np.float64)
    hit = np.zeros(strikers.shape[0], np.int32)
    gap = np.zeros(strikers.shape[0])
    L = lib()
    L.orc_contact_search.restype = C.c_int
    n = L.orc_contact_search(C.c_int64(facets.shape[0]), _p(facets), _p(facet_surface), C.c_int64(strikers.shape[0]), _p(strikers),
                             C.c_int64(x.shape[0]), _p(x), _p(hit), _p(gap))
    assert n == int((hit >= 0).sum())
    return hit, gap


def pairs_from_hits(facets, strikers, striker_area, hit):
    """the pair rows (three facet nodes, striker) and areas of the active strikers, in striker order"""
    act = np.nonzero(hit >= 0)[0]
    facets = np.asarray(facets).reshape(-1, 3)
    pairs = np.concatenate([facets[hit[act]], np.asarray(strikers)[act, None]], axis=1).astype(np.int32)
    return pairs, np.asarray(striker_area)[act]


def contact_force(pairs, area, X, u, v=None, K=0.0, mu=0.0, eps=1e-6, visc=0.0, constKd=1.0):
    """PenaltyContact3DT::RHSDriver over a list of active striker-facet pairs: (nodal forces [nn][3], pairs in contact, deepest penetration)"""
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 4)
    area = np.ascontiguousarray(area, np.float64)
    X = np.ascontiguousarray(X)
    u = np.ascontiguousarray(u)
    f = np.zeros_like(X)
    hmax = C.c_double(0.0)
    L = lib()
    L.orc_contact_force.restype = C.c_int
    vv = None if v is None else np.ascontiguousarray(v)
    n = L.orc_contact_force(C.c_int64(pairs.shape[0]), _p(pairs), _p(area), C.c_double(K), C.c_double(mu), C.c_double(eps), C.c_double(visc),
                            C.c_double(constKd), C.c_int64(X.shape[0]), _p(X), _p(u), _p(vv) if vv is not None else None, _p(f), C.byref(hmax))
    return f, n, hmax.value


def inertial_force(density, mass_type, conn, X, acc, scale=1.0):
    """ContinuumElementT::FormMa summed over the mesh: M a [nn][3] (mass_type 1 consistent, 2 lumped)"""
    conn = np.ascontiguousarray(conn, np.int32)
    f = np.zeros_like(X)
    err = lib().orc_inertial_force(C.c_double(density), int(mass_type), C.c_int64(conn.shape[0]), _p(conn), _p(np.ascontiguousarray(X)),
                                   _p(np.ascontiguousarray(acc)), C.c_double(scale), _p(f))
    assert err == 0
    return f


def assemble_mass(density, mass_type, constM, conn, X, eqnos, rowptr, colind, val):
    """val += constM * M on the CSR of csr_structure (ContinuumElementT::FormMass through ElementLHSDriver)"""
    conn = np.ascontiguousarray(conn, np.int32)
    err = lib().orc_assemble_mass(C.c_double(density), int(mass_type), C.c_double(constM), C.c_int64(conn.shape[0]), _p(conn),
                                  _p(np.ascontiguousarray(X)), _p(eqnos), _p(rowptr), _p(colind), _p(val))
    assert err == 0
    return val
