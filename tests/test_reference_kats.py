"""The reference's own known-answer / property tests for this path (tests/materials/test_ExplJ2Plasticity.cpp,
tests/materials/test_IsotropicT.cpp), restated against the oracle so that the oracle is pinned on them too.  (The reference builds
them with googletest, which needs the network; the assertions, tolerances and load paths below are theirs.)"""
import numpy as np
import pytest


def _steel(oracle):  # test_ExplJ2Plasticity.cpp:33-38 MakeSteel
    E, nu = 200e3, 0.3
    return oracle.material({"type": "explicit_J2", "mu": E / (2 * (1 + nu)), "kappa": E / (3 * (1 - 2 * nu)), "density": 7.85e-9,
                            "sigma_Y": 250.0, "hardening_modulus": 1000.0})


def _h0():
    h = np.zeros(16)
    h[[0, 4, 8]] = 1.0
    return h


def _mises(sig):
    p = sig[:3].sum() / 3.0
    d = sig[:3] - p
    return np.sqrt(1.5 * ((d * d).sum() + 2.0 * (sig[3:] ** 2).sum()))


def test_initialize_history_sets_Fn_to_identity(oracle):  # :43-59
    h = oracle.explicit_solid_history(10)
    assert np.all(h[:, :, [0, 4, 8]] == 1.0) and np.all(h[:, :, 9:] == 0.0) and np.all(h[:, :, [1, 2, 3, 5, 6, 7]] == 0.0)


def test_elastic_bulk_below_yield(oracle):  # :73-102
    mat, h, eps = _steel(oracle), _h0(), 1e-4
    sig = oracle.explicit_material_stress(mat, np.eye(3) * (1.0 + eps), h)
    expected = 3.0 * (200e3 / (3 * (1 - 0.6))) * eps
    assert np.abs(sig[:3] - expected).max() < 1e-6 * expected and np.abs(sig[3:]).max() < 1e-10 and abs(h[15]) < 1e-12


def test_pure_shear_yields_at_mises_criterion(oracle):  # :113-152
    mat, h = _steel(oracle), _h0()
    mu = 200e3 / 2.6
    F = np.eye(3)
    F[0, 1] = 1.5 * 250.0 / (mu * np.sqrt(3.0))
    sig = oracle.explicit_material_stress(mat, F, h)
    assert h[15] > 0.0 and abs(_mises(sig) - (250.0 + 1000.0 * h[15])) < 1e-4 * (250.0 + 1000.0 * h[15])


def test_rigid_rotation_gives_no_plastic_strain_and_bounded_drift(oracle):  # :190-235
    mat, h = _steel(oracle), _h0()
    for step in range(1, 101):
        th = 0.001 * step
        F = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
        sig = oracle.explicit_material_stress(mat, F, h)
    assert np.abs(sig).max() < 30.0 and abs(h[15]) < 1e-10


def test_uniaxial_stretch_hardening_curve(oracle):  # :243-286
    E, nu = 117e3, 0.35
    mat = oracle.material({"type": "explicit_J2", "mu": E / (2 * (1 + nu)), "kappa": E / (3 * (1 - 2 * nu)), "density": 8.96e-9,
                           "sigma_Y": 90.0, "hardening_modulus": 150.0})
    h = _h0()
    for step in range(1, 201):
        lz = 1.0 + 0.10 * step / 200
        sig = oracle.explicit_material_stress(mat, np.diag([1 / np.sqrt(lz), 1 / np.sqrt(lz), lz]), h)
    assert h[15] > 0.0 and abs(_mises(sig) - (90.0 + 150.0 * h[15])) < 0.05 * (90.0 + 150.0 * h[15])


def test_elastic_load_unload_path_dependence(oracle):  # :300-end
    mat, h = _steel(oracle), _h0()
    gamma_peak = 0.5 * 250.0 / (200e3 / 2.6 * np.sqrt(3.0))
    for step in range(1, 101):
        F = np.eye(3)
        F[0, 1] = gamma_peak * step / 100
        sig = oracle.explicit_material_stress(mat, F, h)
    peak_q = abs(sig[5]) * np.sqrt(3.0)
    for step in range(1, 101):
        F = np.eye(3)
        F[0, 1] = gamma_peak * (1.0 - step / 100)
        sig = oracle.explicit_material_stress(mat, F, h)
    assert peak_q > 0 and abs(sig[5]) * np.sqrt(3.0) < 0.05 * peak_q and h[15] == 0.0


def test_isotropic_moduli_relations(oracle):  # tests/materials/test_IsotropicT.cpp:13-35
    E, nu = 210e3, 0.3
    m = oracle.material({"type": "small_strain_StVenant", "E": E, "nu": nu, "density": 1.0})
    assert abs(m.mu - E / (2 * (1 + nu))) < 1.0 and abs(m.lam - E * nu / ((1 + nu) * (1 - 2 * nu))) < 1.0
    assert abs(m.kappa - (m.lam + 2.0 * m.mu / 3.0)) < 1e-9 * m.kappa
    mk = oracle.material({"type": "Simo_isotropic", "mu": m.mu, "kappa": m.kappa, "density": 1.0})
    E_back = 9.0 * mk.kappa * mk.mu / (3.0 * mk.kappa + mk.mu)
    assert abs(E_back - E) < 1.0e3 and abs(mk.lam - m.lam) < 1e-9 * m.lam
    # the harness binding of the product derives the same constants
    from tahoe_b200 import capi
    c = capi.material({"type": "small_strain_StVenant", "E": E, "nu": nu, "density": 1.0})
    assert abs(c.mu - m.mu) < 1e-9 * m.mu and abs(c.lam - m.lam) < 1e-9 * m.lam and abs(c.kappa - m.kappa) < 1e-9 * m.kappa
