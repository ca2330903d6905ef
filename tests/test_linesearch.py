"""Host logic of the nonlinear PCG solver's line search (tahoe_b200/csrc/tb2_linesearch.h) without a device: the library's
proposal / observation state machine against a direct restatement of PCGSolver_LS::Update (PCGSolver_LS.cpp:213-348) on
families of G(s) that drive every branch -- regular secant convergence, sign-preserving brackets, roots beyond max_step
(clamped), negative roots, the trial budget, and the best-step fallback.  The sequences of evaluated steps must be identical."""
import ctypes as C
import math

import numpy as np
import pytest

from tahoe_b200 import capi


def reference_search(G, G0, max_step, abs_tol, rel_tol, max_trials):
    """PCGSolver_LS::Update, restated (test infrastructure): returns (final step, list of evaluated steps)"""
    evaluated = []

    def gvalue(s):
        evaluated.append(s)
        return G(s)

    s_a, G_a, s_b = 0.0, G0, 1.0
    trials = [(s_a, G_a)]
    G_b = gvalue(s_b)
    trials.append((s_b, G_b))
    s_current = s_b
    G_0 = G_b if abs(G_a) > abs(G_b) else G_a
    count, give_up = 2, False
    while True:
        m = (G_a - G_b) / (s_a - s_b)
        b = G_b - m * s_b
        s_new = -b / m
        if s_new > max_step or s_new < 0.0:
            give_up = True
            if s_new > max_step:
                s_new = max_step
                G_new = gvalue(s_new)
                s_current = s_new
                trials.append((s_new, G_new))
                count += 1
            break
        G_new = gvalue(s_new)
        s_current = s_new
        trials.append((s_new, G_new))
        if abs(G_a) > abs(G_new) and abs(G_a) > abs(G_b):
            G_a, s_a = G_new, s_new
        elif abs(G_b) > abs(G_new) and abs(G_b) > abs(G_a):
            G_b, s_b = G_new, s_new
        elif G_b * G_a > 0:
            if G_a * G_new < 0:
                G_a, s_a = G_new, s_new
            elif G_b * G_new < 0:
                G_b, s_b = G_new, s_new
            else:
                give_up = True
        else:
            give_up = True
        count += 1
        if count >= max_trials:
            give_up = True
        if not (abs(G_new) > abs_tol and abs(G_new / G_0) > rel_tol and not give_up):
            break
    if give_up:
        s_best, G_best, best = abs(trials[0][0]), abs(trials[0][1]), 0
        for i in range(1, count):
            s_test, G_test = abs(trials[i][0]), abs(trials[i][1])
            if abs(s_best) < 1.0e-12 or (s_test > 1.0e-12 and G_test < G_best):
                s_best, G_best, best = s_test, G_test, i
        if trials[best][0] != s_current:
            gvalue(trials[best][0])
            s_current = trials[best][0]
    return s_current, evaluated


SLOPE = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def library_search(G, G0, max_step, abs_tol, rel_tol, max_trials):
    evaluated = []

    def cb(s, _):
        evaluated.append(s)
        return G(s)

    L = capi.lib()
    L.tb2_secant_search_host.argtypes = [SLOPE, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_int)]
    final, n = C.c_double(0.0), C.c_int(0)
    assert L.tb2_secant_search_host(SLOPE(cb), None, G0, max_step, abs_tol, rel_tol, max_trials, C.byref(final), C.byref(n)) == 0
    assert n.value == len(evaluated)
    return final.value, evaluated


def families():
    rng = np.random.default_rng(42)
    out = []
    for _ in range(400):
        kind = rng.integers(0, 5)
        a, b, c = rng.standard_normal(3)
        root = rng.uniform(0.05, 4.0)
        if kind == 0:    # linear: one secant step lands on the root (or beyond max_step)
            out.append(lambda s, a=a, root=root: a * (root - s))
        elif kind == 1:  # smooth nonlinear with a root
            out.append(lambda s, a=a, b=b, root=root: a * (root - s) * (1.0 + 0.4 * b * math.tanh(s)))
        elif kind == 2:  # no sign change on [0, max_step]
            out.append(lambda s, a=a, b=b: a * (1.5 + math.cos(b * s)))
        elif kind == 3:  # oscillating: the bracket rules and the trial budget decide
            out.append(lambda s, a=a, b=b, c=c: a * math.sin(3.0 * b * s + c) + 0.1 * c)
        else:            # steep then flat: negative secant roots
            out.append(lambda s, a=a, b=b, root=root: a * (math.exp(-abs(b) * 3.0 * s) - 0.5 * root))
    return out


@pytest.mark.parametrize("max_trials", [0, 2, 3, 10])
def test_state_machine_takes_the_reference_decisions(max_trials):
    n_fallback = n_clamped = 0
    for G in families():
        for max_step, rel_tol in ((2.5, 0.1), (1.2, 0.25), (10.0, 1e-6)):
            G0 = G(0.0)
            if G0 == 0.0:
                continue
            ref = reference_search(G, G0, max_step, 1e-12, rel_tol, max_trials)
            got = library_search(G, G0, max_step, 1e-12, rel_tol, max_trials)
            assert got[1] == ref[1] and got[0] == ref[0]  # same evaluations, bit for bit, and the same final step
            n_clamped += max_step in ref[1]
            n_fallback += len(ref[1]) >= 2 and ref[1][-1] in ref[1][:-1]
    assert n_clamped > 10 and n_fallback > 10  # the sample exercises the clamp and the best-step fallback


def test_linear_slope_converges_in_one_secant_step():
    final, ev = library_search(lambda s: 3.0 * (0.4 - s), 1.2, 2.5, 1e-12, 0.1, 10)
    assert ev[0] == 1.0 and len(ev) == 2 and abs(final - 0.4) < 1e-15
