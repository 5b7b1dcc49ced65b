"""Pins the plain-C restatement (oracle/c, the multi-core CPU baseline of bench.py) to the NumPy
restatement, which is itself pinned to the reference's own outputs (tests/test_oracle.py)."""
import numpy as np
import pytest

from oracle import c_port as C, restate as R

pytestmark = pytest.mark.skipif(not C.available(), reason="oracle/c not built (run __graft_entry__.build())")


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (4, 2), (7, 3), (10, 4)])
def test_c_split_matches_numpy_restatement(n, seed):
    if n >= 4 and n % 2 == 0:
        edges = R.random_regular_edges(n, seed=seed)
    else:
        edges = [(i, i + 1) for i in range(n - 1)]
    prob = R.maxcut_structured(n, edges)
    cp = C.CProblem(prob)
    coeff = np.random.RandomState(seed).normal(0, 1, [len(prob["terms"]), 6])
    ns, dt, ts = R.step_grid(0.1, 1.9, 5)
    u = R.coef_table_plain(coeff, prob["omegas"], prob["T"], ts)
    want = R.evolve_split_structured(prob, u, dt, prob["psi0"])
    got = cp.evolve(u, dt, prob["psi0"].copy())
    np.testing.assert_allclose(got, want, atol=1e-13)
    assert abs(cp.energy(got) - R.energy_diag(prob["m_diag"], want)) < 1e-12


def test_c_split_matches_golden(golden):
    g = golden("split_n8")
    prob = R.maxcut_structured(int(g["n"]), g["edges"].tolist())
    cp = C.CProblem(prob)
    got = cp.evolve(g["u"], float(g["dt"]), prob["psi0"].copy())
    np.testing.assert_allclose(got, g["final"], atol=1e-13)
    grad, en, steps = C.grad_mc(cp, g["coeff"], float(g["s"][0]), int(g["per_step"]), return_energies=True)
    np.testing.assert_allclose(en, g["energies"][0], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(grad, g["grads"][0], rtol=1e-10, atol=1e-12)
    s = float(g["s"][0])
    ps = int(g["per_step"])
    assert steps == int(ps * (s + 1)) + 2 * len(prob["terms"]) * int(ps * (prob["T"] - s + 1))


def test_c_shift_gate_is_the_reference_gate():
    prob = R.maxcut_structured(5, [(0, 1), (1, 2), (2, 3), (3, 4), (0, 4)])
    cp = C.CProblem(prob)
    rng = np.random.RandomState(0)
    phi = rng.normal(size=32) + 1j * rng.normal(size=32)
    for i, term in enumerate(prob["terms"]):
        for sign in (+1, -1):
            np.testing.assert_allclose(cp.shift(i, phi, sign), R.apply_shift_gate(prob, term, phi, sign), atol=1e-15)


def test_host_core_override_ignores_the_launcher_setting():
    """torchrun exports OMP_NUM_THREADS=1; bench.py's CPU arm asks for every core of the host explicitly."""
    import os
    from oracle import c_port as C
    if not C.available():
        import pytest
        pytest.skip("oracle/c/liboracle_c.so not built")
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    C.load().oc_set_num_threads(1)
    assert C.num_threads() == 1
    assert C.use_host_cores() == want
