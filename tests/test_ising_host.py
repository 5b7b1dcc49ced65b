"""Host-side logic of the structured path (no GPU): pulse tables, step grids, angle rows."""
import numpy as np
import pytest

from diffquantum_b200 import pulses as P
from diffquantum_b200 import pulses
from diffquantum_b200.ising import IsingProblem
from oracle import restate as R


def test_step_grid_matches_reference_grid(golden):
    g = golden("demo_bspline_ref")
    n, dt, ts = P.step_grid(0, float(g["T"]), int(g["per_step"]))
    assert n == len(g["ts"]) and dt == float(g["dt"])
    np.testing.assert_array_equal(ts, g["ts"])
    for T0, T in ((0.37, 2.0), (0.0, 0.0049), (1.999, 2.0)):
        a, b = P.step_grid(T0, T, 10), R.step_grid(T0, T, 10)
        assert a[0] == b[0] and a[1] == b[1]
        np.testing.assert_array_equal(a[2], b[2])
    assert P.step_grid(1.2, 0.4, 12, use_abs=True)[0] == 21 and P.step_grid(1.2, 0.4, 12)[0] == 2


def test_u_table_matches_reference_closures(golden):
    for name in ("demo_bspline_ref", "demo_legendre_ref", "h2_vqe_ref"):
        g = golden(name)
        u = P.u_table(g["coeff"], g["omegas"], float(g["T"]), g["ts"], str(g["basis"]))
        np.testing.assert_allclose(u, g["u_tab"], rtol=1e-13, atol=1e-15)


def test_poly_and_fourier_bases_match_reference_closures(golden):
    """generate_u with basis 'poly' / 'Fourier' (sim_plain.py:84-92): the host table and the oracle's scalar restatement
    against the values of the reference's own closures; the estimator is undefined for them in the reference (it raises at
    :178) and is rejected here."""
    for name in ("demo_poly_ref", "demo_fourier_ref"):
        g = golden(name)
        basis = str(g["basis"])
        u = P.u_table(g["coeff"], g["omegas"], float(g["T"]), g["ts"], basis)
        np.testing.assert_allclose(u, g["u_tab"], rtol=1e-13, atol=1e-15)      # np.exp vs math.exp: 1 ulp
        ref = np.array([[R.u_plain(i, t, g["coeff"], g["omegas"], float(g["T"]), basis) for i in range(len(g["omegas"]))] for t in g["ts"]])
        np.testing.assert_array_equal(ref, g["u_tab"])
        assert str(g["estimator_error"]) == "UnboundLocalError"
        with pytest.raises(ValueError):
            P.dudc_table(g["coeff"], g["omegas"], float(g["T"]), 0.3, basis)
        with pytest.raises(ValueError):
            P.dudc_tables(g["coeff"], g["omegas"], float(g["T"]), [0.3, 0.4], basis)


def test_dudc_matches_autograd_restatement():
    rng = np.random.RandomState(3)
    coeff = rng.normal(0, 1, [5, 6])
    om = rng.uniform(1, 3, 5)
    for basis in ("BSpline", "Legendre"):
        d = P.dudc_table(coeff, om, 1.7, 0.61, basis)
        ref = np.array([R.dudc_plain(i, 0.61, coeff, om, 1.7, basis) for i in range(5)])
        np.testing.assert_allclose(d, ref, rtol=1e-14, atol=1e-16)


def test_f_u_table_matches_restated_cc(golden):
    g = golden("diffqc_cc_restated")
    channels, k = [], 0
    for c in g["chan_counts"]:
        channels.append([list(g["chan_flat"][k + i]) for i in range(c)])
        k += c
    for func_type in (0, 1):
        for tag, (T0, T) in zip(("fwd", "bwd"), g["spans"]):
            n, dt, ts = P.step_grid(float(T0), float(T), int(g["per_step"]), use_abs=True)
            u = P.f_u_table(channels, float(g["duration"]), func_type, g["vv"], ts)
            np.testing.assert_allclose(u, g["f%d_%s_u" % (func_type, tag)], rtol=1e-13, atol=1e-15)


def test_maxcut_angle_rows_reproduce_oracle_phases():
    edges = R.random_regular_edges(6, seed=1)
    prob = IsingProblem.maxcut(6, edges)
    ref = R.maxcut_structured(6, edges)
    assert prob.T == ref["T"] and prob.n_zz == len(edges)
    coeff = np.random.RandomState(0).normal(0, 1, [len(prob.terms), 6])
    n, dt, ts = P.step_grid(0, prob.T, 10)
    u = P.u_table(coeff, prob.omegas, prob.T, ts)
    rows = prob.angle_rows(u, dt)
    k = 7
    angle = np.full(64, rows[k, 0])
    for e, (a, b) in enumerate(prob.zz_pairs):
        angle += rows[k, 1 + e] * R.z_diag(6, a) * R.z_diag(6, b)
    want = dt * ref["h0_diag"]
    for i, t in enumerate(ref["terms"]):
        if t[0] == "zz":
            want = want + dt * u[k, i] * R.term_diag(ref, t)
    np.testing.assert_allclose(angle, want, atol=1e-13)
    np.testing.assert_allclose(rows[k, 1 + prob.n_zz:], dt * u[k, len(edges):], atol=1e-16)


def test_batched_grids_and_tables_are_bit_identical_to_per_sample_calls():
    """step_grids / dudc_tables (one vectorised call per gradient batch) against step_grid / dudc_table per sample,
    including the end points s = 0 and s = T (ten zero-length steps, as the reference does at sim_plain.py:123,133)."""
    rng = np.random.RandomState(3)
    T = 2.0
    s = rng.uniform(size=200) * T
    s[0], s[1] = 0.0, T
    for basis, n_basis in (("BSpline", 6), ("Legendre", 5)):
        coeff = rng.normal(size=(7, n_basis))
        om = rng.uniform(1, 4, size=7)
        for T0s, T1s in ((0.0, s), (s, T)):
            n, dt, ts = pulses.step_grids(T0s, T1s, 10)
            a, b = np.broadcast_arrays(np.asarray(T0s, float), np.asarray(T1s, float))
            off = 0
            U = pulses.u_table(coeff, om, T, ts, basis)
            for i in range(len(s)):
                n1, dt1, ts1 = pulses.step_grid(float(a[i]), float(b[i]), 10)
                assert n1 == n[i] and dt1 == dt[i]
                assert np.array_equal(ts1, ts[off:off + n1])
                assert np.array_equal(pulses.u_table(coeff, om, T, ts1, basis), U[off:off + n1])
                off += n1
            assert off == len(ts)
        D = pulses.dudc_tables(coeff, om, T, s, basis)
        for i in range(len(s)):
            assert np.array_equal(D[i], pulses.dudc_table(coeff, om, T, s[i], basis))
    n, dt, ts = pulses.step_grids([1.0, 3.0], [0.5, -1.0], 1)          # negative spans: int() truncation, zero steps allowed
    assert list(n) == [int(1 * ((0.5 - 1.0) + 1)), 0] and len(ts) == int(n.sum())


def test_measurement_noise_consumes_the_global_stream_in_reference_order():
    """is_noisy (sim_plain.py:207-208,217-218): per control, ps_p then ps_m, one normal(scale=|ps|/5) each."""
    en = np.arange(1.0, 13.0).reshape(2, 3, 2)
    np.random.seed(5)
    want = en.copy()
    for b in range(2):
        for i in range(3):
            want[b, i, 0] += np.random.normal(scale=abs(want[b, i, 0]) / 5)
            want[b, i, 1] += np.random.normal(scale=abs(want[b, i, 1]) / 5)
    np.random.seed(5)
    got = pulses.add_measurement_noise(en.copy())
    assert np.array_equal(got, want)


def test_control_order_x_before_zz_is_rejected():
    """the device applies the diagonal factor first, then every X rotation; a list with an X control before
    a ZZ control is a different (non-commuting) product under diffqc.cc:155-164 and must not be accepted silently."""
    import pytest
    from diffquantum_b200.ising import IsingProblem
    with pytest.raises(ValueError):
        IsingProblem(2, [('x', 0), ('zz', 0, 1)], [1.0, 1.0], 1.0)
    with pytest.raises(ValueError):
        IsingProblem(3, [('zz', 0, 1), ('x', 2), ('zz', 1, 2), ('x', 0)], [1.0] * 4, 1.0)
    p = IsingProblem(3, [('zz', 0, 1), ('zz', 1, 2), ('x', 2), ('x', 0), ('x', 2)], [1.0] * 5, 1.0)   # ZZ first: fine
    rows = p.angle_rows(np.array([[0.1, 0.2, 0.3, 0.4, 0.5]]), 0.5)
    np.testing.assert_allclose(rows[0, 1 + p.n_zz:], [0.2, 0.0, 0.4])       # commuting X pulses on one qubit add


def test_interleaved_commuting_order_matches_oracle_semantics():
    """ZZ-then-X lists in any internal order give the oracle's list-order product (everything that is reordered commutes)."""
    from diffquantum_b200.ising import IsingProblem
    from oracle import restate as R
    terms = [('zz', 1, 2), ('zz', 0, 1), ('x', 2), ('x', 0), ('x', 1)]
    p = IsingProblem(3, terms, [1.0] * 5, 1.0, h0_zz={(0, 2): 0.3}, h0_const=0.1)
    u = np.random.RandomState(0).normal(size=(4, 5))
    rows = p.angle_rows(u, 0.25)
    # replay the angle rows with plain numpy: diagonal phase then rotations
    prob = dict(n=3, terms=terms, h0_diag=0.1 + 0.3 * R.z_diag(3, 0) * R.z_diag(3, 2))
    psi0 = np.random.RandomState(1).normal(size=8) + 1j * np.random.RandomState(2).normal(size=8)
    want = R.evolve_split_structured(prob, u, 0.25, psi0)
    psi = psi0.copy()
    for k in range(4):
        ang = rows[k, 0] + sum(rows[k, 1 + e] * R.z_diag(3, a) * R.z_diag(3, b) for e, (a, b) in enumerate(p.zz_pairs))
        psi = psi * np.exp(-1j * ang)
        for q in range(3):
            th = rows[k, 1 + p.n_zz + q]
            v = psi.reshape(1 << q, 2, -1)
            psi = np.stack([np.cos(th) * v[:, 0] - 1j * np.sin(th) * v[:, 1], np.cos(th) * v[:, 1] - 1j * np.sin(th) * v[:, 0]], axis=1).reshape(-1)
    np.testing.assert_allclose(psi, want, atol=1e-14)


def test_angle_rows_indexed_add_equals_the_term_loop():
    """IsingProblem.angle_rows adds every pulse column with one indexed add when no two controls share a pair / qubit; same bits
    as the term-by-term loop, which stays for repeated controls."""
    import numpy as np
    from diffquantum_b200.ising import IsingProblem
    from oracle import restate as R

    def loop(p, u, dt):
        rows = np.zeros((u.shape[0], p.row_len))
        rows[:, 0] = p.h0_const
        rows[:, 1:1 + p.n_zz] = p.h0_zz[None, :]
        for i in range(len(p.terms)):
            col = (1 + p.term_index[i]) if p.term_kind[i] == 0 else (1 + p.n_zz + p.term_index[i])
            rows[:, col] += u[:, i]
        return rows * dt

    prob = IsingProblem.maxcut(12, R.random_regular_edges(12, seed=3))
    u = np.random.default_rng(1).normal(size=(37, len(prob.terms)))
    assert np.array_equal(loop(prob, u, 0.013), prob.angle_rows(u, 0.013))
    rep = IsingProblem(4, [('zz', 0, 1), ('x', 0), ('x', 0), ('x', 1)], [1.0, 0.5, 0.25, 2.0], 1.0)
    u = np.random.default_rng(2).normal(size=(5, 4))
    assert np.array_equal(loop(rep, u, 0.1), rep.angle_rows(u, 0.1))
