"""diffqc.cc:75-135 (my_expit, bspline, f_u): the oracle restatement, the product's NumPy table and the product's
host routine inside the shared library, each against the reference's OWN code compiled from the reference source
(oracle/ref_cc/Makefile -> oracle/_ref/libfu.so; fixture tests/golden/fu_cc_ref.npz made with it)."""
import numpy as np
import pytest

from diffquantum_b200 import pulses
from oracle import ref_cc, restate as R


def cases(g):
    counts, flat = g["chan_counts"], g["chan_flat"]
    channels, k = [], 0
    for c in counts:
        channels.append([list(flat[k + i]) for i in range(c)])
        k += c
    for i in range(int(g["n_cases"])):
        func_type, n_basis = (int(v) for v in g["meta_%d" % i])
        yield channels, func_type, g["vv_%d" % i], g["ts_%d" % i], g["u_%d" % i]


def close(a, b):
    # sums of <= 3 channel terms of size <= 2: a few ulp of 1 absolute, 1e-13 relative
    return np.abs(a - b).max() <= 1e-13 * max(1.0, np.abs(b).max())


def test_fixture_walks_every_branch(golden):
    g = golden("fu_cc_ref")
    assert "REFERENCE" in str(g["source"])
    np.testing.assert_array_equal(g["expit_y"][[0, 1, -3, -1]], [0.0, 1 / (1 + np.exp(32.0)), 1 / (1 + np.exp(-32.0)), 1.0])
    for channels, func_type, vv, ts, u in cases(g):
        assert np.all(u[:, 3] == 0.0)                       # term without channels
        assert np.all(np.isfinite(u))


def test_restatement_matches_compiled_reference(golden):
    g = golden("fu_cc_ref")
    for channels, func_type, vv, ts, u in cases(g):
        mine = np.array([[R.f_u_cc(h, float(t), vv, channels, 2.0, func_type) for h in range(len(channels))] for t in ts])
        assert close(mine, u)
    for x, y in zip(g["expit_x"], g["expit_y"]):
        assert R.expit_cc(float(x)) == y


def test_product_numpy_table_matches_compiled_reference(golden):
    g = golden("fu_cc_ref")
    for channels, func_type, vv, ts, u in cases(g):
        assert close(pulses.f_u_table(channels, 2.0, func_type, vv, ts), u)


def test_product_library_routine_matches_compiled_reference(golden):
    """dq_pulse_f_u_table is the host code dq_dense_trotter evaluates its pulses with (same g++ libm as the reference
    build, same operation order): bit-equal on this toolchain."""
    g = golden("fu_cc_ref")
    for channels, func_type, vv, ts, u in cases(g):
        got = pulses.f_u_table_lib(channels, 2.0, func_type, vv, ts)
        assert close(got, u)
    with pytest.raises(ValueError):
        pulses.f_u_table_lib([[[0.0, 1.0, 0.0, 7.0]]], 2.0, 0, np.zeros((2, 2, 3)), [0.1])


@pytest.mark.skipif(not ref_cc.available(), reason="oracle/_ref/libfu.so not built (needs /root/reference)")
def test_live_compiled_reference_random_inputs():
    rng = np.random.RandomState(5)
    for func_type in (0, 1):
        vv = rng.normal(0, 2, [2, 3, 7])
        channels = [[[0.0, rng.normal(), rng.normal(), float(rng.randint(3))] for _ in range(rng.randint(1, 4))] for _ in range(5)]
        ts = rng.uniform(-0.2, 3.2, size=40)
        ref = ref_cc.f_u_table(channels, 3.0, func_type, vv, ts)
        assert close(pulses.f_u_table_lib(channels, 3.0, func_type, vv, ts), ref)
        assert close(pulses.f_u_table(channels, 3.0, func_type, vv, ts), ref)
        mine = np.array([[R.f_u_cc(h, float(t), vv, channels, 3.0, func_type) for h in range(5)] for t in ts])
        assert close(mine, ref)
        lib = ref_cc.load()
        for b in range(7):
            for x in (-0.1, 0.0, 0.2, 0.55, 1.0, 1.3):
                assert lib.ref_bspline(b, 7, x) == R.bspline_value(b, 7, x)
