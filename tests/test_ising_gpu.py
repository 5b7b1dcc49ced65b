"""GPU parity tests of the structured path through the C ABI: CUDA result vs the oracle
(oracle/restate.py, split step = diffqc.cc:155-164) and vs the committed golden fixtures.
Tolerance: 1e-10 relative on amplitudes, energies and per-sample gradients (BASELINE north_star)."""
import numpy as np
import pytest

import diffquantum_b200 as dq
from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL = 1e-10


def graph_for(n):
    """3-regular when that exists (n even, n >= 4); otherwise a ring with a few chords."""
    if n >= 4 and n % 2 == 0:
        return R.random_regular_edges(n, seed=n)
    edges = [(i, i + 1) for i in range(n - 1)]
    if n >= 3:
        edges.append((0, n - 1))
    edges += [(i, (i + n // 2) % n) for i in range(0, n // 2, 2) if n >= 5]
    return sorted(set(tuple(sorted(e)) for e in edges))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("name", ["split_n4_demo", "split_n6", "split_n8", "split_n12"])
@pytest.mark.parametrize("engine", [0, 1])
def test_split_golden_fixture(golden, name, engine):
    g = golden(name)
    n = int(g["n"])
    prob = dq.IsingProblem.maxcut(n, g["edges"].tolist())
    sim = dq.IsingSimulator(prob, per_step=int(g["per_step"]), engine=engine)
    assert sim.info("engine") == (engine if n >= 12 else 0)        # the fused engines need a 2^12 tile
    psi, en = sim.evolve(g["coeff"], 0, prob.T)
    assert rel(psi[0], g["final"]) < TOL
    assert abs(en[0] - float(g["energy"])) < TOL * abs(float(g["energy"]))
    grads, energies = sim.grad_samples(g["coeff"], g["s"], return_energies=True)
    assert rel(energies, g["energies"]) < TOL
    assert rel(grads, g["grads"]) < TOL


@pytest.mark.parametrize("n,per_step,engine", [(1, 5, 0), (2, 5, 0), (3, 4, 0), (10, 3, 0), (13, 3, 1),
                                               (14, 2, 1), (15, 2, 1), (16, 1, 1), (16, 1, 0), (17, 1, 1), (18, 1, 1),
                                               (20, 1, 1), (19, 1, 1), (12, 3, 1)])
def test_evolve_vs_oracle_live(n, per_step, engine):
    edges = graph_for(n)
    prob = dq.IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    coeff = np.random.RandomState(n).normal(0, 1, [len(prob.terms), 6])
    sim = dq.IsingSimulator(prob, per_step=per_step, engine=engine)
    for (T0, T1) in ((0, prob.T), (0.31, 1.17)):
        ns, dt, ts = R.step_grid(T0, T1, per_step)
        u = R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts)
        want = R.evolve_split_structured(ref, u, dt, ref["psi0"])
        psi, en = sim.evolve(coeff, T0, T1)
        assert rel(psi[0], want) < TOL
        assert abs(en[0] - R.energy_diag(ref["m_diag"], want)) < TOL * max(1.0, abs(en[0]))


def test_nonuniform_psi0_and_batch():
    n = 9
    edges = R.random_regular_edges(n + 1, seed=2)
    edges = [e for e in edges if max(e) < n]
    prob = dq.IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    rng = np.random.RandomState(1)
    psi0 = rng.normal(size=(3, 1 << n)) + 1j * rng.normal(size=(3, 1 << n))
    psi0 /= np.linalg.norm(psi0, axis=1, keepdims=True)
    coeff = rng.normal(0, 1, [len(prob.terms), 6])
    sim = dq.IsingSimulator(prob, per_step=5)
    out, en = sim.evolve(coeff, 0.2, 1.9, psi0=psi0)
    ns, dt, ts = R.step_grid(0.2, 1.9, 5)
    u = R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts)
    for b in range(3):
        want = R.evolve_split_structured(ref, u, dt, psi0[b])
        assert rel(out[b], want) < TOL


def test_torch_device_tensor_handoff():
    import torch
    n = 8
    edges = R.random_regular_edges(n, seed=4)
    prob = dq.IsingProblem.maxcut(n, edges)
    sim = dq.IsingSimulator(prob, per_step=6)
    coeff = np.random.RandomState(2).normal(0, 1, [len(prob.terms), 6])
    host, _ = sim.evolve(coeff, 0, prob.T)
    psi0 = torch.full((1, 1 << n), 1.0 / np.sqrt(2.0 ** n), dtype=torch.complex128, device="cuda")
    dev, _ = sim.evolve(coeff, 0, prob.T, psi0=psi0)
    assert dev.is_cuda and rel(dev.cpu().numpy(), host) < 1e-14


def test_gradient_vs_oracle_live_n10():
    n = 10
    edges = R.random_regular_edges(n, seed=7)
    prob = dq.IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    coeff = np.random.RandomState(9).normal(0, 1, [len(prob.terms), 6])
    sim = dq.IsingSimulator(prob, per_step=6)
    s_list = [0.013, 0.77, 1.991]
    grads, energies = sim.grad_samples(coeff, s_list, return_energies=True)
    for b, s in enumerate(s_list):
        g_ref, e_ref = R.grad_mc_structured(ref, coeff, s, 6, mode="split", return_energies=True)
        assert rel(energies[b], e_ref) < TOL
        assert rel(grads[b], g_ref) < TOL
        # estimator identity: ps = 2 Im <phi| U^+ M U H_i |phi> is independent of r
    g2 = sim.grad_samples(coeff, s_list, r=0.25)
    assert rel(g2, grads) < 1e-9


def test_norm_is_preserved_and_errors_are_python_exceptions():
    prob = dq.IsingProblem.maxcut(6, R.random_regular_edges(6, seed=1))
    sim = dq.IsingSimulator(prob, per_step=10)
    coeff = np.random.RandomState(0).normal(0, 1, [len(prob.terms), 6])
    psi, _ = sim.evolve(coeff, 0, prob.T)
    assert abs(np.linalg.norm(psi[0]) - 1) < 1e-13
    rows = prob.trajectory_rows(coeff, 0, prob.T, 10)
    rows[3, 2] = np.nan
    with pytest.raises(ValueError):
        sim.evolve_rows(rows)
    with pytest.raises(ValueError):
        dq.IsingProblem(3, [("zz", 0, 3)], [1.0], 1.0)


@pytest.mark.parametrize("n,engine", [(12, 1), (15, 1), (16, 1), (17, 1), (20, 1)])
def test_fused_gradients_agree_with_generic_engine(n, engine):
    """Every shifted ket (each ZZ pair and each X qubit, both signs) through the fused passes vs the
    one-kernel-per-term engine on the same device; the generic engine is pinned to the oracle above."""
    edges = graph_for(n)
    prob = dq.IsingProblem.maxcut(n, edges)
    coeff = np.random.RandomState(100 + n).normal(0, 1, [len(prob.terms), 6])
    s_list = [0.41, 1.63]
    ref = dq.IsingSimulator(prob, per_step=2, engine=0).shifted_energies(coeff, s_list)
    sim = dq.IsingSimulator(prob, per_step=2, engine=engine)
    assert sim.info("engine") == engine
    got = sim.shifted_energies(coeff, s_list)
    assert rel(got, ref) < TOL
    if engine == 1:                      # estimator by linearity: n_H + 1 suffix trajectories, same outputs
        sim.set_option("linear", 1)
        lin = sim.shifted_energies(coeff, s_list)
        assert sim.stat("steps") < 0.6 * 2 * len(prob.terms) * 3 * len(s_list) + 50
        sim.set_option("linear", 0)
        assert rel(lin, ref) < TOL
        g_lin = sim.assemble_gradients(coeff, s_list, lin)
        g_ref = sim.assemble_gradients(coeff, s_list, ref)
        assert rel(g_lin, g_ref) < TOL
    # large angles force the unscaled (cos, sin) butterflies
    big = coeff * 6.0
    sim2 = dq.IsingSimulator(prob, per_step=1, engine=engine)
    ref2 = dq.IsingSimulator(prob, per_step=1, engine=0).shifted_energies(big, s_list[:1])
    assert rel(sim2.shifted_energies(big, s_list[:1]), ref2) < TOL
    if engine == 1:
        sim2.set_option("linear", 1)
        assert rel(sim2.shifted_energies(big, s_list[:1]), ref2) < TOL


def test_exact_step_reproduces_the_reference_fixture(golden):
    """The structured path with step='exact' against the reference's OWN outputs (demo_maxcut.py problem,
    SimulatorPlain.trotter / compute_energy_grad_MC run unmodified: tests/golden/demo_bspline_ref.npz)."""
    g = golden("demo_bspline_ref")
    prob = dq.IsingProblem.maxcut(4, [[0, 1], [0, 3], [1, 2], [2, 3]])
    sim = dq.IsingSimulator(prob, per_step=int(g["per_step"]), step="exact")
    psi, en = sim.evolve(g["coeff"], 0, prob.T)
    assert rel(psi[0], g["final"]) < TOL
    assert abs(en[0] - complex(g["energy"]).real) < TOL * abs(complex(g["energy"]))
    for k, s in enumerate(g["s"][:3]):
        phi, _ = sim.evolve(g["coeff"], 0, float(s))
        assert rel(phi[0], g["phis"][k]) < TOL
    grads = sim.grad_samples(g["coeff"], g["s"])
    assert rel(grads, g["grads"]) < TOL


@pytest.mark.parametrize("n,per_step", [(6, 4), (10, 3), (14, 2), (16, 1)])
def test_exact_step_vs_matrix_free_oracle(n, per_step):
    """expm_multiply restatement of sim_plain.py:147 (oracle/restate.py evolve_exact_structured)."""
    edges = graph_for(n)
    prob = dq.IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    coeff = np.random.RandomState(n).normal(0, 1, [len(prob.terms), 6])
    sim = dq.IsingSimulator(prob, per_step=per_step, step="exact")
    assert sim.info("engine") == 0
    ns, dt, ts = R.step_grid(0.3, 1.1, per_step)
    u = R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts)
    want = R.evolve_exact_structured(ref, u, dt, ref["psi0"])
    psi, en = sim.evolve(coeff, 0.3, 1.1)
    assert rel(psi[0], want) < TOL
    assert abs(np.linalg.norm(psi[0]) - 1) < 1e-12
    if n <= 10:
        g_ref, e_ref = R.grad_mc_structured(ref, coeff, 0.9, per_step, mode="exact", return_energies=True)
        grads, energies = sim.grad_samples(coeff, [0.9], return_energies=True)
        assert rel(energies[0], e_ref) < TOL and rel(grads[0], g_ref) < TOL


@pytest.mark.parametrize("name", ["split_ref_n4", "split_ref_n6"])
def test_split_step_matches_reference_executed_fixture(golden, name):
    """CUDA vs output of the reference's OWN product-form lines (sim_plain.py:139,142 enabled in memory by
    oracle/ref_loader.load_sim_plain_split; fixture made by oracle/make_golden.py split_ref)."""
    g = golden(name)
    n = int(g["n"])
    prob = dq.IsingProblem.maxcut(n, g["edges"].tolist())
    sim = dq.IsingSimulator(prob, per_step=int(g["per_step"]))
    psi, en = sim.evolve(g["coeff"], 0, prob.T)
    assert rel(psi[0], g["final"]) < TOL
    assert abs(en[0] - complex(g["energy"]).real) < TOL * abs(complex(g["energy"]))
    part, _ = sim.evolve(g["coeff"], float(g["partial_span"][0]), float(g["partial_span"][1]))
    assert rel(part[0], g["partial"]) < TOL
    for k, s in enumerate(g["s"]):
        phi, _ = sim.evolve(g["coeff"], 0, float(s))
        assert rel(phi[0], g["phis"][k]) < TOL
    assert rel(sim.grad_samples(g["coeff"], g["s"]), g["grads"]) < TOL


def test_headline_config_gradients_vs_c_port():
    """BASELINE configs[3] at its own size: n=20 random 3-regular MaxCut (the bench graph and coefficients),
    per_step=10, fused engine, through IsingSimulator.grad_samples -- against the plain-C port of the reference step
    and estimator (oracle/c, pinned to restate.py, which is pinned to the reference-executed fixtures)."""
    from oracle import c_port as C
    if not C.available():
        pytest.skip("oracle/c/liboracle_c.so not built")
    import networkx as nx
    n, per_step = 20, 10
    g = nx.random_regular_graph(3, n, seed=0)
    edges = sorted(tuple(sorted(e)) for e in g.edges())
    prob = dq.IsingProblem.maxcut(n, edges)
    n_H = len(prob.terms)
    coeff = np.random.default_rng(0).normal(0, 1, [n_H, 6])
    sim = dq.IsingSimulator(prob, per_step=per_step, engine=1)
    assert sim.info("engine") == 1
    s = 1.87                                    # 28 prefix steps, 11 suffix steps per shifted ket
    grads, energies = sim.grad_samples(coeff, [s], return_energies=True)
    C.use_host_cores()
    ref = R.maxcut_structured(n, edges)
    terms = [0, 17, 29, 30, 41, 49]             # three ZZ controls, three X controls (low, middle, high bit)
    g_ref, e_ref, _ = C.grad_mc(C.CProblem(ref), coeff, s, per_step, terms=terms, return_energies=True)
    assert rel(energies[0][terms], e_ref[terms]) < TOL
    assert rel(grads[0][terms], g_ref[terms]) < TOL
    assert np.abs(g_ref[terms]).max() > 1e-3    # not a comparison of zeros


def test_device_resident_training_exact_step_follows_the_reference_run(golden):
    """dq_ising_train with step='exact' on the demo problem (n=4, generic engine): pulses, angle rows, evolution, K=1 gradient
    sample and Adam on the device for 202 epochs -- against the reference's OWN demo_maxcut.py run (np.random.seed(0))."""
    g = golden("demo_training_ref")
    prob = dq.IsingProblem.maxcut(4, [[0, 1], [0, 3], [1, 2], [2, 3]])
    sim = dq.IsingSimulator(prob, per_step=10, step="exact")
    np.random.seed(0)
    dev = dq.EnergyTrainer(sim, n_basis=6, n_epoch=202, lr=2e-2, device_resident=True, ground_energy=-4.0)
    dev.train_energy()
    assert np.abs(np.array(dev.losses_energy) - g["losses_energy"]).max() < 1e-8      # Adam's rounding over 202 steps
    assert np.abs(dev.spectral_coeff.detach().numpy() - g["final_coeff"]).max() < 1e-8
    assert rel(dev.final_state, g["final_state"]) < 1e-8
    assert dev.find_state()[0] in (0b0101, 0b1010)
    assert dev.device_ms > 0


@pytest.mark.parametrize("n,engine,K,n_epoch", [(6, 0, 3, 8), (12, 1, 2, 5), (13, 1, 1, 4)])
def test_device_resident_training_matches_host_loop(n, engine, K, n_epoch):
    """Split step: the device-resident loop against the host-driven EnergyTrainer (NumPy pulse tables, torch Adam) on the same
    sample times; both engines; K > 1 = mean over the epoch's samples; non-uniform start state at n=6."""
    edges = graph_for(n)
    psi0 = None
    if n == 6:
        rs = np.random.RandomState(1)
        psi0 = rs.normal(size=1 << n) + 1j * rs.normal(size=1 << n)
        psi0 /= np.linalg.norm(psi0)
    prob = dq.IsingProblem.maxcut(n, edges)
    prob.psi0 = psi0
    sim = dq.IsingSimulator(prob, per_step=3, engine=engine)
    assert sim.info("engine") == engine
    e0 = float(-len(edges))
    np.random.seed(n)
    a = dq.EnergyTrainer(sim, n_basis=6, n_epoch=n_epoch, lr=5e-2, n_samples=K, device_resident=True, ground_energy=e0)
    a.train_energy()
    np.random.seed(n)
    b = dq.EnergyTrainer(sim, n_basis=6, n_epoch=n_epoch, lr=5e-2, n_samples=K, ground_energy=e0)
    b.train_energy()
    assert np.abs(np.array(a.losses_energy) - np.array(b.losses_energy)).max() < 1e-10
    assert np.abs(a.spectral_coeff.detach().numpy() - b.spectral_coeff.detach().numpy()).max() < 1e-10
    assert rel(a.final_state, b.final_state) < 1e-10
    assert np.abs(np.array(a.losses_energy)).max() > 1e-3


def test_shot_sampling_matches_the_reference_run(golden):
    """sampling_measure=True on the structured path: <Z_a Z_b> of every shifted ket from the device (dq_ising_grad_pairs),
    the reference's draws on the host -- against the reference's own run of the demo problem (exact step, tests/golden/
    demo_sampling_ref.npz: gradients plain and noisy, a direct measurement, an 8-epoch train_energy)."""
    gs = golden("demo_sampling_ref")
    g = golden(str(gs["base"]))
    prob = dq.IsingProblem.maxcut(4, [[0, 1], [0, 3], [1, 2], [2, 3]])
    assert [w for _, w in prob.measure_terms] == list(gs["weights"])
    sim = dq.IsingSimulator(prob, per_step=int(g["per_step"]), step="exact")
    zz = sim.pair_expectations(g["final"])[0]
    dense_zz = [float(np.real(np.vdot(g["final"], np.diag(R.z_diag(4, a) * R.z_diag(4, b)) @ g["final"]))) for a, b in prob.zz_pairs]
    assert np.abs(zz - dense_zz).max() < 1e-13
    np.random.seed(4100)
    assert abs(sim.stochastic_measure(g["final"]) - complex(gs["measure_final"]).real) < 1e-12
    for tag, noisy in (("plain", False), ("noisy", True)):
        for k in range(len(gs["s"])):
            np.random.seed(int(gs["seed0"]) + k)
            s_k = np.random.uniform() * prob.T
            grad = sim.grad_samples(g["coeff"], [s_k], is_noisy=noisy, sampling_measure=True)[0]
            assert rel(grad, gs["grads_" + tag][k]) < TOL
    np.random.seed(int(gs["train_seed"]))
    tr = dq.EnergyTrainer(sim, n_basis=6, n_epoch=int(gs["n_epoch"]), lr=2e-2, sampling_measure=True, ground_energy=-4.0)
    tr.train_energy()
    assert np.abs(np.array(tr.losses_energy) - gs["losses_energy"]).max() < 1e-9
    assert np.abs(tr.spectral_coeff.detach().numpy() - gs["final_coeff"]).max() < 1e-9


def test_pair_expectations_with_the_automatic_layout_and_split_step():
    """n=14 (fused engine would be the default; the pair call runs the per-term kernels): <ZZ> of the shifted kets vs the oracle."""
    n = 14
    edges = graph_for(n)
    prob = dq.IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    coeff = np.random.RandomState(3).normal(0, 1, [len(prob.terms), 6])
    sim = dq.IsingSimulator(prob, per_step=2)
    assert sim.info("engine") == 1 and sim.info("identity_layout") in (0, 1)
    s = 0.77
    zz = sim.shifted_pair_expectations(coeff, [s])[0]
    assert sim.info("engine") == 1                         # back on the fused engine afterwards
    en = sim.shifted_energies(coeff, [s])[0]
    # the observable is -1/2 sum_e (1 - Z_a Z_b): the pair expectations must reproduce the fused engine's energies
    assert np.abs((-0.5 * (len(edges) - zz.sum(axis=2))) - en).max() < 1e-10
    ns, dt, ts = R.step_grid(s, ref["T"], 2)
    phi = R.evolve_split_structured(ref, R.coef_table_plain(coeff, ref["omegas"], ref["T"], R.step_grid(0, s, 2)[2]), R.step_grid(0, s, 2)[1], ref["psi0"])
    for i in (0, len(edges) + 3):
        for k, sign in enumerate((+1, -1)):
            ket = R.evolve_split_structured(ref, R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts), dt, R.apply_shift_gate(ref, ref["terms"][i], phi, sign))
            want = [float(np.sum(np.abs(ket) ** 2 * R.z_diag(n, a) * R.z_diag(n, b))) for a, b in prob.zz_pairs]
            assert np.abs(zz[i, k] - want).max() < 1e-10


@pytest.mark.parametrize("n,step,func_type", [(4, "exact", 0), (6, "split", 1), (13, "split", 1)])
def test_iq_channel_pulse_model_on_the_structured_path(n, step, func_type):
    """IsingSimulator.trotter_cc: the native twin's pulse model f_u (diffqc.cc:95-135, library host routine pinned to the
    compiled reference) driving Pauli-term evolution, forward and backward spans -- against the oracle's restatement of
    diffqc.trotter (oracle/restate.py trotter_cc) on the dense twin (n <= 6) or its structured split step (n = 13)."""
    edges = graph_for(n)
    prob = dq.IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    rng = np.random.RandomState(40 + n)
    n_param, n_basis = 5, 5
    channels = [[[0.0, float(rng.uniform(0.5, 2.0)), float(rng.uniform(0, 2.5)), float(rng.randint(n_param))]
                 for _ in range(1 + i % 2)] for i in range(len(prob.terms))]
    vv = rng.normal(0, 1, [2, n_param, n_basis])
    sim = dq.IsingSimulator(prob, per_step=4, step=step)
    for (T0, T) in ((0.0, 1.3), (1.1, 0.35)):
        psi, en = sim.trotter_cc(channels, 2.0, func_type, vv, T0, T)
        ns, dt, ts = R.step_grid(T0, T, 4, use_abs=True)
        if n <= 6:
            H0, Hs, M = R.maxcut_dense(ref)
            want = R.trotter_cc(H0, Hs, channels, 2.0, func_type, ref["psi0"], T0, T, 4, vv, mode=step)
        else:
            u = np.array([[R.f_u_cc(h, t, vv, channels, 2.0, func_type) for h in range(len(prob.terms))] for t in ts])
            want = R.evolve_split_structured(ref, u, dt, ref["psi0"])
        assert rel(psi[0], want) < TOL
        assert abs(en[0] - R.energy_diag(ref["m_diag"], want)) < TOL * max(1.0, abs(en[0]))


@pytest.mark.parametrize("n", [21, 22])
def test_slice_pass_engine_for_batched_gradients_above_20_qubits(n):
    """n > 20: dq_ising_grad / dq_ising_evolve run the fused slice passes (one Gray-code phase pass + rotation passes of up to
    12 qubits per read/write, engine 2) instead of n + 1 per-term kernels -- against the per-term engine on the device and,
    for a subset of controls, against the plain-C port of the reference step and estimator."""
    from oracle import c_port as C
    edges = graph_for(n)
    prob = dq.IsingProblem.maxcut(n, edges)
    coeff = np.random.RandomState(n).normal(0, 1, [len(prob.terms), 6])
    sim = dq.IsingSimulator(prob, per_step=2)
    assert sim.info("engine") == 2
    s = 0.9
    grads, energies = sim.grad_samples(coeff, [s], return_energies=True)
    psi, en = sim.evolve(coeff, 0.0, 0.7)
    assert abs(np.linalg.norm(psi[0]) - 1) < 1e-12
    gen = dq.IsingSimulator(prob, per_step=2, engine=0)
    assert gen.info("engine") == 0
    psi0, en0 = gen.evolve(coeff, 0.0, 0.7)
    assert rel(psi[0], psi0[0]) < TOL and abs(en[0] - en0[0]) < TOL * abs(en0[0])
    if n == 21:
        e_gen = gen.shifted_energies(coeff, [s])
        assert rel(energies, e_gen) < TOL
    if C.available():
        C.use_host_cores()
        ref = R.maxcut_structured(n, edges)
        terms = [0, len(edges) - 1, len(edges), len(prob.terms) - 1]
        g_ref, e_ref, _ = C.grad_mc(C.CProblem(ref), coeff, s, 2, terms=terms, return_energies=True)
        assert rel(energies[0][terms], e_ref[terms]) < TOL
        assert rel(grads[0][terms], g_ref[terms]) < TOL
