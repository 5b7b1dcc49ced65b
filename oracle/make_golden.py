"""Generate tests/golden/*.npz.  TEST INFRASTRUCTURE ONLY; runs in the build container only.

    python -m oracle.make_golden

Fixtures marked REFERENCE come from the reference's own, unmodified code
(/root/reference/sim_plain.py, demo_maxcut.py) imported behind oracle/standin.  Fixtures
marked RESTATEMENT come from oracle/restate.py for code the reference cannot run here
(diffqc.cc needs Eigen, which is absent; the per-term product form is commented out in the
reference); restate.py itself is pinned to the REFERENCE fixtures by tests/test_oracle.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader, restate as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

DEMO_GRAPH = [[0, 1], [0, 3], [1, 2], [2, 3]]       # demo_maxcut.py:11

# H2 / STO-3G / Jordan-Wigner, R = 0.7414 A (textbook 4-qubit coefficients; the reference lists
# the H2 VQE demo as TODO, README.md:27, and ships no Hamiltonian - SURVEY H10).
H2_TERMS = [
    ("IIII", -0.81261), ("ZIII", 0.171201), ("IZII", 0.171201), ("IIZI", -0.2227965),
    ("IIIZ", -0.2227965), ("ZZII", 0.16862325), ("ZIZI", 0.12054625), ("ZIIZ", 0.165868),
    ("IZZI", 0.165868), ("IZIZ", 0.12054625), ("IIZZ", 0.17434925), ("XXYY", -0.04532175),
    ("XYYX", 0.04532175), ("YXXY", 0.04532175), ("YYXX", -0.04532175),
]
_P = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
      "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.array([[1, 0], [0, -1]], dtype=complex)}


def pauli_string(s):
    return R.multi_kron(*[_P[c] for c in s]).astype(np.complex128)


def h2_problem():
    """Config 2 inputs (builder-supplied): M = H2 Hamiltonian, drift = always-on ZZ chain,
    controls = X_q and Y_q on each of 4 qubits, psi0 = Hartree-Fock |1100>."""
    M = sum(c * pauli_string(s) for s, c in H2_TERMS)
    H0 = 0.5 * sum(pauli_string(s) for s in ("ZZII", "IZZI", "IIZZ"))
    Hs = []
    for q in range(4):
        for p in "XY":
            Hs.append(pauli_string("".join(p if j == q else "I" for j in range(4))))
    psi0 = np.zeros(16, dtype=np.complex128)
    psi0[0b1100] = 1.0
    omegas = np.full(len(Hs), 2.0)
    return dict(M=M, H0=H0, Hs=np.array(Hs), psi0=psi0, omegas=omegas, T=1.5)


def make_sim(sp, n_basis, basis, T, omegas, n_Hs, coeff, per_step=10):
    sim = sp.SimulatorPlain(n_basis=n_basis, basis=basis, n_epoch=1, per_step=per_step)
    sim.T = T
    sim.omegas = list(omegas)
    sim.n_Hs = n_Hs
    sim.spectral_coeff = torch.tensor(coeff, requires_grad=True)
    return sim


def ref_H(sim, qp, H0, Hs, coeff):
    """H list exactly as train_energy builds it: sim_plain.py:272-274."""
    H = [qp.Qobj(H0)]
    for i in range(len(Hs)):
        H.append([qp.Qobj(Hs[i]), sim.generate_u(i, coeff.copy())])
    return H


def golden_dense_reference(sp, qp, name, H0, Hs, M, psi0, omegas, T, basis, n_basis, seed,
                           per_step=10, n_samples=6):
    """REFERENCE: SimulatorPlain.trotter and compute_energy_grad_MC on a dense problem."""
    rng = np.random.RandomState(seed)
    coeff = rng.normal(0, 1, [len(Hs), n_basis])
    sim = make_sim(sp, n_basis, basis, T, omegas, len(Hs), coeff, per_step)
    H = ref_H(sim, qp, H0, Hs, coeff)
    q0 = qp.Qobj(psi0)
    final = sim.trotter(H, q0, 0, T).full().reshape(-1)
    energy = qp.Qobj(M).matrix_element(qp.Qobj(final), qp.Qobj(final))
    n_steps, dt, ts = R.step_grid(0, T, per_step)
    u_tab = np.array([[H[i + 1][1](t, None) for i in range(len(Hs))] for t in ts])
    s_list, grads, phis = [], [], []
    for k in range(n_samples):
        np.random.seed(1000 + k)
        state = np.random.get_state()
        g = sim.compute_energy_grad_MC(qp.Qobj(M), H, q0).numpy().copy()
        np.random.set_state(state)
        s = np.random.uniform() * T           # sim_plain.py:167, same stream position
        s_list.append(s)
        grads.append(g)
        phis.append(sim.trotter(H, q0, 0, s).full().reshape(-1))
    np.savez(os.path.join(OUT, name + ".npz"), H0=H0, Hs=np.array(Hs), M=M, psi0=psi0,
             omegas=np.array(omegas, dtype=float), T=T, n_basis=n_basis, basis=basis,
             per_step=per_step, coeff=coeff, final=final, energy=energy, ts=ts, dt=dt, u_tab=u_tab,
             s=np.array(s_list), grads=np.array(grads), phis=np.array(phis),
             source="REFERENCE sim_plain.py:119-231 run unmodified behind oracle/standin")
    print(name, "energy", energy.real, "|grad0|", np.linalg.norm(grads[0]))


def golden_noisy_estimator(sp, qp, name, base, n_samples=3):
    """REFERENCE: compute_energy_grad_MC with is_noisy=True (sim_plain.py:207-208,217-218) on the inputs of the
    fixture `base`: every shifted energy gets np.random.normal(scale=|ps|/5) added, drawn from the same global
    stream right after the sample time."""
    g = np.load(os.path.join(OUT, base + ".npz"), allow_pickle=False)
    coeff = g["coeff"]
    sim = make_sim(sp, int(g["n_basis"]), str(g["basis"]), float(g["T"]), g["omegas"], len(g["Hs"]), coeff, int(g["per_step"]))
    sim.is_noisy = True
    H = ref_H(sim, qp, g["H0"], list(g["Hs"]), coeff)
    grads, s_list = [], []
    for k in range(n_samples):
        np.random.seed(2000 + k)
        state = np.random.get_state()
        grads.append(sim.compute_energy_grad_MC(qp.Qobj(g["M"]), H, qp.Qobj(g["psi0"])).numpy().copy())
        np.random.set_state(state)
        s_list.append(np.random.uniform() * float(g["T"]))
    np.savez(os.path.join(OUT, name + ".npz"), base=base, s=np.array(s_list), grads=np.array(grads), seed0=2000,
             source="REFERENCE sim_plain.py:156-231 with is_noisy=True, run unmodified behind oracle/standin")
    print(name, "|grad0|", np.linalg.norm(grads[0]))


def golden_demo_training(name):
    """REFERENCE: the shipped demo, demo_maxcut.py, run end to end with np.random.seed(0)."""
    ns = ref_loader.run_demo_maxcut(seed=0)
    sim = ns["sim"]
    np.savez(os.path.join(OUT, name + ".npz"),
             losses_energy=np.array(sim.losses_energy, dtype=float),
             final_coeff=sim.spectral_coeff.detach().numpy(),
             final_state=sim.final_state.full().reshape(-1),
             cut_state=int(ns["state"]), prob=np.array(ns["prob"]),
             H0=ns["H0"].full(), H_cost=ns["H_cost"].full(),
             Hs=np.array([h.full() for h in ns["Hs"]]),
             psi0=ns["superposition"].full().reshape(-1), T=sim.T, omegas=np.array(sim.omegas),
             stdout_tail=ns["__stdout__"].strip().splitlines()[-1],
             source="REFERENCE demo_maxcut.py:1-89, np.random.seed(0), 202 epochs")
    print(name, ns["__stdout__"].strip().splitlines()[-1], sim.losses_energy[-1])


def golden_split(name, n, edges, seed, per_step, n_samples=3):
    """RESTATEMENT: the disabled per-term product (diffqc.cc:155-164).  For n <= 8 the dense
    product of matrix exponentials is evaluated literally and stored next to the structured
    form so tests can pin one against the other."""
    prob = R.maxcut_structured(n, edges)
    rng = np.random.RandomState(seed)
    n_H = len(prob["terms"])
    coeff = rng.normal(0, 1, [n_H, 6])
    n_steps, dt, ts = R.step_grid(0, prob["T"], per_step)
    u = R.coef_table_plain(coeff, prob["omegas"], prob["T"], ts)
    final = R.evolve_split_structured(prob, u, dt, prob["psi0"])
    extra = {}
    if n <= 8:
        H0, Hs, M = R.maxcut_dense(prob)
        extra["final_dense_product"] = R.evolve_split_dense(H0, Hs, u, dt, prob["psi0"])
    s_list = np.random.RandomState(seed + 1).uniform(size=n_samples) * prob["T"]
    grads, ens = [], []
    for s in s_list:
        g, e = R.grad_mc_structured(prob, coeff, float(s), per_step, mode="split",
                                    return_energies=True)
        grads.append(g)
        ens.append(e)
    np.savez(os.path.join(OUT, name + ".npz"), n=n, edges=np.array(edges), coeff=coeff,
             per_step=per_step, T=prob["T"], omegas=prob["omegas"], u=u, dt=dt, final=final,
             energy=R.energy_diag(prob["m_diag"], final), s=s_list, grads=np.array(grads),
             energies=np.array(ens),
             source="RESTATEMENT oracle/restate.py evolve_split_structured (diffqc.cc:155-164)",
             **extra)
    print(name, "energy", R.energy_diag(prob["m_diag"], final))


def golden_split_reference(sp_split, qp, name, n, edges, seed, per_step, n_samples=3):
    """REFERENCE (its own disabled lines, enabled in memory): SimulatorPlain.trotter with sim_plain.py:139,142 active
    and :140,143,145-146,149 off (oracle/ref_loader.py: load_sim_plain_split), on the demo's dense construction of an
    n-qubit MaxCut problem; step grid, pulse closures and compute_energy_grad_MC are the reference's unchanged code."""
    prob = R.maxcut_structured(n, edges)
    H0, Hs, M = R.maxcut_dense(prob)
    rng = np.random.RandomState(seed)
    coeff = rng.normal(0, 1, [len(Hs), 6])
    sim = make_sim(sp_split, 6, "BSpline", prob["T"], prob["omegas"], len(Hs), coeff, per_step)
    H = ref_H(sim, qp, H0, Hs, coeff)
    q0 = qp.Qobj(prob["psi0"])
    final = sim.trotter(H, q0, 0, prob["T"]).full().reshape(-1)
    partial = sim.trotter(H, q0, 0.31, 1.17).full().reshape(-1)
    energy = qp.Qobj(M).matrix_element(qp.Qobj(final), qp.Qobj(final))
    s_list, grads, phis = [], [], []
    for k in range(n_samples):
        np.random.seed(3000 + k)
        state = np.random.get_state()
        g = sim.compute_energy_grad_MC(qp.Qobj(M), H, q0).numpy().copy()
        np.random.set_state(state)
        s = np.random.uniform() * prob["T"]
        s_list.append(s)
        grads.append(g)
        phis.append(sim.trotter(H, q0, 0, s).full().reshape(-1))
    np.savez(os.path.join(OUT, name + ".npz"), n=n, edges=np.array(edges), coeff=coeff, per_step=per_step,
             T=prob["T"], omegas=prob["omegas"], final=final, partial=partial, partial_span=np.array([0.31, 1.17]),
             energy=energy, s=np.array(s_list), grads=np.array(grads), phis=np.array(phis),
             source="REFERENCE sim_plain.py:119-231 with its own commented product-form lines :139,:142 enabled and "
                    ":140,:143,:145-146,:149 disabled in memory (oracle/ref_loader.py load_sim_plain_split), run behind "
                    "oracle/standin")
    print(name, "energy", energy.real, "|grad0|", np.linalg.norm(grads[0]))


def golden_basis_evolution(sp, qp, name, basis, n_basis, seed, per_step=10):
    """REFERENCE: generate_u / trotter with the 'poly' and 'Fourier' bases (sim_plain.py:84-92) on the demo problem, and what
    compute_energy_grad_MC does with them (it stops at :178: coeff_A is only assigned for 'Legendre' and 'BSpline')."""
    demo = R.maxcut_structured(4, DEMO_GRAPH)
    H0, Hs, M = R.maxcut_dense(demo)
    rng = np.random.RandomState(seed)
    coeff = rng.normal(0, 1, [len(Hs), n_basis])
    sim = make_sim(sp, n_basis, basis, demo["T"], demo["omegas"], len(Hs), coeff, per_step)
    H = ref_H(sim, qp, H0, Hs, coeff)
    q0 = qp.Qobj(demo["psi0"])
    final = sim.trotter(H, q0, 0, demo["T"]).full().reshape(-1)
    part = sim.trotter(H, q0, 0.4, 1.3).full().reshape(-1)
    n_steps, dt, ts = R.step_grid(0, demo["T"], per_step)
    u_tab = np.array([[H[i + 1][1](t, None) for i in range(len(Hs))] for t in ts])
    try:
        np.random.seed(1)
        sim.compute_energy_grad_MC(qp.Qobj(M), H, q0)
        err = ""
    except Exception as e:                                   # UnboundLocalError at sim_plain.py:178
        err = type(e).__name__
    np.savez(os.path.join(OUT, name + ".npz"), H0=H0, Hs=np.array(Hs), M=M, psi0=demo["psi0"], omegas=demo["omegas"], T=demo["T"],
             n_basis=n_basis, basis=basis, per_step=per_step, coeff=coeff, final=final, partial=part,
             partial_span=np.array([0.4, 1.3]), ts=ts, dt=dt, u_tab=u_tab, estimator_error=err,
             source="REFERENCE sim_plain.py:73-99,119-153 run unmodified behind oracle/standin")
    print(name, "estimator:", err or "ran", "|final|", np.linalg.norm(final))


def golden_comparators(sp, qp, name):
    """REFERENCE: the comparison methods (sim_plain.py:308-475) behind the stand-in's mesolve (DOP853, rtol 1e-12).
    (i) compute_energy_grad_FD and a noisy one on the demo problem; (ii) train_energy_FD, 2 epochs; (iii) train_fidelity,
    4 epochs over two (initial, target) pairs of a two-qubit problem, plain and noisy."""
    demo = R.maxcut_structured(4, DEMO_GRAPH)
    H0, Hs, M = R.maxcut_dense(demo)
    rng = np.random.RandomState(61)
    coeff = rng.normal(0, 1, [len(Hs), 6])
    out = {}
    for tag, noisy in (("plain", False), ("noisy", True)):
        sim = make_sim(sp, 6, "BSpline", demo["T"], demo["omegas"], len(Hs), coeff)
        sim.is_noisy = noisy
        H = ref_H(sim, qp, H0, Hs, coeff)
        np.random.seed(5000)
        out["fd_grad_" + tag] = sim.compute_energy_grad_FD(qp.Qobj(M), H, qp.Qobj(demo["psi0"])).numpy().copy()
    sim = sp.SimulatorPlain(lr=2e-2, n_basis=6, n_epoch=2)
    sim.T = demo["T"]
    sim.omegas = list(demo["omegas"])
    np.random.seed(5001)
    sim.train_energy_FD(qp.Qobj(M), qp.Qobj(H0), [qp.Qobj(h) for h in Hs], qp.Qobj(demo["psi0"]))
    out["fd_train_losses"] = np.array(sim.losses_energy, dtype=float)
    out["fd_train_coeff"] = sim.spectral_coeff.detach().numpy()
    out["fd_train_final"] = sim.final_state.full().reshape(-1)
    # state transfer on two qubits: drift ZZ, controls X and Y on each qubit; |00> -> Bell-like, |01> -> its partner
    F_H0 = 0.4 * pauli_string("ZZ")
    F_Hs = [pauli_string(s) for s in ("XI", "IX", "YI", "IY")]
    inits = [np.array([1, 0, 0, 0], dtype=complex), np.array([0, 1, 0, 0], dtype=complex)]
    targets = [np.array([1, 0, 0, 1j], dtype=complex) / np.sqrt(2), np.array([0, 1, 1j, 0], dtype=complex) / np.sqrt(2)]
    for tag, noisy in (("plain", False), ("noisy", True)):
        sim = sp.SimulatorPlain(lr=5e-2, n_basis=5, n_epoch=4, is_noisy=noisy)
        sim.T = 1.5
        sim.omegas = [2.0, 2.0, 1.5, 1.5]
        import contextlib, io
        np.random.seed(5002)
        with contextlib.redirect_stdout(io.StringIO()):
            sim.train_fidelity(qp.Qobj(F_H0), [qp.Qobj(h) for h in F_Hs], [qp.Qobj(p) for p in inits], [qp.Qobj(p) for p in targets])
        out["fid_losses_" + tag] = np.array(sim.losses_energy, dtype=float)
        out["fid_coeff_" + tag] = sim.spectral_coeff.detach().numpy()
    np.savez(os.path.join(OUT, name + ".npz"), H0=H0, Hs=np.array(Hs), M=M, psi0=demo["psi0"], omegas=demo["omegas"], T=demo["T"],
             coeff=coeff, delta=1e-3, F_H0=F_H0, F_Hs=np.array(F_Hs), F_inits=np.array(inits), F_targets=np.array(targets),
             F_T=1.5, F_omegas=np.array([2.0, 2.0, 1.5, 1.5]),
             source="REFERENCE sim_plain.py:308-475 run unmodified behind oracle/standin (mesolve = SciPy DOP853 at rtol 1e-12)", **out)
    print(name, "|fd grad|", np.linalg.norm(out["fd_grad_plain"]), "fid losses", out["fid_losses_plain"])


def demo_pauli_m(qp, n_qubit=4, graph=DEMO_GRAPH):
    """sim.Pauli_M exactly as demo_maxcut.py:47-65 builds it (Z-strings of the edges with weight 0.5, the identity with
    weight -len(graph)/2, each with the stand-in's eigenstates())."""
    I = np.array([[1, 0], [0, 1]])
    Z = np.array([[1, 0], [0, -1]])
    II = I
    for i in range(n_qubit - 1):
        II = np.kron(II, I)
    pauli_m = []
    for e in graph:
        curr = Z if 0 in e else I
        for i in range(1, n_qubit):
            curr = np.kron(curr, Z if i in e else I)
        pauli_m.append([curr, 0.5])
    pauli_m.append([II, -0.5 * len(graph)])
    for i in range(len(pauli_m)):
        pauli_m[i].append(qp.Qobj(pauli_m[i][0]).eigenstates())
    return pauli_m


def golden_sampling(sp, qp, name, base, n_samples=3, n_epoch=8):
    """REFERENCE: shot sampling (sampling_measure=True -> stochastic_measure, sim_plain.py:101-117,202-203,212-213,278-279).
    (i) compute_energy_grad_MC on the inputs of fixture `base`, plain and with is_noisy; (ii) train_energy for n_epoch epochs
    on the demo problem.  Pauli_M as demo_maxcut.py:47-65 builds it."""
    g = np.load(os.path.join(OUT, base + ".npz"), allow_pickle=False)
    coeff = g["coeff"]
    pauli_m = demo_pauli_m(qp)
    out = {}
    for tag, noisy in (("plain", False), ("noisy", True)):
        sim = make_sim(sp, int(g["n_basis"]), str(g["basis"]), float(g["T"]), g["omegas"], len(g["Hs"]), coeff, int(g["per_step"]))
        sim.sampling_measure = True
        sim.is_noisy = noisy
        sim.Pauli_M = pauli_m
        H = ref_H(sim, qp, g["H0"], list(g["Hs"]), coeff)
        grads, s_list = [], []
        for k in range(n_samples):
            np.random.seed(4000 + k)
            state = np.random.get_state()
            grads.append(sim.compute_energy_grad_MC(qp.Qobj(g["M"]), H, qp.Qobj(g["psi0"])).numpy().copy())
            np.random.set_state(state)
            s_list.append(np.random.uniform() * float(g["T"]))
        out["grads_" + tag] = np.array(grads)
        out["s"] = np.array(s_list)
    # one direct stochastic_measure call on the final state of the base fixture
    np.random.seed(4100)
    out["measure_final"] = sim.stochastic_measure(qp.Qobj(g["final"]))
    # train_energy with shot sampling
    sim = sp.SimulatorPlain(lr=2e-2, n_basis=6, n_epoch=n_epoch, sampling_measure=True)
    sim.T = float(g["T"])
    sim.omegas = list(g["omegas"])
    sim.Pauli_M = pauli_m
    np.random.seed(7)
    sim.train_energy(qp.Qobj(g["M"]), qp.Qobj(g["H0"]), [qp.Qobj(h) for h in g["Hs"]], qp.Qobj(g["psi0"]))
    np.savez(os.path.join(OUT, name + ".npz"), base=base, seed0=4000, train_seed=7, n_epoch=n_epoch,
             weights=np.array([p[1] for p in pauli_m]), evals=np.array([p[2][0] for p in pauli_m]),
             estates=np.array([[e.full().reshape(-1) for e in p[2][1]] for p in pauli_m]),
             losses_energy=np.array(sim.losses_energy, dtype=float), final_coeff=sim.spectral_coeff.detach().numpy(),
             source="REFERENCE sim_plain.py:101-117,156-231,245-305 with sampling_measure=True, run unmodified behind oracle/standin",
             **out)
    print(name, "|grad0|", np.linalg.norm(out["grads_plain"][0]), "loss[-1]", sim.losses_energy[-1])


def fu_cases():
    """Inputs that walk every branch of diffqc.cc:75-135: both bases, the +-32 cutoff of my_expit, |N| < 1e-6,
    idx rounding (C round: half away from zero), times outside [0, duration] (bump support / Legendre beyond +-1)."""
    rng = np.random.RandomState(31)
    cases = []
    for func_type in (0, 1):
        for n_basis in (3, 5, 6):
            n_param = 4
            vv = rng.normal(0, 1, [2, n_param, n_basis])
            vv[:, 1, :] *= 1e-8                      # |N| < 1e-6  (diffqc.cc:128-129)
            vv[:, 2, :] *= 60.0                      # N > 32      (diffqc.cc:76-77)
            channels = [[[0.0, 1.5, 0.7, 0.0]],
                        [[0.0, 2.0, 0.0, 1.0], [0.0, 0.5, 1.3, 2.0]],
                        [[0.0, 1.0, 2.1, 2.4999], [0.0, -0.8, -3.0, 2.5], [0.0, 0.3, 0.2, 0.5]],
                        []]
            ts = np.concatenate([np.linspace(-0.3, 2.4, 28), [0.0, 2.0, 1.0, 0.5, 1.5]])
            cases.append(dict(func_type=func_type, n_basis=n_basis, vv=vv, channels=channels, duration=2.0, ts=ts))
    return cases


def golden_fu_reference(name):
    """REFERENCE (compiled): f_u of diffqc.cc:95-135 built from the reference source by oracle/ref_cc/Makefile."""
    from oracle import ref_cc
    if not ref_cc.build():
        raise RuntimeError("oracle/_ref/libfu.so could not be built")
    out = {}
    lib = ref_cc.load()
    for i, c in enumerate(fu_cases()):
        out["u_%d" % i] = ref_cc.f_u_table(c["channels"], c["duration"], c["func_type"], c["vv"], c["ts"])
        out["vv_%d" % i] = c["vv"]
        out["ts_%d" % i] = c["ts"]
        out["meta_%d" % i] = np.array([c["func_type"], c["n_basis"]])
    xs = np.array([-40.0, -32.0, -31.999, -1.0, 0.0, 1e-7, 3.0, 31.999, 32.0, 32.001, 50.0])
    out["expit_x"] = xs
    out["expit_y"] = np.array([lib.ref_my_expit(float(x)) for x in xs])
    np.savez(os.path.join(OUT, name + ".npz"), n_cases=len(fu_cases()),
             chan_flat=np.array([ch for c in fu_cases()[0]["channels"] for ch in c]),
             chan_counts=np.array([len(c) for c in fu_cases()[0]["channels"]]), duration=2.0,
             source="REFERENCE diffqc.cc:75-135 compiled from the reference source (oracle/ref_cc/Makefile -> oracle/_ref/libfu.so)",
             **out)
    print(name, "cases", len(fu_cases()))


def golden_diffqc_cc(name):
    """RESTATEMENT of diffqc.set_H/trotter (diffqc.cc:43-135,173-205); Eigen is absent so the
    C++ cannot be built here.  Two-level + coupled-qubit style inputs, both basis types."""
    rng = np.random.RandomState(7)
    X, Y, Z, I = _P["X"], _P["Y"], _P["Z"], _P["I"]
    H0 = 0.3 * np.kron(Z, Z) + 0.2 * np.kron(Z, I)
    Hs = [np.kron(X, I), np.kron(I, X), np.kron(Y, I) + 0.1 * np.kron(I, Y)]
    # channels[h][c] = [unused, omega, w, idx]   (diffqc.cc:108-111)
    channels = [[[0.0, 1.5, 0.7, 0.0]],
                [[0.0, 2.0, 0.0, 1.0], [0.0, 0.5, 1.3, 2.0]],
                [[0.0, 1.0, 2.1, 2.0]]]
    n_param, n_basis = 3, 5
    vv = rng.normal(0, 1, [2, n_param, n_basis])
    vv[:, 1, :] *= 1e-8          # exercises the |N| < 1e-6 branch (diffqc.cc:128-129)
    psi0 = rng.normal(size=4) + 1j * rng.normal(size=4)
    psi0 /= np.linalg.norm(psi0)
    out = {}
    for func_type in (0, 1):
        for (T0, T) in ((0.0, 1.7), (1.2, 0.4)):     # second span is negative: abs() matters
            key = "f%d_%s" % (func_type, "fwd" if T > T0 else "bwd")
            out[key] = R.trotter_cc(H0, Hs, channels, 2.0, func_type, psi0, T0, T, 12, vv)
            n_steps, dt, ts = R.step_grid(T0, T, 12, use_abs=True)
            out[key + "_u"] = np.array([[R.f_u_cc(h, t, vv, channels, 2.0, func_type)
                                         for h in range(3)] for t in ts])
    chan_flat = np.array([c for h in channels for c in h])
    np.savez(os.path.join(OUT, name + ".npz"), H0=H0, Hs=np.array(Hs), chan_flat=chan_flat,
             chan_counts=np.array([len(h) for h in channels]), duration=2.0, vv=vv, psi0=psi0,
             per_step=12, spans=np.array([[0.0, 1.7], [1.2, 0.4]]),
             source="RESTATEMENT oracle/restate.py trotter_cc (diffqc.cc:43-135,173-205)", **out)
    print(name, "done")


def main():
    os.makedirs(OUT, exist_ok=True)
    sp = ref_loader.load_sim_plain()
    qp = ref_loader.load_qutip_standin()
    if sys.argv[1:] == ["noisy"]:                       # only the fixture added last; the others stay as committed
        golden_noisy_estimator(sp, qp, "demo_noisy_ref", "demo_bspline_ref")
        return
    if sys.argv[1:] == ["bases"]:
        golden_basis_evolution(sp, qp, "demo_poly_ref", "poly", 4, seed=51)
        golden_basis_evolution(sp, qp, "demo_fourier_ref", "Fourier", 6, seed=52)
        return
    if sys.argv[1:] == ["comparators"]:
        golden_comparators(sp, qp, "comparators_ref")
        return
    if sys.argv[1:] == ["sampling"]:
        golden_sampling(sp, qp, "demo_sampling_ref", "demo_bspline_ref")
        return
    if sys.argv[1:] == ["fu_ref"]:
        golden_fu_reference("fu_cc_ref")
        return
    if sys.argv[1:] == ["split_ref"]:
        sps = ref_loader.load_sim_plain_split()
        golden_split_reference(sps, qp, "split_ref_n4", 4, DEMO_GRAPH, seed=21, per_step=10)
        golden_split_reference(sps, qp, "split_ref_n6", 6, R.random_regular_edges(6, seed=1), seed=22, per_step=10)
        return
    demo = R.maxcut_structured(4, DEMO_GRAPH)
    H0, Hs, M = R.maxcut_dense(demo)
    golden_dense_reference(sp, qp, "demo_bspline_ref", H0, Hs, M, demo["psi0"], demo["omegas"],
                           demo["T"], "BSpline", 6, seed=1234)
    golden_dense_reference(sp, qp, "demo_legendre_ref", H0, Hs, M, demo["psi0"], demo["omegas"],
                           demo["T"], "Legendre", 5, seed=4321, n_samples=3)
    h2 = h2_problem()
    golden_dense_reference(sp, qp, "h2_vqe_ref", h2["H0"], list(h2["Hs"]), h2["M"], h2["psi0"],
                           h2["omegas"], h2["T"], "BSpline", 6, seed=99, n_samples=4)
    golden_demo_training("demo_training_ref")
    golden_split("split_n4_demo", 4, DEMO_GRAPH, seed=11, per_step=10)
    golden_split("split_n6", 6, R.random_regular_edges(6, seed=1), seed=12, per_step=10)
    golden_split("split_n8", 8, R.random_regular_edges(8, seed=2), seed=13, per_step=7)
    golden_split("split_n12", 12, R.random_regular_edges(12, seed=3), seed=14, per_step=3,
                 n_samples=1)
    golden_diffqc_cc("diffqc_cc_restated")
    golden_noisy_estimator(sp, qp, "demo_noisy_ref", "demo_bspline_ref")
    sps = ref_loader.load_sim_plain_split()
    golden_split_reference(sps, qp, "split_ref_n4", 4, DEMO_GRAPH, seed=21, per_step=10)
    golden_split_reference(sps, qp, "split_ref_n6", 6, R.random_regular_edges(6, seed=1), seed=22, per_step=10)
    golden_fu_reference("fu_cc_ref")
    golden_sampling(sp, qp, "demo_sampling_ref", "demo_bspline_ref")
    golden_basis_evolution(sp, qp, "demo_poly_ref", "poly", 4, seed=51)
    golden_basis_evolution(sp, qp, "demo_fourier_ref", "Fourier", 6, seed=52)
    golden_comparators(sp, qp, "comparators_ref")


if __name__ == "__main__":
    main()
