"""Energy-minimisation loop of the reference (SimulatorPlain.train_energy, sim_plain.py:245-305) on top
of the device estimators — SURVEY 8(f) rank 1.  Host side keeps the reference's exact sequence of
NumPy-RNG draws (one normal block at :259, then one uniform() per epoch inside the estimator at :167)
and its torch.optim.Adam update (:266,291-292), so a seeded run follows the reference's trajectory."""
import numpy as np


class EnergyTrainer(object):
    """backend: DenseSimulator or IsingSimulator (needs .evolve energies and .grad_samples).
    n_samples > 1 averages that many stochastic samples per epoch (the reference uses 1)."""

    def __init__(self, backend, n_basis=6, n_epoch=202, lr=2e-2, n_samples=1, ground_energy=None, is_noisy=False,
                 device_resident=False, sampling_measure=False):
        self.backend = backend
        self.n_basis = n_basis
        self.n_epoch = n_epoch
        self.lr = lr
        self.n_samples = n_samples
        self.ground_energy = ground_energy
        self.is_noisy = is_noisy                       # SimulatorPlain(is_noisy=True): sim_plain.py:283-284, 207-208, 217-218
        # True: the whole loop runs on the device (backend.train_energy_device); the sample times are drawn up front from the
        # same np.random stream the reference consumes one per epoch (nothing else draws from it unless is_noisy)
        self.device_resident = device_resident
        # SimulatorPlain(sampling_measure=True): energies by shot sampling (stochastic_measure, sim_plain.py:101-117,278-279,
        # 202-203,212-213); the backend needs set_measurement(Pauli_M)
        self.sampling_measure = sampling_measure
        self.device_ms = None
        self.losses_energy = []
        self.final_state = None
        self.spectral_coeff = None

    def _n_terms(self):
        b = self.backend
        return b.n_H if hasattr(b, "n_H") else len(b.problem.terms)

    def _T(self):
        b = self.backend
        return b.T if hasattr(b, "T") else b.problem.T

    def _final(self, coeff):
        b = self.backend
        if hasattr(b, "problem"):                       # IsingSimulator
            psi, en = b.evolve(coeff, 0, b.problem.T, psi0=b.problem.psi0)
            return psi[0], float(en[0])
        psi = b.evolve(coeff, 0, b.T)
        return psi, float(b.energy(psi))

    def train_energy(self):
        import torch
        n_H, T = self._n_terms(), self._T()
        coeff = np.random.normal(0, 1e-3, [n_H, self.n_basis])             # sim_plain.py:259
        if self.device_resident:
            if self.is_noisy or self.sampling_measure:
                raise ValueError("is_noisy / sampling_measure interleave host draws with the sample times: use the host loop")
            s_all = np.array([[np.random.uniform() * T for _ in range(self.n_samples)] for _ in range(self.n_epoch)])   # :167
            if self.ground_energy is None and hasattr(self.backend, "problem"):
                raise ValueError("pass ground_energy (min of the observable diagonal) for structured problems")
            c, losses, final = self.backend.train_energy_device(coeff, s_all, lr=self.lr, e0=self.ground_energy)
            self.spectral_coeff = torch.tensor(c, requires_grad=True)
            self.losses_energy = list(losses)
            self.final_state = final
            self.device_ms = getattr(self.backend, "train_device_ms", None)
            if self.device_ms is None and hasattr(self.backend, "stat"):
                self.device_ms = self.backend.stat("kernel_ms")
            return self.spectral_coeff
        self.spectral_coeff = torch.tensor(coeff, requires_grad=True)
        optimizer = torch.optim.Adam([self.spectral_coeff], lr=self.lr)     # :266
        e0 = self.ground_energy
        if e0 is None:
            b = self.backend
            if hasattr(b, "problem"):
                raise ValueError("pass ground_energy (min of the observable diagonal) for structured problems")
            e0 = float(np.linalg.eigvalsh(b.M)[0])                          # M.eigenenergies()[0], :294
        self.losses_energy = []
        for epoch in range(1, self.n_epoch + 1):
            c = self.spectral_coeff.detach().numpy().copy()
            self.final_state, loss_energy = self._final(c)                  # :276-281
            if self.sampling_measure:
                loss_energy = float(np.real(self.backend.stochastic_measure(self.final_state)))   # :278-279
            if self.is_noisy:
                loss_energy += np.random.normal(scale=np.abs(loss_energy) / 5)   # :283-284, before the estimator's draws
            optimizer.zero_grad()
            s_list = [np.random.uniform() * T for _ in range(self.n_samples)]   # :167
            kw = {"sampling_measure": True} if self.sampling_measure else {}
            grads = self.backend.grad_samples(c, s_list, is_noisy=self.is_noisy, **kw)
            self.spectral_coeff.grad = torch.from_numpy(np.asarray(grads).mean(axis=0))
            optimizer.step()                                                # :291-292
            self.losses_energy.append(loss_energy - e0)
        return self.spectral_coeff

    def find_state(self):
        """SimulatorPlain.find_state (sim_plain.py:494-505): (argmax of the outcome probabilities, probabilities)."""
        a = np.asarray(self.final_state).reshape(-1)
        p = a.real ** 2 + a.imag ** 2
        return int(np.argmax(p)), p
