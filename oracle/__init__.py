"""CPU oracle for the diffquantum hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import anything from this package, and only as the
checker or as the timed CPU baseline.  `diffquantum_b200/` never imports it.

Contents
  standin/       minimal qutip / matplotlib / logger stand-ins so that the reference's own
                 `sim_plain.py` / `demo_maxcut.py` import UNMODIFIED (container only).
  ref_loader.py  imports the real reference from /root/reference (absent on the GPU box).
  restate.py     NumPy/SciPy restatement of the reference algorithms (travels to the GPU box):
                   exact step  = sim_plain.py:119-153 / diffqc.cc:173-205 (live code)
                   split step  = diffqc.cc:137-170 / sim_plain.py:139,142 (disabled product form)
                   pulses      = sim_plain.py:52-99, diffqc.cc:75-135
                   estimator   = sim_plain.py:156-231
  c/             plain-C (OpenMP) restatement of the split step and estimator, used as the
                 multi-core CPU baseline in bench.py and cross-checked against restate.py.
  make_golden.py generates tests/golden/*.npz by running the real reference (container only).

Parity status: PINNED against outputs of the reference's own code run in the build container
(the reference ships no tests or golden vectors of its own — SURVEY.md F5); fixtures and the
generating script are committed.  The split-step and large-n restatements have no live reference
code to run (the reference keeps the product form commented out and cannot build a dense matrix
beyond n~13); they are pinned to the reference's dense per-term product at small n instead.
"""
