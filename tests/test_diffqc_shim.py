"""CPU-only checks of the `diffqc` drop-in module's pybind11-compatible conversions (diffqc.cc:210-228)."""
import numpy as np
import pytest


def test_module_surface():
    import diffqc
    for name in ("print_test", "complex_test", "test_eigen", "set_H", "trotter"):
        assert callable(getattr(diffqc, name))
    assert diffqc.__version__ == "dev"


def test_identity_round_trips(capsys):
    import diffqc
    diffqc.print_test()
    assert capsys.readouterr().out == "hello\n"
    out = diffqc.complex_test((1, 2.5, 3j, np.complex128(1 - 1j)))
    assert out == [1 + 0j, 2.5 + 0j, 3j, 1 - 1j] and all(type(v) is complex for v in out)
    assert diffqc.complex_test(np.array([1j, 2])) == [1j, 2 + 0j]
    assert diffqc.test_eigen([[1, 2], (3.5, 4)]) == [[1.0, 2.0], [3.5, 4.0]]
    assert diffqc.test_eigen([]) == []


def test_conversion_failures_are_type_errors():
    import diffqc
    with pytest.raises(TypeError):
        diffqc.complex_test("12")
    with pytest.raises(TypeError):
        diffqc.complex_test(3.0)
    with pytest.raises(TypeError):
        diffqc.test_eigen([["a"]])
    with pytest.raises(TypeError):
        diffqc.set_H([[1, 0], [0, 1]], [], [], "1.0", 0)
    with pytest.raises(TypeError):
        diffqc.set_H([[1, 0], [0, 1]], [], [], 1.0, 0.0)            # func_type is a C++ int
    with pytest.raises(ValueError):
        diffqc.set_H([[1, 0, 0], [0, 1, 0]], [], [], 1.0, 0)        # not square: UB in the reference
    with pytest.raises(TypeError):
        diffqc.trotter([1, 0], 0, 1, 2.5, [[[1.0]], [[1.0]]])       # per_step is a C++ int


def test_trotter_before_set_H_raises():
    from diffquantum_b200 import diffqc as m
    m._state["dim"] = None
    with pytest.raises(RuntimeError):
        m.trotter([1, 0], 0.0, 1.0, 10, [[[1.0]], [[1.0]]])
