"""Drop-in for the reference's pybind11 extension: `import diffqc` (diffqc.cc:210-228) resolves to
the B200-native implementation in diffquantum_b200/diffqc.py when the repo root is on sys.path."""
from diffquantum_b200.diffqc import (__version__, complex_test, print_test, set_device, set_H, test_eigen,  # noqa: F401
                                     trotter)
