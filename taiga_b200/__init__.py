"""taiga_b200: B200-native (sm_100a) prover hot path for anoma/taiga's Halo2/IPA proofs.

The CUDA library (libtaiga_b200.so, C ABI in include/taiga_b200.h) is the product; this package is the thin
Python host side used by the tests and the benchmark.  There is no CPU fallback."""
from . import lib  # noqa: F401
