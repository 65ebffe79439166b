"""Python harness around libmcr_b200.so — the B200 hot path of gpu-montecarlo-risk.

The product is the C-ABI library (include/mcr_b200.h) and the C plugin behind the reference's
src/model_interface.h; this package only binds them for tests, bench.py and smoke().
Importing `mcr_b200.api` loads the shared library and fails loudly if it was not built.
"""
from .synthetic import RiskModel, synthetic_model  # noqa: F401
