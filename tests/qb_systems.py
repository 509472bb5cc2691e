"""Synthetic systems used by the parity tests: re-exported from quandary_b200.systems."""
from quandary_b200.systems import TWO_PI, basis_states, cfg_system, cfg_text, random_alpha, system_params  # noqa: F401
