"""opensbli_b200 -- B200-native execution back end for the OpenSBLI per-timestep solver hot path.

    from opensbli_b200 import B200        # drop-in for `OPSC(alg)` in an OpenSBLI app script
    from opensbli_b200 import Simulation  # run a plan on a GPU through the C ABI
"""
from .plan import validate, to_text, PlanError          # noqa: F401
from .runtime import Simulation, BackendError, load_library, measure_fp64_peak  # noqa: F401


def B200(algorithm, operation_count=False, OPS_diagnostics=1, **kwargs):
    """Back-end entry point with the call shape of the reference's `OPSC(algorithm, operation_count,
    OPS_diagnostics)` (opensbli/code_generation/opsc.py:253)."""
    from .backend import B200 as _B200
    return _B200(algorithm, operation_count=operation_count, OPS_diagnostics=OPS_diagnostics, **kwargs)
