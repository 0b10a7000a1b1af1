"""sapt_b200 — B200-native surrogate-assisted parallel tempering (hot path only).

Host-side mirror of the reference's class surface (Network, ptReplica helpers,
ParallelTempering) over libsapt.so (include/sapt.h). There is no CPU implementation in this
package: without a CUDA device and the built library every entry point raises.
"""
from . import _lib  # noqa: F401
from .ladder import assign_temperatures, default_beta_ladder  # noqa: F401

__all__ = ["Network", "ptReplica", "ParallelTempering", "Problem", "PTEngine", "EngineConfig",
           "assign_temperatures", "default_beta_ladder"]


def __getattr__(name):  # lazy: torch is only needed by the device-facing classes
    if name in ("Problem",):
        from .problem import Problem
        return Problem
    if name in ("PTEngine", "EngineConfig"):
        from . import engine
        return getattr(engine, name)
    if name in ("Network", "ptReplica"):
        from . import network
        return getattr(network, name)
    if name == "ParallelTempering":
        from .parallel_tempering import ParallelTempering
        return ParallelTempering
    raise AttributeError(name)
