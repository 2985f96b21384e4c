"""cnf_ot_b200 -- B200-native kernels for the cnf_ot flow train step.

Host-side mirror of the reference's model / loss / solver interface
(`cnf_ot.models.flows.RQSFlow`, `cnf_ot.mfc.applications`, `cnf_ot.mfc.solvers`)
over the C ABI in `include/cnfot.h` (libcnfot.so, hand-written sm_100a CUDA).
There is no CPU fallback: any compute call without the built library or
without a CUDA device raises.
"""
from .layout import FlowShape, pack, unpack  # noqa: F401

__all__ = ["FlowShape", "pack", "unpack"]
