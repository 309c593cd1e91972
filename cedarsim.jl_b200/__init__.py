"""cedarsim.jl_b200 -- B200-native batched circuit-sweep engine behind CedarSim's sweep API.

Only the sweep hot path lives here: the sweep iterators and
`CircuitSweep` / `dc_` / `tran_` host API (mirror of reference src/sweeps.jl), the netlist
flattener and Verilog-A -> CUDA C generator that feed it, and the ctypes binding of the
C-ABI engine `csrc/libcedarb200.so` (include/cedarb200.h).  There is no CPU fallback:
solving without the CUDA library or without a GPU raises.
"""
from . import flat  # noqa: F401

__all__ = ["flat"]
