"""oblas_b200 -- B200 (sm_100a) implementation of the oblas hot path.

The product is the C-ABI shared library oblas_b200/liboblas_b200.so (sources in
oblas_b200/csrc, headers in include/).  The Python in this package is plumbing
for tests, the benchmark and multi-GPU sharding: ctypes bindings (capi), helpers
that pass torch device buffers / streams to the library (device) and the
partitioning rules for 1/2/4/8 GPUs (shard).
"""
from . import capi  # noqa: F401

__all__ = ["capi", "device", "shard"]
