"""pyphysim_b200 — B200-native (sm_100a CUDA) implementation of pyphysim's per-realization link
hot path behind the reference's own Python API.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/b200phy.h) -> libb200phy.so
  _lib.py          ctypes binding of the C ABI (no CPU fallback: import errors are loud)
  links.py         fused link ops (throughput path): one call = a batch of realizations
  distributed.py   realization sharding over ranks + counter all-reduce (torch.distributed)
  modulators/ channels/ mimo/ util/ simulations/
                   host-side mirror of pyphysim.{modulators,channels,mimo,util,simulations}
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401

SEED_DEFAULT = 0x5EEDB200
