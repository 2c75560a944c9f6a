"""pyarrowspace_b200 -- B200-native build-and-search hot path of pyarrowspace.

``csrc/``          hand-written sm_100a CUDA + the C ABI (include/arrowspace_b200.h)
``_lib.py``        ctypes binding of libarrowspace_b200.so
``api.py``         host-side mirror of the reference's Python surface (src/lib.rs)
``distributed.py`` row-sharded multi-GPU orchestration over torch.distributed
"""
from .api import (ArrowSpace, ArrowSpaceBuilder, GraphLaplacian, PanicException, launch_count,  # noqa: F401
                  set_debug, shard_rows, stat)

__all__ = ["ArrowSpaceBuilder", "ArrowSpace", "GraphLaplacian", "set_debug", "PanicException", "shard_rows"]
