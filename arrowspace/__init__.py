"""Drop-in import name of the reference module (Cargo.toml:7, src/lib.rs:379-386):

    from arrowspace import ArrowSpaceBuilder, ArrowSpace, GraphLaplacian, set_debug

backed by the B200-native implementation in ``pyarrowspace_b200`` (C ABI + sm_100a CUDA).
"""
from pyarrowspace_b200.api import (ArrowSpace, ArrowSpaceBuilder, GraphLaplacian, PanicException,  # noqa: F401
                                   set_debug)

__all__ = ["ArrowSpaceBuilder", "ArrowSpace", "GraphLaplacian", "set_debug"]
