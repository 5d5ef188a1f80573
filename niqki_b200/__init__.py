"""niqki_b200 — B200-native (sm_100a) implementation of the NIQKI hot path.

The product is ``lib/libniqki_b200.so`` (hand-written CUDA behind the C ABI of
``include/niqki_b200.h``) and the C++ host ``bin/niqki_b200`` that keeps the ``niqki`` CLI
contract.  This package is the Python mirror of the reference's ``class Index`` surface
(/root/reference/src/niqki_index.h:35-213) over that C ABI, used by the parity tests and the
benchmark driver.  There is no CPU fallback: importing works anywhere, computing needs the
compiled library and a CUDA device.
"""
from .capi import LIB_PATH, NiqkiError, Params, build_library, lib, library_available  # noqa: F401
from .index import Context, Index  # noqa: F401

__all__ = ["Context", "Index", "Params", "NiqkiError", "lib", "build_library", "library_available", "LIB_PATH"]
