"""krepp_b200 -- B200-native query path of krepp (`dist` / `place`) behind a C ABI.

The product is the shared library built from krepp_b200/csrc (hand-written sm_100a CUDA + C++ host code) and the
`krepp_b200` CLI.  This Python package is a thin ctypes mirror of the reference's IBatch interface
(src/query.hpp:46-97) used by tests and bench.py; it contains no compute and no CPU fallback.
"""
from .capi import (Index, IBatch, LibraryBuilder, Params, KreppError, Reader, build_library, library_path, load_library)  # noqa: F401
