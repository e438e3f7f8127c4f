"""ndarray_b200 — B200-native data-parallel core of rust-ndarray behind ndarray's own API shape.

Everything computes in libndb200.so (hand-written sm_100a CUDA behind a C ABI, include/ndb200.h).
Importing this package does not need a GPU; creating a Context / Array does, and fails loudly
without one — there is no CPU fallback.
"""
from ._ffi import NdbError, ShapeError, KERNEL_NAMES, lib  # noqa: F401
from .array import (Array, Axis, Context, NewAxis, Op, Red, Slice, Zip, azip, default_context, indices, par_azip, s_)  # noqa: F401
from .linalg import dot, general_mat_mul, general_mat_vec_mul  # noqa: F401

__all__ = ["Array", "Axis", "Context", "NewAxis", "Op", "Red", "Slice", "Zip", "default_context", "s_", "azip", "par_azip", "indices", "dot",
           "general_mat_mul", "general_mat_vec_mul", "NdbError", "ShapeError", "KERNEL_NAMES", "lib"]
