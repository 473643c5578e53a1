"""ctypes binding of ``libd4gs.so`` (the C ABI declared in ``include/d4gs.h``).

The product path has NO fallback: if the library is missing or a symbol cannot
be resolved, importing / calling raises.  ``lib()`` loads the in-tree
``deblur4dgs_b200/libd4gs.so`` (build it with ``python -m deblur4dgs_b200.build``
or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libd4gs.so")
_lib = None

P = c_void_p  # device pointer
I, L, F = c_int, c_int64, c_float

# name -> (restype, argtypes); mirrors include/d4gs.h one to one
PROTOTYPES = {
    "d4_version": (c_int, []),
    "d4_last_error": (c_char_p, []),
    "d4_project_fwd": (c_int, [P, L, P, L, P, P, L, P, L, I, I, I, I, F, F, F, F, I, I, I, P, P, P, P, P, P, I, P]),
    "d4_project_bwd": (c_int, [P, L, P, L, P, P, L, P, L, I, I, I, I, F, P, P, P, P, P, P, P, P, P, P]),
    "d4_scan_workspace_bytes": (c_size_t, [L]),
    "d4_exclusive_scan_i32": (c_int, [P, L, P, P, P, c_size_t, P]),
    "d4_tile_n_bits": (c_int, [I]),
    "d4_isect_emit": (c_int, [P, P, P, P, I, I, I, I, I, P, P, P]),
    "d4_sort_workspace_bytes": (c_size_t, [L]),
    "d4_sort_pairs_u64": (c_int, [P, P, P, P, L, I, I, P, c_size_t, POINTER(c_int), P]),
    "d4_tile_sort_capacity": (c_int, []),
    "d4_tile_count": (c_int, [P, P, I, I, I, I, I, P, P]),
    "d4_bucket_emit": (c_int, [P, P, P, I, I, I, I, I, P, P, P, L, I, P]),
    "d4_tile_sort_capacity_max": (c_int, []),
    "d4_scan_counts": (c_int, [P, L, P, P, P, c_size_t, P]),
    "d4_tile_sort_pack_cap": (c_int, [P, P, P, L, I, I, I, I, P, P, P, P, P, P, I, I, P, P, P, I, P]),
    "d4_tile_sort": (c_int, [P, P, L, I, I, I, I, P, P, P]),
    "d4_tile_offsets": (c_int, [P, L, I, I, I, P, P]),
    "d4_blend_fwd": (c_int, [P, P, P, P, L, P, P, I, I, I, I, I, I, I, I, P, P, L, I, P, P, P, P, P, P]),
    "d4_blend_bwd": (c_int, [P, P, P, P, L, P, P, I, I, I, I, I, I, I, I, P, P, L, I, P, P, P, P, P, P, P, P, P, P, P, I, P]),
    "d4_isect_pack": (c_int, [P, P, P, P, I, I, I, I, I, P, P, L, P, P, P]),
    "d4_tile_sort_pack": (c_int, [P, P, L, I, I, I, I, P, P, P, P, P, P, I, I, P, P, P]),
    "d4_slab_hit_words": (c_size_t, [L, L]),
    "d4_blend_fwd_slab": (c_int, [P, P, P, P, L, P, I, I, I, I, I, I, I, I, I, I, P, P, P, P, P, P]),
    "d4_blend_fwd_slab_default_variant": (c_int, []),
    "d4_blend_fwd_slab_variant": (c_int, [P, P, P, P, L, P, I, I, I, I, I, I, I, I, I, I, P, P, P, P, P, I, P]),
    "d4_blend_bwd_slab": (c_int, [P, P, P, P, L, P, I, I, I, I, I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P]),
    "d4_blend_bwd_slab_default_variant": (c_int, []),
    "d4_blend_bwd_slab_variant": (c_int, [P, P, P, P, L, P, I, I, I, I, I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, I, P]),
    "d4_deform_fwd": (c_int, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, P, P, P]),
    "d4_deform_bwd": (c_int, [P, P, P, P, P, P, P, P, P, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P]),
    "d4_compute_transforms_fwd": (c_int, [P, P, P, P, I, I, I, I, P, P]),
    "d4_compute_transforms_bwd": (c_int, [P, P, P, P, I, I, I, I, P, P, P, P, P, P]),
    "d4_camera_interp_fwd": (c_int, [P, P, I, P, P]),
    "d4_camera_interp_bwd": (c_int, [P, P, I, P, P, P, P]),
    "d4_assemble_fwd": (c_int, [P, P, P, P, P, P, P, I, I, I, I, P, P, P, P]),
    "d4_assemble_bwd": (c_int, [P, P, P, P, P, P, I, I, I, I, P, P, P, P, P, P, P, P]),
    "d4_densify_stats": (c_int, [P, P, I, I, F, F, F, P, P, P, P]),
    "d4_band_partial": (c_int, [P, P, P, P, I, I, I, I, I, I, I, I, I, P, P, P]),
    "d4_band_winner": (c_int, [P, P, P, I, I, I, I, I, I, I, I, P, P, P]),
    "d4_band_finalize": (c_int, [P, P, P, I, I, I, I, I, I, I, P, P, P]),
    "d4_band_bwd": (c_int, [P, P, P, I, I, I, I, I, I, I, I, I, P, P, P, P, P]),
    "d4_correlation_fwd": (c_int, [P, P, I, I, I, I, P, P]),
    "d4_correlation_bwd": (c_int, [P, P, P, I, I, I, I, P, P, P]),
    "d4_combine_fwd": (c_int, [P, P, I, L, I, I, I, I, P, P, P, P, P]),
    "d4_combine_bwd": (c_int, [P, P, I, L, I, I, I, P, P, P, P, P]),
}


class D4Error(RuntimeError):
    pass


def lib():
    """Load libd4gs.so; raise loudly when it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise D4Error(
                f"{LIB_PATH} not found: build the CUDA library first "
                "(python -m deblur4dgs_b200.build). There is no CPU fallback for the render path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if handle.d4_version() != 2:
            raise D4Error("libd4gs.so ABI version mismatch")
        _lib = handle
    return _lib


# bench.py sets PROFILE to a dict to collect (start, end) CUDA events around every C-ABI call on
# the launching stream; LAUNCHES = kernels launched per call where that is not 1.
PROFILE = None
LAUNCHES = {"d4_deform_fwd": 2, "d4_deform_bwd": 2, "d4_exclusive_scan_i32": 2, "d4_scan_counts": 2}


FILLS = 0  # torch-side kernels (zero-fills, one tiny copy) issued by the op code: counted so that bench.py reports every launch


def count_fill(n: int = 1):
    global FILLS
    FILLS += n


def call(name: str, *args):
    """Invoke an int-returning entry point; translate a non-zero status into D4Error."""
    prof = PROFILE
    if prof is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib(), name)(*args)
    if prof is not None:
        e1.record()
        prof.setdefault(name, []).append((e0, e1))
    if rc != 0:
        msg = lib().d4_last_error()
        raise D4Error(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")


def check_tensors(*ts, what="deblur4dgs_b200 ops"):
    """Every op enqueues on the CURRENT device's current stream: tensors must be CUDA tensors of that device
    (there is no CPU fallback, and a launch on another device's stream would fault or race)."""
    import torch
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise D4Error(f"{what} need CUDA tensors: there is no CPU fallback")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise D4Error(f"{what}: tensor on cuda:{t.device.index} but the current device is cuda:{cur} "
                          "(wrap the call in torch.cuda.device(...))")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
