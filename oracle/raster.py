"""oracle/raster.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/raster_oracle.c``: the CPU restatement of
``gsplat.rendering.rasterization`` (gsplat==1.1.1, ``packed=False``) as the
reference calls it at ``flow3d/scene_model.py:360-373``.  The Python glue here
restates what gsplat's own Python wrapper does around its kernels (opacity /
colour broadcast over cameras, depth channel for "RGB+ED", expected-depth
normalisation).

PARITY UNPINNED (see the header of raster_oracle.c): gsplat is not available
in this environment and the reference holds no golden vectors for this path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "raster_oracle.c")
_LIB = os.path.join(_HERE, "liboracle_raster.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile raster_oracle.c with gcc (strict fp32, OpenMP)."""
    if force or (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_isect_count.restype = ctypes.c_int64
        _lib.orc_tile_n_bits.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def tile_n_bits(n_tiles: int) -> int:
    return lib().orc_tile_n_bits(int(n_tiles))


# --------------------------------------------------------------------------- #
# individual stages
# --------------------------------------------------------------------------- #
def _cam_stride(a, per_elem, G):
    """a is [G, per_elem] (shared, stride 0) or [C, G, per_elem]."""
    return 0 if a.ndim == 2 else G * per_elem


def project_fwd(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
                far_plane=1e10, radius_clip=0.0):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    viewmats, Ks = _f32(viewmats), _f32(Ks)
    C, G = viewmats.shape[0], scales.shape[0]
    radii = np.zeros((C, G), np.int32)
    means2d = np.zeros((C, G, 2), np.float32)
    depths = np.zeros((C, G), np.float32)
    conics = np.zeros((C, G, 3), np.float32)
    lib().orc_project_fwd(_p(means), ctypes.c_long(_cam_stride(means, 3, G)), _p(quats),
                          ctypes.c_long(_cam_stride(quats, 4, G)), _p(scales), _p(viewmats),
                          ctypes.c_long(16), _p(Ks), ctypes.c_long(9), C, G, int(width), int(height),
                          ctypes.c_float(eps2d), ctypes.c_float(near_plane), ctypes.c_float(far_plane),
                          ctypes.c_float(radius_clip), _p(radii), _p(means2d), _p(depths), _p(conics))
    return radii, means2d, depths, conics


def project_bwd(means, quats, scales, viewmats, Ks, width, height, radii, conics, v_means2d,
                v_depths, v_conics, eps2d=0.3, want_viewmats=True):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    viewmats, Ks = _f32(viewmats), _f32(Ks)
    C, G = viewmats.shape[0], scales.shape[0]
    v_means = np.zeros(means.shape, np.float64)
    v_quats = np.zeros(quats.shape, np.float64)
    v_scales = np.zeros(scales.shape, np.float64)
    v_viewmats = np.zeros((C, 4, 4), np.float64) if want_viewmats else None
    lib().orc_project_bwd(_p(means), ctypes.c_long(_cam_stride(means, 3, G)), _p(quats),
                          ctypes.c_long(_cam_stride(quats, 4, G)), _p(scales), _p(viewmats),
                          ctypes.c_long(16), _p(Ks), ctypes.c_long(9), C, G, int(width), int(height),
                          ctypes.c_float(eps2d), _p(np.ascontiguousarray(radii, dtype=np.int32)),
                          _p(_f32(conics)), _p(_f32(v_means2d)), _p(_f32(v_depths)), _p(_f32(v_conics)),
                          _p(v_means), _p(v_quats), _p(v_scales), _p(v_viewmats))
    return v_means, v_quats, v_scales, v_viewmats


def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True):
    means2d, depths = _f32(means2d), _f32(depths)
    radii = np.ascontiguousarray(radii, dtype=np.int32)
    C, G = radii.shape
    tiles_per_gauss = np.zeros((C, G), np.int32)
    n = lib().orc_isect_count(_p(means2d), _p(radii), C, G, int(tile_size), int(tile_width),
                              int(tile_height), _p(tiles_per_gauss))
    isect_ids = np.zeros((n,), np.int64)
    flatten_ids = np.zeros((n,), np.int32)
    lib().orc_isect_emit(_p(means2d), _p(radii), _p(depths), C, G, int(tile_size), int(tile_width),
                         int(tile_height), _p(isect_ids), _p(flatten_ids))
    if sort:
        tb = tile_n_bits(tile_width * tile_height)
        cb = int(math.floor(math.log2(C))) + 1
        lib().orc_sort_pairs(_p(isect_ids), _p(flatten_ids), ctypes.c_int64(n), 32 + tb + cb)
    return tiles_per_gauss, isect_ids, flatten_ids


def sort_pairs(keys, vals, end_bit=64):
    keys = np.ascontiguousarray(keys, dtype=np.int64).copy()
    vals = np.ascontiguousarray(vals, dtype=np.int32).copy()
    lib().orc_sort_pairs(_p(keys), _p(vals), ctypes.c_int64(keys.shape[0]), int(end_bit))
    return keys, vals


def isect_offset_encode(isect_ids, C, tile_width, tile_height):
    isect_ids = np.ascontiguousarray(isect_ids, dtype=np.int64)
    offsets = np.zeros((C, tile_height, tile_width), np.int32)
    lib().orc_tile_offsets(_p(isect_ids), ctypes.c_int64(isect_ids.shape[0]), C, int(tile_width),
                           int(tile_height), _p(offsets))
    return offsets


def blend_fwd(means2d, conics, opacities, colors, backgrounds, width, height, tile_size,
              isect_offsets, flatten_ids):
    """means2d [C,G,2], conics [C,G,3], opacities [C,G], colors [C,G,D], backgrounds [C,D]|None."""
    means2d, conics, opacities, colors = _f32(means2d), _f32(conics), _f32(opacities), _f32(colors)
    C, G, D = colors.shape
    backgrounds = _f32(backgrounds) if backgrounds is not None else None
    th, tw = isect_offsets.shape[1:]
    out = np.zeros((C, height, width, D), np.float32)
    alphas = np.zeros((C, height, width, 1), np.float32)
    last_ids = np.zeros((C, height, width), np.int32)
    edge = np.zeros((C, height, width), np.uint8)
    flatten_ids = np.ascontiguousarray(flatten_ids, dtype=np.int32)
    isect_offsets = np.ascontiguousarray(isect_offsets, dtype=np.int32)
    lib().orc_blend_fwd(_p(means2d), _p(conics), _p(opacities), _p(colors), _p(backgrounds), C, D,
                        int(width), int(height), int(tile_size), int(tw), int(th), _p(isect_offsets),
                        _p(flatten_ids), ctypes.c_int64(flatten_ids.shape[0]), _p(out), _p(alphas),
                        _p(last_ids), _p(edge))
    return out, alphas, last_ids, edge


def hit_masks(means2d, conics, opacities, width, height, tile_size, isect_offsets, flatten_ids):
    """Per-intersection 8-bit block masks of the forward and a knife-edge flag per intersection
    (orc_hit_masks).  means2d [C,G,2], conics [C,G,3], opacities [C,G]."""
    means2d, conics, opacities = _f32(means2d), _f32(conics), _f32(opacities)
    C = means2d.shape[0]
    th, tw = isect_offsets.shape[1:]
    flatten_ids = np.ascontiguousarray(flatten_ids, dtype=np.int32)
    isect_offsets = np.ascontiguousarray(isect_offsets, dtype=np.int32)
    n = flatten_ids.shape[0]
    masks = np.zeros((n,), np.uint8)
    edge = np.zeros((n,), np.uint8)
    lib().orc_hit_masks(_p(means2d), _p(conics), _p(opacities), C, int(width), int(height), int(tile_size), int(tw),
                        int(th), _p(isect_offsets), _p(flatten_ids), ctypes.c_int64(n), _p(masks), _p(edge))
    return masks, edge


def blend_bwd(means2d, conics, opacities, colors, backgrounds, width, height, tile_size,
              isect_offsets, flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas):
    means2d, conics, opacities, colors = _f32(means2d), _f32(conics), _f32(opacities), _f32(colors)
    C, G, D = colors.shape
    backgrounds = _f32(backgrounds) if backgrounds is not None else None
    th, tw = isect_offsets.shape[1:]
    v_means2d = np.zeros((C, G, 2), np.float64)
    v_conics = np.zeros((C, G, 3), np.float64)
    v_colors = np.zeros((C, G, D), np.float64)
    v_opacities = np.zeros((C, G), np.float64)
    flatten_ids = np.ascontiguousarray(flatten_ids, dtype=np.int32)
    isect_offsets = np.ascontiguousarray(isect_offsets, dtype=np.int32)
    lib().orc_blend_bwd(_p(means2d), _p(conics), _p(opacities), _p(colors), _p(backgrounds), C, G, D,
                        int(width), int(height), int(tile_size), int(tw), int(th), _p(isect_offsets),
                        _p(flatten_ids), ctypes.c_int64(flatten_ids.shape[0]), _p(_f32(render_alphas)),
                        _p(np.ascontiguousarray(last_ids, dtype=np.int32)), _p(_f32(v_render_colors)),
                        _p(_f32(v_render_alphas)), _p(v_means2d), _p(v_conics), _p(v_colors),
                        _p(v_opacities))
    v_backgrounds = None
    if backgrounds is not None:
        # gsplat computes this in Python: sum over pixels of v_colors * (1 - alpha)
        v_backgrounds = (np.asarray(v_render_colors, np.float64) *
                         (1.0 - np.asarray(render_alphas, np.float64))).sum(axis=(1, 2))
    return v_means2d, v_conics, v_colors, v_opacities, v_backgrounds


# --------------------------------------------------------------------------- #
# the operator: gsplat.rendering.rasterization (packed=False)
# --------------------------------------------------------------------------- #
def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16,
                  backgrounds=None, render_mode="RGB"):
    """Forward.  means [G,3] (or [C,G,3]: one set of centres per camera -- the
    sub-exposure batch), quats [G,4] wxyz (or [C,G,4]), scales [G,3], opacities
    [G], colors [G,D0], viewmats [C,4,4], Ks [C,3,3], backgrounds [C,D0]|None.
    Returns (render_colors [C,H,W,D], render_alphas [C,H,W,1], meta)."""
    assert render_mode in ("RGB", "RGB+ED", "RGB+D")
    viewmats, Ks = _f32(viewmats), _f32(Ks)
    C = viewmats.shape[0]
    scales = _f32(scales)
    G = scales.shape[0]
    radii, means2d, depths, conics = project_fwd(means, quats, scales, viewmats, Ks, width, height,
                                                 eps2d, near_plane, far_plane, radius_clip)
    opac = np.broadcast_to(_f32(opacities)[None, :], (C, G)).copy()
    col = _f32(colors)
    if col.ndim == 2:
        col = np.broadcast_to(col[None], (C,) + col.shape)
    bg = _f32(backgrounds) if backgrounds is not None else None
    if render_mode in ("RGB+ED", "RGB+D"):
        col = np.concatenate([col, depths[..., None]], axis=-1)
        if bg is not None:
            bg = np.concatenate([bg, np.zeros((C, 1), np.float32)], axis=-1)
    col = np.ascontiguousarray(col)
    tw = math.ceil(width / float(tile_size))
    th = math.ceil(height / float(tile_size))
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th)
    isect_offsets = isect_offset_encode(isect_ids, C, tw, th)
    acc, alphas, last_ids, edge = blend_fwd(means2d, conics, opac, col, bg, width, height, tile_size,
                                            isect_offsets, flatten_ids)
    render = acc
    if render_mode == "RGB+ED":
        render = acc.copy()
        render[..., -1:] = acc[..., -1:] / np.maximum(alphas, np.float32(1e-10))
    meta = dict(camera_ids=None, gaussian_ids=None, radii=radii, means2d=means2d, depths=depths,
                conics=conics, opacities=opac, tile_width=tw, tile_height=th,
                tiles_per_gauss=tiles_per_gauss, isect_ids=isect_ids, flatten_ids=flatten_ids,
                isect_offsets=isect_offsets, width=width, height=height, tile_size=tile_size,
                n_cameras=C,
                # oracle-only extras
                last_ids=last_ids, edge=edge, _acc=acc, _colors=col, _backgrounds=bg,
                _render_mode=render_mode,
                _inputs=dict(means=_f32(means), quats=_f32(quats), scales=scales, viewmats=viewmats, Ks=Ks,
                             eps2d=eps2d, colors_shape=np.asarray(colors).shape))
    return render, alphas, meta


def rasterization_backward(meta, render_alphas, v_render_colors, v_render_alphas, want_viewmats=True):
    """Backward of :func:`rasterization` for cotangents (v_render_colors
    [C,H,W,D], v_render_alphas [C,H,W,1]).  Returns a dict of float64 grads
    keyed like the operator's inputs, plus ``means2d`` (the screen-space grad
    the reference's densifier reads, trainer.py:975)."""
    mode = meta["_render_mode"]
    acc, col, bg = meta["_acc"], meta["_colors"], meta["_backgrounds"]
    C, H, W, D = acc.shape
    vrc = np.array(v_render_colors, np.float32, copy=True)
    vra = np.array(v_render_alphas, np.float32, copy=True).reshape(C, H, W, 1)
    if mode == "RGB+ED":
        # d = acc_d / max(alpha, 1e-10)
        a = render_alphas.reshape(C, H, W, 1)
        ac = np.maximum(a, np.float32(1e-10))
        vd = vrc[..., -1:].copy()
        vrc[..., -1:] = vd / ac
        vra = vra + np.where(a > 1e-10, -vd * acc[..., -1:] / (ac * ac), 0).astype(np.float32)
    width, height, ts = meta["width"], meta["height"], meta["tile_size"]
    v_means2d, v_conics, v_colors, v_opac, v_bg = blend_bwd(
        meta["means2d"], meta["conics"], meta["opacities"], col, bg, width, height, ts,
        meta["isect_offsets"], meta["flatten_ids"], render_alphas, meta["last_ids"], vrc, vra)
    inp = meta["_inputs"]
    v_depths = np.zeros(meta["depths"].shape, np.float64)
    if mode in ("RGB+ED", "RGB+D"):
        v_depths = v_colors[..., -1]
        v_colors = v_colors[..., :-1]
        if v_bg is not None:
            v_bg = v_bg[..., :-1]
    v_means, v_quats, v_scales, v_viewmats = project_bwd(
        inp["means"], inp["quats"], inp["scales"], inp["viewmats"], inp["Ks"], width, height,
        meta["radii"], meta["conics"], v_means2d, v_depths, v_conics, eps2d=inp["eps2d"],
        want_viewmats=want_viewmats)
    if len(inp["colors_shape"]) == 2:
        v_colors = v_colors.sum(axis=0)
    return dict(means=v_means, quats=v_quats, scales=v_scales, opacities=v_opac.sum(axis=0),
                colors=v_colors, backgrounds=v_bg, viewmats=v_viewmats, means2d=v_means2d,
                conics=v_conics, depths=v_depths)
