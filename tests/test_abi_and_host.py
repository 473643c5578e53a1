"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/d4gs.h
declares (no compute calls without a GPU), ctypes prototypes agree with the header, host-side
logic (channel padding, camera layout, sharding)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "d4gs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|const char \*)\s*(d4_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_exports_every_declared_symbol():
    from deblur4dgs_b200 import _cabi
    fns = _header_functions()
    assert len(fns) >= 17
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in include/d4gs.h but not exported"
    assert set(fns) == set(_cabi.PROTOTYPES), set(fns) ^ set(_cabi.PROTOTYPES)
    for name, n in fns.items():
        assert len(_cabi.PROTOTYPES[name][1]) == n, f"{name}: ctypes prototype has wrong arity"
    l = _cabi.lib()
    assert l.d4_version() == 2
    assert [l.d4_tile_n_bits(x) for x in (1, 2, 576, 1024, 3600)] == [1, 2, 10, 11, 12]
    assert l.d4_sort_workspace_bytes(1 << 20) >= 256 * 512 * 4
    assert l.d4_scan_workspace_bytes(300000) >= 8 * 147


def test_argument_errors_are_reported_not_thrown():
    """Bad arguments come back as a status + message without touching the device."""
    from deblur4dgs_b200 import _cabi
    with pytest.raises(_cabi.D4Error, match="bad sizes"):
        _cabi.call("d4_project_fwd", None, 0, None, 0, None, None, 0, None, 0, 0, 5, 10, 10, 0.3, 0.01, 1e10, 0.0, 16, 1,
                   1, None, None, None, None, None, None, 0, None)
    with pytest.raises(_cabi.D4Error, match="tile_size 16"):
        _cabi.call("d4_blend_fwd", None, None, None, None, 0, None, None, 1, 1, 3, 32, 32, 8, 4, 4, None, None, 0, 0,
                   None, None, None, None, None, None)


def test_no_cpu_fallback():
    from deblur4dgs_b200 import _cabi
    from deblur4dgs_b200.motion import deform_subexposures
    from deblur4dgs_b200.rendering import rasterization
    z = torch.zeros
    with pytest.raises(_cabi.D4Error):
        rasterization(z(2, 3), z(2, 4), torch.ones(2, 3), torch.ones(2), torch.ones(2, 3), torch.eye(4)[None],
                      torch.eye(3)[None], 32, 32)
    with pytest.raises(_cabi.D4Error):
        deform_subexposures(z(2, 3), z(2, 4), z(2, 3), None, None, z(3, 8, 6), z(3, 8, 3), z(2))


def test_host_logic_padding_and_layout():
    from deblur4dgs_b200.rendering import SUPPORTED_D, _cam_layout, _pad_channels
    for d0, depth in [(3, False), (4, True), (16, True), (10, False), (11, True), (20, True)]:
        c, b, pad = _pad_channels(torch.zeros(5, d0), torch.zeros(1, d0), depth)
        assert c.shape[1] + int(depth) in SUPPORTED_D and c.shape[1] == d0 + pad and b.shape[1] == d0 + pad
    G = 7
    C, ms, qs, vs, ks = _cam_layout(torch.zeros(G, 3), torch.zeros(G, 4), torch.zeros(1, 4, 4), torch.zeros(1, 3, 3), G)
    assert (C, ms, qs, vs, ks) == (1, 0, 0, 16, 9)
    C, ms, qs, vs, ks = _cam_layout(torch.zeros(9, G, 3), torch.zeros(9, G, 4), torch.zeros(1, 4, 4), torch.zeros(1, 3, 3), G)
    assert (C, ms, qs, vs, ks) == (9, 21, 28, 0, 0)
    C, ms, qs, vs, ks = _cam_layout(torch.zeros(G, 3), torch.zeros(G, 4), torch.zeros(3, 4, 4), torch.zeros(3, 3, 3), G)
    assert (C, ms, qs, vs, ks) == (3, 0, 0, 16, 9)


def test_synthetic_configs_match_baseline():
    from deblur4dgs_b200.synthetic import CONFIGS, make_scene
    assert CONFIGS["c3"][:5] == (300_000, 1280, 720, 10, 9) and CONFIGS["c2"][:5] == (100_000, 512, 288, 6, 5)
    sc = make_scene(G=1000, width=64, height=48, K=5, N=4, seed=3)
    sc2 = make_scene(G=1000, width=64, height=48, K=5, N=4, seed=3)
    assert torch.equal(sc.fg_means, sc2.fg_means) and sc.num_fg == 300 and sc.colors_all(16).shape == (1000, 16)
    assert sc.times.shape == (4,) and sc.RTs.shape == (4, 3, 4)


def test_checkpoint_loader_reads_reference_layout():
    """Row f2: tests/golden/ckpt_small.pt was written with the reference's own GaussianParams / MotionBases
    state dicts in Trainer.save_checkpoint's layout (tests/golden/make_golden.py)."""
    from deblur4dgs_b200.checkpoint import load_checkpoint, scene_from_state_dict, scene_to_state_dict
    from deblur4dgs_b200.synthetic import make_scene
    path = os.path.join(ROOT, "tests", "golden", "ckpt_small.pt")
    sc, extras = load_checkpoint(path, 96, 64, frame=2, N=5)
    ref = make_scene(G=900, width=96, height=64, K=5, N=3, seed=31)
    for k in ["fg_means", "fg_quats", "fg_scales", "fg_colors", "fg_opacities", "motion_coefs", "bg_means", "bg_quats",
              "bg_scales", "bg_colors", "bg_opacities", "rots", "transls"]:
        assert torch.equal(getattr(sc, k), getattr(ref, k)), k
    assert extras["global_step"] == 1234 and extras["epoch"] == 7
    assert sc.w2c.shape == (1, 4, 4) and abs(float(sc.w2c[0, 0, 3]) - (-0.1 + 0.2 * 2 / 7)) < 1e-6
    assert sc.times.shape == (5,) and abs(float(sc.times[0]) - 1.5) < 1e-6 and sc.RTs.shape == (5, 3, 4)
    sd = scene_to_state_dict(sc)
    sc2 = scene_from_state_dict(sd, 96, 64, frame=0, N=1)
    assert torch.equal(sc2.fg_means, sc.fg_means) and sc2.num_bg == sc.num_bg
    with pytest.raises(KeyError):
        scene_from_state_dict({"fg.params.means": torch.zeros(1, 3)}, 8, 8)


def test_bench_reference_arm_json_contract():
    """bench.py --impl reference runs on the CPU (oracle port) and prints one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "1", "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
