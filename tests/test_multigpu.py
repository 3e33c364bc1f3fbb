"""Runs tests/multigpu_check.py under torchrun when the box has >= 2 GPUs (skipped on the 1-GPU test box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_frames_equal_single_gpu_frame():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


@pytest.mark.gpu
def test_single_process_multi_gpu_api(apertures):
    """lfb_create_multi / lfb_render_ghosts_multi: ONE host thread drives every GPU of the box (1 on the test box: the same code
    path with one device; all of them where there are more), tile-sparse into one page-locked buffer: every frame of a
    sequence equals lfb_render_ghosts bit for bit, for FP32 / STRICT and an uneven job split."""
    import numpy as np
    import torch
    from lens_flare_b200 import capi
    n_dev = torch.cuda.device_count()
    lens = capi.builtin_lens(3, 550.0)
    tex = apertures["pentbig500_14"]
    ref = capi.Engine(0)
    ref.set_lens(lens)
    ref.set_aperture(tex)
    W, H = 1000, 562
    buf = capi.PinnedArray((H, W, 3), np.float64)
    try:
        for n in sorted({1, min(2, n_dev), min(3, n_dev), n_dev}):
            m = capi.MultiEngine(list(range(n)))
            try:
                m.set_lens(lens)
                m.set_aperture(tex)
                buf.array[...] = 0.0
                first = True
                for prec in (capi.FP32, capi.STRICT):
                    p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=80, pair_set=capi.PAIRS_ALL, include_direct=1, precision=prec)
                    for lights in ([capi.make_light(0.45, 0.55, theta=capi.physical_theta(0.45, 0.55))],
                                   [capi.make_light(0.7, 0.3, theta=0.09, radiance=(2.0, 1.0, 0.5)), capi.make_light(0.3, 0.6, theta=0.05)], []):
                        tiles = m.render_ghosts(lights, p, buf.array, out_is_clear=first)
                        first = False
                        want = ref.render_ghosts(lights, p)
                        assert np.array_equal(buf.array, want), (n, prec, len(lights))
                        assert tiles >= 0 and (tiles > 0 or not lights)  # an empty frame still re-zeroes the previous frame's tiles
                st = m.stats()
                assert st["n_devices"] == n and st["call_ms"] > 0
                mine = np.full((H, W, 3), -1.0)  # pageable: full-frame fallback
                assert m.render_ghosts(lights, p, mine, out_is_clear=True) == -1
                assert np.array_equal(mine, ref.render_ghosts(lights, p))
            finally:
                m.close()
    finally:
        buf.free()
        ref.close()
