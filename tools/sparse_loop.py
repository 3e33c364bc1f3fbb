#!/usr/bin/env python
"""A short loop of blocking lfb_render_ghosts_sparse frames (cfg2, the sun alternating between two positions) -- the workload
ncu is pointed at to look at the tile kernel (`-k regex:tiles_kernel`).  python tools/sparse_loop.py [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from lens_flare_b200 import capi
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    tex = bench.load_aperture()
    W, H = 1920, 1080
    eng = capi.Engine(0)
    eng.set_lens(capi.builtin_lens(3, 550.0))
    eng.set_aperture(tex)
    p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=256, pair_set=capi.PAIRS_ALL, include_direct=1)
    la, lb = [bench.make_sun(0.45, 0.55)], [bench.make_sun(0.55, 0.45)]
    out = capi.PinnedArray((H, W, 3), np.float64)
    out.array[...] = 0.0
    for k in range(n):
        t = eng.render_ghosts_sparse(la if k % 2 == 0 else lb, p, out.array, out_is_clear=(k == 0))
        st = eng.stats()
    print("tiles", t, "frame_ms", st["last_frame_ms"], "trace_ms", st["last_trace_ms"])
    out.free()
    eng.close()


if __name__ == "__main__":
    main()
