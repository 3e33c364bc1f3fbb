#!/usr/bin/env python
"""A tiny pass through the tile-sparse host paths (blocking call, frames in flight with staging + paced drain, device finalize) for
compute-sanitizer:   compute-sanitizer --tool memcheck python tools/sanitize_sparse.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from lens_flare_b200 import capi
    z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
    tex = z["pent_11"].astype(np.float32) * np.float32(1.0 / 255.0)
    e = capi.Engine(0, host_write_mbps=20000)
    e.set_lens(capi.builtin_lens(3, 550.0))
    e.set_aperture(tex)
    ok = True
    for (W, H) in ((320, 176), (250, 99)):
        p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=24, pair_set=capi.PAIRS_ALL, include_direct=1)
        mk = lambda x, y: capi.make_light(x, y, theta=capi.physical_theta(x, y))  # noqa: E731
        seq = [[mk(0.45, 0.55)], [mk(0.6, 0.4)], [], [mk(0.45, 0.55), mk(0.3, 0.6)], [mk(0.5, 0.52)]]
        want = [e.render_ghosts(l, p) for l in seq]
        bufs = [capi.PinnedArray((H, W, 3), np.float64) for _ in range(3)]
        for b in bufs:
            b.array[...] = 0
        got = [None] * len(seq)
        for k, l in enumerate(seq):
            s = k % 3
            if k >= 3:
                e.render_ghosts_sparse_end(s)
                got[k - 3] = bufs[s].array.copy()
            e.render_ghosts_sparse_begin(l, p, bufs[s].array, s, out_is_clear=(k < 3))
        for k in range(len(seq) - 3, len(seq)):
            e.render_ghosts_sparse_end(k % 3)
            got[k] = bufs[k % 3].array.copy()
        ok &= all(np.array_equal(g, w) for g, w in zip(got, want))
        one = capi.PinnedArray((H, W, 3), np.float64)
        one.array[...] = 0
        for k, l in enumerate(seq):
            e.render_ghosts_sparse(l, p, one.array, out_is_clear=(k == 0))
            ok &= bool(np.array_equal(one.array, want[k]))
        one.free()
        for b in bufs:
            b.free()
    e.close()
    print("frames equal:", ok)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
