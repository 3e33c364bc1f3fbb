#!/usr/bin/env python
"""Where does a pipelined end-to-end frame (lfb_render_ghosts_sparse_begin / _end, R in flight) spend its host time?

  python tools/e2e_probe.py [--steps 400]

cfg2, suns alternating per slot like bench.py's e2e leg.  Per variant: ms per frame and the host seconds spent inside each
call (perf_counter around the ctypes calls).  Variants: the bench's loop; without the per-frame aperture upload; with a fixed
sun (no job rebuild); both.  A measurement tool, not a bench.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lens_flare_b200 import capi
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--reduce-ctas", default="0")
    ap.add_argument("--slots", default="2,3,4")
    ap.add_argument("--mbps", default="0", help="lfb_options.host_write_mbps values (0 auto, -1 unpaced)")
    args = ap.parse_args()
    tex = bench.load_aperture()
    W, H = 1920, 1080
    lens = capi.builtin_lens(3, 550.0)
    p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=256, pair_set=capi.PAIRS_ALL, include_direct=1)
    la, lb = [bench.make_sun(0.45, 0.55)], [bench.make_sun(0.55, 0.45)]
    la2 = [bench.make_sun(0.45, 0.55, radiance=(1.0001, 1.0, 1.0))]
    ptex = capi.PinnedArray(tex.shape, np.float32)
    ptex.array[...] = tex
    outs = [capi.PinnedArray((H, W, 3), np.float64) for _ in range(4)]
    cases = [(name, upload, moving, int(c), int(R), int(m)) for m in args.mbps.split(",") for c in args.reduce_ctas.split(",") for R in args.slots.split(",")
             for name, upload, moving in (("bench_loop", True, True), ("no_aperture_upload", False, True), ("fixed_sun", True, False), ("neither", False, False), ("fixed_sun_new_tables", False, "rekey"))]
    for name, upload, moving, ctas, R, mbps in cases:
        eng = capi.Engine(0, reduce_ctas=ctas, host_write_mbps=mbps)
        eng.set_lens(lens)
        eng.set_aperture(tex)
        for o in outs:
            o.array[...] = 0.0
        pending, fresh = [False] * R, [True] * R
        t_call = {"end": 0.0, "set_aperture": 0.0, "begin": 0.0}
        timeline = []

        def step(k):
            s = k % R
            if pending[s]:
                t0 = time.perf_counter()
                eng.render_ghosts_sparse_end(s)
                t_call["end"] += time.perf_counter() - t0
                timeline.append(eng.sparse_slot_times(s))
            if upload:
                t0 = time.perf_counter()
                eng.set_aperture(ptex.array)
                t_call["set_aperture"] += time.perf_counter() - t0
            if moving == "rekey":  # the same sun, a hair brighter every other frame: new tables go up, the tiles stay the same
                lights = (la, la2)[k % 2]
            else:
                lights = (la, lb)[(k // R + s) % 2] if moving else la  # every slot's buffer alternates between the suns
            t0 = time.perf_counter()
            eng.render_ghosts_sparse_begin(lights, p, outs[s].array, s, out_is_clear=fresh[s])
            t_call["begin"] += time.perf_counter() - t0
            pending[s], fresh[s] = True, False

        for k in range(20):
            step(k)
        for s in range(R):
            eng.render_ghosts_sparse_end(s)
            pending[s] = False
        for key in t_call:
            t_call[key] = 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(args.steps):
            step(k)
        for s in range(R):
            if pending[s]:
                eng.render_ghosts_sparse_end(s)
                pending[s] = False
        dt = time.perf_counter() - t0
        print(json.dumps({"variant": name, "reduce_ctas": ctas, "slots": R, "host_write_mbps": mbps, "ms_per_frame": dt * 1e3 / args.steps,
                          "host_ms_per_frame_in": {k2: v * 1e3 / args.steps for k2, v in t_call.items()}}), flush=True)
        print("   trace kernels of the last frame (ev_trace0 -> ev_trace1): %.1f us" % (eng.stats()["last_trace_ms"] * 1e3))
        base = timeline[-8][0]
        print("   device timeline of 8 frames (us): [stream reached | traced | tile start | staged | done]",
              " ".join("[%.0f %.0f %.0f %.0f %.0f]" % tuple((v - base) * 1e3 for v in t) for t in timeline[-8:]), flush=True)
        eng.close()
    for o in outs:
        o.free()
    ptex.free()


if __name__ == "__main__":
    main()
