#!/usr/bin/env python
"""The timed region of bench.py's `value` leg and nothing else (N = 1, cfg2): W + K device-resident frames through the same
two-stream SparsePipeline (trace -> tile-sparse finalize into a device frame).  The command ncu's launch list is taken from when
only the step's own kernels should be in it:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv python tools/value_loop.py [K]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lens_flare_b200 import capi
    import bench
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    tex = bench.load_aperture()
    eng, fin = capi.Engine(0), capi.Engine(0, stream_priority=1)
    lens = capi.builtin_lens(3, bench.COATING_NM)
    for e in (eng, fin):
        e.set_lens(lens)
        e.set_aperture(tex)
    p = capi.make_params(capi.MODE_EXACT_GRID, bench.WIDTH, bench.HEIGHT, grid_n=bench.GRID_N, pair_set=capi.PAIRS_ALL, include_direct=1)
    lights = [bench.make_sun(*bench.sun_positions(1)[0])]
    sp = bench.SparsePipeline(eng, fin, p, dev, 3, torch, capi)
    for _ in range(5 + K):
        sp.frame(lights)
    sp.finish()
    torch.cuda.synchronize()
    print("frames", 5 + K, "launches", eng.stats()["kernel_launches"] + fin.stats()["kernel_launches"])
    fin.close()
    eng.close()


if __name__ == "__main__":
    main()
