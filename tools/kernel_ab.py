#!/usr/bin/env python
"""A/B timing of the EXACT_GRID throughput kernels' variants on ONE GPU in ONE process (same box, same clocks).

  python tools/kernel_ab.py [--cfg 2|3] [--reps 30] [--only name,name]

Per variant: the engine's own CUDA events around the trace kernels of a frame (lfb_stats.last_trace_ms), one frame at a time,
L2 flushed (256 MiB memset) before every sample; prints median / min per variant as JSON lines.  A measurement tool, not a bench.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "pairs": dict(kernel_select=1),
    "pairs_alt": dict(kernel_select=1, ctas_per_sm=1),
    "pairs_unstaged": dict(kernel_select=1, ctas_per_sm=2),
    "pairs_landcall": dict(kernel_select=1, ctas_per_sm=3),
    "pairs_peek_l2": dict(kernel_select=1, experiment=1),
    "families_peek_l2": dict(kernel_select=2, experiment=1),
    "pairs_table": dict(kernel_select=1, weights_table=1),
    "pairs_nooverlap": dict(kernel_select=1, prefix_overlap=-1),
    "pairs_nocache": dict(kernel_select=1, prefix_budget_bytes=-1),
    "families": dict(kernel_select=2),
    "families_alt": dict(kernel_select=2, ctas_per_sm=1),
    "families_split2": dict(kernel_select=2, family_split=2),
    "families_table": dict(kernel_select=2, weights_table=1),
    "auto": dict(),
}


def main():
    import torch
    from lens_flare_b200 import capi
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--only", default="")
    ap.add_argument("--precisions", default="fp32,strict")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    tex = bench.load_aperture()
    if args.cfg == 2:
        lens, grid, W, H = capi.builtin_lens(3, 550.0), 256, 1920, 1080
    elif args.cfg == 3:
        lens, grid, W, H = capi.builtin_lens(32, 550.0), 512, 1920, 1080
    else:
        raise SystemExit("--cfg 2 or 3")
    lights = [bench.make_sun(0.45, 0.55)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    names = [n for n in VARIANTS if not args.only or n in args.only.split(",")]
    precs = {"fp32": capi.FP32, "strict": capi.STRICT}
    for name in names:
        eng = capi.Engine(0, **VARIANTS[name])
        eng.set_lens(lens)
        eng.set_aperture(tex)
        for pname in args.precisions.split(","):
            p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=grid, pair_set=capi.PAIRS_ALL, include_direct=1, precision=precs[pname])
            acc = torch.zeros((capi.lib().lfb_accum_bytes(W, H) + 7) // 8, dtype=torch.int64, device=dev)
            _, inter, _ = capi.count_work(lens, p, 1)
            ts = []
            for k in range(args.reps + 3):
                flush.zero_()
                torch.cuda.synchronize()
                eng.render_ghosts_device(lights, p, acc.data_ptr(), clear_first=True)
                eng.sync()
                if k >= 3:
                    ts.append(eng.stats()["last_trace_ms"])
            ts.sort()
            print(json.dumps({"variant": name, "precision": pname, "cfg": args.cfg, "trace_ms_median": ts[len(ts) // 2], "trace_ms_min": ts[0],
                              "trace_ms_p90": ts[int(len(ts) * 0.9)], "interactions_per_s_median": inter / (ts[len(ts) // 2] * 1e-3),
                              "checksum": int(acc[: W * H * 3].sum().item())}), flush=True)
            del acc
        eng.close()


if __name__ == "__main__":
    main()
