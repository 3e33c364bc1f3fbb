"""A small pass over every kernel (all modes, both precisions, rect / frame / starburst paths) at sizes that finish in
seconds under compute-sanitizer:   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_flare_b200 import capi  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
tex = z["pent_11"].astype(np.float32) * np.float32(1.0 / 255.0)
e = capi.Engine(0)
e.set_lens(capi.builtin_lens(3, 550.0))
e.set_aperture(tex)
e.set_starburst_aperture(tex)
lt = [capi.make_light(0.45, 0.55, theta=0.06), capi.make_light(0.9, 0.1, theta=0.12)]
W, H = 97, 61
total = 0.0
for mode in (capi.MODE_REF_QUADS, capi.MODE_PARAXIAL_GRID, capi.MODE_EXACT_GRID):
    for prec in (capi.FP32, capi.FP64):
        for splat in (capi.SPLAT_NEAREST, capi.SPLAT_BILINEAR):
            p = capi.make_params(mode, W, H, grid_n=37, pair_set=capi.PAIRS_ALL if mode else capi.PAIRS_REF, include_direct=int(mode != 0),
                                 precision=prec, splat=splat, px_per_unit=0.3)
            total += e.render_ghosts(lt, p).sum()
            buf = np.zeros((H, W, 4))
            e.render_ghosts_rect(lt, p, buf, stride=32)
            total += buf.sum()
            if mode:
                total += e.dump_rays(lt[0], p, 1, 7, 2)["weight"].sum()
                total += e.render_ghosts(lt, capi.copy_params(p, shard=(1, 3))).sum()
p = capi.make_params(capi.MODE_EXACT_GRID, 640, 360, grid_n=50, pair_set=capi.PAIRS_ALL, include_direct=1, px_per_unit=5.0)
total += e.render_ghosts(lt, p).sum()  # footprint larger than the smem tile: global-atomic fallback
total += e.render_starburst(lt, W, H, 9.0, 1.0).sum()
total += float(e.render_frame_rgba8(lt, p, flare_radius=9.0, base_hdr=np.zeros((360, 640, 3))).sum())
print("sanitize_small ok, checksum", total, e.probe_peaks()["fp32_flops"] > 0)
e.close()
