#!/usr/bin/env python
"""Where do lfb_render_frame_rgba8 and oracle.to_color(host-side sum of the engine's own HDR layers) differ?  Measurement tool."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_flare_b200 import capi  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def main():
    port = ob.PortOracle()
    e = capi.Engine(0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
    ap = {k: z[k].astype(np.float32) * np.float32(1 / 255.0) for k in z.files}
    e.set_lens(capi.builtin_lens(3, 550.0))
    e.set_aperture(ap["pentbig500_14"])
    e.set_starburst_aperture(ap["pent_11"])
    W, H = 640, 360
    rng = np.random.default_rng(2)
    base = rng.uniform(0, 0.2, (H, W, 3))
    for mode in (capi.MODE_REF_QUADS, capi.MODE_EXACT_GRID):
        p = capi.make_params(mode, W, H, grid_n=128, pair_set=capi.PAIRS_ALL if mode else capi.PAIRS_REF, include_direct=int(mode != 0))
        lt = [capi.make_light(0.45, 0.55, radiance=(1.0, 0.9, 0.7))] if mode == 0 else \
             [capi.make_light(0.45, 0.55, theta=capi.physical_theta(0.45, 0.55), radiance=(1.0, 0.9, 0.7))]
        ghosts = e.render_ghosts(lt, p)
        star = e.render_starburst(lt, W, H, 25.0, 1.0)
        star2 = e.render_starburst(lt, W, H, 25.0, 1.0)
        print("mode", mode, "starburst repeatable:", np.array_equal(star, star2))
        acc = base.copy()
        e.render_starburst(lt, W, H, 25.0, 1.0, out=acc, additive=True)
        print("   additive starburst == base + star:", np.array_equal(acc, base + star), "max rel", np.abs(acc - (base + star)).max())
        for name, b, use_star in (("ghosts", None, False), ("base+ghosts", base, False), ("ghosts+star", None, True), ("base+ghosts+star", base, True)):
            got = e.render_frame_rgba8(lt, p, flare_radius=25.0 if use_star else -1.0, flare_intensity=1.0, base_hdr=b)
            for order in ("(g+s)+b", "(b+g)+s"):
                if order == "(g+s)+b":
                    hdr = ghosts + (star if use_star else 0) + (b if b is not None else 0)
                else:
                    hdr = (b if b is not None else 0) + ghosts
                    if use_star:
                        hdr = hdr + star
                want = port.to_color(hdr)
                d = np.abs(got.view(np.uint8).astype(int) - np.ascontiguousarray(want).view(np.uint8).astype(int))
                print("   %-18s host order %s: mismatches %d max %d" % (name, order, int((d > 0).sum()), int(d.max())))


if __name__ == "__main__":
    main()
