#!/usr/bin/env python
"""How often does the device's toColor differ from the oracle's (glibc pow) on random HDR values?  Measurement tool."""
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
    e.set_lens(capi.builtin_lens(3))
    z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
    e.set_aperture(z["pent_11"].astype(np.float32) / 255)
    W, H = 2048, 1024
    p = capi.make_params(capi.MODE_REF_QUADS, W, H)
    rng = np.random.default_rng(11)
    for name, hdr in (("loguniform", 10 ** rng.uniform(-6, 0.3, (H, W, 3))), ("uniform", rng.uniform(0, 0.75, (H, W, 3))),
                      ("float32-valued", rng.uniform(0, 0.75, (H, W, 3)).astype(np.float32).astype(np.float64))):
        got = e.render_frame_rgba8([], p, flare_radius=-1.0, base_hdr=hdr)
        want = port.to_color(hdr)
        g8, w8 = got.view(np.uint8).astype(int), np.ascontiguousarray(want).view(np.uint8).astype(int)
        diff = np.abs(g8 - w8)
        print(name, "mismatches", int((diff > 0).sum()), "of", diff.size, "max", int(diff.max()))
        bad = np.argwhere(diff.reshape(H, W, 4) > 0)[:5]
        for (y, x, c) in bad:
            v = hdr[y, x, c]
            print("   hdr %.17g  device %d  glibc %d" % (v, g8.reshape(H, W, 4)[y, x, c], w8.reshape(H, W, 4)[y, x, c]))


if __name__ == "__main__":
    main()
