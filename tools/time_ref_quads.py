"""Times the literal drop-in mode (REF_QUADS = PathTracer::generate_ghost_buffer) at 1080p: blocking full-frame call,
dirty-rectangle call (host clears the previous rectangle), and the reference itself on one host thread."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lens_flare_b200 import capi  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
tex = z["pentbig500_14"].astype(np.float32) * np.float32(1.0 / 255.0)
e = capi.Engine(0)
e.set_lens(capi.builtin_lens(3))
e.set_aperture(tex)
W, H = 1920, 1080
p = capi.make_params(capi.MODE_REF_QUADS, W, H)
suns = [capi.make_light(0.45, 0.55), capi.make_light(0.3, 0.2)]
out = capi.PinnedArray((H, W, 3), np.float64)
res = {}
for k in range(3):
    e.render_ghosts([suns[k % 2]], p, out=out.array)
t = time.perf_counter()
for k in range(50):
    e.render_ghosts([suns[k % 2]], p, out=out.array)
res["full_frame_ms"] = (time.perf_counter() - t) / 50 * 1e3
res["device_ms"] = e.stats()["last_trace_ms"]
out.array[...] = 0
rect = None
for k in range(3):
    if rect:
        out.array[rect[1]:rect[3] + 1, rect[0]:rect[2] + 1] = 0
    rect = e.render_ghosts_rect([suns[k % 2]], p, out.array)
t = time.perf_counter()
for k in range(50):
    if rect:
        out.array[rect[1]:rect[3] + 1, rect[0]:rect[2] + 1] = 0
    rect = e.render_ghosts_rect([suns[k % 2]], p, out.array)
res["rect_ms"] = (time.perf_counter() - t) / 50 * 1e3
res["rect"] = rect
try:
    from oracle import bindings as ob
    if os.path.exists(ob.REF_SO):
        ref = ob.RefOracle()
        s, _ = ref.time_ghost_buffer(tex, W, H, 0.45, 0.55, suns[0].theta, 5)
        res["reference_cpu_ms_1thread"] = s * 1e3
except Exception as exc:
    res["reference_cpu_ms_1thread"] = repr(exc)
print(json.dumps(res))
