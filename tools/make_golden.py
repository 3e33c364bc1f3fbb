#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled by oracle/Makefile.

Run in the build container (needs /root/reference):   python tools/make_golden.py
Everything written here is an OUTPUT of the reference's own code for seeded/deterministic
inputs (or, for the aperture masks, the reference loader's decode of its own PNG assets):

  apertures.npz      CameraApertureTexture::init (camera.h:26-83) of final_apertures/pent_11.png and
                     pentbig500_14.png as uint8 (value*255, exact), plus total_value and bbox
  ref_vectors.npz    * the prescription matrices Ts / Ls / R_red|green|blue (pathtracer.cpp:539-586)
                     * trace_ray_auto_before/after (:588-689) on a table of (r, theta, i, j, colour)
                     * generate_ghost_buffer (:714-762) frames, stored sparsely (non-zero pixels)
                     * find_sun_pos (:32-64) for a few camera poses
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402

REF_ROOT = os.environ.get("LFB_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

FRAMES = [  # name, aperture, W, H, sun (ns_x, ns_y)
    ("f512_pent11_a", "pent_11", 512, 512, (0.7, 0.6)),
    ("f512_pent11_b", "pent_11", 512, 512, (0.45, 0.55)),
    ("f1080_pentbig_a", "pentbig500_14", 1920, 1080, (0.3, 0.2)),
    ("f1080_pentbig_b", "pentbig500_14", 1920, 1080, (0.45, 0.55)),
    ("f333_pent11_edge", "pent_11", 333, 217, (0.95, 0.08)),  # ghosts clipped by the frame border
]


def main():
    ob.build()
    ref = ob.RefOracle()
    os.makedirs(OUT, exist_ok=True)

    # --- aperture masks through the reference's own loader
    ap = {}
    for name in ("pent_11", "pentbig500_14"):
        tex, total, bbox = ref.load_aperture(os.path.join(REF_ROOT, "final_apertures", name + ".png"))
        u8 = np.rint(tex * 255.0)
        # Color(const uchar*) (CGL/src/color.cpp:16-21): value = byte * float(1/255), in float
        assert np.array_equal(u8.astype(np.float32) * np.float32(1.0 / 255.0), tex), "mask is not byte*inv255"
        ap[name] = u8.astype(np.uint8)
        ap[name + "_total"] = np.float64(total)
        ap[name + "_bbox"] = np.array(bbox, np.int32)
        print(name, tex.shape, "nonzero", int((tex > 0).sum()), "total", total, "bbox", bbox)
    np.savez_compressed(os.path.join(OUT, "apertures.npz"), **ap)

    g = {}
    # --- prescription
    pr = ref.prescription()
    g["presc_T"], g["presc_L"] = pr["T"], pr["L"]
    g["presc_R"] = np.stack(pr["R"])
    g["presc_curvature"], g["presc_ior"] = pr["curvature"], pr["ior"]
    g["sizeof_vector3d"] = np.int32(ref.sizeof_vector3d())

    # --- trace_ray_auto_* table
    rows = []
    rng = np.random.default_rng(184)
    rs = [14.5, -14.5, 7.25, -3.0, 0.5, -0.125, 12.0, -12.0]
    thetas = [0.0, 0.1, -0.3, float(np.float32(np.arctan(0.55 / 0.45))), 0.02]
    pairs = [(i, j) for i in range(9) for j in range(i + 1, 9) if i != 5 and j != 5]
    for (i, j) in pairs:
        which = 1 if i >= 6 else 0  # straddling pairs go through trace_ray_auto_before (it handles any i<j)
        for c in range(3):
            for r in rs + list(rng.uniform(-14.5, 14.5, 3)):
                for th in thetas:
                    r32, t32 = float(np.float32(r)), float(np.float32(th))
                    x, a = ref.trace(which, r32, t32, i, j, c)
                    rows.append((which, i, j, c, r32, t32, x, a))
    g["trace_table"] = np.array(rows, np.float64)
    print("trace rows", len(rows))

    # --- generate_ghost_buffer frames (sparse)
    tex_f = {k: ap[k].astype(np.float32) * np.float32(1.0 / 255.0) for k in ("pent_11", "pentbig500_14")}
    for name, apn, W, H, (ax, ay) in FRAMES:
        ang = float(np.float32(np.arctan(ay / ax)))  # pathtracer.cpp:50
        img = ref.generate_ghost_buffer(tex_f[apn], W, H, ax, ay, ang)
        flat = img.reshape(-1, 3)
        nz = np.flatnonzero((flat != 0).any(axis=1))
        g[name + "_meta"] = np.array([W, H, ax, ay, ang], np.float64)
        g[name + "_idx"] = nz.astype(np.int32)
        g[name + "_val"] = flat[nz]
        print(name, "nonzero px", nz.size, "sum", flat.sum(axis=0), "l2", np.sqrt((flat ** 2).sum()))
    g["frame_names"] = np.array([f[0] for f in FRAMES])
    g["frame_apertures"] = np.array([f[1] for f in FRAMES])

    # --- find_sun_pos
    sun = []
    poses = [
        (np.eye(3), (0, 0, 0), 50.0, 35.0, (0.2, 0.1, -1.0)),
        (np.eye(3), (0, 0, 1), 40.0, 30.0, (-0.1, 0.15, -2.0)),
        (np.array([[0.8, 0, 0.6], [0, 1, 0], [-0.6, 0, 0.8]]), (1, 2, 3), 60.0, 45.0, (0.5, 2.2, 1.0)),
        (np.eye(3), (0, 0, 0), 50.0, 35.0, (5.0, 0.1, -1.0)),  # off screen
    ]
    for c2w, pos, hf, vf, lp in poses:
        n, nx, ny, ang = ref.find_sun_pos(c2w, pos, hf, vf, lp, (0, 0, -1))
        sun.append(list(np.asarray(c2w, float).ravel()) + list(pos) + [hf, vf] + list(lp) + [n, nx, ny, ang])
    g["sun_table"] = np.array(sun, np.float64)
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), **g)

    # --- starburst: raytrace_starburst at sample pixels.  Light 0 = (origin, radiance (1,0,0)) drives the
    # DFT; light 1 sits far off screen with radiance (0,1,0), so channel G = the DFT scalar + a ~1e-8 falloff with
    # negligible noise (the falloff draws from the process-global RNG); columns 3..5 are a separate falloff draw.
    sb = {}
    cfgs = [  # name, aperture, W, H, origin, flare_radius, flare_intensity, n pixels
        ("sb512", "pent_11", 512, 512, (0.7, 0.6), 30.0, 1.0, 160),
        ("sb_odd", "pent_11", 333, 217, (0.4, 0.3), 10.0, 2.5, 120),
        ("sb1080", "pentbig500_14", 1920, 1080, (0.45, 0.55), 50.0, 4.0, 80),
    ]
    for name, apn, W, H, fo, radius, intensity, npx in cfgs:
        rng = np.random.default_rng(len(name))
        ox, oy = int(np.ceil(fo[0] * W)), int(np.ceil(fo[1] * H))
        xs = list(rng.integers(0, W, npx - 40)) + [min(max(ox + d, 0), W - 1) for d in rng.integers(-int(radius), int(radius) + 1, 40)]
        ys = list(rng.integers(0, H, npx - 40)) + [min(max(oy + d, 0), H - 1) for d in rng.integers(-int(radius), int(radius) + 1, 40)]
        xs[0], ys[0] = ox, oy
        out = ref.starburst_multi(tex_f[apn], W, H, [fo, (60.0, -45.0)], [(1, 0, 0), (0, 1, 0)], radius, intensity, xs, ys)
        sb[name + "_meta"] = np.array([W, H, fo[0], fo[1], radius, intensity], np.float64)
        sb[name + "_xy"] = np.array([xs, ys], np.int32)
        sb[name + "_out"] = out
        print(name, "pixels", len(xs), "G range", out[:, 1].min(), out[:, 1].max())
    sb["names"] = np.array([c[0] for c in cfgs])
    sb["apertures"] = np.array([c[1] for c in cfgs])
    sb["far_origin"] = np.array([60.0, -45.0])
    np.savez_compressed(os.path.join(OUT, "starburst.npz"), **sb)
    # the path-traced scene pass: PathTracer::est_radiance_global_illumination of the compiled reference (its own BVHAccel,
    # Triangle, Sphere, BSDFs and lights) on the scenes of tests/scene_fixtures.py, through the pixel centres
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_scene_pass
    sc = {}
    for name, scene, cam, W, H in test_scene_pass.cases():
        sc[name] = ref.scene_radiance(scene, cam, W, H)
        print("scene", name, sc[name].shape, "lit", float((sc[name].sum(axis=2) > 0).mean()), "max", float(sc[name].max()))
    np.savez_compressed(os.path.join(OUT, "scene.npz"), **sc)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
