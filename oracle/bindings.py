"""ctypes bindings for the two CPU oracles -- TEST INFRASTRUCTURE ONLY.

  RefOracle   oracle/_ref/libref_oracle.so : the UNMODIFIED reference sources compiled by
              oracle/Makefile (present here and, as a built file, on the GPU box).
  PortOracle  oracle/_build/liblf_oracle.so: the C restatement oracle/lf_oracle.c.

Only tests/, tools/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; nothing under lens_flare_b200/ does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)) if os.path.dirname(HERE) not in sys.path else None
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(HERE, "_build", "liblf_oracle.so")

# struct layouts / enums of include/lfb200.h are shared with the product's binding (the oracle
# may depend on the product's declarations; the product never imports the oracle)
from lens_flare_b200.capi import (  # noqa: E402,F401
    F32x3, F64x3, FP32, FP64, MAX_LAMBDA, MAX_SURFACES, MODE_EXACT_GRID, MODE_PARAXIAL_GRID, MODE_REF_QUADS, PAIRS_ALL,
    PAIRS_REF, RAY_HIT_DTYPE, RAY_MISSED, RAY_OFF_SENSOR, RAY_STOPPED, RAY_TIR, RAY_VIGNETTED, REF_GHOST_DTYPE,
    SPLAT_BILINEAR, SPLAT_NEAREST, Lens, Light, Params, copy_params, lights_array, make_light, make_params)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def build(which=("port", "ref")):
    """Compile the oracles (ref only when /root/reference is present)."""
    ref_root = os.environ.get("LFB_REFERENCE", "/root/reference")
    for w in which:
        if w == "ref" and not os.path.isdir(ref_root):
            continue
        subprocess.run(["make", "-s", "-C", HERE, w, f"REF={ref_root}"], check=True)


class RefOracle:
    """The compiled reference (oracle/_ref)."""

    def __init__(self, path=REF_SO):
        self.lib = L = C.CDLL(path)
        L.ref_trace.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.ref_generate_ghost_buffer.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_double, C.c_double, C.c_float, C.POINTER(C.c_double)]
        L.ref_time_ghost_buffer.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.c_double, C.c_float, C.c_int, C.POINTER(C.c_double)]
        L.ref_time_ghost_buffer.restype = C.c_double
        L.ref_time_trace_grid.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]
        L.ref_time_trace_grid.restype = C.c_double
        L.ref_load_aperture.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_find_sun_pos.argtypes = [C.POINTER(C.c_double)] * 2 + [C.c_double, C.c_double] + \
            [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_float)]
        L.ref_starburst.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double)]

    def starburst_multi(self, tex, W, H, origins, radiances, flare_radius, flare_intensity, xs, ys):
        """-> (n, 6): raytrace_starburst rgb, then a separate calculate_irradiance_falloff draw rgb."""
        tex = np.ascontiguousarray(tex, np.float32)
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        fo = np.ascontiguousarray(origins, np.float64).reshape(-1, 2)
        rad = np.ascontiguousarray(radiances, np.float64).reshape(-1, 3)
        out = np.zeros((xs.size, 6))
        f = self.lib.ref_starburst_multi
        f.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                      C.POINTER(C.c_double)]
        f(_fp(tex), tex.shape[1], tex.shape[0], W, H, fo.shape[0], _dp(fo), _dp(rad), flare_radius, flare_intensity,
          xs.ctypes.data_as(C.POINTER(C.c_int)), ys.ctypes.data_as(C.POINTER(C.c_int)), xs.size, _dp(out))
        return out

    def scene_radiance(self, scene, cam, W, H):
        """PathTracer::est_radiance_global_illumination of the compiled reference over its own BVH, per pixel centre."""
        self.lib.ref_scene_radiance.argtypes = SCENE_ARGTYPES
        keep, out, args = _scene_args(scene, cam, W, H)
        assert self.lib.ref_scene_radiance(*args) == 0
        return out

    def to_color(self, hdr):
        hdr = np.ascontiguousarray(hdr, np.float64)
        out = np.zeros(hdr.shape[:2], np.uint32)
        self.lib.ref_to_color.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p]
        self.lib.ref_to_color(_dp(hdr), hdr.shape[1], hdr.shape[0], out.ctypes.data)
        return out

    def sizeof_vector3d(self):
        return self.lib.ref_sizeof_vector3d()

    def prescription(self):
        t, l, rr, rg, rb = (np.zeros((9, 4)) for _ in range(5))
        curv = np.zeros(10, np.float32)
        n = np.zeros((3, 9), np.float32)
        self.lib.ref_prescription(_dp(t), _dp(l), _dp(rr), _dp(rg), _dp(rb), _fp(curv), _fp(n[0]), _fp(n[1]), _fp(n[2]))
        return dict(T=t, L=l, R=[rr, rg, rb], curvature=curv, ior=n)

    def trace(self, which, r, theta, i, j, colour):
        out = (C.c_double * 2)()
        self.lib.ref_trace(which, r, theta, i, j, colour, out)
        return out[0], out[1]

    def generate_ghost_buffer(self, tex, W, H, ns_x, ns_y, angle):
        tex = np.ascontiguousarray(tex, np.float32)
        out = np.zeros((H, W, 3))
        rc = self.lib.ref_generate_ghost_buffer(_fp(tex), tex.shape[1], tex.shape[0], W, H, ns_x, ns_y, angle, _dp(out))
        assert rc == 0
        return out

    def time_ghost_buffer(self, tex, W, H, ns_x, ns_y, angle, reps):
        tex = np.ascontiguousarray(tex, np.float32)
        cs = C.c_double()
        s = self.lib.ref_time_ghost_buffer(_fp(tex), tex.shape[1], tex.shape[0], W, H, ns_x, ns_y, angle, reps, C.byref(cs))
        return s, cs.value

    def time_trace_grid(self, N, theta, ncol, pairset, nthreads):
        cs, rays = C.c_double(), C.c_double()
        s = self.lib.ref_time_trace_grid(N, theta, ncol, pairset, nthreads, C.byref(cs), C.byref(rays))
        return s, rays.value, cs.value

    def load_aperture(self, path):
        out = np.zeros(4096 * 4096 // 16, np.float32)
        w, h, tot = C.c_int(), C.c_int(), C.c_double()
        bbox = (C.c_int * 4)()
        rc = self.lib.ref_load_aperture(path.encode(), _fp(out), out.size, C.byref(w), C.byref(h), C.byref(tot), bbox)
        assert rc == 0
        return out[: w.value * h.value].reshape(h.value, w.value).copy(), tot.value, tuple(bbox)

    def find_sun_pos(self, c2w, cam_pos, hfov, vfov, light_pos, light_dir, radiance=(1, 1, 1)):
        a = [np.ascontiguousarray(v, np.float64) for v in (c2w, cam_pos, light_pos, light_dir, radiance)]
        ns = np.zeros(2)
        ang = C.c_float()
        n = self.lib.ref_find_sun_pos(_dp(a[0]), _dp(a[1]), hfov, vfov, _dp(a[2]), _dp(a[3]), _dp(a[4]), _dp(ns), C.byref(ang))
        return n, ns[0], ns[1], ang.value

    def starburst(self, tex, W, H, fo, radiance, flare_radius, flare_intensity, xs, ys):
        tex = np.ascontiguousarray(tex, np.float32)
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        rad = np.ascontiguousarray(radiance, np.float64)
        out = np.zeros((xs.size, 6))
        self.lib.ref_starburst(_fp(tex), tex.shape[1], tex.shape[0], W, H, fo[0], fo[1], _dp(rad), flare_radius,
                               flare_intensity, xs.ctypes.data_as(C.POINTER(C.c_int)),
                               ys.ctypes.data_as(C.POINTER(C.c_int)), xs.size, _dp(out))
        return out


def _scene_args(scene, cam, W, H):
    """ctypes arguments of ref_scene_radiance / lfo_scene_radiance (layouts: oracle/ref_shim.cpp)."""
    tp = np.ascontiguousarray(scene["tri_pos"], np.float64)
    tn = np.ascontiguousarray(scene["tri_nrm"], np.float64)
    tm = np.ascontiguousarray(scene["tri_mat"], np.int32)
    sp = np.ascontiguousarray(scene["spheres"], np.float64).reshape(-1, 4)
    sm = np.ascontiguousarray(scene["sph_mat"], np.int32)
    ma = np.ascontiguousarray(scene["mats"], np.float64)
    li = np.ascontiguousarray(scene["lights"], np.float64)
    ca = np.ascontiguousarray(cam, np.float64)
    out = np.zeros((H, W, 3))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
    keep = (tp, tn, tm, sp, sm, ma, li, ca)
    return keep, out, [_dp(tp), _dp(tn), ip(tm), tm.size, _dp(sp), ip(sm), sm.size, _dp(ma), ma.shape[0], _dp(li), li.shape[0], _dp(ca), W, H, _dp(out)]


SCENE_ARGTYPES = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int,
                  C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_double)]


class PortOracle:
    """The C restatement (oracle/lf_oracle.c)."""

    def scene_radiance(self, scene, cam, W, H):
        """The path-traced scene pass (emission + direct lighting of delta lights) per pixel centre: (H, W, 3)."""
        self.lib.lfo_scene_radiance.argtypes = SCENE_ARGTYPES
        keep, out, args = _scene_args(scene, cam, W, H)
        assert self.lib.lfo_scene_radiance(*args) == 0
        return out

    def __init__(self, path=PORT_SO):
        self.lib = L = C.CDLL(path)
        LP, LiP, PP = C.POINTER(Lens), C.POINTER(Light), C.POINTER(Params)
        L.lfo_builtin_lens.argtypes = [LP, C.c_int, C.c_float]
        L.lfo_prescription.argtypes = [LP, C.c_int] + [C.POINTER(C.c_double)] * 3
        L.lfo_trace_ray_auto.argtypes = [LP, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.lfo_generate_ghost_buffer.argtypes = [LP, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                                C.c_double, C.c_float, C.POINTER(C.c_double), C.c_void_p, C.c_int]
        L.lfo_paraxial_system.argtypes = [LP, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.lfo_trace_grid.argtypes = [LP, C.POINTER(C.c_float), C.c_int, C.c_int, LiP, PP, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.lfo_render.argtypes = [LP, C.POINTER(C.c_float), C.c_int, C.c_int, LiP, C.c_int, PP, C.POINTER(C.c_double), C.c_void_p]
        L.lfo_reflectance.argtypes = [C.c_double] * 5
        L.lfo_reflectance.restype = C.c_double
        L.lfo_time_render.argtypes = [LP, C.POINTER(C.c_float), C.c_int, C.c_int, LiP, C.c_int, PP, C.c_int, C.POINTER(C.c_double)]
        L.lfo_time_render.restype = C.c_double

    def builtin_lens(self, n_lambda=3, coating_lambda0_nm=0.0):
        lens = Lens()
        assert self.lib.lfo_builtin_lens(C.byref(lens), n_lambda, coating_lambda0_nm) == 0
        return lens

    def prescription(self, lens, lam):
        t, l, r = (np.zeros((lens.n_surfaces, 4)) for _ in range(3))
        self.lib.lfo_prescription(C.byref(lens), lam, _dp(t), _dp(l), _dp(r))
        return t, l, r

    def trace_ray_auto(self, lens, lam, which, r, theta, i, j):
        out = (C.c_double * 2)()
        self.lib.lfo_trace_ray_auto(C.byref(lens), lam, which, r, theta, i, j, out)
        return out[0], out[1]

    def generate_ghost_buffer(self, lens, tex, W, H, ns_x, ns_y, angle, want_ghosts=False):
        tex = np.ascontiguousarray(tex, np.float32)
        out = np.zeros((H, W, 3))
        cap = 64 * lens.n_lambda
        ghosts = np.zeros(cap, REF_GHOST_DTYPE)
        n = self.lib.lfo_generate_ghost_buffer(C.byref(lens), _fp(tex), tex.shape[1], tex.shape[0], W, H, ns_x, ns_y,
                                               angle, _dp(out), ghosts.ctypes.data, cap)
        return (out, ghosts[:n]) if want_ghosts else out

    def paraxial_system(self, lens, lam, i, j, physical_backward=0):
        cross = np.zeros((3, 4))
        full = np.zeros(4)
        nc = self.lib.lfo_paraxial_system(C.byref(lens), lam, i, j, physical_backward, _dp(cross), _dp(full))
        return cross[:nc], full

    def trace_grid(self, lens, tex, light, params, i, j, lam):
        tex = np.ascontiguousarray(tex, np.float32)
        out = np.zeros(params.grid_n * params.grid_n, RAY_HIT_DTYPE)
        rc = self.lib.lfo_trace_grid(C.byref(lens), _fp(tex), tex.shape[1], tex.shape[0], C.byref(light), C.byref(params),
                                     i, j, lam, out.ctypes.data)
        assert rc == 0
        return out

    def render(self, lens, tex, lights, params, want_accum=False):
        tex = np.ascontiguousarray(tex, np.float32)
        out = np.zeros((params.height, params.width, 3))
        acc = np.zeros((params.height, params.width, 3), np.int64) if want_accum else None
        rc = self.lib.lfo_render(C.byref(lens), _fp(tex), tex.shape[1], tex.shape[0], lights_array(lights), len(lights),
                                 C.byref(params), _dp(out), acc.ctypes.data if want_accum else None)
        assert rc == 0
        return (out, acc) if want_accum else out

    def render_mt(self, lens, tex, lights, params, nthreads=None):
        """render() on all host threads (the same bits: integer sums)."""
        tex = np.ascontiguousarray(tex, np.float32)
        out = np.zeros((params.height, params.width, 3))
        f = self.lib.lfo_render_mt
        f.argtypes = [C.POINTER(Lens), C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(Light), C.c_int, C.POINTER(Params), C.c_int, C.POINTER(C.c_double)]
        rc = f(C.byref(lens), _fp(tex), tex.shape[1], tex.shape[0], lights_array(lights), len(lights), C.byref(params), nthreads or os.cpu_count() or 1,
               _dp(out))
        assert rc == 0
        return out

    def starburst_pixels(self, tex, W, H, origins, radiances, flare_radius, flare_intensity, xs, ys):
        """-> (dft scalar (n,), deterministic falloff (n, 3))"""
        tex = np.ascontiguousarray(tex, np.float32)
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        fo = np.ascontiguousarray(origins, np.float64).reshape(-1, 2)
        rad = np.ascontiguousarray(radiances, np.float64).reshape(-1, 3)
        dft, fall = np.zeros(xs.size), np.zeros((xs.size, 3))
        f = self.lib.lfo_starburst_pixels
        f.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
        rc = f(_fp(tex), tex.shape[1], tex.shape[0], W, H, fo.shape[0], _dp(fo), _dp(rad), flare_radius, flare_intensity,
               xs.ctypes.data_as(C.POINTER(C.c_int)), ys.ctypes.data_as(C.POINTER(C.c_int)), xs.size, _dp(dft), _dp(fall))
        assert rc == 0
        return dft, fall

    def to_color(self, hdr):
        hdr = np.ascontiguousarray(hdr, np.float64)
        out = np.zeros(hdr.shape[:2], np.uint32)
        self.lib.lfo_to_color.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p]
        self.lib.lfo_to_color(_dp(hdr), hdr.shape[1], hdr.shape[0], out.ctypes.data)
        return out

    def reflectance(self, n0, n2, cos0, lambda0, lam):
        return self.lib.lfo_reflectance(n0, n2, cos0, lambda0, lam)

    def time_render(self, lens, tex, lights, params, nthreads):
        tex = np.ascontiguousarray(tex, np.float32)
        cs = C.c_double()
        s = self.lib.lfo_time_render(C.byref(lens), _fp(tex), tex.shape[1], tex.shape[0], lights_array(lights),
                                     len(lights), C.byref(params), nthreads, C.byref(cs))
        return s, cs.value
