"""ctypes binding of the C ABI declared in include/lfb200.h (liblfb200.so).

This is the only module that talks to the native library; everything above it
(`pathtracer.py`, `sharding.py`, bench.py, tests) goes through these calls.  The
library is built in-tree by `lens_flare_b200/csrc/Makefile` (see `build()`); it is
never looked up in site-packages, and loading fails loudly when it is missing --
there is no Python or CPU fallback for any compute entry point.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblfb200.so")
CSRC = os.path.join(HERE, "csrc")

MAX_SURFACES = 16
MAX_LAMBDA = 64
ABI_VERSION = 3

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE, ERR_NOMEM = 0, -1, -2, -3, -4, -5
MODE_REF_QUADS, MODE_PARAXIAL_GRID, MODE_EXACT_GRID = 0, 1, 2
PAIRS_REF, PAIRS_ALL = 0, 1
F32x3, F64x3 = 0, 1
FP32, FP64, STRICT = 0, 1, 2
SPLAT_NEAREST, SPLAT_BILINEAR = 0, 1
RAY_MISSED, RAY_VIGNETTED, RAY_TIR, RAY_STOPPED, RAY_OFF_SENSOR = 1, 2, 4, 8, 16

# every symbol include/lfb200.h declares (tests check the library exports each one)
SYMBOLS = (
    "lfb_abi_version", "lfb_create", "lfb_create_ex", "lfb_exec_stats", "lfb_set_scene", "lfb_render_scene", "lfb_render_composite_rgba8", "lfb_create_multi", "lfb_destroy_multi", "lfb_multi_set_lens", "lfb_multi_set_aperture", "lfb_render_ghosts_multi", "lfb_multi_stats", "lfb_render_ghosts_sparse", "lfb_render_ghosts_sparse_begin", "lfb_render_ghosts_sparse_end", "lfb_sparse_slot_times", "lfb_tile_state_bytes", "lfb_finalize_tiles_device", "lfb_reduce_tiles_peers", "lfb_tile_stage_bytes", "lfb_reduce_tiles_peers_staged", "lfb_drain_tiles", "lfb_destroy", "lfb_last_error", "lfb_builtin_lens", "lfb_set_lens",
    "lfb_set_aperture", "lfb_render_ghosts", "lfb_render_ghosts_rect", "lfb_render_ghosts_async", "lfb_dump_rays", "lfb_ref_ghosts", "lfb_accum_bytes", "lfb_stream",
    "lfb_render_ghosts_device", "lfb_finalize_device", "lfb_sync", "lfb_reduce_finalize_peers", "lfb_peer_barrier", "lfb_count_work", "lfb_list_jobs", "lfb_stats",
    "lfb_host_alloc", "lfb_host_free", "lfb_host_register", "lfb_host_unregister", "lfb_host_device_pointer", "lfb_finalize_clear_device", "lfb_probe_peaks", "lfb_set_starburst_aperture", "lfb_render_starburst", "lfb_render_frame_rgba8",
)


class Lens(C.Structure):
    """lfb_lens"""
    _fields_ = [
        ("n_surfaces", C.c_int32), ("stop_index", C.c_int32), ("n_lambda", C.c_int32), ("reserved0", C.c_int32),
        ("curvature", C.c_float * MAX_SURFACES), ("thickness", C.c_float * MAX_SURFACES),
        ("semi_aperture", C.c_float * MAX_SURFACES), ("coating_lambda0_nm", C.c_float * MAX_SURFACES),
        ("ior", (C.c_float * MAX_SURFACES) * MAX_LAMBDA), ("lambda_nm", C.c_float * MAX_LAMBDA),
        ("rgb_weight", (C.c_float * 3) * MAX_LAMBDA),
        ("entrance_half_height", C.c_double), ("stop_half_height", C.c_double),
        ("stop_half_height_neg", C.c_double),
    ]


class Light(C.Structure):
    """lfb_light"""
    _fields_ = [("ns_x", C.c_double), ("ns_y", C.c_double), ("theta", C.c_float), ("radiance", C.c_float * 3),
                ("distance", C.c_double)]


class Params(C.Structure):
    """lfb_params"""
    _fields_ = [
        ("mode", C.c_int32), ("pair_set", C.c_int32), ("include_direct", C.c_int32), ("grid_n", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32), ("precision", C.c_int32), ("splat", C.c_int32),
        ("fixed_point_bits", C.c_int32), ("physical_backward", C.c_int32),
        ("shard_index", C.c_int32), ("shard_count", C.c_int32),
        ("px_per_unit", C.c_float), ("physical_mapping", C.c_int32), ("reserved", C.c_float * 2),
    ]


class Options(C.Structure):
    """lfb_options (lfb_create_ex): zeros = defaults."""
    _fields_ = [
        ("struct_size", C.c_int32), ("stream_priority", C.c_int32), ("kernel_select", C.c_int32), ("family_split", C.c_int32),
        ("ctas_per_sm", C.c_int32), ("prefix_overlap", C.c_int32), ("starburst_lattice", C.c_int32), ("starburst_cache", C.c_int32),
        ("reduce_ctas", C.c_int32), ("collect_stats", C.c_int32), ("prefix_budget_bytes", C.c_int64), ("weights_table", C.c_int32), ("experiment", C.c_int32), ("host_write_mbps", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


def make_options(**fields):
    o = Options()
    o.struct_size = C.sizeof(Options)
    for k, v in fields.items():
        setattr(o, k, v)
    return o


class Scene(C.Structure):
    """lfb_scene: plain arrays (layouts in include/lfb200.h)."""
    _fields_ = [("tri_pos", C.POINTER(C.c_double)), ("tri_nrm", C.POINTER(C.c_double)), ("tri_mat", C.POINTER(C.c_int32)),
                ("n_tri", C.c_int32), ("n_sph", C.c_int32), ("spheres", C.POINTER(C.c_double)), ("sph_mat", C.POINTER(C.c_int32)),
                ("materials", C.POINTER(C.c_double)), ("lights", C.POINTER(C.c_double)), ("n_mat", C.c_int32), ("n_lights", C.c_int32)]


class Camera(C.Structure):
    """lfb_camera: the pinhole camera of Camera::generate_ray (camera.cpp:278-305)."""
    _fields_ = [("pos", C.c_double * 3), ("c2w", C.c_double * 9), ("hfov_deg", C.c_double), ("vfov_deg", C.c_double),
                ("nclip", C.c_double), ("fclip", C.c_double)]


def make_camera(cam16):
    """cam16 = pos xyz, c2w rows, hFov, vFov (degrees), nClip, fClip (tests/scene_fixtures.py, oracle/ref_shim.cpp)."""
    c = Camera()
    a = [float(v) for v in cam16]
    c.pos[:] = a[0:3]
    c.c2w[:] = a[3:12]
    c.hfov_deg, c.vfov_deg, c.nclip, c.fclip = a[12:16]
    return c


RAY_HIT_DTYPE = np.dtype([("x_s", "f8"), ("y_s", "f8"), ("x_ap", "f8"), ("y_ap", "f8"), ("px", "f8"),
                          ("py", "f8"), ("weight", "f8"), ("flags", "u4"), ("pad", "u4")])
REF_GHOST_DTYPE = np.dtype([("i", "i4"), ("j", "i4"), ("colour", "i4"), ("pad", "i4"), ("r1", "f8"),
                            ("r2", "f8"), ("verts", "f4", (4, 2)), ("scale", "f4"), ("shift", "f4")])


class LfbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"liblfb200 error {code}: {message}")
        self.code = code


def make_light(ns_x, ns_y, theta=None, radiance=(1.0, 1.0, 1.0), distance=0.0):
    """theta=None -> the reference's angle_to_sun = atan(ns_y/ns_x) (pathtracer.cpp:50).  distance > 0: a point light
    that many lens units in front of the first surface (0: directional, the reference's only flare source)."""
    lt = Light()
    lt.distance = distance
    lt.ns_x, lt.ns_y = ns_x, ns_y
    lt.theta = float(np.float32(np.arctan(ns_y / ns_x))) if theta is None else theta
    lt.radiance[:] = radiance
    return lt


def physical_theta(ns_x, ns_y, hfov_deg=50.0, vfov_deg=35.0):
    """Off-axis angle (radians) of a distant light seen at normalised screen position (ns_x, ns_y) by a camera with
    the given fields of view: the inverse of Camera::analyze_world_coord (camera.cpp:245-273).  The reference feeds
    its paraxial tracer atan(ns_y/ns_x) instead (pathtracer.cpp:50) -- a screen-space angle of up to 90 degrees that a
    linear ABCD model tolerates but real refraction does not (almost every ray is vignetted or totally reflected);
    the ray-grid modes therefore take the physical angle."""
    import math
    tx = (2.0 * ns_x - 1.0) * math.tan(0.5 * math.radians(hfov_deg))
    ty = (2.0 * ns_y - 1.0) * math.tan(0.5 * math.radians(vfov_deg))
    return float(np.float32(math.atan(math.hypot(tx, ty))))


def make_params(mode, width, height, grid_n=0, pair_set=PAIRS_REF, precision=FP32, splat=SPLAT_BILINEAR,
                include_direct=0, physical_backward=0, bits=0, px_per_unit=0.0, shard=(0, 0)):
    p = Params()
    p.mode, p.pair_set, p.include_direct, p.grid_n = mode, pair_set, include_direct, grid_n
    p.width, p.height, p.precision, p.splat = width, height, precision, splat
    p.fixed_point_bits, p.physical_backward = bits, physical_backward
    p.shard_index, p.shard_count = shard
    p.px_per_unit = px_per_unit
    return p


def copy_params(p, **changes):
    q = Params.from_buffer_copy(bytes(p))
    for k, v in changes.items():
        if k == "shard":
            q.shard_index, q.shard_count = v
        else:
            setattr(q, k, v)
    return q


def lights_array(lights):
    arr = (Light * max(len(lights), 1))()
    for k, lt in enumerate(lights):
        arr[k] = lt
    return arr


def build(verbose=False):
    """Compile liblfb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j4"] + ([] if verbose else ["-s"])
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib():
    """Load liblfb200.so (once).  Raises if the native library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                          f"`make -C {CSRC}`. The engine has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    LP, LiP, PP, vp = C.POINTER(Lens), C.POINTER(Light), C.POINTER(Params), C.c_void_p
    L.lfb_abi_version.restype = C.c_int
    L.lfb_create.argtypes = [C.POINTER(vp), C.c_int]
    L.lfb_create_ex.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Options)]
    L.lfb_exec_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.lfb_destroy.argtypes = [vp]
    L.lfb_destroy.restype = None
    L.lfb_last_error.restype = C.c_char_p
    L.lfb_builtin_lens.argtypes = [LP, C.c_int, C.c_float]
    L.lfb_set_lens.argtypes = [vp, LP]
    L.lfb_set_aperture.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.c_int]
    L.lfb_render_ghosts.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_size_t, C.c_int, C.c_int]
    L.lfb_render_ghosts_async.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_size_t, C.c_int]
    L.lfb_render_ghosts_rect.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
    L.lfb_render_ghosts_sparse.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.lfb_render_ghosts_sparse_begin.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_size_t, C.c_int, C.c_int, C.c_int]
    L.lfb_render_ghosts_sparse_end.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.lfb_sparse_slot_times.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.lfb_tile_state_bytes.argtypes = [C.c_int, C.c_int]
    L.lfb_tile_state_bytes.restype = C.c_size_t
    L.lfb_tile_stage_bytes.restype = C.c_size_t
    L.lfb_tile_stage_bytes.argtypes = [C.c_int, C.c_int, C.c_size_t]
    L.lfb_reduce_tiles_peers_staged.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, PP, vp, C.c_size_t, C.c_int, vp, vp]
    L.lfb_drain_tiles.argtypes = [vp, PP, vp, C.c_size_t, vp, vp]
    L.lfb_finalize_tiles_device.argtypes = [vp, vp, PP, vp, C.c_size_t, C.c_int, vp]
    L.lfb_reduce_tiles_peers.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, PP, vp, C.c_size_t, C.c_int, vp]
    L.lfb_create_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(Options)]
    L.lfb_destroy_multi.argtypes = [vp]
    L.lfb_destroy_multi.restype = None
    L.lfb_multi_set_lens.argtypes = [vp, LP]
    L.lfb_multi_set_aperture.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.c_int]
    L.lfb_render_ghosts_multi.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.lfb_multi_stats.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.lfb_set_scene.argtypes = [vp, C.POINTER(Scene)]
    L.lfb_render_scene.argtypes = [vp, C.POINTER(Camera), C.c_int, C.c_int, vp, C.c_size_t, C.c_int, C.c_int]
    L.lfb_render_composite_rgba8.argtypes = [vp, C.POINTER(Camera), LiP, C.c_int, PP, C.c_double, C.c_double, vp, C.c_int]
    L.lfb_dump_rays.argtypes = [vp, LiP, PP, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]
    L.lfb_ref_ghosts.argtypes = [vp, vp, C.c_int]
    L.lfb_accum_bytes.argtypes = [C.c_int, C.c_int]
    L.lfb_accum_bytes.restype = C.c_size_t
    L.lfb_stream.argtypes = [vp]
    L.lfb_stream.restype = vp
    L.lfb_render_ghosts_device.argtypes = [vp, LiP, C.c_int, PP, vp, C.c_int]
    L.lfb_finalize_device.argtypes = [vp, vp, PP, vp, C.c_size_t, C.c_int]
    L.lfb_finalize_clear_device.argtypes = [vp, vp, PP, vp, C.c_size_t, C.c_int]
    L.lfb_sync.argtypes = [vp]
    L.lfb_peer_barrier.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_uint64]
    L.lfb_reduce_finalize_peers.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, vp, PP, vp, C.c_size_t, C.c_int]
    L.lfb_count_work.argtypes = [LP, PP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.lfb_list_jobs.argtypes = [LP, PP, C.c_int, C.POINTER(C.c_int32), C.c_int]
    L.lfb_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.lfb_host_alloc.argtypes = [C.c_size_t]
    L.lfb_host_alloc.restype = vp
    L.lfb_host_free.argtypes = [vp]
    L.lfb_host_free.restype = None
    L.lfb_host_register.argtypes = [vp, C.c_size_t]
    L.lfb_host_register.restype = C.c_int
    L.lfb_host_unregister.argtypes = [vp]
    L.lfb_host_unregister.restype = C.c_int
    L.lfb_host_device_pointer.argtypes = [vp]
    L.lfb_host_device_pointer.restype = vp
    L.lfb_set_starburst_aperture.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.c_int]
    L.lfb_render_starburst.argtypes = [vp, LiP, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, C.c_size_t, C.c_int, C.c_int]
    L.lfb_render_frame_rgba8.argtypes = [vp, LiP, C.c_int, PP, C.c_double, C.c_double, vp, vp, C.c_int]
    L.lfb_probe_peaks.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    if L.lfb_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.lfb_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise LfbError(rc, lib().lfb_last_error().decode())
    return rc


def builtin_lens(n_lambda=3, coating_lambda0_nm=0.0):
    """The reference's hard-coded prescription (pathtracer.cpp:539-556) as an lfb_lens."""
    lens = Lens()
    check(lib().lfb_builtin_lens(C.byref(lens), n_lambda, coating_lambda0_nm))
    return lens


def count_work(lens, params, n_lights):
    rays, inter, jobs = C.c_double(), C.c_double(), C.c_int()
    check(lib().lfb_count_work(C.byref(lens), C.byref(params), n_lights, C.byref(rays), C.byref(inter), C.byref(jobs)))
    return rays.value, inter.value, jobs.value


def list_jobs(lens, params, n_lights):
    n = check(lib().lfb_list_jobs(C.byref(lens), C.byref(params), n_lights, None, 0))
    out = np.zeros((max(n, 1), 4), np.int32)
    check(lib().lfb_list_jobs(C.byref(lens), C.byref(params), n_lights, out.ctypes.data_as(C.POINTER(C.c_int32)), n))
    return out[:n]


class PinnedArray:
    """A numpy view of cudaHostAlloc'ed memory (lfb_host_alloc)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().lfb_host_alloc(self.nbytes)
        if not self.ptr:
            raise LfbError(ERR_NOMEM, "lfb_host_alloc failed")
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().lfb_host_free(self.ptr)
            self.ptr = None


class MultiEngine:
    """Owner of one lfb_multi: several GPUs of one node driven by THIS process and thread (lfb_create_multi)."""

    def __init__(self, device_ids, **options):
        self._h = C.c_void_p()
        ids = (C.c_int * len(device_ids))(*device_ids)
        check(lib().lfb_create_multi(C.byref(self._h), ids, len(device_ids), C.byref(make_options(**options)) if options else None))

    def close(self):
        if self._h:
            lib().lfb_destroy_multi(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_lens(self, lens):
        check(lib().lfb_multi_set_lens(self._h, C.byref(lens)))

    def set_aperture(self, texels):
        tex = np.ascontiguousarray(texels, np.float32)
        check(lib().lfb_multi_set_aperture(self._h, tex.ctypes.data_as(C.POINTER(C.c_float)), tex.shape[1], tex.shape[0]))

    def render_ghosts(self, lights, params, out, elem=F64x3, stride=None, out_is_clear=False):
        """Tile-sparse frame into `out` (page-locked host array kept between frames), sharded over the devices; returns the
        number of tiles written (-1: pageable `out`, full-frame copy)."""
        if stride is None:
            stride = out.strides[1]
        n = C.c_int()
        check(lib().lfb_render_ghosts_multi(self._h, lights_array(lights), len(lights), C.byref(params), out.ctypes.data, stride, elem,
                                            int(out_is_clear), C.byref(n)))
        return n.value

    def stats(self):
        ms, n = (C.c_float * 4)(), C.c_int()
        check(lib().lfb_multi_stats(self._h, ms, C.byref(n)))
        return dict(trace_ms=ms[0], reduce_ms=ms[1], call_ms=ms[2], enqueue_ms=ms[3], n_devices=n.value)


class Engine:
    """Owner of one lfb_engine (one CUDA device)."""

    def __init__(self, device_id=-1, **options):
        """options: fields of lfb_options (stream_priority=1, kernel_select=2, collect_stats=1, ...)."""
        self._h = C.c_void_p()
        if options:
            check(lib().lfb_create_ex(C.byref(self._h), device_id, C.byref(make_options(**options))))
        else:
            check(lib().lfb_create(C.byref(self._h), device_id))
        self.lens = None

    def close(self):
        if self._h:
            lib().lfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_lens(self, lens):
        check(lib().lfb_set_lens(self._h, C.byref(lens)))
        self.lens = lens

    def set_aperture(self, texels):
        tex = np.ascontiguousarray(texels, np.float32)
        check(lib().lfb_set_aperture(self._h, tex.ctypes.data_as(C.POINTER(C.c_float)), tex.shape[1], tex.shape[0]))

    def render_ghosts(self, lights, params, out=None, elem=F64x3, stride=None, additive=False):
        """-> (H, W, 3) array (float64 for F64x3, float32 for F32x3); `out` may be a caller buffer."""
        dt = np.float64 if elem == F64x3 else np.float32
        if out is None:
            out = np.empty((params.height, params.width, 3), dt)
            stride = out.strides[1]
        elif stride is None:
            stride = out.strides[1] if out.ndim == 3 else (24 if elem == F64x3 else 12)
        check(lib().lfb_render_ghosts(self._h, lights_array(lights), len(lights), C.byref(params),
                                      out.ctypes.data, stride, elem, int(additive)))
        return out

    def render_ghosts_async(self, lights, params, out, elem=F64x3, stride=None):
        """Enqueue a frame into `out` (pinned host array) without waiting; complete after sync()."""
        if stride is None:
            stride = out.strides[1]
        check(lib().lfb_render_ghosts_async(self._h, lights_array(lights), len(lights), C.byref(params), out.ctypes.data, stride, elem))

    def render_ghosts_rect(self, lights, params, out, elem=F64x3, stride=None):
        """Dirty-rectangle form: writes only the bounding rectangle of the frame's deposits into `out` (which the
        caller keeps clear elsewhere) and returns it as (x0, y0, x1, y1) inclusive, or None for an empty frame."""
        if stride is None:
            stride = out.strides[1]
        rect = (C.c_int * 4)()
        check(lib().lfb_render_ghosts_rect(self._h, lights_array(lights), len(lights), C.byref(params),
                                           out.ctypes.data, stride, elem, rect))
        return None if rect[2] < rect[0] else tuple(rect)

    def render_ghosts_sparse(self, lights, params, out, elem=F64x3, stride=None, out_is_clear=False):
        """Tile-sparse frame into `out` (page-locked host array the caller keeps between frames); returns the number of
        16 x 16 tiles written (-1: `out` is pageable and the full frame was copied)."""
        if stride is None:
            stride = out.strides[1]
        n = C.c_int()
        check(lib().lfb_render_ghosts_sparse(self._h, lights_array(lights), len(lights), C.byref(params), out.ctypes.data, stride, elem,
                                             int(out_is_clear), C.byref(n)))
        return n.value

    def render_ghosts_sparse_begin(self, lights, params, out, slot, elem=F64x3, stride=None, out_is_clear=False):
        """Enqueue a tile-sparse frame into `out` (page-locked, the slot's own buffer) and return; collect it with
        render_ghosts_sparse_end(slot).  Two slots: the tile writes of one frame overlap the trace of the next."""
        if stride is None:
            stride = out.strides[1]
        check(lib().lfb_render_ghosts_sparse_begin(self._h, lights_array(lights), len(lights), C.byref(params), out.ctypes.data, stride, elem,
                                                   int(out_is_clear), int(slot)))

    def render_ghosts_sparse_end(self, slot):
        n = C.c_int()
        check(lib().lfb_render_ghosts_sparse_end(self._h, int(slot), C.byref(n)))
        return n.value

    def sparse_slot_times(self, slot):
        """(reached, traced, tile kernel start, staged, done) of the slot's last collected frame, ms since the engine was created."""
        ms = (C.c_float * 5)()
        check(lib().lfb_sparse_slot_times(self._h, int(slot), ms))
        return tuple(ms)

    def finalize_tiles_device(self, accum_ptr, params, out_ptr, stride, elem, state_ptr):
        check(lib().lfb_finalize_tiles_device(self._h, accum_ptr, C.byref(params), out_ptr, stride, elem, state_ptr))

    def reduce_tiles_peers_staged(self, accum_ptrs, rank, params, out_ptr, stride, elem, state_ptr, stage_ptr):
        arr = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        check(lib().lfb_reduce_tiles_peers_staged(self._h, arr, len(accum_ptrs), rank, C.byref(params), out_ptr, stride, elem, state_ptr, stage_ptr))

    def drain_tiles(self, params, out_ptr, stride, state_ptr, stage_ptr):
        check(lib().lfb_drain_tiles(self._h, C.byref(params), out_ptr, stride, state_ptr, stage_ptr))

    def reduce_tiles_peers(self, accum_ptrs, rank, params, out_ptr, stride, elem, state_ptr):
        arr = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        check(lib().lfb_reduce_tiles_peers(self._h, arr, len(accum_ptrs), rank, C.byref(params), out_ptr, stride, elem, state_ptr))

    def set_scene(self, scene):
        """scene: dict of arrays (tri_pos [n,3,3], tri_nrm [n,3,3], tri_mat [n], spheres [m,4], sph_mat [m], mats [k,6], lights [l,7])."""
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        tp = np.ascontiguousarray(scene["tri_pos"], np.float64)
        tn = np.ascontiguousarray(scene["tri_nrm"], np.float64)
        tm = np.ascontiguousarray(scene["tri_mat"], np.int32)
        sp = np.ascontiguousarray(scene["spheres"], np.float64).reshape(-1, 4)
        sm = np.ascontiguousarray(scene["sph_mat"], np.int32)
        ma = np.ascontiguousarray(scene["mats"], np.float64).reshape(-1, 6)
        li = np.ascontiguousarray(scene["lights"], np.float64).reshape(-1, 7)
        sc = Scene(tp.ctypes.data_as(dp), tn.ctypes.data_as(dp), tm.ctypes.data_as(ip), tm.size, sm.size, sp.ctypes.data_as(dp), sm.ctypes.data_as(ip),
                   ma.ctypes.data_as(dp), li.ctypes.data_as(dp), ma.shape[0], li.shape[0])
        check(lib().lfb_set_scene(self._h, C.byref(sc)))

    def render_scene(self, camera, width, height, out=None, elem=F64x3, additive=False):
        """The path-traced scene pass (emission + direct lighting) for the whole frame: (H, W, 3)."""
        if out is None:
            out = np.zeros((height, width, 3), np.float64 if elem == F64x3 else np.float32)
        check(lib().lfb_render_scene(self._h, C.byref(camera), width, height, out.ctypes.data, out.strides[1], elem, int(additive)))
        return out

    def render_composite_rgba8(self, camera, lights, params, flare_radius=-1.0, flare_intensity=1.0, out=None, flip=False):
        """scene pass + ghosts + [starburst] -> toColor -> (H, W) uint32: BASELINE config 5 in one call."""
        if out is None:
            out = np.empty((params.height, params.width), np.uint32)
        check(lib().lfb_render_composite_rgba8(self._h, C.byref(camera), lights_array(lights), len(lights), C.byref(params), flare_radius,
                                               flare_intensity, out.ctypes.data, int(flip)))
        return out

    def set_starburst_aperture(self, texels):
        tex = np.ascontiguousarray(texels, np.float32)
        check(lib().lfb_set_starburst_aperture(self._h, tex.ctypes.data_as(C.POINTER(C.c_float)), tex.shape[1], tex.shape[0]))

    def render_starburst(self, lights, width, height, flare_radius, flare_intensity, out=None, elem=F64x3, additive=False):
        """The whole frame of PathTracer::raytrace_starburst: (H, W, 3)."""
        if out is None:
            out = np.empty((height, width, 3), np.float64 if elem == F64x3 else np.float32)
        check(lib().lfb_render_starburst(self._h, lights_array(lights), len(lights), width, height, flare_radius, flare_intensity,
                                         out.ctypes.data, out.strides[1], elem, int(additive)))
        return out

    def render_frame_rgba8(self, lights, params, flare_radius=-1.0, flare_intensity=1.0, base_hdr=None, out=None, flip=False):
        """[base] + ghosts + [starburst] -> toColor -> (H, W) uint32 0xFFBBGGRR."""
        if out is None:
            out = np.empty((params.height, params.width), np.uint32)
        base = None if base_hdr is None else np.ascontiguousarray(base_hdr, np.float64)
        check(lib().lfb_render_frame_rgba8(self._h, lights_array(lights), len(lights), C.byref(params), flare_radius, flare_intensity,
                                           None if base is None else base.ctypes.data, out.ctypes.data, int(flip)))
        return out

    def dump_rays(self, light, params, i, j, lam):
        out = np.zeros(params.grid_n * params.grid_n, RAY_HIT_DTYPE)
        check(lib().lfb_dump_rays(self._h, C.byref(light), C.byref(params), i, j, lam, out.ctypes.data, out.size))
        return out

    def ref_ghosts(self):
        out = np.zeros(64 * MAX_LAMBDA, REF_GHOST_DTYPE)
        n = check(lib().lfb_ref_ghosts(self._h, out.ctypes.data, out.size))
        return out[:n].copy()

    @property
    def stream(self):
        return lib().lfb_stream(self._h)

    def render_ghosts_device(self, lights, params, accum_ptr, clear_first=True):
        check(lib().lfb_render_ghosts_device(self._h, lights_array(lights), len(lights), C.byref(params),
                                             accum_ptr, int(clear_first)))

    def finalize_device(self, accum_ptr, params, out_ptr, stride, elem, clear=False):
        """accum -> pixels on the engine stream; clear=True also leaves the accumulators zeroed (one kernel)."""
        fn = lib().lfb_finalize_clear_device if clear else lib().lfb_finalize_device
        check(fn(self._h, accum_ptr, C.byref(params), out_ptr, stride, elem))

    def peer_barrier(self, flag_ptrs, rank, epoch):
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        check(lib().lfb_peer_barrier(self._h, arr, len(flag_ptrs), rank, epoch))

    def reduce_finalize_peers(self, accum_ptrs, rank, params, out_ptr, stride, elem, multicast_ptr=None):
        arr = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        check(lib().lfb_reduce_finalize_peers(self._h, arr, len(accum_ptrs), rank, multicast_ptr, C.byref(params), out_ptr, stride, elem))

    def sync(self):
        check(lib().lfb_sync(self._h))

    def stats(self):
        n, t, f = C.c_uint64(), C.c_float(), C.c_float()
        check(lib().lfb_stats(self._h, C.byref(n), C.byref(t), C.byref(f)))
        return dict(kernel_launches=n.value, last_trace_ms=t.value, last_frame_ms=f.value)

    def exec_stats(self):
        """Executed work of the last EXACT_GRID frame (engine created with collect_stats=1)."""
        out = (C.c_uint64 * 4)()
        check(lib().lfb_exec_stats(self._h, out))
        return dict(steps=out[0], ray_pairs_started=out[1], ray_pairs_landed=out[2], families=bool(out[3]))

    def probe_peaks(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(lib().lfb_probe_peaks(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(fp32_flops=a.value, mufu_ops=b.value, sm_clock_hz=c.value)
