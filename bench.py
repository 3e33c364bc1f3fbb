#!/usr/bin/env python
"""bench.py -- the headline benchmark of the lens-flare ghost path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our sm_100a engine
  python bench.py --impl reference [--gpus N] [--steps K] ...     the reference's CPU code (oracle/_ref)
  N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A STEP is one flare frame of BASELINE config 2 per light: RGB (3 wavelengths), a 256x256 ray grid per ghost, every
two-reflection ghost of the built-in lens (28 glass-glass pairs) plus the direct path, 1920x1080 sensor,
final_apertures/pentbig500_14.png at the stop, EXACT_GRID physics (sphere/plane intersection, vector Snell with n(lambda),
Fresnel + quarter-wave coating at 550 nm, aperture lookup, bilinear fixed-point splat), FP32.  With N GPUs the frame has N
suns (weak scaling: one light's worth of ghosts per GPU); the (light x pair x wavelength) jobs are dealt to the ranks, every
rank splats into its own u64 accumulators, and the frame is summed TILE-SPARSE over NVLink peer memory: each rank reduces +
converts its interleaved share of the dirty 16x16 tiles and stores the pixels into rank 0's frame (no collective call; --reduce
nccl / peer select the dense round-1 paths for comparison).

  value     ray-surface interactions / s (NOMINAL: rays x I(i,j), the unit BASELINE.json defines), whole job, frame
            description resident in HBM, device-timed: B brackets of EXACTLY K back-to-back frames each (CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks per bracket); the MEDIAN bracket is reported,
            all of them are listed
  e2e       the same metric through the reference-facing C-ABI call with HOST buffers every step: the aperture mask goes
            host -> device and the Vector3D[] frame comes back (lfb_render_ghosts_sparse: the device writes the frame's dirty
            tiles straight into the caller's page-locked HDRImageBuffer storage and re-zeroes the previous frame's)
  strict    the same pipeline with LFB_STRICT (FP64 geometry): what north_star's 1e-5-lens-unit per-ray bar costs
  parity    this run's GPU frames / rays against the double-precision oracle (and, N > 1, the sharded frame against the
            unsharded one)
"""
import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# declared algorithmic cost of one ray-surface interaction in EXACT_GRID
FLOP_PER_INTERACTION = 64.0
MUFU_PER_INTERACTION = 3.0
GRID_N, WIDTH, HEIGHT = 256, 1920, 1080
COATING_NM = 550.0
N_BRACKETS = 7
WORKLOAD = ("cfg2: RGB (3 wavelengths) x 256x256 ray grid per ghost x (28 two-reflection ghost pairs + direct path) per light, "
            "1920x1080 sensor, pentbig500_14 aperture, EXACT_GRID (sphere/plane hit, Snell, Fresnel + 550 nm quarter-wave coating), "
            "bilinear fixed-point splat")


def physical_theta(ns_x, ns_y, hfov_deg=50.0, vfov_deg=35.0):
    """Off-axis angle of a distant light at normalised screen position (ns_x, ns_y): the inverse of
    Camera::analyze_world_coord (camera.cpp:245-273); float32 like lfb_light.theta.  (Same formula as capi.physical_theta;
    restated here so that the reference arm does not load the product library.)"""
    tx = (2.0 * ns_x - 1.0) * math.tan(0.5 * math.radians(hfov_deg))
    ty = (2.0 * ns_y - 1.0) * math.tan(0.5 * math.radians(vfov_deg))
    return float(np.float32(math.atan(math.hypot(tx, ty))))


def nominal_interactions(n_lights, include_direct=True, grid_n=GRID_N, n_lambda=3, n_surfaces=9, stop=5):
    """I(i, j) = 2 (j - i) + n + 1 per ray over all glass-glass pairs (478 for the built-in lens), n + 1 for the
    direct path; x rays per ghost x wavelengths x lights.  Pure Python (lfb_count_work computes the same)."""
    glass = [k for k in range(n_surfaces) if k != stop]
    per_ray = sum(2 * (j - i) + n_surfaces + 1 for a, i in enumerate(glass) for j in glass[a + 1:])
    if include_direct:
        per_ray += n_surfaces + 1
    return float(per_ray) * grid_n * grid_n * n_lambda * n_lights


def sun_positions(n):
    """Deterministic suns.  The first is SURVEY 8d's (0.45, 0.55); the others sit at the SAME off-axis angle (0.0562 rad
    through the 50 x 35 degree camera) at azimuths 2 pi k / n around the optical axis, so that every light is the same
    amount of work (weak scaling measures the system, not the lights) and none sits at the exact screen centre (the
    reference NaNs there, pathtracer.cpp:414)."""
    ex, ey = math.tan(math.radians(25.0)), math.tan(math.radians(17.5))
    tx0, ty0 = (2 * 0.45 - 1) * ex, (2 * 0.55 - 1) * ey
    rad, phi0 = math.hypot(tx0, ty0), math.atan2(ty0, tx0)
    pts = [(0.45, 0.55)]
    for k in range(1, n):
        phi = phi0 + 2 * math.pi * k / n
        pts.append((0.5 * (rad * math.cos(phi) / ex + 1), 0.5 * (rad * math.sin(phi) / ey + 1)))
    return pts


def make_sun(x, y, **kw):
    """A sun at normalised screen position (x, y) seen through a 50 x 35 degree camera: the physical off-axis angle,
    not the reference's screen-space atan(y/x) that kills every exactly-traced ray."""
    from lens_flare_b200 import capi
    return capi.make_light(x, y, theta=physical_theta(x, y), **kw)


def load_aperture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "apertures.npz"))
    return z["pentbig500_14"].astype(np.float32) * np.float32(1.0 / 255.0)


class ClockSampler:
    """Polls SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as exc:  # NVML missing: report that instead of inventing clocks
            self.nv, self.err = None, str(exc)

    def _run(self):
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def report(self):
        self._stop.set()
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + self.err}
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def spread(xs):
    xs = sorted(xs)
    return {"median": statistics.median(xs), "min": xs[0], "max": xs[-1], "n": len(xs)}


# ------------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU code for this path, all host threads.  It does NOT load the product
# library: the work count and the light's angle are computed in Python above.
# ------------------------------------------------------------------------------------------------
def reference_trace_step(ref, theta, threads):
    """One step of the reference arm: trace_ray_auto_before/after (pathtracer.cpp:588-689) called per ray
    and per axis over cfg2's grid (256^2 x 28 pairs x RGB), std::threads over ghosts.  -> seconds"""
    s, _, _ = ref.time_trace_grid(GRID_N, theta, 3, 28, threads)
    return s


def load_ref_oracle():
    """oracle/_ref through ctypes only (oracle.bindings imports the product's struct declarations, which this arm avoids)."""
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_time_trace_grid.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ref_time_trace_grid.restype = C.c_double

    class Ref:
        def time_trace_grid(self, n, theta, ncol, pairset, nthreads):
            rays, chk = C.c_double(), C.c_double()
            s = L.ref_time_trace_grid(n, theta, ncol, pairset, nthreads, C.byref(rays), C.byref(chk))
            return s, rays.value, chk.value
    return Ref()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    inter = nominal_interactions(1, include_direct=False)  # the reference has no direct path
    theta = physical_theta(0.45, 0.55)
    ref = load_ref_oracle()
    if ref is not None:
        kind = "reference"
        step = lambda: reference_trace_step(ref, theta, threads)  # noqa: E731
        sample = (f"one step = trace_ray_auto_before/after per ray and axis over 256^2 x 28 pairs x RGB "
                  f"({inter:.3g} interactions), {threads} std::threads")
    else:  # the compiled reference is missing: the oracle port's paraxial frame (the other place bench.py may execute oracle/)
        from lens_flare_b200 import capi
        from oracle import bindings as ob
        if not os.path.exists(ob.PORT_SO):
            ob.build(("port",))
        port, tex, lens = ob.PortOracle(), load_aperture(), capi.builtin_lens(3)
        kind = "port"
        pp = capi.make_params(capi.MODE_PARAXIAL_GRID, WIDTH, HEIGHT, grid_n=GRID_N, pair_set=capi.PAIRS_ALL, precision=capi.FP64)
        step = lambda: port.time_render(lens, tex, [make_sun(0.45, 0.55)], pp, threads)[0]  # noqa: E731
        sample = f"one step = oracle port PARAXIAL_GRID frame, 256^2 x 28 pairs x RGB, {threads} pthreads"
    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = inter * args.steps / total
    line = {
        "impl": "reference", "metric": "ray_surface_interactions_per_s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD + " [reference arm: the reference's paraxial ABCD tracer on the same grid; it has no "
                                          "exact physics and no direct path -- the same-physics CPU number is cpu_baseline.port_exact "
                                          "of our arm's line]"},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": threads, "kind": kind, "sample": sample, "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "repo_libraries_loaded": sorted({l.split()[-1][len(ROOT) + 1:] for l in open("/proc/self/maps") if ROOT in l and ".so" in l}),
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# CPU legs reported beside our line (rank 0, N = 1)
# ------------------------------------------------------------------------------------------------
def cpu_baseline_block(port_seconds, inter_frame, sample_steps=1):
    """kind 'reference' = the reference's trace_ray_auto_* per ray (oracle/_ref); port_exact = the oracle's EXACT_GRID frame
    (same physics as our kernels; its seconds come from the parity leg, which renders that frame anyway)."""
    from lens_flare_b200 import capi
    from oracle import bindings as ob
    threads = os.cpu_count() or 1
    inter_pairs = nominal_interactions(1, include_direct=False)
    theta = physical_theta(0.45, 0.55)
    out = {}
    if os.path.exists(ob.REF_SO):
        ref = ob.RefOracle()
        best = min(reference_trace_step(ref, theta, threads) for _ in range(max(1, sample_steps) + 1))
        out.update(value=inter_pairs / best, unit="interactions/s", cores=threads, kind="reference",
                   sample=f"trace_ray_auto_before/after per ray and axis over cfg2's grid (256^2 x 28 pairs x RGB = "
                          f"{inter_pairs:.3g} interactions), {threads} std::threads, best of {max(1, sample_steps) + 1}",
                   seconds=best)
        s1, _ = ref.time_ghost_buffer(load_aperture(), WIDTH, HEIGHT, 0.45, 0.55, capi.make_light(0.45, 0.55).theta, 5)
        out["reference_frame_ms_1thread"] = s1 * 1e3  # PathTracer::generate_ghost_buffer, 1080p (13 quads x RGB)
    out["cpu_model"] = cpu_model()
    exact = dict(value=inter_frame / port_seconds, unit="interactions/s", cores=threads, kind="port",
                 sample=f"oracle/lf_oracle.c EXACT_GRID, one full cfg2 frame ({inter_frame:.3g} interactions), {threads} pthreads",
                 seconds=port_seconds, frame_ms=port_seconds * 1e3)
    if out.get("kind"):
        out["port_exact"] = exact
    else:
        out.update(exact)
    return out


def parity_block(eng, lens, tex, light, params):
    """This run's GPU frames and rays against the double-precision oracle, at cfg2's full size.
    image: relative L2 of the FP32 and STRICT frames against oracle/lf_oracle.c's frame.
    per ray: sensor hit positions (lens units) of every ray of all 28 ghosts + the direct path at the green wavelength
    (29 x 65 536 rays) against lfo_trace_grid."""
    from lens_flare_b200 import capi
    from oracle import bindings as ob
    if not os.path.exists(ob.PORT_SO):
        ob.build(("port",))
    port = ob.PortOracle()
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    port_seconds, _ = port.time_render(lens, tex, [light], params, threads)  # the timed multi-threaded frame (cpu_baseline.port_exact)
    want = port.render(lens, tex, [light], params)                           # the frame itself (single thread)
    oracle_s = time.perf_counter() - t0
    out = {"oracle": "oracle/lf_oracle.c EXACT_GRID (double); parity of this physics is UNPINNED by the reference, which has none (SURVEY 8a a13)",
           "oracle_seconds": oracle_s}
    norm = float(np.sqrt((want ** 2).sum()))
    for name, prec in (("fp32", capi.FP32), ("strict", capi.STRICT)):
        got = eng.render_ghosts([light], capi.copy_params(params, precision=prec))
        out[name] = {"image_rel_l2": float(np.sqrt(((got - want) ** 2).sum()) / norm), "nonzero_pixels": int((got.sum(axis=2) > 0).sum())}
    pairs = [(i, j) for i in range(9) for j in range(i + 1, 9) if i != 5 and j != 5] + [(-1, -1)]
    geo = ob.RAY_MISSED | ob.RAY_VIGNETTED | ob.RAY_TIR
    stats = {"fp32": [], "strict": []}
    mism = {"fp32": 0, "strict": 0}
    n_rays = 0
    for (i, j) in pairs:
        w = port.trace_grid(lens, tex, light, params, i, j, 1)
        n_rays += w.size
        for name, prec in (("fp32", capi.FP32), ("strict", capi.STRICT)):
            g = eng.dump_rays(light, capi.copy_params(params, precision=prec), i, j, 1)
            same = (g["flags"] & geo) == (w["flags"] & geo)
            mism[name] += int((~same).sum())
            ok = same & ~np.isnan(w["x_s"]) & ~np.isnan(g["x_s"])
            stats[name].append(np.hypot(g["x_s"][ok] - w["x_s"][ok], g["y_s"][ok] - w["y_s"][ok]))
    for name in stats:
        d = np.concatenate(stats[name])
        out[name]["per_ray_lens_units"] = {"p50": float(np.median(d)), "p99": float(np.quantile(d, 0.99)), "max": float(d.max()), "rays_compared": int(d.size)}
        out[name]["flags_mismatch"] = mism[name] / n_rays
    out["rays"] = n_rays
    out["tolerances"] = "north_star: per-ray 1e-5 lens units, image 1e-3 relative L2"
    return out, port_seconds


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def dirty_quadrants(*frames):
    """8 x 8 sensor quadrants that hold a non-zero pixel in ANY of the (H, W, 3) frames: what the tile kernels write (a frame's own
    non-zero quadrants plus those the output buffer's previous frame left)."""
    nz = None
    for f in frames:
        m = f.sum(axis=2) != 0
        H, W = m.shape
        pad = np.zeros(((H + 7) // 8 * 8, (W + 7) // 8 * 8), bool)
        pad[:H, :W] = m
        q = pad.reshape(pad.shape[0] // 8, 8, pad.shape[1] // 8, 8).any(axis=(1, 3))
        nz = q if nz is None else (nz | q)
    return int(nz.sum())


class SparsePipeline:
    """N = 1: R rotating (accumulators, output frame, tile state) sets on TWO streams:
        stream A (engine)                       trace frame k into accum[k % R] (clear_first = 0: the finalize left it clear)
        stream B (finalize engine, high prio)   lfb_finalize_tiles_device of frame k -> out[k % R]
    so the tile kernel of frame k runs under the trace of frame k+1 (which also has the engine's own third stream running its
    forward sweeps).  Stage B of frame k waits for stage A of frame k; stage A of frame k + R waits for stage B of frame k."""

    def __init__(self, eng, fin, params, dev, n_buffers, torch, capi):
        self.eng, self.fin, self.params, self.capi, self.torch = eng, fin, params, capi, torch
        W, H = params.width, params.height
        acc_words = (capi.lib().lfb_accum_bytes(W, H) + 7) // 8
        st_words = (capi.lib().lfb_tile_state_bytes(W, H) + 3) // 4
        self.accums = [torch.zeros((acc_words,), dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.outs = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev) for _ in range(n_buffers)]
        self.states = [torch.zeros((st_words,), dtype=torch.int32, device=dev) for _ in range(n_buffers)]
        self.A = torch.cuda.ExternalStream(eng.stream, device=dev)
        self.B = torch.cuda.ExternalStream(fin.stream, device=dev)
        self.traced = [torch.cuda.Event() for _ in range(n_buffers)]
        self.finalized = [torch.cuda.Event() for _ in range(n_buffers)]
        self.fin_valid = [False] * n_buffers
        self.k = 0
        torch.cuda.synchronize(dev)

    def frame(self, lights):
        b = self.k % len(self.accums)
        self.k += 1
        if self.fin_valid[b]:
            self.A.wait_event(self.finalized[b])  # the buffer's previous frame has been converted (and the buffer left clear)
        self.eng.render_ghosts_device(lights, self.params, self.accums[b].data_ptr(), clear_first=False)
        self.traced[b].record(self.A)
        self.B.wait_event(self.traced[b])
        out = self.outs[b]
        self.fin.finalize_tiles_device(self.accums[b].data_ptr(), self.params, out.data_ptr(), out.stride(1) * out.element_size(), self.capi.F32x3,
                                       self.states[b].data_ptr())
        self.finalized[b].record(self.B)
        self.fin_valid[b] = True
        return b

    def finish(self):
        self.A.wait_stream(self.B)  # the timing events are recorded on A


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lens_flare_b200 import capi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus} (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side waits that must not occupy the GPUs

    K, W_ = args.steps, max(args.warmup, 3)
    tex = load_aperture()
    lens = capi.builtin_lens(3, COATING_NM)
    eng = capi.Engine(local)
    eng.set_lens(lens)
    eng.set_aperture(tex)
    n_lights = world
    params = capi.make_params(capi.MODE_EXACT_GRID, WIDTH, HEIGHT, grid_n=GRID_N, pair_set=capi.PAIRS_ALL, include_direct=1,
                              precision=capi.FP32, splat=capi.SPLAT_BILINEAR)
    lights_a = [make_sun(x, y) for x, y in sun_positions(n_lights)]
    lights_b = [make_sun(1.0 - x, 1.0 - y) for x, y in sun_positions(n_lights)]  # e2e alternates frames (same off-axis angle)
    rays_frame, inter_frame, jobs_frame = capi.count_work(lens, params, n_lights)
    assert inter_frame == nominal_interactions(n_lights), "bench.py's nominal count disagrees with lfb_count_work"
    N_BUF = 3  # rotating accumulator / output / state sets
    fin = capi.Engine(local, stream_priority=args.fin_priority, reduce_ctas=args.reduce_ctas)  # the tile finalize / reduce runs on a second, high-priority stream
    fin.set_lens(lens)
    fin.set_aperture(tex)
    A = torch.cuda.ExternalStream(eng.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build_pipeline(p):
        """-> (frames(n), finish(), result(b))"""
        if world == 1:
            sp = SparsePipeline(eng, fin, p, dev, N_BUF, torch, capi)

            def frames(n, lights=lights_a):
                b = 0
                for _ in range(n):
                    b = sp.frame(lights)
                return b
            return frames, sp.finish, (lambda b: sp.outs[b]), sp
        if args.reduce == "sparse":
            ps = sharding.PeerSparse(eng, p, rank, world, dev, dist.group.WORLD, n_buffers=N_BUF, finalize_engine=fin)

            def frames(n, lights=lights_a):
                b = 0
                ps.begin()
                for _ in range(n):
                    b = ps.frame(lights, owner=0)
                return b
            return frames, ps.finish, ps.result, ps
        if args.reduce == "nccl":
            sh = sharding.ShardedFlare(eng, p, rank, world, dev, n_buffers=N_BUF, finalize_engine=fin)
            outs = [torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32, device=dev) for _ in range(N_BUF)]

            def frames(n, lights=lights_a):
                sh.begin()
                for k in range(n):
                    sh.frame(lights, out=outs[k % N_BUF], elem=capi.F32x3, reduce_dst=0)
                return (n - 1) % N_BUF
            return frames, sh.join, (lambda b: outs[b]), sh
        pf = sharding.PeerFlare(eng, p, rank, world, dev, dist.group.WORLD, n_buffers=N_BUF, use_multicast=(args.reduce == "multicast"),
                                finalize_engine=fin)
        if args.reduce == "multicast" and not pf.mc:
            raise SystemExit("--reduce multicast: the symmetric allocation has no NVSwitch multicast binding on this box")

        def frames(n, lights=lights_a):
            b = 0
            pf.begin()
            for _ in range(n):
                b = pf.frame(lights, owner=0)
            return b
        return frames, pf.finish, pf.result, pf

    def timed_brackets(frames, finish, n_brackets, K=K, lights=None):
        """n_brackets x (EXACTLY K frames between a barrier + synchronize on both sides); per bracket the device time of the
        engine stream's events, max over ranks.  -> (list of ms per bracket, host enqueue ms per frame)"""
        out, enq = [], []
        for _ in range(n_brackets):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev_stream = A if world == 1 else torch.cuda.current_stream(dev)  # N = 1: everything runs on the engine's stream
            e0.record(ev_stream)
            th0 = time.perf_counter()
            if lights is None:
                frames(K)
            else:
                frames(K, lights)
            finish()  # N > 1: makes torch's current stream wait for the pipeline's streams
            enq.append((time.perf_counter() - th0) * 1e3 / K)
            e1.record(ev_stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.append(float(t[0]))
        return out, statistics.median(enq)

    clocks = ClockSampler(local)
    frames, finish, result, pipe = build_pipeline(params)
    if world == 1:
        A.wait_stream(torch.cuda.current_stream(dev))
    frames(W_)
    finish()
    barrier()
    # kernels of ours per frame, counted on one frame
    l0 = eng.stats()["kernel_launches"] + (fin.stats()["kernel_launches"])
    frames(1)
    finish()
    barrier()
    launches_per_frame = eng.stats()["kernel_launches"] + (fin.stats()["kernel_launches"]) - l0
    with clocks:
        brackets_ms, host_enqueue_ms = timed_brackets(frames, finish, N_BRACKETS)
    dev_ms = statistics.median(brackets_ms)
    value = inter_frame * K / (dev_ms * 1e-3)

    # ---- N > 1: the sharded frame against the unsharded one (outside the timed region) --------------------------------
    parity = {}
    if world > 1:
        b = frames(1)
        finish()
        barrier()
        same = None
        if rank == 0:
            whole = torch.from_numpy(eng.render_ghosts(lights_a, params, elem=capi.F32x3)).to(dev)
            same = bool(torch.equal(result(b), whole))
        parity["equals_single_gpu"] = same
        parity["equals_single_gpu_how"] = ("rank 0 renders the same %d-light frame unsharded (lfb_render_ghosts) and compares it bit for bit "
                                           "with the frame the %d ranks reduced" % (n_lights, world))
        barrier()

    # ---- the dominant kernel's own duration: the engine's CUDA events around the trace kernels, one frame at a time, L2 flushed
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    trace_ms = []
    for _ in range(12):
        flush.zero_()
        barrier()
        frames(1)
        finish()
        barrier()
        trace_ms.append(eng.stats()["last_trace_ms"])
    del flush
    trace_ms = trace_ms[2:]
    _, inter_rank, jobs_rank = capi.count_work(lens, capi.copy_params(params, shard=(rank, world) if world > 1 else (0, 0)), n_lights)

    # ---- what the kernels executed (the counting build, one frame, outside every timed region) ---------------------------
    executed = None
    if rank == 0:
        se = capi.Engine(local, collect_stats=1)
        se.set_lens(lens)
        se.set_aperture(tex)
        acc = sharding.accum_tensor(params, dev)
        se.render_ghosts_device(lights_a, capi.copy_params(params, shard=(0, world) if world > 1 else (0, 0)), acc.data_ptr(), clear_first=True)
        executed = se.exec_stats()
        se.close()
        del acc

    # ---- strict leg: the same pipeline, FP64 geometry ------------------------------------------------------------------------
    strict = None
    if not args.no_strict:
        ps_ = capi.copy_params(params, precision=capi.STRICT)
        del pipe
        frames_s, finish_s, result_s, pipe_s = build_pipeline(ps_)
        frames_s(W_)
        finish_s()
        barrier()
        sb, _ = timed_brackets(frames_s, finish_s, 3)
        st_ms = []
        for _ in range(6):
            barrier()
            frames_s(1)
            finish_s()
            barrier()
            st_ms.append(eng.stats()["last_trace_ms"])
        strict = {"value": inter_frame * K / (statistics.median(sb) * 1e-3), "unit": "interactions/s", "ms_per_step": statistics.median(sb) / K,
                  "brackets_ms_per_step": [x / K for x in sb], "kernel_ms": statistics.median(st_ms[1:]),
                  "what": "LFB_STRICT: the same kernels with FP64 positions / directions (FP32 weights), the precision that meets north_star's "
                          "1e-5-lens-unit per-ray bar (see parity.strict)"}
        del pipe_s
        frames, finish, result, pipe = build_pipeline(params)  # back to FP32 for the legs below
        frames(W_)
        finish()
        barrier()

    # ---- e2e: the reference-facing call with host buffers, every step ---------------------------------------------------------
    pinned_tex = capi.PinnedArray(tex.shape, np.float32)
    pinned_tex.array[...] = tex
    job_bytes = 304 + 50 * 64  # sizeof(lfb::Job) + LFB_MAX_STEPS * sizeof(lfb::StepF): re-uploaded when the lights change
    h2d = tex.nbytes * world + (jobs_frame + 3 * n_lights) * job_bytes  # ghost jobs + one prefix slot per (light, lambda)
    e2e_extra = {}
    if world == 1:
        pinned_out = capi.PinnedArray((HEIGHT, WIDTH, 3), np.float64)  # HDRImageBuffer::data layout (Vector3D, 24 B), page-locked
        pinned_out.array[...] = 0.0
        tiles_seen = []

        def e2e_step(k):
            eng.set_aperture(pinned_tex.array)  # this step's input, host -> device
            tiles_seen.append(eng.render_ghosts_sparse(lights_a if k % 2 == 0 else lights_b, params, pinned_out.array, elem=capi.F64x3,
                                                       out_is_clear=(k == 0 and not tiles_seen)))

        for k in range(W_ + (W_ % 2)):
            e2e_step(k)
        e2e_brackets = []
        with clocks:
            for _ in range(N_BRACKETS):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(K):
                    e2e_step(k)
                torch.cuda.synchronize()
                e2e_brackets.append((time.perf_counter() - t0) * 1e3)
        blocking_ms = statistics.median(e2e_brackets) / K
        blocking_brackets = e2e_brackets
        e2e_dev = eng.stats()  # device-side split of the last call: trace kernels | whole call (trace + tile kernel's writes into host memory)
        e2e_extra["e2e_device_split_ms"] = {"trace_kernels": e2e_dev["last_trace_ms"], "whole_call_on_device": e2e_dev["last_frame_ms"],
                                            "note": "of one BLOCKING call; the rest of its ms_per_step is the 1 MB aperture upload + its synchronize, launch latency and the ctypes calls"}
        tiles = statistics.median(tiles_seen[-K:])
        # bytes that cross PCIe per frame: the 8 x 8 quadrants (of the dirty 16 x 16 tiles) that are non-zero in this frame or were in
        # the buffer's previous frame (the other sun), 24 bytes per pixel
        d2h = dirty_quadrants(eng.render_ghosts(lights_a, params), eng.render_ghosts(lights_b, params)) * 64 * 24 + 4
        e2e_extra["e2e_blocking_call"] = {"value": inter_frame / (blocking_ms * 1e-3), "unit": "interactions/s", "ms_per_step": blocking_ms,
                                          "brackets_ms_per_step": [t / K for t in blocking_brackets], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                                          "api": "lfb_render_ghosts_sparse, one blocking call per frame (trace, then the tile kernel's PCIe-bound stores, then return)"}
        # the headline e2e: the same frames with FOUR IN FLIGHT (lfb_render_ghosts_sparse_begin / _end, four page-locked host frames
        # in rotation): a frame's paced tile writes (PCIe-bound) overlap the next frames' traces (SM-bound); every frame is
        # collected in host memory by _end before its buffer is used again
        R_SLOTS = 4
        outs = [pinned_out] + [capi.PinnedArray((HEIGHT, WIDTH, 3), np.float64) for _ in range(R_SLOTS - 1)]
        for o in outs:
            o.array[...] = 0.0
        pending = [False] * R_SLOTS
        fresh = [True] * R_SLOTS
        tiles2 = []

        def pipe_drain():
            for s in range(R_SLOTS):
                if pending[s]:
                    tiles2.append(eng.render_ghosts_sparse_end(s))
                    pending[s] = False

        # slot s takes the frames k = s (mod R); the suns are dealt so that EACH slot's buffer alternates between the two suns, like
        # the single buffer of the blocking measurement: every frame also re-zeroes the tiles of the buffer's previous sun
        def pipe_lights(k):
            return (lights_a, lights_b)[(k // R_SLOTS + k % R_SLOTS) % 2]

        def pipe_step(k):
            s = k % R_SLOTS
            if pending[s]:
                tiles2.append(eng.render_ghosts_sparse_end(s))  # frame k - 2 is complete in outs[s]
            eng.set_aperture(pinned_tex.array)  # this step's input, host -> device
            eng.render_ghosts_sparse_begin(pipe_lights(k), params, outs[s].array, s, elem=capi.F64x3, out_is_clear=fresh[s])
            pending[s], fresh[s] = True, False

        for k in range(W_ + 2 * R_SLOTS):
            pipe_step(k)
        pipe_drain()
        e2e_brackets = []
        with clocks:
            for _ in range(N_BRACKETS):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(K):
                    pipe_step(k)
                pipe_drain()
                e2e_brackets.append((time.perf_counter() - t0) * 1e3)
        e2e_ms = statistics.median(e2e_brackets) / K
        tiles = statistics.median(tiles2[-K:])
        # every collected frame is the frame: the last two, against the full-frame call
        for s in range(R_SLOTS):
            kk = max(k for k in range(K) if k % R_SLOTS == s) if K > s else -1
            if kk >= 0:
                parity["e2e_pipelined_frame_%d_equals_full_frame_call" % s] = bool(np.array_equal(outs[s].array, eng.render_ghosts(pipe_lights(kk), params)))
        for o in outs[1:]:
            o.free()
        pinned_out.array[...] = 0.0
        tiles_seen.clear()
        # the frame that landed in host memory is the frame (checked once, outside the timed region)
        eng.render_ghosts_sparse(lights_a, params, pinned_out.array, elem=capi.F64x3, out_is_clear=True)
        parity["e2e_frame_equals_full_frame_call"] = bool(np.array_equal(pinned_out.array, eng.render_ghosts(lights_a, params)))
        e2e_api = ("lfb_render_ghosts_sparse_begin / _end, %d frames in flight (F64x3, stride 24 = HDRImageBuffer layout, %d page-locked host "
                   "frames in rotation): the device writes the non-zero 8x8 quadrants of each frame's dirty 16x16 tiles (median %d tiles of 8160, incl. the re-zeroed "
                   "ones of the buffer's previous frame) into the caller's memory, paced below the PCIe rate, while the next frames are traced; "
                   "e2e_blocking_call is the one-call-per-frame form" % (R_SLOTS, R_SLOTS, int(tiles)))
        # the same frame through the full-frame blocking call (every pixel crosses PCIe: round 1's e2e)
        ts = []
        for k in range(min(K, 20) + 2):
            t0 = time.perf_counter()
            eng.set_aperture(pinned_tex.array)
            eng.render_ghosts(lights_a if k % 2 == 0 else lights_b, params, out=pinned_out.array, elem=capi.F64x3)
            ts.append((time.perf_counter() - t0) * 1e3)
        e2e_extra["e2e_full_frame"] = {"value": inter_frame / (statistics.median(ts[2:]) * 1e-3), "unit": "interactions/s", "ms_per_step": statistics.median(ts[2:]),
                                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": HEIGHT * WIDTH * 24,
                                       "api": "lfb_render_ghosts (every pixel converted and copied: 49.8 MB over PCIe)"}
        pinned_out.free()
    else:
        # N > 1: every rank writes its share of the dirty tiles straight into ONE page-locked host frame that all ranks have
        # mapped (a POSIX shared-memory segment registered by each process): N PCIe links in parallel, no staging on rank 0
        from multiprocessing import shared_memory
        frame_bytes = HEIGHT * WIDTH * 24
        name = [None]
        shm = None
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=N_BUF * frame_bytes)
            name[0] = shm.name
        dist.broadcast_object_list(name, src=0, group=cpu_group)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
            try:  # rank 0 owns (and unlinks) the segment: keep this process's resource tracker from trying again at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        host = np.frombuffer(shm.buf, dtype=np.float64, count=N_BUF * HEIGHT * WIDTH * 3).reshape(N_BUF, HEIGHT, WIDTH, 3)
        if rank == 0:
            host[...] = 0.0
        dist.barrier(group=cpu_group)
        L = capi.lib()
        if L.lfb_host_register(host.ctypes.data, host.nbytes) != capi.OK:
            raise SystemExit("cudaHostRegister of the shared host frame failed: " + L.lfb_last_error().decode())
        base = L.lfb_host_device_pointer(host.ctypes.data)
        del pipe
        drn = None if args.no_drain else capi.Engine(local, stream_priority=1)  # its stream carries the paced host writes
        pe = sharding.PeerSparse(eng, params, rank, world, dev, dist.group.WORLD, n_buffers=N_BUF, finalize_engine=fin,
                                 host_out_ptrs=[base + b * frame_bytes for b in range(N_BUF)], drain_engine=drn, host_stride=24)

        # two frames in flight per rank: before frame k is enqueued, frame k - 2 is collected -- complete in the host frame on EVERY
        # rank (PeerSparse.wait_frame: this rank passed the device-side barrier that each rank reaches after its reduce of that frame)
        def e2e_step(k):
            if pe.k >= 2:
                pe.wait_frame(pe.k - 2)
            eng.set_aperture(pinned_tex.array)
            return pe.frame(lights_a if k % 2 == 0 else lights_b, owner=0, elem=capi.F64x3, stride=24)

        def e2e_drain():
            pe.finish()
            torch.cuda.synchronize()

        pe.begin()
        for k in range(W_ + (W_ % 2)):
            e2e_step(k)
        e2e_drain()
        e2e_brackets, blocking_brackets = [], []
        with clocks:
            for _ in range(N_BRACKETS):
                barrier()
                t0 = time.perf_counter()
                for k in range(K):
                    b = e2e_step(k)
                e2e_drain()
                barrier()
                t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_brackets.append(float(t[0]))
            for _ in range(3):  # the same frames, one at a time (the previous rounds' e2e): every step waits for its own frame
                barrier()
                t0 = time.perf_counter()
                for k in range(K):
                    eng.set_aperture(pinned_tex.array)
                    b = pe.frame(lights_a if k % 2 == 0 else lights_b, owner=0, elem=capi.F64x3, stride=24)
                    e2e_drain()
                barrier()
                t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                blocking_brackets.append(float(t[0]))
        e2e_ms = statistics.median(e2e_brackets) / K
        e2e_extra["e2e_blocking_call"] = {"ms_per_step": statistics.median(blocking_brackets) / K, "unit": "interactions/s",
                                          "value": inter_frame / (statistics.median(blocking_brackets) / K * 1e-3),
                                          "api": "the same frames one at a time: every step waits for its own frame on every rank"}
        eng.set_aperture(pinned_tex.array)
        b = pe.frame(lights_a, owner=0, elem=capi.F64x3, stride=24)
        e2e_drain()
        barrier()
        if rank == 0:
            whole64 = eng.render_ghosts(lights_a, params)
            parity["e2e_host_frame_equals_single_gpu"] = bool(np.array_equal(host[b], whole64))
            nz_quads = dirty_quadrants(whole64, eng.render_ghosts(lights_b, params))
        else:
            nz_quads = 0
        d2h = nz_quads * 64 * 24
        e2e_api = ("PeerSparse with shared page-locked host frames (3 in rotation, two frames in flight): every rank's lfb_reduce_tiles_peers "
                   "leaves its share of the dirty tiles in a device staging buffer and lfb_drain_tiles copies it into host memory over the rank's "
                   "own PCIe link, paced just under the link rate; a frame is collected -- complete on every rank -- before the frame after "
                   "next is enqueued (d2h_bytes_per_step counts the 8x8 quadrants that are non-zero in the frame or in "
                   "the buffer's previous frame)")
        del pe
        barrier()
        if drn is not None:
            drn.close()
        L.lfb_host_unregister(host.ctypes.data)
        del host
        shm.close()
        dist.barrier(group=cpu_group)
        if rank == 0:
            shm.unlink()
    e2e_value = inter_frame / (e2e_ms * 1e-3)

    # ---- BASELINE configs 3 and 4: ONE frame, strong-scaled over the N ranks through the same pipeline (device time, max over
    # ranks; the reduced frame checked against the unsharded one on rank 0) -- north_star's "near-linear scaling on the
    # spectral / many-light configs" as numbers the driver's 1/2/4/8-GPU runs carry
    other_configs = None
    if not args.no_configs and (world == 1 or args.reduce == "sparse"):
        other_configs = {}
        lattice = [(0.1 + 0.8 * a / 7, 0.1 + 0.8 * b / 7) for b in range(8) for a in range(8)]
        try:
            del pipe
        except NameError:
            pass
        for cname, n_lam, grid_c, Wc, Hc, suns, ppu in (("cfg3", 32, 512, 1920, 1080, [(0.45, 0.55)], 0.0), ("cfg4", 3, 1024, 3840, 2160, lattice, 0.8)):
            lens_c = capi.builtin_lens(n_lam, COATING_NM)
            eng.set_lens(lens_c)
            pc = capi.make_params(capi.MODE_EXACT_GRID, Wc, Hc, grid_n=grid_c, pair_set=capi.PAIRS_ALL, include_direct=1, px_per_unit=ppu)
            lights_c = [make_sun(x, y) for x, y in suns]
            _, inter_c, jobs_c = capi.count_work(lens_c, pc, len(lights_c))
            frames_c, finish_c, result_c, pipe_c = build_pipeline(pc)
            frames_c(2, lights_c)
            finish_c()
            barrier()
            tb, _ = timed_brackets(frames_c, finish_c, 3, K=1, lights=lights_c)
            b = frames_c(1, lights_c)
            finish_c()
            barrier()
            same = None
            if rank == 0:
                whole = torch.from_numpy(eng.render_ghosts(lights_c, pc, elem=capi.F32x3)).to(dev)
                same = bool(torch.equal(result_c(b), whole))
                del whole
            other_configs[cname] = {"ms_per_frame": statistics.median(tb), "interactions_per_s": inter_c / (statistics.median(tb) * 1e-3),
                                    "interactions_per_frame": inter_c, "jobs": jobs_c, "lights": len(lights_c), "wavelengths": n_lam, "grid": grid_c,
                                    "sensor": [Wc, Hc], "equals_single_gpu_frame": same, "scaling": "strong: ONE frame over %d GPU(s)" % world}
            del pipe_c, frames_c, finish_c, result_c
            torch.cuda.empty_cache()
            barrier()
        eng.set_lens(lens)

    # ---- N > 1: the single-process form (lfb_create_multi: one host thread drives all GPUs), rank 0 alone ----------------------
    single_process = None
    if world > 1:
        if rank == 0:
            try:
                m = capi.MultiEngine(list(range(world)))
                m.set_lens(lens)
                m.set_aperture(tex)
                buf = capi.PinnedArray((HEIGHT, WIDTH, 3), np.float64)
                buf.array[...] = 0.0
                ts, st = [], None
                for k in range(12):
                    t0 = time.perf_counter()
                    m.render_ghosts(lights_a if k % 2 == 0 else lights_b, params, buf.array, out_is_clear=(k == 0))
                    ts.append((time.perf_counter() - t0) * 1e3)
                m.render_ghosts(lights_a, params, buf.array)
                st = m.stats()
                single_process = {"api": "lfb_create_multi + lfb_render_ghosts_multi: ONE host thread, %d GPUs, event-ordered streams, tiles written "
                                         "straight into the caller's page-locked buffer" % world,
                                  "ms_per_frame_blocking_call": statistics.median(ts[2:]), "value": inter_frame / (statistics.median(ts[2:]) * 1e-3),
                                  "unit": "interactions/s", "stage_ms": st,
                                  "equals_single_gpu": bool(np.array_equal(buf.array, eng.render_ghosts(lights_a, params)))}
                buf.free()
                m.close()
            except Exception as exc:  # reported, never fatal: the headline numbers are above
                single_process = {"error": repr(exc)[:300]}
        dist.barrier(group=cpu_group)

    # ---- N = 1 extras: the displayable frame and the starburst -------------------------------------------------------
    e2e_rgba8 = starburst = None
    if world == 1:
        pinned_rgba = capi.PinnedArray((HEIGHT, WIDTH), np.uint32)

        def rgba_step(k):
            eng.set_aperture(pinned_tex.array)
            eng.render_frame_rgba8(lights_a if k % 2 == 0 else lights_b, params, flare_radius=-1.0, out=pinned_rgba.array)

        for k in range(W_):
            rgba_step(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            rgba_step(k)
        torch.cuda.synchronize()
        rgba_s = time.perf_counter() - t0
        e2e_rgba8 = {"value": inter_frame * K / rgba_s, "unit": "interactions/s", "ms_per_step": rgba_s / K * 1e3,
                     "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": HEIGHT * WIDTH * 4,
                     "api": "lfb_render_frame_rgba8: ghosts -> HDRImageBuffer::toColor -> ImageBuffer RGBA8 on the device (the displayable frame)"}
        pinned_rgba.free()
        out64 = capi.PinnedArray((HEIGHT, WIDTH, 3), np.float64)
        eng.set_starburst_aperture(tex)
        for _ in range(3):
            eng.render_starburst(lights_a, WIDTH, HEIGHT, 50.0, 1.0, out=out64.array)
        ts = []
        for _ in range(10):
            eng.render_starburst(lights_a, WIDTH, HEIGHT, 50.0, 1.0, out=out64.array)
            ts.append(eng.stats()["last_trace_ms"])
        eng.set_starburst_aperture(tex)  # a new mask: the next frame recomputes the lattice spectrum
        eng.render_starburst(lights_a, WIDTH, HEIGHT, 50.0, 1.0, out=out64.array)
        starburst = {"gpu_frame_ms_device": statistics.median(ts), "gpu_frame_ms_device_first_frame_of_a_mask": eng.stats()["last_trace_ms"],
                     "api": "lfb_render_starburst, 1920x1080, pentbig500_14 (bbox 340x350 texels); the lattice spectrum |F| is kept per mask"}
        out64.free()

    peaks = eng.probe_peaks() if rank == 0 else None
    line = None
    if rank == 0:
        k_ms = statistics.median(trace_ms)
        ach = inter_rank * FLOP_PER_INTERACTION / (k_ms * 1e-3)
        mufu_ach = inter_rank * MUFU_PER_INTERACTION / (k_ms * 1e-3)
        frac = ach / peaks["fp32_flops"]
        exec_ratio = executed["steps"] / inter_rank if executed else None
        roofline = {
            "bound": "fp32", "kernel": "lfb::xt::ghost_kernel<float,12,128> + lfb::xt::prefix_kernel<float> (exact_trace.cuh)",
            "achieved": ach / 1e12, "peak": peaks["fp32_flops"] / 1e12, "unit": "TFLOP/s", "frac": frac,
            "frac_is": "NOMINAL: declared 64 FLOP x nominal interactions (rays x I(i,j), BASELINE.json's unit) / kernel time / FFMA peak -- "
                       "algorithmic throughput, NOT pipe occupancy; see executed_* and issue_slot_frac",
            "traffic": 24.03e6, "traffic_note": "bytes per launch: dram__bytes_read.sum 24.03 MB + dram__bytes_write.sum 0 of ghost_kernel in the committed capture "
                                                "profiles/r2_ghost_v16_ncu_summary.txt (cold L2: the prefix cache it reads). The kernel has no algorithmic HBM stream -- rays "
                                                "come from their grid index, the 28 MB prefix cache and the sensor atomics live in L2 -- so there is no byte roofline to compare with",
            "kernel_ms": k_ms, "kernel_ms_spread": spread(trace_ms), "interactions_per_launch": inter_rank, "flop_per_interaction": FLOP_PER_INTERACTION,
            "executed_steps": executed["steps"] if executed else None,
            "executed_ray_pairs_started": executed["ray_pairs_started"] if executed else None,
            "executed_ray_pairs_landed": executed["ray_pairs_landed"] if executed else None,
            "executed_over_nominal": exec_ratio,
            "executed_flop": executed["steps"] * FLOP_PER_INTERACTION if executed else None,
            "frac_executed": frac * exec_ratio if executed else None,
            "executed_note": "surface steps the kernels actually ran (lfb_exec_stats, the counting build): mirror-image ray pairs share one trace, "
                             "the forward sweep is traced once per (light, lambda), rays stop where they die",
            "issue_slot_frac": 0.68, "issue_slot_source": "ncu --set full, profiles/r2_ghost_v16_ncu_summary.txt (Issue Slots Busy 68.2 %, 75.4 us, 24 MB of DRAM reads; stalls: long scoreboard 26 %, wait 16 %, barrier 11 %)",
            "peak_source": "measured live on this GPU by lfb_probe_peaks (register-only FFMA chains); MEASURED_PEAKS.json holds no "
                           "FP32 figure. The trace is scalar FP32/MUFU math: neither 'hbm' nor 'tensor' bounds it",
            "mufu": {"achieved_gops": mufu_ach / 1e9, "peak_gops": peaks["mufu_ops"] / 1e9, "frac": mufu_ach / peaks["mufu_ops"],
                     "mufu_per_interaction": MUFU_PER_INTERACTION},
            "sm_clock_mhz_during_probe": peaks["sm_clock_hz"] / 1e6,
        }
        line = {
            "metric": "ray_surface_interactions_per_s", "value": value, "unit": "interactions/s", "n_gpus": world,
            "steps": K, "warmup": W_, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "lights": n_lights, "sun": "ns=(0.45,0.55) through a 50x35 deg camera -> off-axis angle %.4f rad" % lights_a[0].theta,
                       "jobs_per_frame": jobs_frame, "rays_per_frame": rays_frame, "interactions_per_frame": inter_frame,
                       "l2": "the frame's working set (prefix cache 28 MB, dirty accumulator tiles) is L2-resident by design; the timed loop rotates "
                             "over 3 accumulator / output / state sets; L2 is flushed (256 MiB memset) before each kernel-duration sample",
                       "timing": "%d brackets of exactly K frames back to back (barrier + synchronize on both sides, CUDA events, max over ranks per "
                                 "bracket): value and ms_per_step are the MEDIAN bracket, brackets_ms_per_step lists all" % N_BRACKETS,
                       "multi_gpu": "jobs (light x pair x wavelength) dealt to ranks (whole (light, lambda) groups); " + (
                           "single GPU" if world == 1 else
                           {"sparse": "tile-sparse reduce + finalize over NVLink peer memory (lfb_reduce_tiles_peers: each rank sums its share of the "
                                      "dirty tiles and stores the pixels into rank 0's frame), device-side barrier, no collective call",
                            "nccl": "one NCCL int64 sum-reduce of the whole frame to rank 0",
                            "peer": "dense fused reduce + finalize over NVLink peer memory", "multicast": "dense fused reduce + finalize through NVSwitch multicast"}[args.reduce])},
            "frame_ms_1080p": dev_ms / K, "brackets_ms_per_step": [x / K for x in brackets_ms], "host_enqueue_ms_per_step": host_enqueue_ms,
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "brackets_ms_per_step": [x / K for x in e2e_brackets], "api": e2e_api},
            "strict": strict,
            "parity": parity,
            "single_process": single_process,
            "other_configs": other_configs,
            "starburst": starburst,
            "e2e_rgba8": e2e_rgba8,
            "gpu_launches": int(launches_per_frame * K),
            "gpu_launches_per_frame": int(launches_per_frame),
            "clocks": clocks.report(),
        }
        line.update(e2e_extra)
        if world == 1 and not args.no_cpu:
            pb, port_seconds = parity_block(eng, lens, tex, lights_a[0], params)
            line["parity"].update(pb)
            line["cpu_baseline"] = cpu_baseline_block(port_seconds, inter_frame)
        print(json.dumps(line))
    pinned_tex.free()
    fin.close()
    eng.close()
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    return 0


def main():
    # libraries (NCCL's version banner, torch warnings) may write to fd 1: keep stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and oracle-parity legs")
    ap.add_argument("--no-strict", action="store_true", help="skip the LFB_STRICT leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 3 / config 4 single-frame legs")
    ap.add_argument("--fin-priority", type=int, default=1, help="stream priority of the tile finalize / reduce engine (1: highest, 0: default)")
    ap.add_argument("--no-drain", action="store_true", help="N > 1 e2e: the reduce kernels store into the host frame themselves (unpaced) instead of staging + paced drain")
    ap.add_argument("--reduce-ctas", type=int, default=0, help="CTAs of the cross-GPU tile reduce kernel (0: the library default)")
    ap.add_argument("--reduce", default="sparse", choices=["sparse", "nccl", "peer", "multicast"],
                    help="N > 1: tile-sparse reduce + finalize over NVLink peer memory (default), or round 1's dense paths: one NCCL int64 "
                         "reduce of the whole frame / the dense fused kernel over peer memory / the same through NVSwitch multicast")
    args = ap.parse_args()
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
