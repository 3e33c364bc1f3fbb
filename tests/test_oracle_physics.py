"""EXACT_GRID physics does not exist in the reference (its refract/reflect/Fresnel are empty
stubs, src/pathtracer/advanced_bsdf.cpp:52-169): PARITY UNPINNED by the reference.  The oracle's
restatement is pinned by analytic invariants instead (mode 3)."""
import numpy as np

from oracle import bindings as ob


def tiny_tex():
    return np.ones((8, 8), np.float32)


def test_uncoated_normal_incidence_fresnel(port):
    for n0, n2 in [(1.0, 1.5), (1.652, 1.0), (1.6, 1.7)]:
        assert np.isclose(port.reflectance(n0, n2, 1.0, 0.0, 550.0), ((n0 - n2) / (n0 + n2)) ** 2, rtol=1e-14)


def test_quarter_wave_coating_ideal_index_kills_reflection(port):
    # n1 = sqrt(n0 n2) >= 1.38 -> exact quarter-wave match at lambda0, normal incidence
    n0, n2 = 1.0, 1.38 ** 2 + 0.3
    assert port.reflectance(n0, n2, 1.0, 550.0, 550.0) < 1e-28
    # with the MgF2 floor (n1 = 1.38) the textbook single-layer value
    n0, n2 = 1.0, 1.52
    n1 = 1.38
    want = ((n0 * n2 - n1 * n1) / (n0 * n2 + n1 * n1)) ** 2
    assert np.isclose(port.reflectance(n0, n2, 1.0, 550.0, 550.0), want, rtol=1e-12)


def test_reflectance_bounds_and_tir(port):
    rng = np.random.default_rng(3)
    for _ in range(2000):
        n0, n2 = rng.uniform(1.0, 1.8, 2)
        c = rng.uniform(0.0, 1.0)
        for lam0 in (0.0, 550.0):
            R = port.reflectance(n0, n2, c, lam0, rng.uniform(400, 700))
            assert 0.0 <= R <= 1.0 + 1e-12
    assert port.reflectance(1.6, 1.0, 0.3, 0.0, 550.0) == 1.0  # beyond the critical angle
    assert port.reflectance(1.5, 1.5, 0.7, 550.0, 550.0) == 0.0  # no interface


def test_brewster_angle_p_polarisation(port):
    n0, n2 = 1.0, 1.5
    thB = np.arctan(n2 / n0)
    c0 = np.cos(thB)
    c2 = np.sqrt(1 - (n0 / n2) ** 2 * (1 - c0 * c0))
    rs = (n0 * c0 - n2 * c2) / (n0 * c0 + n2 * c2)
    assert np.isclose(port.reflectance(n0, n2, c0, 0.0, 550.0), 0.5 * rs * rs, rtol=1e-12)  # rp = 0


def test_exact_converges_to_paraxial(port):
    """As entrance height and angle -> 0 the exact trace tends to the physically-consistent ABCD
    system with an O(h^3) residual.  (The reference's R_k^-1 on the backward legs,
    pathtracer.cpp:607-608, is NOT the physical backward refraction -- the curvature sign is not
    flipped -- and does not converge: measured 6-88 % off for pairs with a glass surface between
    the two reflections; hence physical_backward=1 here.)"""
    tex = tiny_tex()
    errs = []
    for P in (1.0, 0.5, 0.25):
        lens = port.builtin_lens(3)
        lens.entrance_half_height = P
        lens.stop_half_height = 100.0  # mask never clips
        for k in range(9):
            lens.semi_aperture[k] = 50.0
        lt = ob.make_light(0.6, 0.55, theta=0.02 * P)
        worst = 0.0
        for (i, j) in [(0, 1), (2, 4), (6, 8), (1, 7), (-1, -1)]:
            pe = ob.make_params(ob.MODE_EXACT_GRID, 64, 64, grid_n=4)
            pp = ob.make_params(ob.MODE_PARAXIAL_GRID, 64, 64, grid_n=4, physical_backward=1)
            he = port.trace_grid(lens, tex, lt, pe, i, j, 1)
            hp = port.trace_grid(lens, tex, lt, pp, i, j, 1)
            assert not he["flags"].any() & ~ob.RAY_OFF_SENSOR
            scale = max(np.abs(hp["x_s"]).max(), np.abs(hp["y_s"]).max())
            worst = max(worst, np.abs(he["x_s"] - hp["x_s"]).max() / scale, np.abs(he["y_s"] - hp["y_s"]).max() / scale)
        errs.append(worst)
    # relative residual is O(h^2): halving h divides it by ~4
    assert errs[0] < 0.15 and errs[2] < 0.01
    assert errs[1] < errs[0] / 3.0 and errs[2] < errs[1] / 3.0


def test_energy_weights(port, apertures):
    """Weights are products of Fresnel terms: in (0,1] for the direct path, tiny (two reflections)
    for ghosts; coating lowers ghost weights; vignetted / stopped rays carry zero weight."""
    tex = apertures["pent_11"]
    lt = ob.make_light(0.6, 0.55, theta=0.05)
    p = ob.make_params(ob.MODE_EXACT_GRID, 512, 512, grid_n=32)
    bare, coated = port.builtin_lens(3), port.builtin_lens(3, 550.0)
    d1 = port.trace_grid(bare, tiny_tex(), lt, p, -1, -1, 1)  # open stop: weight = product of 8 transmittances
    live = d1["weight"] > 0
    assert live.any() and (d1["weight"][live] <= 1).all() and (d1["weight"][live] > 0.5).all()
    d = port.trace_grid(bare, tex, lt, p, -1, -1, 1)
    assert (d["weight"] <= d1["weight"]).all()  # the mask only attenuates
    dead = (d["flags"] & (ob.RAY_MISSED | ob.RAY_VIGNETTED | ob.RAY_TIR | ob.RAY_STOPPED)) != 0
    assert (d["weight"][dead] == 0).all()
    gb = port.trace_grid(bare, tex, lt, p, 0, 4, 1)
    gc = port.trace_grid(coated, tex, lt, p, 0, 4, 1)
    both = (gb["weight"] > 0) & (gc["weight"] > 0)
    assert both.any()
    assert (gb["weight"][both] < 0.02).all()
    assert gc["weight"][both].mean() < 0.25 * gb["weight"][both].mean()


def test_dispersion_model_reproduces_rgb_anchors(port):
    """n(lambda) passes through the reference's R,G,B indices at 650/550/450 nm and is monotone."""
    rgb = port.builtin_lens(3)
    # 30 samples on [400,700]: sample l has lambda = 405 + 10 l -> 650 (l=24.5) is not on the grid,
    # so evaluate the fit through a lens whose samples hit the anchors: n_lambda = 6 -> 425..675 step 50
    six = port.builtin_lens(6)
    lam = np.array(six.lambda_nm[:6])
    assert np.allclose(lam, [425, 475, 525, 575, 625, 675])
    # monotone normal dispersion between the anchors for every glass
    for k in range(9):
        n = np.array([six.ior[l][k] for l in range(6)])
        if rgb.ior[0][k] == 1.0:
            assert (n == 1.0).all()
        else:
            assert (np.diff(n) <= 1e-7).all()  # index falls with wavelength
    # the anchors themselves: n_lambda = 30 puts samples at 405 + 10 l nm -> 455, 545/555, 645/655 bracket them
    thirty = port.builtin_lens(30)
    lam30 = np.array(thirty.lambda_nm[:30])
    for k in (0, 1, 3, 6, 7):
        n = np.array([thirty.ior[l][k] for l in range(30)])
        for a, lam_a in enumerate((650.0, 550.0, 450.0)):
            lo, hi = np.searchsorted(lam30, lam_a) - 1, np.searchsorted(lam30, lam_a)
            assert min(n[lo], n[hi]) - 1e-6 <= rgb.ior[a][k] <= max(n[lo], n[hi]) + 1e-6
    # rgb weights of the spectral lens sum to 1 per channel
    w = np.array([[six.rgb_weight[l][c] for c in range(3)] for l in range(6)])
    assert np.allclose(w.sum(0), 1.0, atol=1e-6)


def test_render_accumulators_match_double_image(port, apertures):
    lens = port.builtin_lens(3)
    lt = ob.make_light(0.45, 0.55)
    for mode in (ob.MODE_PARAXIAL_GRID, ob.MODE_EXACT_GRID):
        for splat in (ob.SPLAT_NEAREST, ob.SPLAT_BILINEAR):
            p = ob.make_params(mode, 128, 96, grid_n=24, splat=splat, px_per_unit=0.1)
            img, acc = port.render(lens, apertures["pent_11"], [lt], p, want_accum=True)
            assert acc.any()
            assert np.array_equal(img, acc.astype(np.float64) * 2.0 ** -40)


def test_sharded_oracle_sums_to_whole(port, apertures):
    lens = port.builtin_lens(3)
    lights = [ob.make_light(0.45, 0.55), ob.make_light(0.7, 0.3, radiance=(0.5, 1.0, 2.0))]
    p = ob.make_params(ob.MODE_EXACT_GRID, 96, 64, grid_n=16, pair_set=ob.PAIRS_ALL, include_direct=1, px_per_unit=0.1)
    _, whole = port.render(lens, apertures["pent_11"], lights, p, want_accum=True)
    for n in (2, 3, 8):
        tot = np.zeros_like(whole)
        for r in range(n):
            _, a = port.render(lens, apertures["pent_11"], lights, ob.copy_params(p, shard=(r, n)), want_accum=True)
            tot += a
        assert np.array_equal(tot, whole)
    # contiguous blocks of (light, lambda) groups, like the engine's list (tests/test_host_api.py): with as many shards as lights,
    # shard r is exactly light r's frame
    for r in range(2):
        _, a = port.render(lens, apertures["pent_11"], lights, ob.copy_params(p, shard=(r, 2)), want_accum=True)
        _, alone = port.render(lens, apertures["pent_11"], [lights[r]], p, want_accum=True)
        assert alone.any() and np.array_equal(a, alone), r


# ---------------------------------------------------------------------------------------------
# point lights (lfb_light.distance; SURVEY 8f-3 -- the reference's flares only fire for DirectionalLight,
# pathtracer.cpp:35, so this too is pinned by invariants)
# ---------------------------------------------------------------------------------------------
def _open_lens(port, P=14.5):
    lens = port.builtin_lens(3)
    lens.entrance_half_height = P
    lens.stop_half_height = 100.0
    for k in range(9):
        lens.semi_aperture[k] = 60.0
    return lens


def test_point_light_through_empty_lens_is_a_straight_line(port):
    """All indices 1: no refraction, transmittance 1.  The ray from the point source through entrance point (x, y) hits
    the sensor plane at (x, y) + Z * (vx, vy) / vz with v = (x/D + sin t, y/D, cos t), carrying |v|^-3."""
    lens = _open_lens(port)
    for lam in range(3):
        for k in range(9):
            lens.ior[lam][k] = 1.0
    Z = sum(float(lens.thickness[k]) for k in range(9))  # the vertices are sums of the float thicknesses, in double
    theta, D, N = 0.07, 180.0, 12
    lt = ob.make_light(0.6, 0.55, theta=theta, distance=D)
    p = ob.make_params(ob.MODE_EXACT_GRID, 64, 64, grid_n=N)
    h = port.trace_grid(lens, tiny_tex(), lt, p, -1, -1, 1)
    idx = np.arange(N * N)
    x = -14.5 + ((idx % N) + 0.5) * (29.0 / N)
    y = -14.5 + ((idx // N) + 0.5) * (29.0 / N)
    th = float(np.float32(theta))
    vx, vy, vz = x / D + np.sin(th), y / D, np.cos(th)
    order = np.lexsort((np.round(h["x_s"], 6), np.round(h["y_s"], 6)))  # hits come in grid order whatever it is: compare as sets
    want_x, want_y = x + Z * vx / vz, y + Z * vy / vz
    worder = np.lexsort((np.round(want_x, 6), np.round(want_y, 6)))
    assert np.allclose(h["x_s"][order], want_x[worder], rtol=0, atol=1e-9)
    assert np.allclose(h["y_s"][order], want_y[worder], rtol=0, atol=1e-9)
    q = vx * vx + vy * vy + vz * vz
    assert np.allclose(h["weight"][order], (q ** -1.5)[worder], rtol=1e-12)


def test_point_light_limits(port, apertures):
    """distance 0, negative or infinite = directional (bit for bit); a very distant point light converges to it; a near
    one does not."""
    tex = apertures["pent_11"]
    lens = port.builtin_lens(3, 550.0)
    p = ob.make_params(ob.MODE_EXACT_GRID, 512, 512, grid_n=24)
    base = port.trace_grid(lens, tex, ob.make_light(0.6, 0.55, theta=0.06), p, 6, 8, 1)
    for d in (0.0, -5.0, float("inf")):
        same = port.trace_grid(lens, tex, ob.make_light(0.6, 0.55, theta=0.06, distance=d), p, 6, 8, 1)
        assert same.tobytes() == base.tobytes()
    far = port.trace_grid(lens, tex, ob.make_light(0.6, 0.55, theta=0.06, distance=1e12), p, 6, 8, 1)
    ok = ~np.isnan(base["x_s"])
    assert ok.sum() > 100 and np.array_equal(np.isnan(far["x_s"]), ~ok)
    assert np.abs(far["x_s"][ok] - base["x_s"][ok]).max() < 1e-7 and np.allclose(far["weight"], base["weight"], rtol=1e-9)
    near = port.trace_grid(lens, tex, ob.make_light(0.6, 0.55, theta=0.06, distance=150.0), p, 6, 8, 1)
    both = ok & ~np.isnan(near["x_s"])
    assert np.abs(near["x_s"][both] - base["x_s"][both]).max() > 1.0
    # whole frames: the near light's ghosts are a different picture, the distant one's the same picture
    pr = ob.make_params(ob.MODE_EXACT_GRID, 256, 256, grid_n=32, pair_set=ob.PAIRS_ALL, include_direct=1)
    f0 = port.render(lens, tex, [ob.make_light(0.6, 0.55, theta=0.06)], pr)
    f1 = port.render(lens, tex, [ob.make_light(0.6, 0.55, theta=0.06, distance=1e9)], pr)
    f2 = port.render(lens, tex, [ob.make_light(0.6, 0.55, theta=0.06, distance=150.0)], pr)
    nrm = np.linalg.norm(f0)
    assert nrm > 0 and np.linalg.norm(f1 - f0) / nrm < 1e-5 and np.linalg.norm(f2 - f0) / nrm > 1e-2


def test_point_light_exact_converges_to_paraxial(port):
    """Same O(h^2) relative residual as the directional bundle: the paraxial point light enters with angle
    (theta + x/D, y/D)."""
    tex = tiny_tex()
    errs = []
    for P in (1.0, 0.5, 0.25):
        lens = _open_lens(port, P)
        lt = ob.make_light(0.6, 0.55, theta=0.02 * P, distance=120.0)
        worst = 0.0
        for (i, j) in [(0, 1), (2, 4), (6, 8), (1, 7), (-1, -1)]:
            pe = ob.make_params(ob.MODE_EXACT_GRID, 64, 64, grid_n=4)
            pp = ob.make_params(ob.MODE_PARAXIAL_GRID, 64, 64, grid_n=4, physical_backward=1)
            he = port.trace_grid(lens, tex, lt, pe, i, j, 1)
            hp = port.trace_grid(lens, tex, lt, pp, i, j, 1)
            scale = max(np.abs(hp["x_s"]).max(), np.abs(hp["y_s"]).max())
            worst = max(worst, np.abs(he["x_s"] - hp["x_s"]).max() / scale, np.abs(he["y_s"] - hp["y_s"]).max() / scale)
        errs.append(worst)
    assert errs[0] < 0.15 and errs[2] < 0.01
    assert errs[1] < errs[0] / 3.0 and errs[2] < errs[1] / 3.0


def test_physical_mapping_puts_the_direct_image_on_the_light(port, apertures):
    """lfb_params.physical_mapping = 1 (ours; the reference's mapping -- origin at the sun pixel, rotation atan mod pi,
    pathtracer.cpp:412-430 -- mirrors the ghosts of a sun left of centre through the sun): with px_per_unit matched to the
    camera (W / (2 f tan(hFov/2)), f = the lens's focal length from its own ABCD system, square pixels) the direct,
    unreflected image of a distant light lands on the light's own pixel (Camera::analyze_world_coord, camera.cpp:245-273),
    in all four quadrants -- and with the reference's mapping it does not."""
    import math
    from lens_flare_b200 import capi
    lens = port.builtin_lens(3)
    tex = apertures["pentbig500_14"]
    W, H, hfov = 1920, 1080, 50.0
    vfov = 2 * math.degrees(math.atan(math.tan(math.radians(hfov / 2)) * H / W))  # square pixels
    _, full = port.paraxial_system(lens, 1, -1, -1)
    assert abs(full[0]) < 1e-3  # the sensor sits in the focal plane (green): x_s = f * theta
    f = full[1]
    ppu = W / (2 * f * math.tan(math.radians(hfov / 2)))
    for ns in ((0.62, 0.58), (0.38, 0.58), (0.38, 0.42), (0.62, 0.40), (0.5, 0.6), (0.41, 0.5)):
        lt = capi.make_light(ns[0], ns[1], theta=capi.physical_theta(ns[0], ns[1], hfov, vfov))
        for phys, on_target in ((1, True), (0, False)):
            p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=24, precision=capi.FP64, px_per_unit=ppu)
            p.physical_mapping = phys
            hits = port.trace_grid(lens, tex, lt, p, -1, -1, 1)
            ok = hits["weight"] > 0
            assert ok.sum() > 50
            px, py = hits["px"][ok].mean(), hits["py"][ok].mean()
            err = math.hypot(px - ns[0] * W, py - ns[1] * H)
            if on_target:
                assert err < 1.5, (ns, px, py)       # within the lens's own aberrations / distortion (pixels)
            else:
                assert err > 20, (ns, px, py)
