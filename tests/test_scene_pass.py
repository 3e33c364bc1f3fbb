"""The path-traced scene pass (BASELINE config 5 on a substitute scene: dae/dragon.dae is missing from the
reference tree).  Pinning chain: compiled reference (oracle/_ref: PathTracer::est_radiance_global_illumination over the
reference's own BVHAccel / Triangle / Sphere / BSDFs / lights) -> golden frames (tests/golden/scene.npz, tools/make_golden.py)
-> oracle restatement (oracle/lf_oracle.c, brute force) -> device kernel (lens_flare_b200/csrc/scene.cu, own BVH)."""
import os

import numpy as np
import pytest

import scene_fixtures as sf
from lens_flare_b200 import capi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene.npz")


def cases():
    cam = sf.default_camera()
    yield "flat", sf.pyramid_scene(3, False), cam, 192, 108
    yield "smooth", sf.pyramid_scene(4, True), sf.look_at((-6.0, 2.5, 6.0), (1.0, 0.6, -1.5)), 160, 90
    inside = sf.pyramid_scene(2, False)
    inside["spheres"] = np.array([[0.0, 1.0, 6.0, 4.5]])   # the camera sits INSIDE this sphere (Sphere::test's t2 branch)
    inside["sph_mat"] = np.array([3], np.int32)
    yield "inside_sphere", inside, sf.look_at((0.5, 2.0, 7.0), (0.0, 0.8, -1.0)), 96, 54


def test_oracle_scene_pass_vs_reference_golden(port):
    z = np.load(GOLDEN)
    for name, scene, cam, W, H in cases():
        want = z[name]
        got = port.scene_radiance(scene, cam, W, H)
        assert want.shape == (H, W, 3) and (want.sum(axis=2) > 0).mean() > (0.05 if name == "inside_sphere" else 0.3)
        assert np.abs(got - want).max() <= 1e-13, name


def test_oracle_scene_pass_vs_compiled_reference(port, ref):
    for name, scene, cam, W, H in cases():
        a, b = ref.scene_radiance(scene, cam, W, H), port.scene_radiance(scene, cam, 0 + W, H)
        assert np.abs(a - b).max() <= 1e-13, name
    # shadows exist (a light is blocked somewhere) and the emissive panel shows its radiance
    _, scene, cam, W, H = next(cases())
    lit = port.scene_radiance(scene, cam, W, H)
    no_blockers = dict(scene, tri_pos=scene["tri_pos"][:2], tri_nrm=scene["tri_nrm"][:2], tri_mat=scene["tri_mat"][:2],
                       spheres=np.zeros((0, 4)), sph_mat=np.zeros(0, np.int32))
    open_ground = port.scene_radiance(no_blockers, cam, W, H)
    assert (open_ground.sum(axis=2) > lit.sum(axis=2) + 1e-6).any()
    assert np.isclose(lit.max(), 4.0)


@pytest.mark.gpu
def test_gpu_scene_pass_vs_reference_golden_and_oracle(engine, port):
    """The device kernel (own BVH, FP64 in the reference's operation order) against the compiled reference's golden frames:
    <= 1e-12 on every pixel; at 1080p against the oracle; F32x3 / additive layouts."""
    z = np.load(GOLDEN)
    for name, scene, cam, W, H in cases():
        engine.set_scene(scene)
        got = engine.render_scene(capi.make_camera(cam), W, H)
        assert np.abs(got - z[name]).max() <= 1e-12, (name, np.abs(got - z[name]).max())
    name, scene, cam, _, _ = next(cases())
    engine.set_scene(scene)
    W, H = 1920, 1080
    want = port.scene_radiance(scene, cam, W, H)
    got = engine.render_scene(capi.make_camera(cam), W, H)
    assert np.abs(got - want).max() <= 1e-12
    assert engine.stats()["last_trace_ms"] < 5.0
    f32 = engine.render_scene(capi.make_camera(cam), W, H, elem=capi.F32x3)
    assert np.array_equal(f32, want.astype(np.float32)) or np.abs(f32 - want).max() <= 1e-6
    acc = np.full((H, W, 3), 0.25)
    engine.render_scene(capi.make_camera(cam), W, H, out=acc, additive=True)
    assert np.abs(acc - (want + 0.25)).max() <= 1e-12
    with pytest.raises(capi.LfbError):
        engine.set_scene(dict(scene, lights=np.array([[2, 1, 1, 1, 0, -1, 0]], float)))  # an area light: not supported


@pytest.mark.gpu
def test_gpu_composite_config5(engine, port, apertures):
    """BASELINE config 5 on the substitute scene: path-traced scene + spectral flare (8 wavelengths, coated, EXACT_GRID) +
    starburst at 1080p, composited and tone-mapped on the device in ONE call, bit-exact against the oracle's toColor of
    (scene + ghosts) + starburst built from the separately verified layers."""
    name, scene, cam, _, _ = next(cases())
    lens = capi.builtin_lens(8, 550.0)
    engine.set_lens(lens)
    engine.set_aperture(apertures["pentbig500_14"])
    engine.set_starburst_aperture(apertures["pentbig500_14"])
    engine.set_scene(scene)
    W, H = 1920, 1080
    camera = capi.make_camera(cam)
    lt = [capi.make_light(0.62, 0.7, theta=capi.physical_theta(0.62, 0.7, cam[12], cam[13]), radiance=(1.0, 0.944, 0.544))]
    p = capi.make_params(capi.MODE_EXACT_GRID, W, H, grid_n=128, pair_set=capi.PAIRS_ALL, include_direct=1)
    base = engine.render_scene(camera, W, H)
    ghosts = engine.render_ghosts(lt, p)
    star = engine.render_starburst(lt, W, H, 30.0, 1.0)
    got = engine.render_composite_rgba8(camera, lt, p, flare_radius=30.0, flare_intensity=1.0)
    want = port.to_color((base + ghosts) + star)
    assert np.array_equal(got, want)
    assert np.array_equal(engine.render_composite_rgba8(camera, lt, p, flip=True), port.to_color(base + ghosts)[::-1])
    assert ((got & 0xFFFFFF) != 0).mean() > 0.3
    engine.set_lens(capi.builtin_lens(3))
