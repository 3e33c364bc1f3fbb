"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/lfb200.h declares, host-only entry points behave, and compute entry points refuse to run
without a device (there is no CPU fallback).  No kernels are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from lens_flare_b200 import capi, pathtracer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "lfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(native_lib):
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(native_lib, n), f"liblfb200.so does not export {n}"
    assert sorted(capi.SYMBOLS) == names
    assert native_lib.lfb_abi_version() == capi.ABI_VERSION


def test_struct_layouts_match_header(native_lib):
    # sizes the C compiler sees (oracle/lf_oracle.c shares the header) vs ctypes
    assert C.sizeof(capi.Lens) == 16 + 4 * 4 * 16 + 4 * 64 * 16 + 4 * 64 + 4 * 64 * 3 + 3 * 8
    assert C.sizeof(capi.Light) == 40
    assert C.sizeof(capi.Params) == 64
    assert C.sizeof(capi.Options) == 80
    assert capi.RAY_HIT_DTYPE.itemsize == 64
    assert capi.REF_GHOST_DTYPE.itemsize == 72


def test_builtin_lens_is_the_reference_prescription(native_lib, golden, port):
    lens = capi.builtin_lens(3)
    assert (lens.n_surfaces, lens.stop_index, lens.n_lambda) == (9, 5, 3)
    assert np.array_equal(np.array(lens.curvature[:9], np.float32), golden["presc_curvature"][:9])
    for c in range(3):
        assert np.array_equal(np.array(lens.ior[c][:9], np.float32), golden["presc_ior"][c])
    assert np.array_equal(np.array(lens.thickness[:9], np.float64), golden["presc_T"][:, 1])
    # identical bytes to the oracle's table, for RGB and for a coated spectral lens
    assert bytes(lens) == bytes(port.builtin_lens(3))
    assert bytes(capi.builtin_lens(32, 550.0)) == bytes(port.builtin_lens(32, 550.0))
    with pytest.raises(capi.LfbError):
        capi.builtin_lens(0)
    with pytest.raises(capi.LfbError):
        capi.builtin_lens(65)


def test_count_work_matches_survey(native_lib):
    """I(i,j) = 2(j-i)+10; sum over the 13 reference pairs = 178, over all 28 = 478."""
    lens = capi.builtin_lens(1)
    p13 = capi.make_params(capi.MODE_EXACT_GRID, 64, 64, grid_n=1)
    p28 = capi.make_params(capi.MODE_EXACT_GRID, 64, 64, grid_n=1, pair_set=capi.PAIRS_ALL)
    p28d = capi.make_params(capi.MODE_EXACT_GRID, 64, 64, grid_n=1, pair_set=capi.PAIRS_ALL, include_direct=1)
    assert capi.count_work(lens, p13, 1) == (13.0, 178.0, 13)
    assert capi.count_work(lens, p28, 1) == (28.0, 478.0, 28)
    assert capi.count_work(lens, p28d, 1) == (29.0, 488.0, 29)
    # config 2 of BASELINE.json: RGB, 256^2 per ghost
    cfg2 = capi.make_params(capi.MODE_EXACT_GRID, 1920, 1080, grid_n=256)
    rays, inter, jobs = capi.count_work(capi.builtin_lens(3), cfg2, 1)
    assert (rays, inter, jobs) == (39 * 65536.0, 178 * 3 * 65536.0, 39)
    ref = capi.make_params(capi.MODE_REF_QUADS, 512, 512)
    assert capi.count_work(capi.builtin_lens(3), ref, 1) == (78.0, 2 * 3 * 178.0, 39)


def test_job_list_order_and_sharding(native_lib):
    lens = capi.builtin_lens(3)
    p = capi.make_params(capi.MODE_EXACT_GRID, 64, 64, grid_n=8, pair_set=capi.PAIRS_ALL, include_direct=1)
    full = capi.list_jobs(lens, p, 2)
    assert full.shape == (2 * 29 * 3, 4)
    # reference order: before-stop pairs, after-stop pairs (pathtracer.cpp:735-762), colours innermost
    first_light = full[full[:, 0] == 0]
    assert [tuple(r[1:3]) for r in first_light[3:3 + 39:3]] == \
        [(i, j) for i in range(5) for j in range(i + 1, 5)] + [(6, 7), (6, 8), (7, 8)]
    assert (first_light[:3, 1] == -1).all()
    for n in (2, 3, 4, 8):
        seen = []
        loads = []
        for r in range(n):
            sh = capi.list_jobs(lens, capi.copy_params(p, shard=(r, n)), 2)
            seen += [tuple(x) for x in sh]
            loads.append(sum(2 * (j - i) + 10 if i >= 0 else 10 for _, i, j, _ in sh))
        assert sorted(seen) == sorted(tuple(x) for x in full)          # a partition
        # 6 (light, lambda) groups: whole groups as far as they divide evenly among the shards (identical cost each), the
        # jobs of the remaining groups one by one, longest first -- never more than one job's worth of imbalance
        assert max(loads) - min(loads) <= (0 if 6 % n == 0 else 26), (n, loads)
    # one RGB light on two GPUs used to be dealt 2 groups : 1 group
    loads = [sum(2 * (j - i) + 10 if i >= 0 else 10 for _, i, j, _ in capi.list_jobs(lens, capi.copy_params(p, shard=(r, 2)), 1)) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 26 and sum(loads) == 3 * 488
    groups0 = {(l, lam) for l, _, _, lam in capi.list_jobs(lens, capi.copy_params(p, shard=(0, 2)), 1)}
    assert (0, 0) in groups0 and (0, 1) not in groups0  # group 0 whole on rank 0, group 1 whole on rank 1, group 2 split
    # contiguous blocks: with as many lights as shards every shard gets ALL wavelengths of exactly one light (its deposits, and
    # the tiles the cross-GPU reduce fetches from it, then stay around that light), in the engine and in the oracle alike
    for n in (2, 4, 8):
        for r in range(n):
            sh = capi.list_jobs(lens, capi.copy_params(p, shard=(r, n)), n)
            assert {int(x[0]) for x in sh} == {r} and len(sh) == 29 * 3, (n, r)
    with pytest.raises(capi.LfbError):
        capi.list_jobs(lens, capi.copy_params(p, shard=(2, 2)), 1)


def test_parameter_validation(native_lib):
    lens = capi.builtin_lens(3)
    for bad in (dict(grid_n=0), dict(width=0), dict(mode=7), dict(precision=3), dict(splat=9), dict(fixed_point_bits=60)):
        with pytest.raises(capi.LfbError) as e:
            capi.count_work(lens, capi.copy_params(capi.make_params(capi.MODE_EXACT_GRID, 8, 8, grid_n=4), **bad), 1)
        assert e.value.code == capi.ERR_INVALID


def test_no_cpu_fallback(native_lib):
    """Without a CUDA device the engine refuses to exist; with one this test does nothing."""
    h = C.c_void_p()
    rc = native_lib.lfb_create(C.byref(h), 0)
    if rc == 0:
        native_lib.lfb_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == capi.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in native_lib.lfb_last_error()
    pt = pathtracer.PathTracer()
    pt.camera = pathtracer.Camera()
    pt.camera.ghost_aperture_texture = pathtracer.CameraApertureTexture().init_from_bytes(np.full((4, 4), 255, np.uint8))
    pt.set_frame_size(8, 8)
    pt.axis_ray, pt.angle_to_sun = (0.4, 0.6), 0.9
    with pytest.raises(capi.LfbError):
        pt.generate_ghost_buffer()


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every exported function of include/lfb200.h to the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [name for name in capi.SYMBOLS if name not in doc]
    assert not missing, missing


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under lens_flare_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lens_flare_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"(from|import)\s+oracle\b|#include[^\n]*oracle|liblf_oracle|libref_oracle|-l\S*oracle", text), f"{f} imports / includes / links the oracle"


def test_find_sun_pos_matches_reference(golden):
    """PathTracer::find_sun_pos + Camera::analyze_world_coord + DirectionalLight ctor vs the compiled reference."""
    for row in golden["sun_table"]:
        cam = pathtracer.Camera(row[:9].reshape(3, 3), row[9:12], row[12], row[13])
        pt = pathtracer.PathTracer()
        pt.camera = cam
        pt.lights = [pathtracer.DirectionalLight((1, 1, 1), row[14:17], (0, 0, -1))]
        pt.find_sun_pos()
        n, nx, ny, ang = row[17:]
        assert len(pt.flare_origins) == int(n)
        if n:
            assert pt.axis_ray == (nx, ny)
            assert pt.angle_to_sun == ang
        else:
            assert pt.axis_ray == (0.0, 0.0)


def test_find_sun_pos_point_lights():
    """Ours: a PointLight is projected like a DirectionalLight's posLight and carries its distance to the
    camera in lens units into lfb_light.distance; directional lights keep distance 0."""
    pt = pathtracer.PathTracer(mode=capi.MODE_EXACT_GRID)
    pt.camera = pathtracer.Camera(pos=(0.0, 0.0, 1.0))
    pt.lights = [pathtracer.DirectionalLight((1, 1, 1), (0.05, 0.02, 2.0), (0, 0, -1)),
                 pathtracer.PointLight((3, 2, 1), (0.1, -0.05, -0.5)),
                 pathtracer.PointLight((1, 1, 1), (50.0, 0.0, -0.5))]  # off screen
    pt.find_sun_pos()
    assert len(pt.flare_origins) == 2 and pt.flare_distance[0] == 0.0
    assert np.isclose(pt.flare_distance[1], 1000.0 * np.linalg.norm([0.1, -0.05, -1.5]))
    want = pt.camera.analyze_world_coord(np.array([0.1, -0.05, -0.5]))
    assert pt.flare_origins[1] == want and pt.axis_ray == want
    lights = pt._lfb_lights()
    assert [lt.distance for lt in lights] == pt.flare_distance
    assert lights[1].theta == capi.physical_theta(want[0], want[1]) and tuple(lights[1].radiance) == (3.0, 2.0, 1.0)


def test_aperture_texture_loader(apertures, tmp_path):
    """CameraApertureTexture::init (camera.h:26-83) semantics: red byte * float(1/255), total, bbox."""
    from PIL import Image
    for name in ("pent_11", "pentbig500_14"):
        t = pathtracer.CameraApertureTexture().init_from_bytes(apertures[name + "_u8"])
        assert np.array_equal(t.aperture, apertures[name])
        assert np.isclose(t.total_value, apertures[name + "_total"], rtol=1e-12)
        assert (t.min_x, t.min_y, t.max_x, t.max_y) == apertures[name + "_bbox"]
    # through a PNG file, gray and RGBA
    u8 = apertures["pent_11_u8"]
    Image.fromarray(u8, "L").save(tmp_path / "g.png")
    rgba = np.stack([u8, u8 // 2, u8 // 3, np.full_like(u8, 255)], -1)
    Image.fromarray(rgba, "RGBA").save(tmp_path / "c.png")
    for f in ("g.png", "c.png"):
        t = pathtracer.CameraApertureTexture().init(str(tmp_path / f))
        assert np.array_equal(t.aperture, apertures["pent_11"])
    with pytest.raises(IOError):
        pathtracer.CameraApertureTexture().init(str(tmp_path / "missing.png"))


# ---- the C++ facade (lens_flare_b200/host): host-only parts -------------------------------------
def _flare_demo():
    import subprocess
    host = os.path.join(ROOT, "lens_flare_b200", "host")
    subprocess.run(["make", "-s", "-C", host], check=True)
    return os.path.join(host, "flare_demo")


def test_cpp_png_reader_matches_reference_loader(native_lib, apertures, tmp_path):
    """lfb::CameraApertureTexture::init (own PNG decoder on zlib) yields the same mask statistics as the reference's
    loader (golden fixture), for gray, RGB, RGBA and palette PNGs with every scanline filter PIL chooses."""
    import json
    import subprocess
    from PIL import Image
    exe = _flare_demo()
    u8 = apertures["pentbig500_14_u8"]
    files = {}
    Image.fromarray(u8, "L").save(tmp_path / "gray.png")
    Image.fromarray(np.stack([u8, u8 // 2, 255 - u8], -1), "RGB").save(tmp_path / "rgb.png")
    Image.fromarray(np.stack([u8, u8 // 2, u8 // 3, np.full_like(u8, 200)], -1), "RGBA").save(tmp_path / "rgba.png", compress_level=9)
    Image.fromarray(u8, "L").convert("P").save(tmp_path / "pal.png")
    for name in ("gray.png", "rgb.png", "rgba.png", "pal.png"):
        out = subprocess.run([exe, "--png-info", str(tmp_path / name)], check=True, capture_output=True, text=True).stdout
        info = json.loads(out)
        assert (info["w"], info["h"]) == (500, 500), name
        assert info["byte_sum"] == int(u8.astype(np.int64).sum()), name
        assert tuple(info["bbox"]) == apertures["pentbig500_14_bbox"], name
        assert np.isclose(info["total"], apertures["pentbig500_14_total"], rtol=1e-12), name
    bad = subprocess.run([exe, "--png-info", str(tmp_path / "missing.png")], capture_output=True, text=True)
    assert bad.returncode == 1 and "cannot open" in bad.stderr


def test_cpp_png_writer_save_image(native_lib, apertures, tmp_path):
    """lfb::save_image (RaytracedRenderer::save_image, raytraced_renderer.cpp:717-755): rows bottom-up, alpha forced
    to 255, RGBA8 PNG that a conforming decoder (PIL) and our own reader return byte for byte."""
    import json
    import subprocess
    from PIL import Image
    exe = _flare_demo()
    u8 = apertures["pent_11_u8"][:123, :77].copy()  # ragged, non-square
    Image.fromarray(u8, "L").save(tmp_path / "in.png")
    subprocess.run([exe, "--png-copy", str(tmp_path / "in.png"), str(tmp_path / "out.png")], check=True)
    img = Image.open(tmp_path / "out.png")
    assert img.mode == "RGBA" and img.size == (u8.shape[1], u8.shape[0])
    got = np.asarray(img)
    flipped = u8[::-1]
    want = np.stack([flipped, flipped // 2, 255 - flipped, np.full_like(u8, 255)], -1)
    np.testing.assert_array_equal(got, want)
    info = json.loads(subprocess.run([exe, "--png-info", str(tmp_path / "out.png")], check=True, capture_output=True, text=True).stdout)
    assert (info["w"], info["h"]) == (u8.shape[1], u8.shape[0]) and info["byte_sum"] == int(u8.astype(np.int64).sum())
    bad = subprocess.run([exe, "--png-copy", str(tmp_path / "in.png"), str(tmp_path / "no_such_dir" / "o.png")], capture_output=True, text=True)
    assert bad.returncode == 1 and "cannot write" in bad.stderr


def test_header_is_plain_c(native_lib, tmp_path):
    """include/lfb200.h compiles as strict C99 and the library links and runs from a plain C program."""
    import subprocess
    exe = tmp_path / "abi_from_c"
    lib_dir = os.path.join(ROOT, "lens_flare_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_from_c.c"), "-o", str(exe), "-L" + lib_dir, "-llfb200", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "abi 3 jobs 87 rays 5701632 interactions 95944704" in out
    assert "sizeof(lens)=5416 light=40 params=64 hit=64" in out
    assert "lfb_create -> 0" in out or "lfb_create -> -2" in out
