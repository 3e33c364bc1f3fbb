"""Host-side mirror of the reference's call surface for the ghost path.

Same names, argument meaning and error behaviour as the reference (paths under the reference
tree), so parity tests read like its call sites:

  CameraApertureTexture.init(path)       src/pathtracer/camera.h:26-83
  Camera.analyze_world_coord(p)          src/pathtracer/camera.cpp:245-273
  DirectionalLight(rad, posLight, dir)   src/scene/light.cpp:11-16
  PathTracer.set_frame_size(w, h)        src/pathtracer/pathtracer.cpp:66-69
  PathTracer.find_sun_pos()              src/pathtracer/pathtracer.cpp:32-64
  PathTracer.generate_ghost_buffer()     src/pathtracer/pathtracer.cpp:714-762
  PathTracer.ghost_buffer                src/pathtracer/pathtracer.h:54 (HDRImageBuffer, util/image.h:105-242)

All pixel and ray work happens in liblfb200.so (capi.Engine); this module only carries the
frame description across the C ABI.  find_sun_pos is the reference's per-light host
arithmetic (a dozen double operations per light) and stays on the host, as in the reference.
"""
import math

import numpy as np

from . import capi


class CameraApertureTexture:
    """camera.h:18-88: PNG -> float mask (red channel * float(1/255)), total_value, bbox of texels > 0."""

    def __init__(self):
        self.width = self.height = 0
        self.aperture = None
        self.total_value = 0.0
        self.min_x = self.min_y = self.max_x = self.max_y = 0

    def init(self, path):
        from PIL import Image  # decode only; any conforming PNG decoder matches lodepng's RGBA8 output
        try:
            rgba = np.asarray(Image.open(path).convert("RGBA"))
        except Exception as exc:  # the reference prints and carries on (camera.h:40-44); we refuse
            raise IOError(f"[Camera] Failed to load aperture file {path}: {exc}") from exc
        return self.init_from_bytes(rgba[:, :, 0])

    def init_from_bytes(self, red_u8):
        red_u8 = np.ascontiguousarray(red_u8, np.uint8)
        self.height, self.width = red_u8.shape
        # Color(const uchar*) CGL/src/color.cpp:16-21: byte * float(1/255) in float
        self.aperture = red_u8.astype(np.float32) * np.float32(1.0 / 255.0)
        self.total_value = float(self.aperture.astype(np.float64).sum())
        ys, xs = np.nonzero(self.aperture > 0)
        if xs.size:
            self.min_x, self.max_x, self.min_y, self.max_y = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
        else:
            self.min_x = self.min_y = self.width
            self.max_x = self.max_y = -1
        return self


class DirectionalLight:
    """light.cpp:11-16 (note the sign flips the reference applies)."""

    def __init__(self, rad, posLight, lightDir):
        self.radiance = np.asarray(rad, np.float64)
        self.posLight = -np.asarray(posLight, np.float64)
        d = np.asarray(lightDir, np.float64)
        self.dirToLight = -d / np.linalg.norm(d)


class Camera:
    """The members of the reference camera the ghost path reads (camera.h:93-200)."""

    def __init__(self, c2w=None, pos=(0.0, 0.0, 0.0), hFov=50.0, vFov=35.0):
        self.c2w = np.eye(3) if c2w is None else np.asarray(c2w, np.float64).reshape(3, 3)
        self.pos = np.asarray(pos, np.float64)
        self.hFov, self.vFov = float(hFov), float(vFov)
        self.aperture_texture = None
        self.ghost_aperture_texture = None

    def analyze_world_coord(self, pos_world):
        """camera.cpp:245-273: world position -> normalised screen coordinates (ns_x, ns_y)."""
        edge_x = math.tan(0.5 * (self.hFov * (math.pi / 180.0)))
        edge_y = math.tan(0.5 * (self.vFov * (math.pi / 180.0)))
        pos_camera = self.c2w.T @ (np.asarray(pos_world, np.float64) - self.pos)
        pos_image = pos_camera / abs(pos_camera[2])
        return ((pos_image[0] / edge_x) + 1) / 2.0, ((pos_image[1] / edge_y) + 1) / 2.0


class HDRImageBuffer:
    """util/image.h:105-242: w, h and data[x + y*w] = (r, g, b) doubles (Vector3D without AVX)."""

    def __init__(self, w=0, h=0):
        self.resize(w, h)

    def resize(self, w, h):
        self.w, self.h = int(w), int(h)
        self.data = np.zeros((self.h, self.w, 3), np.float64)

    def clear(self):
        self.data[...] = 0

    def get_pixel_value(self, x, y):
        return self.data[y, x]


class PointLight:
    """scene/light.h:49-59, light.cpp:47-48.  The reference's flare code skips it (pathtracer.cpp:35 only takes
    DirectionalLight); here it becomes an lfb_light with a finite distance."""

    def __init__(self, rad, pos):
        self.radiance = np.asarray(rad, np.float64)
        self.position = np.asarray(pos, np.float64)


class PathTracer:
    """The ghost-path members of CGL::PathTracer, backed by the CUDA engine.

    mode / grid_n / pair_set select what generate_ghost_buffer renders: REF_QUADS reproduces
    the reference bit for bit; the grid modes trace the ray bundles the north star asks for.
    """

    def __init__(self, device_id=-1, lens=None, mode=capi.MODE_REF_QUADS, grid_n=256, pair_set=capi.PAIRS_REF,
                 precision=capi.FP32, splat=capi.SPLAT_BILINEAR, include_direct=False):
        self._device_id = device_id
        self._lens = lens if lens is not None else capi.builtin_lens(3)
        self._engine = None  # created on first use: the host-side members work without a device
        self.camera = None
        self.lights = []              # scene->lights
        self.flare_origins = []
        self.flare_radiance = []
        self.flare_distance = []      # lens units per flare origin; 0 = directional (ours: the reference has no point flares)
        self.scene_unit = 1000.0      # lens units (mm) per scene unit, for PointLight distances
        self.axis_ray = (0.0, 0.0)
        self.angle_to_sun = 0.0
        self.ghost_buffer = HDRImageBuffer()
        self._frame = (0, 0)
        self.mode, self.grid_n, self.pair_set = mode, grid_n, pair_set
        self.precision, self.splat, self.include_direct = precision, splat, include_direct
        self._uploaded_texture = None

    @property
    def engine(self):
        if self._engine is None:
            self._engine = capi.Engine(self._device_id)  # raises LfbError(ERR_NO_DEVICE) without a GPU
            self._engine.set_lens(self._lens)
        return self._engine

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def set_frame_size(self, width, height):
        self._frame = (int(width), int(height))

    def find_sun_pos(self):
        """pathtracer.cpp:32-64.  Every on-screen directional light is recorded; like the reference,
        axis_ray / angle_to_sun end up describing the LAST one."""
        for light in self.lights:
            if isinstance(light, DirectionalLight):
                where, distance = light.posLight, 0.0
            elif isinstance(light, PointLight):
                where = light.position
                distance = float(np.linalg.norm(where - np.asarray(self.camera.pos, np.float64))) * self.scene_unit
            else:
                continue
            ns_x, ns_y = self.camera.analyze_world_coord(where)
            if 0 <= ns_x <= 1 and 0 <= ns_y <= 1:
                self.flare_origins.append((ns_x, ns_y))
                self.flare_radiance.append(light.radiance)
                self.flare_distance.append(distance)
                self.angle_to_sun = float(np.float32(math.atan(ns_y / ns_x))) if ns_x != 0 else float(np.float32(math.pi / 2))
                self.axis_ray = (ns_x, ns_y)

    def _lfb_lights(self):
        if self.mode == capi.MODE_REF_QUADS:
            return [capi.make_light(self.axis_ray[0], self.axis_ray[1], theta=self.angle_to_sun)]
        if not self.flare_origins:  # axis_ray set by hand, as the reference's GUI code paths can
            self.flare_origins, self.flare_radiance = [self.axis_ray], [np.ones(3)]
        dist = list(self.flare_distance) + [0.0] * (len(self.flare_origins) - len(self.flare_distance))
        out = []
        for (nx, ny), rad, d in zip(self.flare_origins, self.flare_radiance, dist):
            if self.mode == capi.MODE_EXACT_GRID:  # real refraction needs the real off-axis angle
                theta = capi.physical_theta(nx, ny, self.camera.hFov, self.camera.vFov)
            else:  # the reference's angle_to_sun (pathtracer.cpp:50)
                theta = float(np.float32(math.atan(ny / nx))) if nx != 0 else float(np.float32(math.pi / 2))
            out.append(capi.make_light(nx, ny, theta=theta, radiance=tuple(float(v) for v in rad), distance=d))
        return out

    def raytrace_starburst_frame(self, flare_radius=30.0, flare_intensity=1.0, out=None, additive=False):
        """PathTracer::raytrace_starburst (pathtracer.cpp:947-1004) + calculate_irradiance_falloff (:1030-1052) for every
        pixel of the frame at once; camera.aperture_texture is the mask.  -> (H, W, 3) float64."""
        tex = self.camera.aperture_texture
        if tex is None or tex.aperture is None:
            raise RuntimeError("camera.aperture_texture is not loaded")
        if not self.flare_origins:
            raise RuntimeError("no flare origin: run find_sun_pos() first (the reference dereferences flare_origins[0])")
        if getattr(self, "_uploaded_star", None) is not tex:
            self.engine.set_starburst_aperture(tex.aperture)
            self._uploaded_star = tex
        w, h = self._frame
        lights = [capi.make_light(nx, ny, theta=0.0, radiance=tuple(float(v) for v in rad))
                  for (nx, ny), rad in zip(self.flare_origins, self.flare_radiance)]
        return self.engine.render_starburst(lights, w, h, flare_radius, flare_intensity, out=out, additive=additive)

    def generate_ghost_buffer(self):
        """pathtracer.cpp:714-762: clear + resize ghost_buffer to the frame size, draw every ghost."""
        w, h = self._frame
        self.ghost_buffer.resize(w, h)
        if self.axis_ray[0] == 0 and self.axis_ray[1] == 0:  # :724-726
            return
        tex = self.camera.ghost_aperture_texture
        if tex is None or tex.aperture is None:
            raise RuntimeError("camera.ghost_aperture_texture is not loaded")
        if self._uploaded_texture is not tex:
            self.engine.set_aperture(tex.aperture)
            self._uploaded_texture = tex
        params = capi.make_params(self.mode, w, h, grid_n=self.grid_n, pair_set=self.pair_set, precision=self.precision,
                                  splat=self.splat, include_direct=int(self.include_direct))
        self.engine.render_ghosts(self._lfb_lights(), params, out=self.ghost_buffer.data, elem=capi.F64x3)
