"""Deterministic scenes for the path-traced scene pass, as plain arrays (the COLLADA loader is out of scope).

BASELINE config 5 names dae/dragon.dae, which is not in the reference tree (.MISSING_LARGE_BLOBS:1); the substitute follows
dae/pyramid.dae -- its camera (xfov 39.6 deg, 16:9), its directional sun (1, 0.944, 0.544) and two point lights, its sandy
lambert materials -- with our own geometry: pyramids on a ground plane, two spheres, a small emissive panel.
Array layouts: oracle/ref_shim.cpp ref_scene_radiance."""
import math

import numpy as np


def _tri(out_p, out_n, out_m, a, b, c, mat, normals=None):
    a, b, c = (np.asarray(v, float) for v in (a, b, c))
    n = np.cross(b - a, c - a)
    n = n / np.linalg.norm(n)
    out_p.append([a, b, c])
    out_n.append(normals if normals is not None else [n, n, n])
    out_m.append(mat)


def pyramid_scene(n_pyramids=3, smooth=False):
    P, N, M = [], [], []
    # ground (material 0): two triangles
    g = 12.0
    _tri(P, N, M, (-g, 0, -g), (g, 0, g), (g, 0, -g), 0)
    _tri(P, N, M, (-g, 0, -g), (-g, 0, g), (g, 0, g), 0)
    # pyramids (materials 1..): square base, apex above
    for k in range(n_pyramids):
        cx, cz, s, h = -4.0 + 4.0 * k, -1.0 - 1.5 * k, 1.6 - 0.3 * k, 2.4 - 0.5 * k
        base = [(cx - s, 0, cz - s), (cx + s, 0, cz - s), (cx + s, 0, cz + s), (cx - s, 0, cz + s)]
        apex = (cx, h, cz)
        for q in range(4):
            a, b = base[q], base[(q + 1) % 4]
            if smooth:  # vertex normals pointing away from the axis: exercises the normal interpolation
                na = np.array([a[0] - cx, 0.6 * s, a[2] - cz]); nb = np.array([b[0] - cx, 0.6 * s, b[2] - cz]); nc = np.array([0.0, 1.0, 0.0])
                _tri(P, N, M, b, a, apex, 1 + k % 4, normals=[nb / np.linalg.norm(nb), na / np.linalg.norm(na), nc])
            else:
                _tri(P, N, M, b, a, apex, 1 + k % 4)
    # an emissive panel (material 5) facing the camera
    _tri(P, N, M, (5.0, 0.5, -6.0), (6.5, 0.5, -6.0), (6.5, 1.5, -6.0), 5)
    _tri(P, N, M, (5.0, 0.5, -6.0), (6.5, 1.5, -6.0), (5.0, 1.5, -6.0), 5)
    spheres = np.array([[1.8, 0.7, 1.5, 0.7], [-2.2, 0.45, 2.2, 0.45]], float)
    sph_mat = np.array([2, 4], np.int32)
    mats = np.array([[0.651406, 0.4072403, 0.1559266, 0, 0, 0], [0.287441, 0.1274377, 0.03189605, 0, 0, 0],
                     [0.5775806, 0.3005437, 0.08865562, 0, 0, 0], [0.4286906, 0.2232279, 0.1046165, 0, 0, 0],
                     [0.7605246, 0.456411, 0.223228, 0, 0, 0], [0, 0, 0, 4.0, 3.0, 1.5]], float)
    lights = np.array([[0, 1.0, 0.944, 0.544, -0.45, -0.8, -0.35],          # directional: the direction the light travels
                       [1, 3.859998, 3.119712, 1.426067, -3.0, 3.5, 3.0],   # point lights
                       [1, 10.0, 10.0, 10.0, 6.0, 6.0, 4.0]], float)
    return dict(tri_pos=np.array(P, float), tri_nrm=np.array(N, float), tri_mat=np.array(M, np.int32), spheres=spheres, sph_mat=sph_mat,
                mats=mats, lights=lights)


def look_at(eye, target, hfov_deg=39.59775, aspect=16.0 / 9.0, nclip=0.1, fclip=100.0):
    """cam[16] = pos, c2w rows (camera looks down -z, camera.cpp:278-305), hFov, vFov in degrees, clips."""
    eye, target = np.asarray(eye, float), np.asarray(target, float)
    zc = eye - target
    zc /= np.linalg.norm(zc)
    xc = np.cross([0.0, 1.0, 0.0], zc)
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    c2w = np.stack([xc, yc, zc], axis=1)
    vfov = 2 * math.degrees(math.atan(math.tan(math.radians(hfov_deg) / 2) / aspect))
    return np.concatenate([eye, c2w.reshape(-1), [hfov_deg, vfov, nclip, fclip]])


def default_camera():
    return look_at((0.5, 3.2, 9.5), (0.0, 0.8, -1.0))
