"""Small seeded scenes shared by the parity tests (built with the oracle's emission path)."""
import importlib

import numpy as np

from conftest import PKG_NAME
from oracle import orc

synth = importlib.import_module(PKG_NAME + ".synth")


def make_volume(dims, fmt, seed):
    if fmt == "u8":
        return synth.volume_u8(dims, seed)
    if fmt == "u16":
        return synth.volume_u16(dims, seed)
    return synth.volume_f32(dims, seed)


def directional_light(n_side, direction=(0.3, -0.5, 0.8), radiance=(1.0, 0.9, 0.8)):
    """light samples + intersections for a directional light fitted to the unit-cube proxy"""
    d = synth.normalize(direction)
    # a point far behind the volume along -d, as baseLightToPackedLight's position would be
    plane_point = np.array([0.5, 0.5, 0.5], np.float32) - 2.0 * d
    o, u, v = orc.fit_light_plane(synth.CUBE_VERTICES, plane_point, d)
    area = float(np.float32(np.linalg.norm(u)) * np.float32(np.linalg.norm(v)))
    n = n_side * n_side
    samples = orc.sample_uniform2d(float(n_side), float(n_side), n)
    ls = orc.light_sample_directional(samples, radiance, d, o, u, v, area)
    isect = orc.light_mesh_intersect(synth.CUBE_VERTICES, synth.CUBE_INDICES, ls)
    return dict(samples=samples, light_samples=ls, isect=isect, dir=d, origin=o, u=u, v=v, area=area,
                radiance=radiance, n=n)


def point_light(n_side, position=(0.5, 0.5, -1.0), radiance=(1.0, 1.0, 1.0)):
    n = n_side * n_side
    samples = orc.sample_uniform2d(float(n_side), float(n_side), n)
    ls = orc.light_sample_point(samples, radiance, position)
    isect = orc.light_mesh_intersect(synth.CUBE_VERTICES, synth.CUBE_INDICES, ls)
    return dict(samples=samples, light_samples=ls, isect=isect, n=n, position=position, radiance=radiance)


def rng_states(n, seed=0):
    return orc.rng_seed_streams(orc.rng_host_base_offsets(seed, n))
