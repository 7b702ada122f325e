"""Seeded synthetic inputs for tests and benchmarks: heterogeneous volumes,
the workspace transfer function, the scene proxy cube and light set-ups.  numpy only; every
random draw comes from splitmix64 so that a (seed, shape) pair names one input everywhere."""
from __future__ import annotations

import numpy as np


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n uint64 outputs of splitmix64 started at `seed`."""
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, n + 1, dtype=np.uint64))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int) -> np.ndarray:
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _interp_matrix(n_out: int, n_in: int) -> np.ndarray:
    """(n_out, n_in) linear interpolation matrix"""
    x = np.linspace(0, n_in - 1, n_out)
    i0 = np.minimum(np.floor(x).astype(int), n_in - 2)
    a = x - i0
    m = np.zeros((n_out, n_in), np.float32)
    m[np.arange(n_out), i0] = 1 - a
    m[np.arange(n_out), i0 + 1] = a
    return m


def volume_field(dims, seed: int, t: float = 0.0, n_blobs: int = 12, noise_res: int = 9) -> np.ndarray:
    """float32 field in [0,1], shape (nz, ny, nx): Gaussian blobs (on closed orbits in t, period 1)
    plus smooth value noise."""
    nx, ny, nz = dims
    r = uniform01(seed, n_blobs * 8 + noise_res ** 3)
    b = r[: n_blobs * 8].reshape(n_blobs, 8)
    noise = r[n_blobs * 8:].reshape(noise_res, noise_res, noise_res).astype(np.float32)
    xs = (np.arange(nx, dtype=np.float32) + 0.5) / nx
    ys = (np.arange(ny, dtype=np.float32) + 0.5) / ny
    zs = (np.arange(nz, dtype=np.float32) + 0.5) / nz
    gx = np.empty((n_blobs, nx), np.float32)
    gy = np.empty((n_blobs, ny), np.float32)
    gz = np.empty((n_blobs, nz), np.float32)
    amp = np.empty(n_blobs, np.float32)
    for k in range(n_blobs):
        cx, cy, cz, rad, a, orad, oph, _ = b[k]
        ang = 2 * np.pi * (t + oph)
        cx = 0.15 + 0.7 * cx + 0.12 * orad * np.cos(ang)
        cy = 0.15 + 0.7 * cy + 0.12 * orad * np.sin(ang)
        cz = 0.15 + 0.7 * cz
        s = 0.04 + 0.08 * rad
        gx[k] = np.exp(-0.5 * ((xs - cx) / s) ** 2)
        gy[k] = np.exp(-0.5 * ((ys - cy) / s) ** 2)
        gz[k] = np.exp(-0.5 * ((zs - cz) / s) ** 2)
        amp[k] = 0.35 + 0.65 * a
    f = np.einsum("k,kz,ky,kx->zyx", amp, gz, gy, gx, optimize=True).astype(np.float32)
    mz, my, mx = _interp_matrix(nz, noise_res), _interp_matrix(ny, noise_res), _interp_matrix(nx, noise_res)
    nfield = np.einsum("zc,yb,xa,cba->zyx", mz, my, mx, noise, optimize=True)
    f = 0.93 * np.minimum(f, 1.0) + 0.07 * nfield
    return np.clip(f, 0.0, 1.0).astype(np.float32)


def volume_u8(dims, seed):
    return np.ascontiguousarray(np.rint(volume_field(dims, seed) * 255.0).astype(np.uint8))


def volume_u16(dims, seed):
    return np.ascontiguousarray(np.rint(volume_field(dims, seed) * 65535.0).astype(np.uint16))


def volume_f32(dims, seed, t=0.0):
    return np.ascontiguousarray(volume_field(dims, seed, t))


# The six transfer-function points of workspaces/CorrelatedPhotonMappingSingleVolume.inv:662-687
WS_TF_POINTS = [
    (0.01686747, (1.0, 0.59633785, 0.24313726, 0.0)),
    (0.036445361, (0.90980393, 0.49831387, 0.29256123, 0.0)),
    (0.073654622, (0.93725491, 0.58783853, 0.48149845, 0.0)),
    (0.22178316, (0.6156863, 0.25906768, 0.10623331, 0.18884119)),
    (0.28514057, (0.93725491, 0.1506981, 0.25557336, 0.39484981)),
    (0.67068273, (0.10786603, 0.61843288, 0.65490198, 0.53218883)),
]


def rasterise_tf(points=WS_TF_POINTS, width: int = 1024) -> np.ndarray:
    """(width, 4) float32 RGBA: piecewise-linear between points, constant outside."""
    pos = np.array([p[0] for p in points], np.float64)
    col = np.array([p[1] for p in points], np.float64)
    x = np.arange(width, dtype=np.float64) / (width - 1)
    out = np.stack([np.interp(x, pos, col[:, c]) for c in range(4)], axis=1)
    return np.ascontiguousarray(out.astype(np.float32))


def dense_tf(alpha: float, width: int = 256) -> np.ndarray:
    """constant-opacity transfer function (homogeneous medium tests)"""
    tf = np.zeros((width, 4), np.float32)
    tf[:, :3] = 1.0
    tf[:, 3] = alpha
    return tf


# Scene proxy: the unit cube in texture space, 8 vertices / 12 triangles
CUBE_VERTICES = np.array([[x, y, z] for z in (0.0, 1.0) for y in (0.0, 1.0) for x in (0.0, 1.0)], np.float32)
CUBE_INDICES = np.array([
    0, 2, 1, 1, 2, 3,  # z = 0
    4, 5, 6, 5, 7, 6,  # z = 1
    0, 1, 4, 1, 5, 4,  # y = 0
    2, 6, 3, 3, 6, 7,  # y = 1
    0, 4, 2, 2, 4, 6,  # x = 0
    1, 3, 5, 3, 7, 5,  # x = 1
], np.int32)


def normalize(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def volume_field_torch(dims, seed: int, t: float = 0.0, device="cpu", n_blobs: int = 12, noise_res: int = 9):
    """volume_field() evaluated with torch on `device` (benchmark-sized volumes: a 512^3 step takes
    seconds in numpy).  Same blobs, orbits and noise lattice; values agree with volume_field() to
    fp32 rounding, which is all a benchmark input needs.  Returns a (nz, ny, nx) float32 tensor."""
    import torch
    nx, ny, nz = dims
    r = uniform01(seed, n_blobs * 8 + noise_res ** 3)
    b = r[: n_blobs * 8].reshape(n_blobs, 8)
    noise = torch.tensor(r[n_blobs * 8:].reshape(noise_res, noise_res, noise_res), dtype=torch.float32, device=device)
    ax = [(torch.arange(n, dtype=torch.float32, device=device) + 0.5) / n for n in (nx, ny, nz)]
    cx, cy, cz, rad, a, orad, oph = (b[:, k] for k in range(7))
    ang = 2 * np.pi * (t + oph)
    centre = [0.15 + 0.7 * cx + 0.12 * orad * np.cos(ang), 0.15 + 0.7 * cy + 0.12 * orad * np.sin(ang), 0.15 + 0.7 * cz]
    s = torch.tensor(0.04 + 0.08 * rad, dtype=torch.float32, device=device)[:, None]
    g = [torch.exp(-0.5 * ((ax[k][None, :] - torch.tensor(centre[k], dtype=torch.float32, device=device)[:, None]) / s) ** 2)
         for k in range(3)]
    amp = torch.tensor(0.35 + 0.65 * a, dtype=torch.float32, device=device)
    f = torch.einsum("kz,ky,kx->zyx", g[2] * amp[:, None], g[1], g[0])
    m = [torch.tensor(_interp_matrix(n, noise_res), device=device) for n in (nx, ny, nz)]
    nf = torch.einsum("zc,cyx->zyx", m[2], torch.einsum("yb,cbx->cyx", m[1], torch.einsum("xa,cba->cbx", m[0], noise)))
    f = 0.93 * torch.clamp(f, max=1.0) + 0.07 * nf
    return torch.clamp(f, 0.0, 1.0).contiguous()
