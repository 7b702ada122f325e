"""Sequence players of the uniformgridcl module: org.inviwo.UniformGrid3DPlayerProcessor
(ugc/processors/uniformgrid3dplayerprocessor.cpp:87-152) and org.inviwo.VolumeSequencePlayer
(ugc/processors/volumesequenceplayer.cpp:94-180): time -> index bookkeeping (host), interpolation on the device."""
import ctypes as C
import importlib

import numpy as np
import pytest

from conftest import PKG_NAME


@pytest.fixture(scope="module")
def host():
    return importlib.import_module(PKG_NAME + ".host")


def _fma_mix(x, y, a):
    """x + (y - x) * a with one rounding at the end (mix(): fma), via float64"""
    return (x.astype(np.float64) + (y - x).astype(np.float64) * np.float64(np.float32(a))).astype(np.float32)


def test_player_contract(host):
    got = host.describe_processors()
    assert got["org.inviwo.UniformGrid3DPlayerProcessor"] == ({"Sequence", "InterpolatedData"},
                                                              {"time", "selectedSequenceIndex", "timePerElement", "frameRate", "playSequence"})
    assert got["org.inviwo.VolumeSequencePlayer"] == ({"volumeSequence", "InterpolatedVolume"},
                                                      {"time", "selectedSequenceIndex", "timePerVolume", "volumesPerSecond", "playSequence"})


def test_player_clock_follows_the_reference_timer(host):
    """onSequenceTimerEvent: time += (1000 / frameRate) / 1000 with INTEGER milliseconds (frame rate 30 -> 33 ms, not 33.3),
    wrapped by subtracting the maximum (n - 1) * timePerElement; index = floor(time / timePerElement) % n + 1"""
    n, tpe, rate, ticks = 5, 0.25, 30, 200
    t_out, i_out = np.zeros(ticks, np.float32), np.zeros(ticks, np.int32)
    assert host.lib().cpmh_player_clock(n, C.c_float(tpe), rate, ticks, t_out.ctypes.data_as(C.c_void_p), i_out.ctypes.data_as(C.c_void_p)) == 0
    t, tmax, dt = np.float32(0), np.float32((n - 1) * tpe), np.float32(np.float32(1000 // rate) / np.float32(1000))
    for k in range(ticks):
        t = np.float32(t + dt)
        if t > tmax:
            t = np.float32(t - tmax)
        assert t_out[k] == t, k
        assert i_out[k] == int(np.floor(np.float32(t / np.float32(tpe)))) % n + 1, k
    assert set(i_out.tolist()) == {1, 2, 3, 4}        # the last element is reached only as the "next" of n - 1


@pytest.mark.gpu
def test_grid_player_interpolates_and_ping_pongs(host, torch_cuda):
    rng = np.random.default_rng(4)
    n_grids, n_cells, tpe = 4, 4099, 0.5
    grids = rng.random((n_grids, n_cells)).astype(np.float32)
    times = np.array([0.0, 0.1, 0.49, 0.5, 0.75, 1.3, 1.5], np.float32)
    out = np.zeros((len(times), n_cells), np.float32)
    idx, buf = np.zeros(len(times), np.int32), np.zeros(len(times), np.int32)
    rc = host.lib().cpmh_player_grids_f32(grids.ctypes.data_as(C.c_void_p), n_grids, C.c_size_t(n_cells), C.c_float(tpe),
                                          times.ctypes.data_as(C.c_void_p), len(times), out.ctypes.data_as(C.c_void_p),
                                          idx.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))
    assert rc == 0, host.lib().cpmh_last_error()
    for k, t in enumerate(times):
        whole = np.floor(np.float32(t / np.float32(tpe)))
        w = np.float32(np.float32(t / np.float32(tpe)) - whole)
        step = int(whole) % n_grids
        assert idx[k] == step + 1
        want = _fma_mix(grids[step], grids[(step + 1) % n_grids], w)
        assert np.abs(out[k].view(np.int32) - want.view(np.int32)).max() <= 1, k        # (float64 emulation of the fma: <= 1 ulp)
    assert buf.tolist() == [k % 2 for k in range(len(times))]        # two output grids alternate


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["u8", "f32"])
def test_volume_player_interpolates(host, cpm, synth, torch_cuda, fmt):
    dims = (24, 20, 16)
    make = synth.volume_u8 if fmt == "u8" else (lambda d, s: synth.volume_f32(d, s, 0.1 * s))
    vols = np.stack([make(dims, s) for s in (1, 2, 3)])
    times = np.array([0.0, 0.3, 1.0, 1.6], np.float32)
    out = np.zeros((len(times),) + vols.shape[1:], vols.dtype)
    idx = np.zeros(len(times), np.int32)
    rc = host.lib().cpmh_player_volumes(vols.ctypes.data_as(C.c_void_p), 3, (C.c_int * 3)(*dims), cpm.CPM_FMT_U8 if fmt == "u8" else cpm.CPM_FMT_F32,
                                        C.c_float(1.0), times.ctypes.data_as(C.c_void_p), len(times), out.ctypes.data_as(C.c_void_p),
                                        idx.ctypes.data_as(C.c_void_p))
    assert rc == 0, host.lib().cpmh_last_error()
    for k, t in enumerate(times):
        step = int(np.floor(t)) % 3
        w = np.float32(t - np.floor(t))
        assert idx[k] == step + 1
        a, b = vols[step], vols[(step + 1) % 3]
        if fmt == "f32":
            want = _fma_mix(a.reshape(-1), b.reshape(-1), w).reshape(a.shape)
            assert np.abs(out[k].view(np.int32) - want.view(np.int32)).max() <= 1
        else:
            m = a.astype(np.float64) / 255 + (b.astype(np.float64) / 255 - a.astype(np.float64) / 255) * float(w)
            want = np.rint(np.clip(m, 0, 1) * 255)
            # float32 on the device, float64 here: they may round differently only where the mix lands on a tie (x.5)
            assert np.abs(out[k].astype(np.int32) - want.astype(np.int32)).max() <= 1
            mism = out[k] != want
            assert np.all(np.abs(np.abs(m * 255 - np.floor(m * 255)) - 0.5)[mism] < 1e-3)
            assert mism.mean() < 0.1
