""".u3d uniform-grid sequences (ugc/uniformgrid3dreader.cpp:59-183, ugc/uniformgrid3dwriter.cpp:47-102): header
text as the reference's writer emits it, round trips, and the reader's tolerance for the alternative keys."""
import importlib

import numpy as np
import pytest

from conftest import PKG_NAME


@pytest.fixture(scope="module")
def host():
    return importlib.import_module(PKG_NAME + ".host")


def test_u3d_header_text_and_round_trip(host, tmp_path):
    rng = np.random.default_rng(3)
    grids = rng.random((3, 4, 5, 6), dtype=np.float32)
    model = np.arange(16, dtype=np.float32).reshape(4, 4) / 8      # column-major: model[c][r]
    world = np.eye(4, dtype=np.float32)
    world[3, :3] = (1.5, -2.0, 0.25)                                 # translation column
    p = tmp_path / "seq.u3d"
    host.u3d_write(p, grids, cell=(8, 4, 2), model=model, world=world)
    text = p.read_text().splitlines()
    assert text[0] == "RawFile: seq.raw"
    assert text[1] == "Resolution: 6 5 4 3"
    assert text[2] == "Format: FLOAT32"
    # rows of the matrix, first row on the key's line (the reference writes transpose(columns))
    assert text[3].startswith("ModelMatrix:") and [float(x) for x in text[3].split(":")[1].split()] == list(model[:, 0])
    assert [float(x) for x in text[4].split()] == list(model[:, 1])
    assert text[7].startswith("WorldMatrix:")
    assert text[11] == "CellDimensions: 8 4 2"
    assert (tmp_path / "seq.raw").stat().st_size == grids.nbytes
    assert np.array_equal(np.fromfile(tmp_path / "seq.raw", np.float32), grids.reshape(-1))
    back, cell, m, w = host.u3d_read(p)
    assert np.array_equal(back, grids) and cell == (8, 4, 2)
    assert np.array_equal(m, model) and np.array_equal(w, world)


def test_u3d_minmax_grids_and_alternative_keys(host, tmp_path):
    rng = np.random.default_rng(4)
    mm = rng.integers(0, 65536, size=(2, 3, 3, 3, 2), dtype=np.uint16)
    p = tmp_path / "mm.u3d"
    host.u3d_write(p, mm)
    assert "Format: Vec2UINT16" in p.read_text()
    back, cell, m, w = host.u3d_read(p)
    assert back.dtype == np.uint16 and np.array_equal(back, mm) and cell == (8, 8, 8)
    assert np.array_equal(m, np.eye(4, dtype=np.float32))
    # a hand-written header: comments, blank lines, ObjectFileName / Dimensions spellings, mixed case keys
    (tmp_path / "hand.raw").write_bytes(mm.tobytes())
    (tmp_path / "hand.u3d").write_text("# a comment\n\nObjectFileName: hand.raw\nDIMENSIONS: 3 3 3 2   # trailing comment\n"
                                       "format: Vec2UINT16\n// another comment\nCellDimensions: 4 4 4\n")
    back2, cell2, _, _ = host.u3d_read(tmp_path / "hand.u3d")
    assert np.array_equal(back2, mm) and cell2 == (4, 4, 4)


def test_u3d_errors(host, tmp_path):
    with pytest.raises(host.HostError):
        host.u3d_read(tmp_path / "missing.u3d")
    (tmp_path / "nores.u3d").write_text("RawFile: x.raw\nFormat: FLOAT32\n")
    with pytest.raises(host.HostError, match="Resolution"):
        host.u3d_read(tmp_path / "nores.u3d")
    (tmp_path / "nofmt.u3d").write_text("RawFile: x.raw\nResolution: 2 2 2 1\n")
    with pytest.raises(host.HostError, match="Format"):
        host.u3d_read(tmp_path / "nofmt.u3d")
    (tmp_path / "short.raw").write_bytes(b"\0" * 8)
    (tmp_path / "short.u3d").write_text("RawFile: short.raw\nResolution: 2 2 2 1\nFormat: FLOAT32\n")
    with pytest.raises(host.HostError):
        host.u3d_read(tmp_path / "short.u3d")
    with pytest.raises(ValueError):
        host.u3d_write(tmp_path / "bad.u3d", np.zeros((2, 2, 2), np.float32))


@pytest.mark.gpu
def test_network_exports_sequence_grids(host, cpm, synth, torch_cuda, tmp_path):
    """UniformGrid3DExport of the resident sequence: the written min-max grids are the ones cpm_volume_minmax made"""
    from oracle import orc
    dims, T = (32, 32, 32), 3
    vols = [synth.volume_f32(dims, 4, t / T) for t in range(T)]
    net = host.Network(dims, cpm.CPM_FMT_F32, 32, [(0.3, -0.5, 0.8)], with_importance_grid=True)
    net.set_transfer_function(synth.WS_TF_POINTS)
    net.set_sequence_host(vols)
    net.export_sequence_grids(0, tmp_path / "minmax.u3d")
    net.export_sequence_grids(1, tmp_path / "diff.u3d")
    net.close()
    mm, cell, _, _ = host.u3d_read(tmp_path / "minmax.u3d")
    assert mm.shape == (T, 4, 4, 4, 2) and cell == (8, 8, 8)
    for t in range(T):
        assert np.array_equal(mm[t], orc.volume_minmax(vols[t], 8))
    diff, _, _, _ = host.u3d_read(tmp_path / "diff.u3d")
    assert diff.shape == (T, 4, 4, 4) and diff.max() > 0
