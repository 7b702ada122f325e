"""MWC64X: oracle pinned against the reference's own sources; CUDA bit-exact against the oracle."""
import ctypes as C
import json

import numpy as np
import pytest

from conftest import GOLDEN

A = 4294883355
M = 18446383549859758079
BASEID = 4077358422479273989


def golden():
    return json.loads((GOLDEN / "mwc64x.json").read_text())


# ---------------------------------------------------------------- CPU: oracle vs reference --
def test_constants():
    assert M == A * 2 ** 32 - 1                      # rng/cl/random.cl:46-47


def test_oracle_matches_golden_reference_vectors(orc):
    g = golden()
    base = orc.rng_host_base_offsets(0, 64)
    assert base[:, 0].tolist() == g["seed0_base"]
    state = orc.rng_seed_streams(base.copy())
    assert state.tolist() == g["seed0_state"]
    s = state.copy()
    for k in range(3):
        r = orc.rng_uniform(s, 1)[:, 0]
        assert [float(x).hex() for x in r] == g["seed0_random01_hex"][k]
    assert s.tolist() == g["seed0_state_after3"]
    for gap in (1, 12345, 1 << 40, (1 << 63) + 5):
        st = np.array([[b, 0] for b in g["gap_bases"]], np.uint32)
        orc.rng_seed_streams(st, gap=gap)
        assert st.tolist() == g[f"gap_{gap}"]
    for item in g["steps"]:
        x, c = C.c_uint32(item["x"]), C.c_uint32(item["c"])
        for want in item["seq"]:
            orc.lib().orc_rng_step(C.byref(x), C.byref(c))
            assert [x.value, c.value] == want


def test_oracle_matches_reference_library_live(orc):
    """Same comparison against oracle/_ref (the reference sources compiled here) on fresh inputs."""
    ref = orc.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libmwc64x_ref.so not built (reference tree absent)")
    rng = np.random.default_rng(7)
    n = 257
    base = np.zeros((n, 2), np.uint32)
    base[:, 0] = rng.integers(0, 2 ** 31, n)
    want = base.copy()
    ref.ref_generate_random_state(want.ctypes.data_as(C.c_void_p), n)
    got = orc.rng_seed_streams(base.copy())
    assert np.array_equal(got, want)
    o_ref = np.zeros(n, np.float32)
    ref.ref_random_number_generator(want.ctypes.data_as(C.c_void_p), n, o_ref.ctypes.data_as(C.c_void_p))
    o = orc.rng_uniform(got, 1)[:, 0]
    assert np.array_equal(o.view(np.uint32), o_ref.view(np.uint32))
    assert np.array_equal(got, want)


def test_skip_ahead_identity(orc):
    """x -> x*A mod M is one MWC step, so SeedStreams(dist = d + 1) == Step(SeedStreams(dist = d)).
    Ties the seeding code to the step function without any external vector."""
    s0 = np.array([[1000, 0]], np.uint32)
    s1 = np.array([[1001, 0]], np.uint32)
    orc.rng_seed_streams(s0, gap=0)
    orc.rng_seed_streams(s1, gap=0)
    x, c = C.c_uint32(int(s0[0, 0])), C.c_uint32(int(s0[0, 1]))
    orc.lib().orc_rng_step(C.byref(x), C.byref(c))
    assert [x.value, c.value] == s1[0].tolist()
    # and the closed form: state = BASEID * A^dist mod M split as (x/A, x%A)
    v = BASEID * pow(A, 1000, M) % M
    assert s0[0].tolist() == [v // A, v % A]


def test_random01_range(orc):
    st = orc.rng_seed_streams(orc.rng_host_base_offsets(0, 4096))
    r = orc.rng_uniform(st, 16)
    assert r.min() >= 0.0 and r.max() <= 1.0
    assert abs(float(r.mean()) - 0.5) < 0.01


# ---------------------------------------------------------------- GPU: CUDA vs oracle -------
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 31, 4096, 100_003])
def test_cuda_seed_streams_bit_exact(cpm, orc, ctx, torch_cuda, n):
    torch = torch_cuda
    base = cpm.capi.rng_host_base_offsets(0, n)
    want = orc.rng_seed_streams(base.copy())
    st = torch.from_numpy(base.view(np.int32)).cuda()
    ctx.rng_seed_streams(st, n)
    ctx.sync()
    got = st.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_cuda_seed_streams_golden(cpm, ctx, torch_cuda):
    """straight against the vectors produced by the reference's own source"""
    torch = torch_cuda
    g = golden()
    base = cpm.capi.rng_host_base_offsets(0, 64)
    st = torch.from_numpy(base.view(np.int32)).cuda()
    ctx.rng_seed_streams(st, 64)
    out = torch.empty(64, dtype=torch.float32, device="cuda")
    ctx.sync()
    assert st.cpu().numpy().view(np.uint32).tolist() == g["seed0_state"]
    for k in range(3):
        ctx.rng_uniform(st, out, 1)
        ctx.sync()
        assert [float(x).hex() for x in out.cpu().numpy()] == g["seed0_random01_hex"][k]
    assert st.cpu().numpy().view(np.uint32).tolist() == g["seed0_state_after3"]
    for gap in (1, 12345, 1 << 40, (1 << 63) + 5):
        b = np.array([[v, 0] for v in g["gap_bases"]], np.uint32)
        s = torch.from_numpy(b.view(np.int32)).cuda()
        ctx.rng_seed_streams(s, len(b), gap=gap)
        ctx.sync()
        assert s.cpu().numpy().view(np.uint32).tolist() == g[f"gap_{gap}"]


@pytest.mark.gpu
def test_cuda_seed_streams_sharded(cpm, orc, ctx, torch_cuda):
    """first_stream: a GPU seeding streams [k, k+m) gets exactly the slice of the full set"""
    torch = torch_cuda
    n, k, m = 5000, 1234, 777
    base = cpm.capi.rng_host_base_offsets(0, n)
    want = orc.rng_seed_streams(base.copy())[k:k + m]
    st = torch.from_numpy(base[k:k + m].copy().view(np.int32)).cuda()
    ctx.rng_seed_streams(st, m, first_stream=k)
    ctx.sync()
    assert np.array_equal(st.cpu().numpy().view(np.uint32), want)


@pytest.mark.gpu
def test_cuda_uniform_bit_exact(cpm, orc, ctx, torch_cuda):
    torch = torch_cuda
    n, per = 10_000, 7
    state = orc.rng_seed_streams(orc.rng_host_base_offsets(0, n))
    st = torch.from_numpy(state.copy().view(np.int32)).cuda()
    out = torch.empty(n * per, dtype=torch.float32, device="cuda")
    ctx.rng_uniform(st, out, per)
    ctx.sync()
    want = orc.rng_uniform(state, per)
    assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(n, per), want.view(np.uint32))
    assert np.array_equal(st.cpu().numpy().view(np.uint32), state)
