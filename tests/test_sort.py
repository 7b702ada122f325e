"""Selection stage: threshold / count / iota and the onesweep radix sort (bit-exact, stable)."""
import numpy as np
import pytest


def importance_like_keys(synth, n, seed, frac=0.1):
    """the real key distribution: most photons keep 0x7FFFFFFF, a fraction carries
    0x7FFFFFFF - ceil(100*importance) (ppm/cl/photonrecomputationdetector.cl:156)"""
    u = synth.uniform01(seed, 2 * n)
    keys = np.full(n, 0x7FFFFFFF, np.uint32)
    sel = u[:n] < frac
    imp = np.ceil(100.0 * -np.log(1.0 - u[n:][sel]) * 3.0).astype(np.uint32)
    keys[sel] = np.uint32(0x7FFFFFFF) - imp
    return keys


def key_sets(synth, n, seed):
    yield "uniform", (synth.splitmix64(seed, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    yield "importance", importance_like_keys(synth, n, seed + 1)
    yield "constant", np.full(n, 0x7FFFFFFF, np.uint32)
    yield "few_values", ((synth.splitmix64(seed + 2, n) % np.uint64(5)) * np.uint64(0x01010101)).astype(np.uint32)
    yield "descending", np.arange(n, 0, -1).astype(np.uint32)


# ------------------------------------------------------------------------- CPU: oracle pinned --
def test_oracle_sorts_match_numpy_stable(orc, synth):
    for n in (1, 2, 17, 4096, 100_003):
        for name, keys in key_sets(synth, n, 2):
            perm = np.argsort(keys, kind="stable").astype(np.uint32)
            for fn in (orc.radix_sort, orc.merge_sort):
                k, v = keys.copy(), np.arange(n, dtype=np.uint32)
                fn(k, v)
                assert np.array_equal(k, keys[perm]), (name, n)
                assert np.array_equal(v, perm), (name, n)
                k2 = keys.copy()
                fn(k2)
                assert np.array_equal(k2, keys[perm])


def test_oracle_max_bits(orc, synth):
    keys = (synth.splitmix64(9, 5000) & np.uint64(0xFFF)).astype(np.uint32)
    perm = np.argsort(keys, kind="stable").astype(np.uint32)
    k, v = keys.copy(), np.arange(5000, dtype=np.uint32)
    orc.radix_sort(k, v, max_bits=12)   # 3 four-bit passes: odd -> copy back path
    assert np.array_equal(v, perm)


# ------------------------------------------------------------------------- GPU ------------------
def _dev(torch, a):
    return torch.from_numpy(a.view(np.int32)).cuda()


def _host(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 33, 4095, 4096, 4097, 8191, 8192, 8193, 65_536, 1_000_003])
def test_cuda_radix_sort_pairs_bit_exact(orc, synth, ctx, torch_cuda, n):
    torch = torch_cuda
    for name, keys in key_sets(synth, n, 2):
        want_k, want_v = keys.copy(), np.arange(n, dtype=np.uint32)
        orc.radix_sort(want_k, want_v)
        k, v = _dev(torch, keys), _dev(torch, np.arange(n, dtype=np.uint32))
        tk, tv = torch.empty_like(k), torch.empty_like(v)
        ctx.radix_sort(k, v, tk, tv)
        ctx.sync()
        assert np.array_equal(_host(k), want_k), (name, n)
        assert np.array_equal(_host(v), want_v), (name, n)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 1000, 262_144 + 17])
def test_cuda_radix_sort_keys_only_and_max_bits(orc, synth, ctx, torch_cuda, n):
    torch = torch_cuda
    keys = (synth.splitmix64(4, n) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    for bits in (0, 32, 24, 20, 16, 9, 8, 3):
        kk = keys if bits in (0, 32) else (keys & np.uint32((1 << bits) - 1))
        want = np.sort(kk, kind="stable")
        k = _dev(torch, kk)
        tk = torch.empty_like(k)
        ctx.radix_sort(k, None, tk, None, max_bits=bits)
        ctx.sync()
        assert np.array_equal(_host(k), want), bits
        # with values as well (odd pass counts exercise the copy-back)
        perm = np.argsort(kk, kind="stable").astype(np.uint32)
        k, v = _dev(torch, kk), _dev(torch, np.arange(n, dtype=np.uint32))
        tk, tv = torch.empty_like(k), torch.empty_like(v)
        ctx.radix_sort(k, v, tk, tv, max_bits=bits)
        ctx.sync()
        assert np.array_equal(_host(v), perm), bits


@pytest.mark.gpu
@pytest.mark.parametrize("n", [(1 << 24) + 5])
def test_cuda_radix_sort_large(orc, synth, ctx, torch_cuda, n):
    """many tiles in flight: the chained scan's look-back windows cross hundreds of predecessors"""
    torch = torch_cuda
    for name, keys in key_sets(synth, n, 7):
        if name in ("few_values", "descending"):
            continue
        want_k, want_v = keys.copy(), np.arange(n, dtype=np.uint32)
        orc.radix_sort(want_k, want_v)
        k, v = _dev(torch, keys), _dev(torch, np.arange(n, dtype=np.uint32))
        tk, tv = torch.empty_like(k), torch.empty_like(v)
        for _ in range(3):     # repeated runs reuse the status scratch
            k.copy_(_dev(torch, keys)); v.copy_(_dev(torch, np.arange(n, dtype=np.uint32)))
            ctx.radix_sort(k, v, tk, tv)
        ctx.sync()
        assert np.array_equal(_host(k), want_k), name
        assert np.array_equal(_host(v), want_v), name


@pytest.mark.gpu
def test_cuda_radix_sort_prefix_of_buffer(orc, synth, ctx, torch_cuda):
    """sort #2 of the reference sorts only the first nRecomputed indices of a larger buffer
    (ppm/processor/progressivephotontracercl.cpp:467-473)"""
    torch = torch_cuda
    n, m = 50_000, 12_345
    idx = (synth.splitmix64(6, n) % np.uint64(1 << 22)).astype(np.uint32)
    d = _dev(torch, idx)
    t = torch.empty_like(d)
    ctx.radix_sort(d, None, t, None, n=m)
    ctx.sync()
    out = _host(d)
    assert np.array_equal(out[:m], np.sort(idx[:m]))
    assert np.array_equal(out[m:], idx[m:])


@pytest.mark.gpu
def test_cuda_radix_sort_errors(cpm, ctx, torch_cuda):
    torch = torch_cuda
    k = torch.zeros(16, dtype=torch.int32, device="cuda")
    with pytest.raises(cpm.CpmError):      # clogs: "elements is zero"
        ctx.radix_sort(k, None, k.clone(), None, n=0)
    with pytest.raises(cpm.CpmError):      # clogs: "maxBits is too large"
        ctx.radix_sort(k, None, k.clone(), None, max_bits=33)


@pytest.mark.gpu
def test_cuda_threshold_count_iota(orc, synth, ctx, torch_cuda):
    torch = torch_cuda
    for n in (1, 1000, 1_234_567):
        keys = importance_like_keys(synth, n, 12)
        d = _dev(torch, keys)
        out = torch.empty_like(d)
        ctx.threshold(d, 2147483647, out)
        io = torch.empty_like(d)
        ctx.iota(io)
        ctx.sync()
        assert np.array_equal(_host(out), (keys < 2147483647).astype(np.uint32))
        assert np.array_equal(_host(io), np.arange(n, dtype=np.uint32))
        want = orc.count_below(keys, 2147483647)
        assert ctx.reduce_sum_i32(out) == want
        io2 = torch.zeros_like(d)
        assert ctx.count_below(d, 2147483647, io2) == want
        assert np.array_equal(_host(io2), np.arange(n, dtype=np.uint32))
        assert ctx.count_below(d, 2147483647) == want
    neg = torch.tensor([-5, 3, -1], dtype=torch.int32, device="cuda")
    assert ctx.reduce_sum_i32(neg) == -3


@pytest.mark.gpu
def test_cuda_select_below_equals_sort_select_sort(orc, synth, ctx, torch_cuda):
    """cpm_select_below == threshold/count + sort ids by importance + cut at the count + keys-only id sort
    (ppm/processor/progressivephotontracercl.cpp:318-473 with a budget that covers every invalid photon)"""
    torch = torch_cuda
    for n, seed in ((1, 1), (31, 2), (4096, 3), (4097, 4), (1_234_567, 5), (1 << 22, 6)):
        keys = importance_like_keys(synth, n, seed)
        # reference pipeline on the oracle
        cnt = orc.count_below(keys, 2147483647)
        sk, ids = keys.copy(), np.arange(n, dtype=np.uint32)
        orc.radix_sort(sk, ids)
        sel = np.ascontiguousarray(ids[:cnt])
        if cnt:
            orc.radix_sort(sel, None)
        d = _dev(torch, keys)
        out = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        got = ctx.select_below(d, 2147483647, out)
        ctx.sync()
        assert got == cnt
        h = _host(out)
        assert np.array_equal(h[:cnt], sel)
        assert (h[cnt:] == 0xFFFFFFFF).all()          # entries past the count are untouched
    # all / none selected
    d = _dev(torch, np.zeros(5000, np.uint32))
    out = torch.zeros(5000, dtype=torch.int32, device="cuda")
    assert ctx.select_below(d, 1, out) == 5000 and np.array_equal(_host(out), np.arange(5000, dtype=np.uint32))
    assert ctx.select_below(d, 0, out) == 0
