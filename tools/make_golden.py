"""Generate tests/golden/*.json.

mwc64x.json comes from the REFERENCE's own MWC64X sources compiled for the host
(oracle/_ref/libmwc64x_ref.so, built by oracle/Makefile from /root/reference).  Run in the
build container (the reference tree does not exist on the GPU box); the JSON is committed.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402


def mwc64x():
    ref = orc.ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libmwc64x_ref.so missing: run `make -C oracle ref` first")
    out = {"source": "reference rng/cl/{skip_mwc,random,randstategen,randomnumbergenerator}.cl compiled via oracle/Makefile"}
    # (a) the reference pipeline: srand(0); base_i = rand(); GenerateRandomState; 3 x random_01
    n = 64
    base = orc.rng_host_base_offsets(0, n)
    state = base.copy()
    ref.ref_generate_random_state(state.ctypes.data_as(C.c_void_p), n)
    out["seed0_base"] = base[:, 0].tolist()
    out["seed0_state"] = state.tolist()
    s = state.copy()
    draws = []
    for _ in range(3):
        o = np.zeros(n, np.float32)
        ref.ref_random_number_generator(s.ctypes.data_as(C.c_void_p), n, o.ctypes.data_as(C.c_void_p))
        draws.append([float(x).hex() for x in o])
    out["seed0_random01_hex"] = draws
    out["seed0_state_after3"] = s.tolist()
    # (b) per-stream gap variant with awkward bases
    bases = np.array([[0, 0], [1, 0], [2147483647, 0], [4294967295, 0], [123456789, 0]], np.uint32)
    for gap in (1, 12345, 1 << 40, (1 << 63) + 5):
        st = bases.copy()
        ref.ref_generate_per_stream_random_state(st.ctypes.data_as(C.c_void_p), C.c_uint64(gap), len(st))
        out[f"gap_{gap}"] = st.tolist()
    out["gap_bases"] = bases[:, 0].tolist()
    # (c) raw steps from corner states
    steps = []
    for x, c in ((0, 1), (1, 0), (0xFFFFFFFF, 0xFFFEB81A), (0x12345678, 0x9ABCDEF0 % 4294883355)):
        xx, cc = C.c_uint32(x), C.c_uint32(c)
        seq = []
        for _ in range(4):
            ref.ref_step(C.byref(xx), C.byref(cc))
            seq.append([xx.value, cc.value])
        steps.append({"x": x, "c": c, "seq": seq})
    out["steps"] = steps
    (ROOT / "tests" / "golden").mkdir(parents=True, exist_ok=True)
    (ROOT / "tests" / "golden" / "mwc64x.json").write_text(json.dumps(out, indent=1))
    print("wrote tests/golden/mwc64x.json")


if __name__ == "__main__":
    mwc64x()
