"""The sharded path on two (or more) real GPUs: C-ABI communicator, peer / multimem exchange kernels, sharded ingest.
Skipped below 2 GPUs (the round-end box has one); run it with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_two_gpu_worker(torch_cuda):
    n = torch_cuda.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-6000:])
    assert r.returncode == 0 and "MGPU_OK" in r.stdout


def test_comm_symbols_and_id():
    """no GPU needed: the library exports the communicator ABI and NCCL resolves at run time"""
    import importlib
    sys.path.insert(0, str(ROOT))
    cpm = importlib.import_module("correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200")
    lib = cpm.lib()
    for name in ("cpm_comm_unique_id", "cpm_comm_init", "cpm_comm_init_nccl", "cpm_comm_destroy", "cpm_comm_split", "cpm_allreduce_lightvol",
                 "cpm_allreduce_lightvol_begin", "cpm_allreduce_lightvol_end", "cpm_allgather_photons", "cpm_allgather_volume",
                 "cpm_comm_upload_volume_sharded", "cpm_comm_barrier", "cpm_comm_transport"):
        assert hasattr(lib, name), name
    try:
        uid = cpm.capi.comm_unique_id()
    except cpm.capi.CpmError:
        pytest.skip("no libnccl.so.2 on this machine")
    assert len(uid) == 128 and uid != bytes(128)
