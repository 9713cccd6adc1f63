"""Multi-GPU parity on real GPUs (needs >= 2 visible devices; the 1-GPU round-end run skips it and the
world_size-2 gloo test in tests/test_multi_rank_gloo.py covers the host logic)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    import ctypes
    try:
        lib = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    n = ctypes.c_int(0)
    return n.value if lib.cuInit(0) == 0 and lib.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0


def test_point_and_frame_sharding_equal_single_gpu(gpu):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "MULTI_GPU_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
