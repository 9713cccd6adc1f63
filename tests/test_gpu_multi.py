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


def test_lost_peer_fails_the_render_call(gpu, cpu_oracle):
    """One GPU is enough for this: two renderers of ONE process attach to each other (same-process peers use the plain
    device addresses, no CUDA IPC) and only one of them renders.  Its merge kernel waits for a peer that never comes,
    gives up after peer_timeout_ms, and the render call — not a side channel — reports RTR_ERR_COMM.  A second attach
    needs a fresh export on every rank."""
    import numpy as np
    import scenes
    case = scenes.CASES["small_160x96"]
    rec = cpu_oracle.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    calib = gpu.CameraCalibration()
    calib.setIntrinsicsMatrix(case.K)
    calib.setWidth(case.W)
    calib.setHeight(case.H)
    P = case.W * case.H
    a, b = gpu.ProjectCloud.from_packed(rec[: case.n // 2]), gpu.ProjectCloud.from_packed(rec[case.n // 2:])
    for pc in (a, b):
        pc.set_camera(calib, case.poses[0])
    blobs = a.peer_export() + b.peer_export()
    a.peer_attach(blobs, 0, 2)
    b.peer_attach(blobs, 1, 2)
    with pytest.raises(gpu.RtrError) as e:          # the same blobs again: stale protocol state, refused
        a.peer_attach(blobs, 0, 2)
    assert e.value.code == gpu.RTR_ERR_STATE
    a.set_option("peer_timeout_ms", 300)
    color, depth = np.zeros(P * 3, np.uint8), np.zeros(P, np.float32)
    with pytest.raises(gpu.RtrError) as e:
        a.computeFilteredRGBD(calib, case.poses[0], color, depth)
    assert e.value.code == gpu.RTR_ERR_COMM
    assert a.get_option("peer_error") == 0          # reported once, through the call
    a.peer_detach()
    b.peer_detach()
    # detached: a plain single-GPU frame of the shard again
    assert a.computeFilteredRGBD(calib, case.poses[0], color, depth) == 1
    a.close()
    b.close()
