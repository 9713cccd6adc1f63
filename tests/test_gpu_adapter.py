"""The header-only C++ adapter (include/rtr_b200/project_cloud.hpp) on a GPU: a C++ program written like the reference's
example (tests/cpp/adapter_main.cpp: reference-style unordered_map<int, Block> grid -> fromGrid -> computeRGBD /
computeFilteredRGBD with cv::Mat outputs -> computeFull with a caller-supplied neural stage) is compiled against the
in-tree library and must produce the frames of the committed golden (= the reference's own outputs on a B200)."""
import os
import subprocess

import numpy as np
import pytest

import scenes
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cpp_adapter_renders_the_golden_frames(gpu, cpu_oracle, tmp_path):
    case = scenes.CASES["c1_640x480"]
    g = np.load(os.path.join(ROOT, "tests", "golden", case.name + ".npz"))
    rec = cpu_oracle.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    (tmp_path / "cloud.bin").write_bytes(rec.tobytes())
    (tmp_path / "pose.bin").write_bytes(np.ascontiguousarray(case.poses[0], np.float64).tobytes())
    libdir = os.path.dirname(gpu.LIB_PATH)
    exe = tmp_path / "adapter_main"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", f"-I{os.path.join(ROOT, 'include')}", f"-I{os.path.join(ROOT, 'oracle', 'stubs')}",
                    os.path.join(ROOT, "tests", "cpp", "adapter_main.cpp"), "-o", str(exe), f"-L{libdir}", "-lrtr_b200", f"-Wl,-rpath,{libdir}"],
                   check=True)
    K = case.K
    res = subprocess.run([str(exe), str(tmp_path / "cloud.bin"), str(case.W), str(case.H), repr(float(K[0, 0])), repr(float(K[1, 1])),
                          repr(float(K[0, 2])), repr(float(K[1, 2])), str(tmp_path / "pose.bin"), str(tmp_path / "out")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ADAPTER_OK" in res.stdout, res.stdout + res.stderr
    P = case.W * case.H
    for name, key, dt in (("raw_color", "raw_color_host", np.uint8), ("raw_depth", "raw_depth_host", np.uint32),
                          ("flt_color", "flt_color_host", np.uint8), ("flt_depth", "flt_depth_host", np.uint32)):
        a = np.fromfile(tmp_path / f"out_{name}.bin", dtype=dt)
        assert scenes.sha(a) == bytes(g[f"f0_{key}_sha"]).decode(), f"adapter {name} differs from the reference's frame"
    # computeFull with the identity "network": colour = saturate(rint(255 * tensor planes 0..2)) = the filtered image
    # wherever a pixel is kept (planes hold half(image / 255)), 0 where it is masked
    full = np.fromfile(tmp_path / "out_full_color.bin", dtype=np.uint8).reshape(P, 3)
    flt = np.fromfile(tmp_path / "out_flt_color.bin", dtype=np.uint8).reshape(P, 3)
    assert full.shape == flt.shape and np.abs(full.astype(int) - flt.astype(int)).max() <= 1
