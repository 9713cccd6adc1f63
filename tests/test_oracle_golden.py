"""Pin the CPU oracle to the REFERENCE: tests/golden/*.npz hold outputs of the reference's own CUDA
code (compiled unmodified, run on a B200 by tests/golden/make_golden.py).  Re-create each case's
cloud from its seed, apply the stored projection tap (the points where MUFU.RCP and a correctly
rounded divide disagree), run the oracle, compare with what the reference produced."""
import os

import numpy as np
import pytest

import scenes
from conftest import ROOT

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
AVAILABLE = [n for n in scenes.CASES if os.path.exists(os.path.join(GOLDEN_DIR, n + ".npz"))]


def test_golden_files_are_committed():
    assert AVAILABLE, "no golden vectors in tests/golden (run tests/golden/make_golden.py on a B200)"
    for n in scenes.CASES:
        assert n in AVAILABLE, f"golden for {n} missing"


@pytest.mark.parametrize("name", AVAILABLE)
def test_oracle_reproduces_reference_outputs(cpu_oracle, name):
    case = scenes.CASES[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    rec = cpu_oracle.synth_packed(case.seed, case.n, 0, case.n, case.hall, case.n_boxes)
    assert int(g["n_frames"][0]) == len(case.poses)
    taps = []
    for fi, E in enumerate(case.poses):
        cam = g[f"f{fi}_cam_proj"]
        # a1: the host-side K4*E product equals what the reference's glm expression uploaded
        assert np.array_equal(cpu_oracle.cam_proj(case.K, E), cam)
        pix, zb = cpu_oracle.project(rec, cam, case.W, case.H)
        idx = g[f"f{fi}_tap_idx"]
        # documented pixel-boundary ties: the CPU divide and MUFU.RCP round a few points differently
        assert len(idx) <= max(20, 2e-3 * case.n), f"{len(idx)} projection ties is more than the MUFU.RCP error explains"
        if len(idx):
            both = (pix[idx] >= 0) & (g[f"f{fi}_tap_pix"] >= 0)
            du = np.abs(pix[idx][both] % case.W - g[f"f{fi}_tap_pix"][both] % case.W)
            dv = np.abs(pix[idx][both] // case.W - g[f"f{fi}_tap_pix"][both] // case.W)
            assert (np.maximum(du, dv) <= 1).all(), "a tie moved a point by more than one pixel"
            assert np.array_equal(zb[idx][both], g[f"f{fi}_tap_zb"][both]), "depth bits never depend on the divide"
        pix[idx], zb[idx] = g[f"f{fi}_tap_pix"], g[f"f{fi}_tap_zb"]
        taps.append((pix, zb))
    frames = scenes.oracle_frames(cpu_oracle, case, rec, taps)
    for fi, f in enumerate(frames):
        for k in scenes.OUTPUT_KEYS:
            assert scenes.sha(f[k]) == bytes(g[f"f{fi}_{k}_sha"]).decode(), f"{name} frame {fi} {k}: oracle != reference"
            if case.full:
                assert np.array_equal(f[k], g[f"f{fi}_{k}"])
