"""The C-ABI library loads and exports every symbol include/rtr_b200.h declares; without a GPU it
fails loudly instead of falling back to anything."""
import ctypes
import os
import re

from conftest import ROOT, has_gpu


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rtr_b200.h")).read() + open(os.path.join(ROOT, "include", "rtr_b200_io.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("rtr_create", "rtr_upload_cloud_xyz_bgr", "rtr_set_intrinsics", "rtr_set_pose_w2c", "rtr_render_rgbd",
              "rtr_render_filtered", "rtr_render_tensor", "rtr_destroy", "rtr_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"librtr_b200.so does not export {s}"
    assert set(declared_symbols()) == set(pkg.API) | set(pkg.API_IO), "python binding and headers disagree"


def test_no_cpu_fallback(pkg):
    if has_gpu():
        return
    lib = pkg.load_library()
    h = ctypes.c_void_p()
    rc = lib.rtr_create(0, ctypes.byref(h))
    assert rc == pkg.RTR_ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.rtr_last_error(None)
    try:
        pkg.ProjectCloud()
        raise AssertionError("ProjectCloud() must raise without a GPU")
    except pkg.RtrError as e:
        assert e.code == pkg.RTR_ERR_CUDA


def test_product_does_not_reference_the_oracle():
    pkgdir = os.path.join(ROOT, "real-time-neural-rendering-of-lidar-point-clouds_b200")
    for base, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "oracle/" not in src.replace("oracle/rtr_oracle.c:rtro_synth_packed", "") \
                    or f == "rtr_synth_common.h", f"{f} references the oracle"


def test_headers_compile_as_c99_and_cpp17_and_link(pkg, tmp_path):
    """include/rtr_b200.h is plain C; the header-only C++ adapter (reference method names) compiles
    and links against the in-tree library; without a GPU the constructor throws (no fallback)."""
    import subprocess
    libdir = os.path.dirname(pkg.LIB_PATH)
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "rtr_b200.h"\nint main(void) { return rtr_version() == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(c), "-o", str(tmp_path / "t_c"),
                    f"-L{libdir}", "-lrtr_b200", f"-Wl,-rpath,{libdir}"], check=True)
    assert subprocess.run([str(tmp_path / "t_c")]).returncode == 0
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "rtr_b200/project_cloud.hpp"\n#include <cstdio>\n'
                   'int main() { try { rtr_b200::ProjectCloud pc(0); rtr_b200::Intrinsics k; double E[16] = {1,0,0,0,0,1,0,0,0,0,1,0,0,0,0,1};\n'
                   '  auto net = [](void* t, int, int) -> const void* { return t; };\n'
                   '  if (pc.computeFull(k, E, nullptr, nullptr, net) > 0) return 4;   // no cloud uploaded: must fail, not crash\n'
                   '  return pc.computeRGBD(k, E, nullptr, nullptr) == -1 ? 0 : 3; }\n'
                   '  catch (const std::exception& e) { std::puts(e.what()); return 2; } }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", f"-I{inc}", str(cpp), "-o", str(tmp_path / "t_cpp"),
                    f"-L{libdir}", "-lrtr_b200", f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([str(tmp_path / "t_cpp")], capture_output=True, text=True)
    assert res.returncode == (0 if has_gpu() else 2)
    if not has_gpu():
        assert "no CPU fallback" in res.stdout
