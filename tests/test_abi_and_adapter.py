"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/orb_b200.h
declares, refuses to compute without a GPU (no CPU fallback), and the C++ adapter header compiles and links against
the OpenCV-compat shim."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vo():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "vo_slam_test_b200", "lib", "libvoslam_b200.so")):
        g.build()
    import vo_slam_test_b200 as v
    return v


def test_library_exports_every_declared_symbol(vo):
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b((?:orbx|hamm)_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = vo.load_library()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/orb_b200.h must compile as C99 (no C++-isms, no torch / CUDA types)."""
    src = tmp_path / "c_hdr.c"
    src.write_text('#include "orb_b200.h"\nint main(void) { orbx_params p; orbx_frame_view v; (void)p; (void)v; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "orb_b200.h")).read(), flags=re.S)
    assert "torch" not in hdr and "cuda_runtime" not in hdr and "at::" not in hdr and "#include <cuda" not in hdr


def test_no_cpu_fallback(vo):
    import numpy as np
    if vo.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(vo.OrbError):
        vo.ORBextractor()
    with pytest.raises(vo.OrbError):
        vo.Matcher(0.7).knn2(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))


def test_round2_entry_points_without_a_gpu(vo):
    """The run-time switch of the Hamming variant is pure host state; the resident-frame entry points refuse null handles and,
    without a GPU, fail loudly instead of computing anything on the CPU."""
    import ctypes as C
    import numpy as np
    L = vo.lib()
    prev = L.hamm_set_variant(1)
    assert prev in (0, 1) and L.hamm_set_variant(-1) == 1          # any other argument only queries
    assert L.hamm_set_variant(0) == 1 and L.hamm_set_variant(prev) == 0
    h = C.c_void_p(); n = C.c_int(0)
    img = np.zeros((480, 640), np.uint8)
    cam = vo.camera(500.0, 500.0, 320.0, 240.0)
    assert L.orbx_frame_create(None, C.byref(cam), C.c_void_p(img.ctypes.data), 640, 480, 640, None, 0, C.byref(h), C.byref(n)) == -1
    assert L.orbx_frame_size(None, C.byref(n)) == -1 and L.orbx_frame_destroy(None) == 0
    assert L.orbx_search_by_projection_frame_h(None, None, None, 15.0, 40.0, 0, 0, 1, None, None) == -1
    assert L.orbx_frame_upload(None, 0, C.byref(h)) == -1
    if vo.device_count() == 0:
        kps = np.zeros(4, vo.KP_DTYPE); desc = np.zeros((4, 32), np.uint8); ur = np.full(4, -1.0, np.float32); sf = np.ones(8, np.float32)
        v = vo.api._FrameView()
        v.kps = kps.ctypes.data; v.desc = desc.ctypes.data; v.uright = ur.ctypes.data; v.n = 4
        v.xmin, v.xmax, v.ymin, v.ymax = 0.0, 640.0, 0.0, 480.0
        v.scale_factors = sf.ctypes.data; v.nlevels = 8; v.occupied0 = None
        assert L.orbx_frame_upload(C.byref(v), 0, C.byref(h)) == -2 and not h.value          # ORBX_ERR_CUDA, no handle
        p = C.c_void_p()
        assert L.orbx_host_alloc(4096, 0, C.byref(p)) == -2 and not p.value


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vo_slam_test_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("# oracle/", ""), f


def _build_adapter_check(tmp_path):
    exe = str(tmp_path / "adapter_check")
    cmd = ["g++", "-std=c++11", "-I" + os.path.join(ROOT, "oracle", "compat"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "tools", "adapter_check.cpp"), "-L" + os.path.join(ROOT, "vo_slam_test_b200", "lib"),
           "-lvoslam_b200", "-ldl", "-Wl,-rpath," + os.path.join(ROOT, "vo_slam_test_b200", "lib"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapter_header_compiles_and_links(vo, tmp_path):
    exe = _build_adapter_check(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter" in r.stdout


@pytest.mark.gpu
def test_adapter_equals_oracle_on_gpu(vo, tmp_path):
    """The C++ ORBextractor adapter (cv::Mat with step != cols in, std::vector<cv::KeyPoint> + descriptor Mat out) against
    the oracle port and, where oracle/_ref is built, the reference's own class: every keypoint and descriptor byte."""
    import oracle
    oracle.build()
    exe = _build_adapter_check(tmp_path)
    port = os.path.join(ROOT, "oracle", "liborbport.so")
    ref = os.path.join(ROOT, "oracle", "_ref", "liborbref_parity.so")
    args = [exe, port] + ([ref] if os.path.exists(ref) else [])
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter_check: all equal (4 cases" in r.stdout and r.stdout.count("adapter parity ok vs oracle port") == 4, r.stdout
    if os.path.exists(ref):
        assert r.stdout.count("the reference's own ORBextractor class") == 4, r.stdout
