"""GPU parity of the alternative kernel selections.  The default build runs the TMA-staged kernels (resize, blur,
orientation/descriptor patches, FAST tile); the library falls back to global-load kernels when a level cannot be
described by a tensor map (unaligned caller strides) or when an ORBX_* switch asks for it.  The switches are read once per
process, so every selection runs in its own interpreter and is compared bit for bit with the CPU oracle there."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import numpy as np
import oracle
import vo_slam_test_b200 as v
from vo_slam_test_b200 import synth
for seed, H, W, nf in [(42, 480, 640, 1000), (5, 479, 641, 500), (9, 270, 480, 400)]:
    img = synth.make_frame(seed, H, W)
    ex = v.ORBextractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    rk, rd = oracle.Port(nf).extract(img)
    assert len(kps) == len(rk), (H, W, len(kps), len(rk))
    for name in ["x", "y", "size", "angle", "response", "octave", "class_id"]:
        assert np.array_equal(kps[name].view(np.uint32), rk[name].view(np.uint32)), (H, W, name)
    assert np.array_equal(desc, rd), (H, W)
    ex.close()
print("variant ok")
"""


@pytest.mark.parametrize("env", [
    {"ORBX_NO_TMA": "1"},                                                   # no tensor maps at all: every fallback kernel
    {"ORBX_RESIZE_TMA": "0", "ORBX_BLUR_TMA": "0", "ORBX_OD_TMA": "0"},     # maps exist, global-load kernels selected
    {"ORBX_RESIZE_TMA": "0", "ORBX_RESIZE_WALK": "0", "ORBX_BLUR_WALK": "0", "ORBX_BLUR_TMA": "0", "ORBX_FAST_WARP": "0"},
    {"ORBX_PDL": "2"},                                                      # every kernel of the chunk chain launched dependent
    {"ORBX_PDL": "0"},
], ids=["no_tma", "global_load_walks", "first_generation_kernels", "pdl_all", "pdl_off"])
def test_kernel_selection_keeps_parity(env):
    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_fused_pyramid_kernel_in_a_subprocess():
    """ORBX_PYR_FUSED=1 (read once per process): ONE launch builds the whole pyramid, tiles of level l wait on completion
    counters of level l-1 instead of kernel boundaries.  Opt-in (it measured slower than the seven dependent launches); the
    frames, a 12-frame chunk and the stage taps must still equal the oracle bit for bit."""
    import os, subprocess, sys
    code = r'''
import numpy as np, oracle, vo_slam_test_b200 as vo
from vo_slam_test_b200 import synth
P = oracle.Port()
ex = vo.ORBextractor()
imgs = synth.make_sequence(12, seed=31)
kk, dd, cc = ex.extract_batch(imgs)
assert ex.launch_count() == 5, ex.launch_count()          # pyramid + FAST + quadtree + blur + orient/desc
for f in (0, 5, 11):
    rk, rd = P.extract(imgs[f])
    assert cc[f] == len(rk) and np.array_equal(kk[f, :cc[f]], rk) and np.array_equal(dd[f, :cc[f]], rd)
k1, d1 = ex(imgs[3])
rk, rd, lv = P.extract(imgs[3], want_levels=True)
assert np.array_equal(k1, rk) and np.array_equal(d1, rd)
for l in range(1, 8):
    assert np.array_equal(ex.debug_level(0, l), lv[l]), l
big = synth.make_frame(3, 1080, 1920)
ex2 = vo.ORBextractor(2000)
k2, d2 = ex2(big)
rk, rd = oracle.Port(2000).extract(big)
assert np.array_equal(k2, rk) and np.array_equal(d2, rd)
print("fused pyramid ok")
'''
    env = dict(os.environ, ORBX_PYR_FUSED="1", ORBX_CHUNK="12")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and "fused pyramid ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
