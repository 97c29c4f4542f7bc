"""GPU parity of the opt-in integer-tensor-pipe Hamming kernel (ORBX_HAMM_MMA=1: descriptor bits as +-1 bytes through
mma.sync IMMA.16832, exact integers) against the CPU oracle: ragged frame pairs, empty frames, self pairs (distance-0 ties,
duplicates), counts that are not multiples of the 128-query / 64-train tiles.  The switch is read once per process, so the
comparison runs in a child process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.path.insert(0, %r)
import numpy as np, torch
import oracle
import vo_slam_test_b200 as vo
from vo_slam_test_b200 import api, synth
rng = np.random.default_rng(11)
cap = 1100
counts = np.array([1006, 1100, 0, 1, 63, 64, 65, 129, 1000, 257, 1024, 7], np.int32)
F = len(counts)
desc = rng.integers(0, 256, (F, cap, 32), dtype=np.uint8)
for f in range(1, F):                       # consecutive frames share near-duplicates, exact duplicates and repeated rows
    n0, n1 = counts[f - 1], counts[f]
    if n0 == 0 or n1 == 0:
        continue
    src = rng.integers(0, n0, n1)
    desc[f, :n1] = synth.flip_bits(desc[f - 1, src], rng.integers(0, 60, n1), rng)
    desc[f, :n1:7] = desc[f - 1, src[::7]]
desc[8, 500:520] = desc[8, 100:120]         # duplicates inside one frame: first index must win
qf = np.array(list(range(F - 1)) + list(range(F)) + [0, 10, 8], np.int32)
tf = np.array(list(range(1, F)) + list(range(F)) + [10, 0, 1], np.int32)
dev = torch.device("cuda", 0)
d_desc = torch.from_numpy(desc).to(dev); d_cnt = torch.from_numpy(counts).to(dev)
d_qf = torch.from_numpy(qf).to(dev); d_tf = torch.from_numpy(tf).to(dev)
P_ = len(qf)
o = [torch.full((P_, cap), -9, dtype=torch.int32, device=dev) for _ in range(3)] + [torch.full((P_, cap), 9, dtype=torch.uint8, device=dev)]
api.knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), cap, d_qf.data_ptr(), d_tf.data_ptr(), P_, 50, 0.7,
                      o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
got = [x.cpu().numpy() for x in o]
P = oracle.Port()
acc = 0
for p in range(P_):
    nq, nt = counts[qf[p]], counts[tf[p]]
    w = P.knn2(desc[qf[p], :nq], desc[tf[p], :nt], 50, 0.7)
    for a, b, name in zip(got, w, ["idx", "d1", "d2", "ok"]):
        assert np.array_equal(a[p, :nq], b), (p, int(qf[p]), int(tf[p]), name)
        assert (a[p, nq:] == (9 if name == "ok" else -9)).all(), "rows beyond the query count must stay untouched"
    acc += int(w[3].sum())
assert acc > 2000
# one query set against a long train set (hamm_knn2: splits + merge), same sizes as tests/test_gpu_hamming.py
for nq, nt in [(1000, 1000), (1, 1), (130, 7), (257, 100000), (1000, 300001), (5, 0), (300, 64), (256, 65)]:
    r2 = np.random.default_rng(nq * 7 + nt)
    t = synth.make_descriptors(max(nt, 1), seed=nt)[:nt]
    q = synth.make_descriptors(nq, seed=nq + 1)
    if nt > 10:
        q = synth.flip_bits(t[r2.integers(0, nt, nq)], r2.integers(0, 60, nq), r2)
        t[r2.integers(0, nt, 20)] = t[r2.integers(0, nt, 20)]
    want = P.knn2(q, t, 50, 0.7, nthreads=8)
    got = vo.Matcher(0.7).knn2(q, t, th=50)
    for a, b, name in zip(got, want, ["idx", "d1", "d2", "ok"]):
        assert np.array_equal(a, b), (nq, nt, name)
t = np.repeat(synth.make_descriptors(1, seed=2), 5000, axis=0)      # all-equal rows: first index wins, d2 == d1
i_, a_, b_, _ = vo.Matcher(0.7).knn2(synth.make_descriptors(64, seed=1), t)
assert (i_ == 0).all() and np.array_equal(a_, b_)
print("mma pairs ok", acc)
'''


@pytest.mark.parametrize("mma", ["1", "0"])
def test_pairs_kernel_variants_equal_oracle(mma):
    env = dict(os.environ, ORBX_HAMM_MMA=mma)
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "mma pairs ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
