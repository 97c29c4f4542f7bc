"""Regenerate tests/golden/*.npz.  Run in the authoring container only (needs cv2 4.13.0 and
/root/reference for oracle/_ref):   python tests/golden/make_golden.py

What is pinned, and by what:
  cv2_primitives.npz   outputs of the REAL OpenCV 4.13.0 (cv2 wheel) for resize / GaussianBlur / FAST /
                       fastAtan2 on small seeded inputs (inputs stored too, so no generator drift).
  libm_sincos.npz      glibc 2.39 sinf/cosf bit patterns on 4096 angles (the descriptor rotation).
  orb_*.npz            keypoints + descriptors of the reference's own ORBextractor.cpp compiled in
                       place (oracle/_ref/liborbref_parity.so) on synthetic frames; the frame is
                       regenerated from its seed and checked against the stored sha256.
"""
import ctypes
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

import oracle  # noqa: E402
from vo_slam_test_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)


def fast_cv2(img, th):
    k = cv2.FastFeatureDetector_create(th, True).detect(img)
    return np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in k], np.int32).reshape(-1, 3)


def undistort_golden():
    """cv2.undistortPoints(pts, K, D, None, K) -- the call of frame.cpp:58 -- on keypoint-like coordinates."""
    rng = np.random.default_rng(77)
    N = 6000
    pts = np.stack([rng.uniform(0, 640, N), rng.uniform(0, 480, N)], 1).astype(np.float32)
    pts[: N // 2] = np.floor(pts[: N // 2])                       # level-0 keypoints are integer valued
    pts[:4] = [[0, 0], [639, 479], [318.6, 255.3], [320, 240]]
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)     # TUM fr1 (config/example.yaml)
    sets = {"tum1": [0.2624, -0.9531, -0.0054, 0.0026, 1.1633], "four": [-0.28, 0.07, 0.0002, 0.00002],
            "rational": [0.1, -0.2, 0.001, 0.002, 0.05, 0.01, 0.02, 0.003],
            "strong": [-6.0, 2.0, 0.01, -0.02, 0.5]}              # drives icdist < 0 for points far from the centre
    d = {"pts": pts, "K": K}
    for name, D in sets.items():
        D = np.array(D, np.float32)
        d["D_" + name] = D
        d["out_" + name] = cv2.undistortPoints(pts.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "cv2_undistort.npz"), cv2_version=cv2.__version__, **d)


def main():
    undistort_golden()
    rng = np.random.default_rng(1234)
    d = {}
    # resize: noise image and a structured one, odd sizes
    src = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    d["resize_src"] = src
    d["resize_dst"] = cv2.resize(src, (109, 81), interpolation=cv2.INTER_LINEAR)
    frame = synth.make_frame(5, 120, 160)
    d["frame"] = frame
    d["frame_resize"] = cv2.resize(frame, (133, 100), interpolation=cv2.INTER_LINEAR)
    # blur
    d["blur_noise"] = cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    d["blur_frame"] = cv2.GaussianBlur(frame, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    # FAST on ROIs (non-contiguous views) and whole image, both thresholds
    d["fast20_frame"] = fast_cv2(frame, 20)
    d["fast7_frame"] = fast_cv2(frame, 7)
    roi = frame[30:73, 40:83]
    d["fast20_roi"] = fast_cv2(roi, 20)
    d["fast7_roi"] = fast_cv2(roi, 7)
    d["fast7_noise"] = fast_cv2(src, 7)
    # fastAtan2 on integer-valued moments
    yx = rng.integers(-200000, 200001, (2000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [5, 5], [0, 7], [7, 0], [-3, 0], [0, -3], [-5, -5], [1, -200000]]
    d["atan_yx"] = yx
    d["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(OUT, "cv2_primitives.npz"), cv2_version=cv2.__version__, **d)

    libm = ctypes.CDLL("libm.so.6")
    libm.sinf.restype = ctypes.c_float; libm.sinf.argtypes = [ctypes.c_float]
    libm.cosf.restype = ctypes.c_float; libm.cosf.argtypes = [ctypes.c_float]
    deg = rng.uniform(0, 360, 4096).astype(np.float32)
    deg[:4] = [0, 90, 180, 359.99997]
    rad = (deg * np.float32(np.float32(np.pi / np.float32(180.0)))).astype(np.float32)
    factor = np.float32(3.1415926535897932384626433832795 / float(np.float32(180.0)))
    rad = deg * factor
    s = np.array([libm.sinf(float(a)) for a in rad], np.float32)
    c = np.array([libm.cosf(float(a)) for a in rad], np.float32)
    np.savez_compressed(os.path.join(OUT, "libm_sincos.npz"), rad=rad, sin=s, cos=c)

    cases = [("orb_640x480_seed42", 42, 480, 640, 1000), ("orb_320x240_seed7", 7, 240, 320, 300),
             ("orb_752x480_seed9", 9, 480, 752, 1200), ("orb_1280x720_seed3", 3, 720, 1280, 1500)]
    for name, seed, H, W, nf in cases:
        img = synth.make_frame(seed, H, W)
        R = oracle.Ref(nf)
        kps, desc = R.extract(img)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, H=H, W=W, nfeatures=nf,
                            sha256=hashlib.sha256(img.tobytes()).hexdigest(), kps=kps, desc=desc)
        print(name, len(kps))


if __name__ == "__main__":
    main()
