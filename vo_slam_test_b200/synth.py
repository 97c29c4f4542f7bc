"""Deterministic synthetic inputs (numpy only) shared by tests and bench.py.

The reference ships no data (its dataset is external: config/example.yaml:4), so every workload in
BASELINE.json is synthetic.  Frames are corner-rich on purpose (stress for the quadtree stage) and
contain one flat band so that some FAST cells come back empty and take the minThFAST retry
(ORBextractor.cpp:820-824).
"""
import numpy as np


def _canvas(rng, H, W):
    """Low-frequency background + random grey rectangles, uint8 (H, W)."""
    gh, gw = H // 8 + 2, W // 8 + 2
    coarse = rng.integers(0, 256, (gh, gw)).astype(np.float32)
    # separable linear up-sampling of the coarse grid (numpy only)
    ys = np.linspace(0, gh - 1.001, H); xs = np.linspace(0, gw - 1.001, W)
    y0 = ys.astype(np.int64); x0 = xs.astype(np.int64)
    fy = (ys - y0)[:, None].astype(np.float32); fx = (xs - x0)[None, :].astype(np.float32)
    a = coarse[y0][:, x0]; b = coarse[y0][:, x0 + 1]; c = coarse[y0 + 1][:, x0]; d = coarse[y0 + 1][:, x0 + 1]
    img = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    img = img * 0.5 + 64
    nrect = (H * W) // 1500
    x = rng.integers(0, W, nrect); y = rng.integers(0, H, nrect)
    w = rng.integers(6, 41, nrect); h = rng.integers(6, 41, nrect)
    g = rng.integers(0, 256, nrect)
    for i in range(nrect):
        img[y[i]:y[i] + h[i], x[i]:x[i] + w[i]] = g[i]
    return img


def make_frame(seed, H=480, W=640, flat_band=True):
    """One synthetic grayscale frame, uint8 (H, W)."""
    rng = np.random.default_rng(seed)
    img = _canvas(rng, H, W)
    img = img + rng.integers(-4, 5, (H, W))
    if flat_band:
        y0 = int(H * 0.55)
        img[y0:y0 + max(H // 10, 40), :] = 97
    return np.clip(img, 0, 255).astype(np.uint8)


def make_sequence(nframes, seed=0, H=480, W=640, step=2, out=None):
    """A panning 'video': frame i is a window sliding `step` px per frame over one wide canvas, plus
    per-frame sensor noise.  Consecutive frames overlap, so frame-to-frame matching finds real
    correspondences.  Returns uint8 (nframes, H, W) (written into `out` if given)."""
    rng = np.random.default_rng(seed)
    period = 2048  # canvas columns; the pan wraps around
    canvas = _canvas(rng, H, period + W)
    canvas[:, period:] = canvas[:, :W]  # make the wrap seamless
    y0 = int(H * 0.55)
    canvas[y0:y0 + max(H // 10, 40), :] = 97
    noise = rng.integers(-4, 5, (H, period + W + 64)).astype(np.float32)
    if out is None:
        out = np.empty((nframes, H, W), np.uint8)
    for i in range(nframes):
        x = (i * step) % period
        nx = (i * 37) % 64
        f = canvas[:, x:x + W] + noise[:, nx:nx + W] * (1 if (i // 64) % 2 == 0 else -1)
        np.clip(f, 0, 255, out=f)
        out[i] = f.astype(np.uint8)
    return out


def make_descriptors(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def flip_bits(desc, nflips, rng):
    """Copy of 32-byte descriptors with `nflips[i]` random bit positions toggled in row i."""
    out = desc.copy()
    for i in range(len(out)):
        k = int(nflips[i])
        if k:
            pos = rng.choice(256, size=k, replace=False)
            np.bitwise_xor.at(out[i], pos // 8, (1 << (pos % 8)).astype(np.uint8))
    return out
