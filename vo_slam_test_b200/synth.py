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


def make_sequence(nframes, seed=0, H=480, W=640, step=2, out=None, start=0):
    """A panning 'video': frame i is a window sliding `step` px per frame over one wide canvas, plus
    per-frame sensor noise.  Consecutive frames overlap, so frame-to-frame matching finds real
    correspondences.  Returns uint8 (nframes, H, W) (written into `out` if given) = frames [start, start + nframes)
    of the sequence (a rank of a sharded run generates only its own block).  The pan wraps after period / step = 1024
    frames and the noise pattern after 128, so frames i and i + 1024 are identical: a 4096-frame batch holds 1024 distinct
    frames (no kernel caches by content, so timing is unaffected; bench.py states it in `config`)."""
    rng = np.random.default_rng(seed)
    period = 2048  # canvas columns; the pan wraps around
    canvas = _canvas(rng, H, period + W)
    canvas[:, period:] = canvas[:, :W]  # make the wrap seamless
    y0 = int(H * 0.55)
    canvas[y0:y0 + max(H // 10, 40), :] = 97
    noise = rng.integers(-4, 5, (H, period + W + 64)).astype(np.float32)
    if out is None:
        out = np.empty((nframes, H, W), np.uint8)
    for k in range(nframes):
        i = start + k
        x = (i * step) % period
        nx = (i * 37) % 64
        f = canvas[:, x:x + W] + noise[:, nx:nx + W] * (1 if (i // 64) % 2 == 0 else -1)
        np.clip(f, 0, 255, out=f)
        out[k] = f.astype(np.uint8)
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


def make_projection_case(kps, desc, scale_factors, m, seed=0, W=640, H=480, stereo=False, local=False):
    """BASELINE config 4: `m` synthetic map points projected into a frame whose extractor output is (kps, desc).

    Each point is a perturbed copy of a random frame feature: position jittered by a few pixels (some fall outside
    the image), octave within +-1, descriptor with U[0,60] random bit flips, random observation flag.  Returns
    (frame_dict, points_dict) in the layout the oracle and the product API share."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    nl = len(scale_factors)
    src = rng.integers(0, n, m)
    u = (kps["x"][src] + rng.normal(0, 4.0, m)).astype(np.float32)
    v = (kps["y"][src] + rng.normal(0, 4.0, m)).astype(np.float32)
    far = rng.random(m) < 0.05                       # 5 %: anywhere, including outside the image
    u[far] = rng.uniform(-40, W + 40, far.sum()).astype(np.float32)
    v[far] = rng.uniform(-40, H + 40, far.sum()).astype(np.float32)
    octave = np.clip(kps["octave"][src] + rng.integers(-1, 2, m), 0, nl - 1).astype(np.int32)
    mp_desc = flip_bits(desc[src], rng.integers(0, 61, m), rng)
    frame = dict(kps=kps, desc=desc, bounds=(0.0, float(W), 0.0, float(H)), scale_factors=np.asarray(scale_factors, np.float32),
                 uright=(kps["x"] - rng.uniform(5, 40, n)).astype(np.float32) if stereo else np.full(n, -1.0, np.float32),
                 occupied0=(rng.random(n) < 0.03).astype(np.uint8))
    z = rng.uniform(0.5, 8.0, m).astype(np.float32)
    z[rng.random(m) < 0.02] *= -1                    # behind the camera
    pts = dict(valid=(rng.random(m) > 0.03).astype(np.uint8), u=u, v=v, desc=mp_desc,
               has_obs=(rng.random(m) < 0.6).astype(np.uint8))
    if local:
        pts.update(ur=(u - 40.0 / np.abs(z)).astype(np.float32), level=octave,
                   view_cos=rng.uniform(0.99, 1.0, m).astype(np.float32))
    else:
        pts.update(invz=(1.0 / z).astype(np.float32), octave=octave,
                   angle=((kps["angle"][src] + rng.normal(0, 6.0, m)) % 360.0).astype(np.float32))
    return frame, pts


def make_bow_side(desc, angle, valid=None, node_bits=9, seed=0, shuffle=True):
    """A synthetic DBoW3 FeatureVector for a descriptor set: vocabulary node = the first `node_bits` bits of the
    descriptor (so near-duplicate descriptors mostly share a node, like words of a real vocabulary), returned as the
    CSR the C ABI takes: node ids ascending (std::map order), feature indices in insertion order (ascending index,
    optionally shuffled to prove that the in-node order is honoured)."""
    rng = np.random.default_rng(seed)
    n = len(desc)
    bits = np.unpackbits(desc[:, :4], axis=1)[:, :node_bits].astype(np.uint32)
    node = (bits << np.arange(node_bits - 1, -1, -1, dtype=np.uint32)).sum(1).astype(np.uint32)
    ids = np.unique(node)
    start = [0]; feat = []
    for v in ids:
        f = np.flatnonzero(node == v)
        if shuffle and len(f) > 1 and rng.random() < 0.3:
            f = rng.permutation(f)
        feat.append(f); start.append(start[-1] + len(f))
    return dict(desc=desc, angle=np.asarray(angle, np.float32), valid=np.ones(n, np.uint8) if valid is None else valid,
                node_ids=ids, group_start=np.asarray(start, np.int32),
                feat_idx=np.concatenate(feat).astype(np.int32) if feat else np.zeros(0, np.int32))
