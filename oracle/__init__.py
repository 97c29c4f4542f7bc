"""oracle -- TEST INFRASTRUCTURE, not product code.

CPU oracles for the ORB front-end hot path of guisongchen/vo_slam_test:

* ``Port``  -- this repo's own C++ restatement (oracle/orb_port.cpp, built to oracle/liborbport.so).
* ``Ref``   -- the reference's OWN ``src/ORBextractor.cpp`` compiled in place against oracle/compat
              (oracle/_ref/liborbref*.so; built only where /root/reference exists, prebuilt files
              travel to the GPU box).
* oracle/_ref/libmatcherref.so -- the reference's OWN ``src/matcher.cpp`` compiled in place against its own
              matcher.h and the stand-in object types of oracle/compat_myslam; linked by the C++ check program
              of tests/test_matcher_adapter.py (not loaded from Python).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product (``vo_slam_test_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)


def build(verbose=False):
    """Compile the oracle libraries (idempotent).  The reference build runs only if /root/reference exists."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ptr(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("ini_th", C.c_int),
                ("min_th", C.c_int)]


class _SbpFrameIn(C.Structure):
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p), ("n", C.c_int),
                ("xmin", C.c_float), ("xmax", C.c_float), ("ymin", C.c_float), ("ymax", C.c_float),
                ("scale_factors", C.c_void_p), ("nlevels", C.c_int), ("occupied0", C.c_void_p),
                ("m", C.c_int), ("valid", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("invz", C.c_void_p),
                ("octave", C.c_void_p), ("angle", C.c_void_p), ("mp_desc", C.c_void_p), ("has_obs", C.c_void_p),
                ("radius", C.c_float), ("bf", C.c_float), ("forward", C.c_int), ("backward", C.c_int),
                ("check_rot", C.c_int)]


class _SbpLocalIn(C.Structure):
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p), ("n", C.c_int),
                ("xmin", C.c_float), ("xmax", C.c_float), ("ymin", C.c_float), ("ymax", C.c_float),
                ("scale_factors", C.c_void_p), ("nlevels", C.c_int), ("occupied0", C.c_void_p),
                ("m", C.c_int), ("valid", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("ur", C.c_void_p),
                ("level", C.c_void_p), ("view_cos", C.c_void_p), ("mp_desc", C.c_void_p), ("has_obs", C.c_void_p),
                ("th_radius", C.c_float), ("ratio", C.c_float)]


class _BowSide(C.Structure):
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("angle", C.c_void_p), ("valid", C.c_void_p), ("ngroups", C.c_int),
                ("node_ids", C.c_void_p), ("group_start", C.c_void_p), ("feat_idx", C.c_void_p)]


def _bow_side(side, keep, struct=_BowSide):
    def a(x, dt):
        y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
    s = struct()
    s.n = len(side["desc"]); s.desc = a(side["desc"], np.uint8); s.angle = a(side["angle"], np.float32)
    s.valid = a(side["valid"], np.uint8); s.ngroups = len(side["node_ids"]); s.node_ids = a(side["node_ids"], np.uint32)
    s.group_start = a(side["group_start"], np.int32); s.feat_idx = a(side["feat_idx"], np.int32)
    return s


class _TriSide(C.Structure):
    _fields_ = [("side", _BowSide), ("kps", C.c_void_p), ("uright", C.c_void_p)]


class Port:
    """ctypes view of oracle/liborbport.so (this repo's CPU restatement)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        path = os.path.join(HERE, "liborbport.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.p = _Params(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        L = self.lib
        L.port_fast_atan2.restype = C.c_float
        L.port_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.port_ic_angle.restype = C.c_float
        L.port_ic_angle.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.port_descriptor.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.port_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.port_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.port_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.port_fast_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.port_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.port_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p]
        L.port_extract_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_int]
        L.port_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.port_grid_build.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                      C.c_void_p]
        L.port_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.port_sincosf.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.port_hamming.argtypes = [C.c_void_p, C.c_void_p]

    # ---- tables / geometry -------------------------------------------------------------------
    def tables(self):
        n = self.nlevels
        sc = np.zeros(n, np.float32); inv = np.zeros(n, np.float32); nf = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        self.lib.port_tables(C.byref(self.p), _ptr(sc), _ptr(inv), _ptr(nf), _ptr(um))
        return sc, inv, nf, um

    def level_size(self, W, H, level):
        w = C.c_int(); h = C.c_int()
        self.lib.port_level_size(C.byref(self.p), W, H, level, C.byref(w), C.byref(h))
        return w.value, h.value

    # ---- primitives --------------------------------------------------------------------------
    def resize(self, src, dw, dh):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.empty((dh, dw), np.uint8)
        self.lib.port_resize(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dw)
        return dst

    def blur(self, src):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.empty_like(src)
        self.lib.port_blur(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dst.strides[0])
        return dst

    def fast(self, img, th, nms=True):
        assert img.dtype == np.uint8 and img.strides[1] == 1
        cap = img.shape[0] * img.shape[1]
        out = np.empty((max(cap, 1), 3), np.int32)
        n = self.lib.port_fast(C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], th, int(nms),
                               _ptr(out), cap)
        return out[:n].copy()

    def fast_cells(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = img.size // 2 + 16
        out = np.empty((cap, 3), np.int32)
        n = self.lib.port_fast_cells(_ptr(img), img.shape[1], img.shape[0], img.strides[0], self.p.ini_th,
                                     self.p.min_th, _ptr(out), cap)
        return out[:n].copy()

    def octree(self, cand, region_w, region_h, N):
        cand = np.ascontiguousarray(cand, np.int32).reshape(-1, 3)
        cap = N + 64
        sel = np.empty(cap, np.int32)
        n = self.lib.port_octree(_ptr(cand), len(cand), region_w, region_h, N, _ptr(sel), cap)
        assert n <= cap
        return sel[:n].copy()

    def fast_atan2(self, y, x):
        return self.lib.port_fast_atan2(float(y), float(x))

    def sincosf(self, a):
        s = C.c_float(); c = C.c_float()
        self.lib.port_sincosf(C.c_float(a), C.byref(s), C.byref(c))
        return s.value, c.value

    def ic_angle(self, img, x, y):
        img = np.ascontiguousarray(img, np.uint8)
        return self.lib.port_ic_angle(_ptr(img), img.strides[0], int(x), int(y))

    def descriptor(self, blurred, x, y, angle):
        blurred = np.ascontiguousarray(blurred, np.uint8)
        out = np.empty(32, np.uint8)
        self.lib.port_descriptor(_ptr(blurred), blurred.strides[0], int(x), int(y), C.c_float(angle), _ptr(out))
        return out

    # ---- full extractor ----------------------------------------------------------------------
    def extract(self, img, want_levels=False):
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = self.p.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        lv = None
        if want_levels:
            tot = sum(np.prod(self.level_size(W, H, l)) for l in range(self.nlevels))
            lv = np.empty(int(tot), np.uint8)
        n = self.lib.port_extract(C.byref(self.p), _ptr(img), W, H, img.strides[0], _ptr(kps), _ptr(desc), cap,
                                  _ptr(lv) if lv is not None else None)
        if n < 0:
            raise RuntimeError("port_extract error %d" % n)
        assert n <= cap
        if want_levels:
            levels, off = [], 0
            for l in range(self.nlevels):
                w, h = self.level_size(W, H, l)
                levels.append(lv[off:off + w * h].reshape(h, w)); off += w * h
            return kps[:n].copy(), desc[:n].copy(), levels
        return kps[:n].copy(), desc[:n].copy()

    def extract_batch(self, imgs, nthreads):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        B, H, W = imgs.shape
        cap = self.p.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); cnt = np.zeros(B, np.int32)
        self.lib.port_extract_batch(C.byref(self.p), _ptr(imgs), B, W, H, _ptr(kps), _ptr(desc), cap, _ptr(cnt),
                                    nthreads)
        return kps, desc, cnt

    # ---- matcher -----------------------------------------------------------------------------
    def hamming(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return self.lib.port_hamming(_ptr(a), _ptr(b))

    def knn2(self, q, t, th=50, ratio=0.7, nthreads=1):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        Q = len(q)
        idx = np.empty(Q, np.int32); d1 = np.empty(Q, np.int32); d2 = np.empty(Q, np.int32); ok = np.empty(Q, np.uint8)
        self.lib.port_knn2(_ptr(q), Q, _ptr(t), len(t), th, C.c_float(ratio), _ptr(idx), _ptr(d1), _ptr(d2), _ptr(ok),
                           nthreads)
        return idx, d1, d2, ok

    def grid_build(self, kps, xmin, xmax, ymin, ymax):
        kps = np.ascontiguousarray(kps)
        start = np.empty(64 * 48 + 1, np.int32); ids = np.empty(max(len(kps), 1), np.int32)
        n = self.lib.port_grid_build(_ptr(kps), len(kps), xmin, xmax, ymin, ymax, _ptr(start), _ptr(ids))
        return start, ids[:n].copy()

    def frame_finish(self, kps, cam, depth=None):
        """frame.cpp:36-133 + :72-97 on one frame.  cam: dict(fx, fy, cx, cy, dist, bf, bounds=(xmin, xmax, ymin, ymax))."""
        kps = np.ascontiguousarray(kps)
        if kps.dtype != KP_DTYPE:
            kps = np.ascontiguousarray(kps, np.float32).view(KP_DTYPE).reshape(-1)
        n = len(kps)
        cs = camera_struct(cam)
        un = np.empty_like(kps); ur = np.empty(n, np.float32); dp = np.empty(n, np.float32)
        start = np.empty(64 * 48 + 1, np.int32); ids = np.empty(max(n, 1), np.int32)
        H, W, step = (0, 0, 0)
        if depth is not None:
            assert depth.dtype == np.float32 and depth.strides[1] == 4
            H, W = depth.shape; step = depth.strides[0]
        self.lib.port_frame_finish.restype = C.c_int
        self.lib.port_frame_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        m = self.lib.port_frame_finish(C.addressof(cs), _ptr(kps), n, depth.ctypes.data if depth is not None else None, W, H,
                                       step, _ptr(un), _ptr(ur), _ptr(dp), _ptr(start), _ptr(ids))
        return un, ur, dp, start, ids[:m].copy()

    def features_in_area(self, kps, bounds, u, v, r, min_level, max_level):
        kps = np.ascontiguousarray(kps)
        out = np.empty(max(len(kps), 1), np.int32)
        n = self.lib.port_features_in_area(_ptr(kps), len(kps), bounds[0], bounds[1], bounds[2], bounds[3], u, v, r,
                                           min_level, max_level, _ptr(out), len(out))
        return out[:n].copy()

    def sbp_frame(self, frame, pts, radius, bf=40.0, forward=False, backward=False, check_rot=True):
        """frame: dict(kps, desc, uright, bounds, scale_factors, occupied0); pts: dict(valid,u,v,invz,octave,angle,desc,has_obs)."""
        keep = []
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        s = _SbpFrameIn()
        s.kps = a(frame["kps"], KP_DTYPE); s.desc = a(frame["desc"], np.uint8); s.uright = a(frame["uright"], np.float32)
        s.n = len(frame["kps"])
        s.xmin, s.xmax, s.ymin, s.ymax = [float(b) for b in frame["bounds"]]
        s.scale_factors = a(frame["scale_factors"], np.float32); s.nlevels = len(frame["scale_factors"])
        s.occupied0 = a(frame["occupied0"], np.uint8)
        s.m = len(pts["u"]); s.valid = a(pts["valid"], np.uint8); s.u = a(pts["u"], np.float32)
        s.v = a(pts["v"], np.float32); s.invz = a(pts["invz"], np.float32); s.octave = a(pts["octave"], np.int32)
        s.angle = a(pts["angle"], np.float32); s.mp_desc = a(pts["desc"], np.uint8); s.has_obs = a(pts["has_obs"], np.uint8)
        s.radius = radius; s.bf = bf; s.forward = int(forward); s.backward = int(backward); s.check_rot = int(check_rot)
        assign = np.empty(max(s.n, 1), np.int32)
        self.lib.port_sbp_frame.restype = C.c_int
        cnt = self.lib.port_sbp_frame(C.byref(s), _ptr(assign))
        return assign[:s.n].copy(), cnt

    def search_for_triangulation(self, a, b, F12, epipole, scale2, th_low=50, check_rot=True):
        """a, b: BoW side dicts + 'kps' (KP_DTYPE) + 'uright'."""
        keep = []
        def tri(d):
            t = _TriSide(); t.side = _bow_side(d, keep)
            k = np.ascontiguousarray(d["kps"], KP_DTYPE); u = np.ascontiguousarray(d["uright"], np.float32); keep.extend([k, u])
            t.kps = k.ctypes.data; t.uright = u.ctypes.data
            return t
        ta, tb = tri(a), tri(b)
        F = np.ascontiguousarray(F12, np.float64).reshape(9); sc = np.ascontiguousarray(scale2, np.float32)
        match = np.empty(max(ta.side.n, 1), np.int32)
        self.lib.port_search_for_triangulation.restype = C.c_int
        self.lib.port_search_for_triangulation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                                           C.c_int, C.c_int, C.c_void_p]
        cnt = self.lib.port_search_for_triangulation(C.byref(ta), C.byref(tb), _ptr(F), epipole[0], epipole[1], _ptr(sc), th_low,
                                                     int(check_rot), _ptr(match))
        return match[:ta.side.n].copy(), cnt

    def medoid(self, desc, start):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
        best = np.empty(max(len(start) - 1, 1), np.int32)
        self.lib.port_medoid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.lib.port_medoid(_ptr(desc), _ptr(start), len(start) - 1, _ptr(best))
        return best[:len(start) - 1].copy()

    def search_by_bow(self, a, b, mode, ratio=0.7, th_low=50, check_rot=True):
        """a, b: dict(desc, angle, valid, node_ids, group_start, feat_idx); mode 0 KF->Frame, 1 KF->KF."""
        keep = []
        sa = _bow_side(a, keep); sb = _bow_side(b, keep)
        n_out = sb.n if mode == 0 else sa.n
        match = np.empty(max(n_out, 1), np.int32)
        self.lib.port_search_by_bow.restype = C.c_int
        self.lib.port_search_by_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
        cnt = self.lib.port_search_by_bow(C.byref(sa), C.byref(sb), mode, ratio, th_low, int(check_rot), _ptr(match))
        return match[:n_out].copy(), cnt

    def _sbp_frame_struct(self, frame, pts, radius, check_rot, keep):
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        s = _SbpFrameIn()
        s.kps = a(frame["kps"], KP_DTYPE); s.desc = a(frame["desc"], np.uint8); s.uright = a(frame["uright"], np.float32)
        s.n = len(frame["kps"])
        s.xmin, s.xmax, s.ymin, s.ymax = [float(b) for b in frame["bounds"]]
        s.scale_factors = a(frame["scale_factors"], np.float32); s.nlevels = len(frame["scale_factors"])
        s.occupied0 = a(frame["occupied0"], np.uint8)
        s.m = len(pts["u"]); s.valid = a(pts["valid"], np.uint8); s.u = a(pts["u"], np.float32)
        s.v = a(pts["v"], np.float32); s.invz = a(pts["invz"], np.float32); s.octave = a(pts["octave"], np.int32)
        s.angle = a(pts["angle"], np.float32); s.mp_desc = a(pts["desc"], np.uint8); s.has_obs = a(pts["has_obs"], np.uint8)
        s.radius = radius; s.bf = 0.0; s.forward = 0; s.backward = 0; s.check_rot = int(check_rot)
        return s

    def sbp_reloc(self, frame, pts, radius, dist_threshold, check_rot=True):
        keep = []
        s = self._sbp_frame_struct(frame, pts, radius, check_rot, keep)
        assign = np.empty(max(s.n, 1), np.int32)
        self.lib.port_sbp_reloc.restype = C.c_int
        self.lib.port_sbp_reloc.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        cnt = self.lib.port_sbp_reloc(C.byref(s), dist_threshold, _ptr(assign))
        return assign[:s.n].copy(), cnt

    def sbp_sim3(self, keyframe, pts, th):
        keep = []
        s = self._sbp_frame_struct(keyframe, pts, 0.0, False, keep)
        assign = np.empty(max(s.n, 1), np.int32)
        self.lib.port_sbp_sim3.restype = C.c_int
        self.lib.port_sbp_sim3.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        cnt = self.lib.port_sbp_sim3(C.byref(s), int(th), _ptr(assign))
        return assign[:s.n].copy(), cnt

    def window_argmin(self, keyframe, pts, th_radius, dist_threshold, chi2=False):
        keep = []
        s = self._sbp_frame_struct(keyframe, pts, 0.0, False, keep)
        best = np.empty(max(s.m, 1), np.int32)
        self.lib.port_window_argmin.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]
        self.lib.port_window_argmin(C.byref(s), th_radius, dist_threshold, int(chi2), _ptr(best))
        return best[:s.m].copy()

    def search_by_sim3(self, kf1, pts12, kf2, pts21, th):
        keep = []
        s12 = self._sbp_frame_struct(kf2, pts12, 0.0, False, keep)
        s21 = self._sbp_frame_struct(kf1, pts21, 0.0, False, keep)
        match = np.empty(max(s12.m, 1), np.int32)
        self.lib.port_search_by_sim3.restype = C.c_int
        self.lib.port_search_by_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        found = self.lib.port_search_by_sim3(C.byref(s12), C.byref(s21), th, _ptr(match))
        return match[:s12.m].copy(), found

    def sbp_local(self, frame, pts, th_radius, ratio):
        """pts: dict(valid,u,v,ur,level,view_cos,desc,has_obs)."""
        keep = []
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        s = _SbpLocalIn()
        s.kps = a(frame["kps"], KP_DTYPE); s.desc = a(frame["desc"], np.uint8); s.uright = a(frame["uright"], np.float32)
        s.n = len(frame["kps"])
        s.xmin, s.xmax, s.ymin, s.ymax = [float(b) for b in frame["bounds"]]
        s.scale_factors = a(frame["scale_factors"], np.float32); s.nlevels = len(frame["scale_factors"])
        s.occupied0 = a(frame["occupied0"], np.uint8)
        s.m = len(pts["u"]); s.valid = a(pts["valid"], np.uint8); s.u = a(pts["u"], np.float32)
        s.v = a(pts["v"], np.float32); s.ur = a(pts["ur"], np.float32); s.level = a(pts["level"], np.int32)
        s.view_cos = a(pts["view_cos"], np.float32); s.mp_desc = a(pts["desc"], np.uint8)
        s.has_obs = a(pts["has_obs"], np.uint8)
        s.th_radius = th_radius; s.ratio = ratio
        assign = np.empty(max(s.n, 1), np.int32)
        self.lib.port_sbp_local.restype = C.c_int
        cnt = self.lib.port_sbp_local(C.byref(s), _ptr(assign))
        return assign[:s.n].copy(), cnt


def best_march():
    """Highest x86-64 micro-architecture level of THIS host among the prebuilt "-march=native" stand-ins (oracle/Makefile):
    "v4" (AVX-512), "v3" (AVX2) or None."""
    try:
        flags = set(next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split(":")[1].split())
    except Exception:       # noqa: BLE001
        return None
    v3 = {"avx", "avx2", "bmi1", "bmi2", "fma", "movbe", "popcnt", "f16c", "abm"} <= flags | ({"abm"} if "lzcnt" in flags else set())
    v4 = v3 and {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    for lvl, okay in (("v4", v4), ("v3", v3)):
        if okay and os.path.exists(os.path.join(HERE, "_ref", "liborbref_%s.so" % lvl)):
            return lvl
    return None


class Ref:
    """ctypes view of the reference's own extractor, compiled in place (oracle/_ref).

    parity=True  -> liborbref_parity.so (monotonic allocator: deterministic quadtree tie-break)
    parity=False -> liborbref.so        (normal allocator, reference flags: the CPU timing baseline)
    """

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, parity=True, march=None):
        name = "liborbref_parity.so" if parity else ("liborbref_%s.so" % march if march else "liborbref.so")
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.orbref_create.restype = C.c_void_p
        L.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbref_destroy.argtypes = [C.c_void_p]
        L.orbref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_int]
        L.orbref_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbref_octree.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orbref_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbref_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.args = (nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = L.orbref_create(*self.args)

    def __del__(self):
        try:
            self.lib.orbref_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sc = np.zeros(n, np.float32); inv = np.zeros(n, np.float32); nf = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        self.lib.orbref_tables(self.h, _ptr(sc), _ptr(inv), _ptr(nf), _ptr(um))
        return sc, inv, nf, um

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.lib.orbref_extract(self.h, _ptr(img), W, H, img.strides[0], _ptr(kps), _ptr(desc), cap)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def pyramid_level(self, level):
        w = C.c_int(); h = C.c_int()
        self.lib.orbref_pyramid_level(self.h, level, None, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        self.lib.orbref_pyramid_level(self.h, level, _ptr(out), C.byref(w), C.byref(h))
        return out

    def octree(self, cand, region_w, region_h, N):
        cand = np.ascontiguousarray(cand, np.int32).reshape(-1, 3)
        cap = N + 64
        out = np.empty((cap, 3), np.int32)
        n = self.lib.orbref_octree(self.h, _ptr(cand), len(cand), region_w, region_h, N, _ptr(out), cap)
        assert n <= cap
        return out[:n].copy()

    def extract_batch(self, imgs, nthreads, keep_outputs=True):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        B, H, W = imgs.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        cnt = np.zeros(B, np.int32)
        if keep_outputs:
            kps = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8)
            self.lib.orbref_extract_batch(*self.args, _ptr(imgs), B, W, H, _ptr(kps), _ptr(desc), cap, _ptr(cnt), nthreads)
            return kps, desc, cnt
        self.lib.orbref_extract_batch(*self.args, _ptr(imgs), B, W, H, None, None, cap, _ptr(cnt), nthreads)
        return None, None, cnt


class CameraStruct(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 8),
                ("ndist", C.c_int), ("bf", C.c_float), ("xmin", C.c_float), ("xmax", C.c_float), ("ymin", C.c_float),
                ("ymax", C.c_float)]


def camera_struct(cam):
    cs = CameraStruct()
    cs.fx, cs.fy, cs.cx, cs.cy, cs.bf = cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam.get("bf", 0.0)
    d = list(cam.get("dist", []))
    assert len(d) <= 8
    for i in range(8):
        cs.dist[i] = d[i] if i < len(d) else 0.0
    cs.ndist = len(d)
    cs.xmin, cs.xmax, cs.ymin, cs.ymax = cam["bounds"]
    return cs


class RefMatcher:
    """The reference's OWN distance function (src/matcher.cpp:1240-1256, compiled in place: oracle/_ref/libmatcherref.so) inside
    the reference's matching loop shape (matcher.cpp:481-507), threaded over queries: the matcher leg of bench.py's CPU arm."""

    def __init__(self, march=None):
        path = os.path.join(HERE, "_ref", "libmatcherref_%s.so" % march if march else "libmatcherref.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)          # raises OSError where the reference was never compiled (callers fall back to the port)
        self.lib.refm_knn2.restype = None
        self.lib.refm_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]

    def knn2(self, q, t, th=50, ratio=0.7, nthreads=1):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        Q = len(q)
        idx = np.empty(Q, np.int32); d1 = np.empty(Q, np.int32); d2 = np.empty(Q, np.int32); ok = np.empty(Q, np.uint8)
        self.lib.refm_knn2(_ptr(q), Q, _ptr(t), len(t), th, ratio, _ptr(idx), _ptr(d1), _ptr(d2), _ptr(ok), nthreads)
        return idx, d1, d2, ok
