"""ctypes binding of include/orb_b200.h + the Python mirror of the reference operator API."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

TH_HIGH = 100   # matcher.cpp:11
TH_LOW = 50     # matcher.cpp:12


class OrbError(RuntimeError):
    pass


def lib_path():
    """In-tree build of the C-ABI library; ORBX_LIB overrides it (A/B experiments with alternative builds)."""
    return os.environ.get("ORBX_LIB") or os.path.join(HERE, "lib", "libvoslam_b200.so")


_lib = None


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("ini_th_fast", C.c_int),
                ("min_th_fast", C.c_int), ("device", C.c_int)]


class _FrameView(C.Structure):
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p), ("n", C.c_int),
                ("xmin", C.c_float), ("xmax", C.c_float), ("ymin", C.c_float), ("ymax", C.c_float),
                ("scale_factors", C.c_void_p), ("nlevels", C.c_int), ("occupied0", C.c_void_p)]


class _SbpFramePoints(C.Structure):
    _fields_ = [("m", C.c_int), ("valid", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("invz", C.c_void_p),
                ("octave", C.c_void_p), ("angle", C.c_void_p), ("desc", C.c_void_p), ("has_obs", C.c_void_p)]


class _SbpLocalPoints(C.Structure):
    _fields_ = [("m", C.c_int), ("valid", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("ur", C.c_void_p),
                ("level", C.c_void_p), ("view_cos", C.c_void_p), ("desc", C.c_void_p), ("has_obs", C.c_void_p)]


class _BowSide(C.Structure):
    _fields_ = [("n", C.c_int), ("desc", C.c_void_p), ("angle", C.c_void_p), ("valid", C.c_void_p), ("ngroups", C.c_int),
                ("node_ids", C.c_void_p), ("group_start", C.c_void_p), ("feat_idx", C.c_void_p)]


class _TriSide(C.Structure):
    _fields_ = [("side", _BowSide), ("kps", C.c_void_p), ("uright", C.c_void_p)]


class _Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 8),
                ("ndist", C.c_int), ("bf", C.c_float), ("xmin", C.c_float), ("xmax", C.c_float), ("ymin", C.c_float),
                ("ymax", C.c_float)]


def load_library():
    """Load libvoslam_b200.so.  Raises OrbError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise OrbError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the CUDA library is the only implementation; there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, i32, f32, sz, ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_device_count.argtypes = [vp]
    L.orbx_host_alloc.argtypes = [sz, i32, vp]
    L.orbx_host_free.argtypes = [vp]
    L.orbx_create.argtypes = [vp, vp]
    L.orbx_frame_create.argtypes = [vp, vp, vp, i32, i32, sz, vp, sz, vp, vp]
    L.orbx_frame_size.argtypes = [vp, vp]
    L.orbx_frame_get.argtypes = [vp, vp, vp, vp, vp, vp, i32]
    L.orbx_frame_destroy.argtypes = [vp]
    L.orbx_frame_grid.argtypes = [vp, vp, vp, i32]
    L.orbx_frame_upload.argtypes = [vp, i32, vp]
    L.orbx_search_by_projection_frame_h.argtypes = [vp, vp, vp, f32, f32, i32, i32, i32, vp, vp]
    L.orbx_search_by_projection_local_h.argtypes = [vp, vp, vp, f32, f32, vp, vp]
    L.orbx_search_by_projection_reloc_h.argtypes = [vp, vp, vp, f32, f32, i32, vp, vp]
    L.orbx_search_by_bow_h.argtypes = [vp, vp, vp, f32, i32, i32, vp, vp]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_get_levels.argtypes = [vp, vp]
    L.orbx_scale_factors.argtypes = [vp, vp, i32]
    L.orbx_inv_scale_factors.argtypes = [vp, vp, i32]
    L.orbx_features_per_level.argtypes = [vp, vp, i32]
    L.orbx_max_keypoints.argtypes = [vp, vp]
    L.orbx_extract.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32, vp]
    L.orbx_extract_batch.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp]
    L.orbx_extract_batch_device.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp, vp]
    L.orbx_launch_count.argtypes = [vp, vp]
    L.orbx_extract_match_batch.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp, i32, f32, vp, vp, vp, vp]
    L.orbx_extract_match_batch_device.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp, i32, f32, vp, vp, vp, vp, vp]
    L.orbx_profile_stages.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, vp, i32, vp, vp, vp]
    L.orbx_debug_level.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    L.orbx_debug_candidates.argtypes = [vp, i32, i32, vp, i32, vp]
    L.orbx_debug_selected.argtypes = [vp, i32, i32, vp, i32, vp]
    L.hamm_knn2.argtypes = [vp, i32, vp, ll, i32, f32, vp, vp, vp, vp, i32]
    L.hamm_knn2_device.argtypes = [vp, i32, vp, ll, i32, f32, vp, vp, vp, vp, vp, sz, vp]
    L.hamm_knn2_workspace_bytes.restype = sz
    L.hamm_knn2_workspace_bytes.argtypes = [i32, ll]
    L.hamm_knn2_pairs_device.argtypes = [vp, vp, i32, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp]
    L.hamm_knn2_merge_device.argtypes = [vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp]
    L.hamm_launch_count.restype = ll
    L.hamm_set_variant.argtypes = [i32]
    L.hamm_set_dynamic.argtypes = [i32]
    L.hamm_exchange_bytes.restype = sz
    L.hamm_exchange_bytes.argtypes = [i32, i32]
    L.hamm_exchange_alloc.argtypes = [i32, i32, i32, vp, vp]
    L.hamm_exchange_open.argtypes = [i32, vp, vp]
    L.hamm_exchange_close.argtypes = [vp]
    L.hamm_exchange_free.argtypes = [vp]
    L.hamm_knn2_sharded_device.argtypes = [vp, i32, vp, ll, ll, i32, f32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    L.hamm_knn2_sharded_phases_device.argtypes = [vp, i32, vp, ll, ll, i32, f32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp, i32]
    L.orbx_grid_build.argtypes = [vp, i32, f32, f32, f32, f32, vp, vp, i32]
    L.orbx_search_by_projection_frame.argtypes = [vp, vp, f32, f32, i32, i32, i32, vp, vp, i32]
    L.orbx_search_by_projection_local.argtypes = [vp, vp, f32, f32, vp, vp, i32]
    L.orbx_search_by_bow.argtypes = [vp, vp, i32, f32, i32, i32, vp, vp, i32]
    L.orbx_search_for_triangulation.argtypes = [vp, vp, vp, f32, f32, vp, i32, i32, i32, vp, vp, i32]
    L.orbx_window_argmin.argtypes = [vp, vp, f32, f32, i32, vp, i32]
    L.orbx_search_by_sim3.argtypes = [vp, vp, vp, vp, f32, vp, vp, i32]
    L.orbx_medoid_descriptors.argtypes = [vp, vp, i32, vp, i32]
    L.orbx_search_by_projection_reloc.argtypes = [vp, vp, f32, f32, i32, vp, vp, i32]
    L.orbx_search_by_projection_sim3.argtypes = [vp, vp, i32, vp, vp, i32]
    L.orbx_frame_finish.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, sz, sz, vp, vp, vp, vp, vp, i32]
    L.orbx_frame_finish_device.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, sz, sz, vp, vp, vp, vp, vp, i32, vp]
    _lib = L
    return L


def lib():
    return load_library()


def _check(rc):
    if rc != 0:
        raise OrbError("libvoslam_b200 error %d: %s" % (rc, load_library().orbx_last_error().decode()))


def device_count():
    n = C.c_int(0)
    load_library().orbx_device_count(C.byref(n))
    return n.value


class HostBuffer:
    """Pinned host staging memory from orbx_host_alloc (optionally write-combined) as a numpy uint8 array `.array`."""

    def __init__(self, shape, write_combined=False):
        n = int(np.prod(shape))
        self.ptr = C.c_void_p()
        _check(load_library().orbx_host_alloc(n, 1 if write_combined else 0, C.byref(self.ptr)))
        self.array = np.ctypeslib.as_array((C.c_uint8 * n).from_address(self.ptr.value)).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            load_library().orbx_host_free(self.ptr)
            self.ptr = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (ORBextractor.h:45-108): same constructor arguments, call operator and getters."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        p = _Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device)
        _check(self._lib.orbx_create(C.byref(p), C.byref(self._h)))
        self.nlevels = nlevels
        self.device = device
        cap = C.c_int()
        _check(self._lib.orbx_max_keypoints(self._h, C.byref(cap)))
        self.max_keypoints = cap.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.orbx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getters (ORBextractor.h:63-75) ---------------------------------------------------------
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        out = np.zeros(self.nlevels, np.float32)
        _check(self._lib.orbx_scale_factors(self._h, _p(out), self.nlevels))
        return out

    def GetScaleFactor(self):
        return float(self.GetScaleFactors()[1]) if self.nlevels > 1 else 1.0

    def GetInverseScaleFactors(self):
        out = np.zeros(self.nlevels, np.float32)
        _check(self._lib.orbx_inv_scale_factors(self._h, _p(out), self.nlevels))
        return out

    def features_per_level(self):
        out = np.zeros(self.nlevels, np.int32)
        _check(self._lib.orbx_features_per_level(self._h, _p(out), self.nlevels))
        return out

    # -- operator() (ORBextractor.cpp:1051) ------------------------------------------------------
    def __call__(self, image, mask=None):
        """(keypoints[KP_DTYPE], descriptors[n,32] uint8); `mask` is ignored like in the reference (ORBextractor.h:58)."""
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise OrbError("image must be 8-bit single channel (assert at ORBextractor.cpp:1058)")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        H, W = image.shape
        cap = self.max_keypoints
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(self._lib.orbx_extract(self._h, C.c_void_p(image.ctypes.data), W, H, image.strides[0], _p(kps), _p(desc),
                                      cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """images: uint8 (B, H, W) host array -> (kps (B,cap), desc (B,cap,32), counts (B,))."""
        images = np.ascontiguousarray(images, np.uint8)
        B, H, W = images.shape
        cap = self.max_keypoints
        kps = np.zeros((B, cap), KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        counts = np.zeros(B, np.int32)
        _check(self._lib.orbx_extract_batch(self._h, _p(images), B, W, H, W, W * H, _p(kps), _p(desc), cap, _p(counts)))
        return kps, desc, counts

    def extract_batch_device(self, d_imgs, nframes, W, H, row_stride, frame_stride, d_kps, d_desc, cap, d_counts, stream=0):
        """Raw device pointers (ints); asynchronous on `stream`."""
        _check(self._lib.orbx_extract_batch_device(self._h, C.c_void_p(d_imgs), nframes, W, H, row_stride, frame_stride,
                                                   C.c_void_p(d_kps), C.c_void_p(d_desc), cap, C.c_void_p(d_counts),
                                                   C.c_void_p(stream)))

    def extract_match_batch_device(self, d_imgs, nframes, W, H, row_stride, frame_stride, d_kps, d_desc, cap, d_counts, th, ratio,
                                   d_midx, d_md1, d_md2, d_mok, stream=0):
        """Raw device pointers (ints); asynchronous on `stream`: extraction + frame-to-frame top-2, matching pipelined per chunk."""
        _check(self._lib.orbx_extract_match_batch_device(self._h, C.c_void_p(d_imgs), nframes, W, H, row_stride, frame_stride,
                                                         C.c_void_p(d_kps), C.c_void_p(d_desc), cap, C.c_void_p(d_counts), th, ratio,
                                                         C.c_void_p(d_midx), C.c_void_p(d_md1), C.c_void_p(d_md2), C.c_void_p(d_mok),
                                                         C.c_void_p(stream)))

    def extract_match_batch(self, imgs_ptr, nframes, W, H, kps_ptr, desc_ptr, cap, counts_ptr, th, ratio, midx_ptr,
                            md1_ptr, md2_ptr, mok_ptr, row_stride=None, frame_stride=None):
        """Host pointers (ints; pinned memory recommended): pipelined H2D -> extract -> frame-to-frame top-2 -> D2H."""
        row_stride = W if row_stride is None else row_stride
        frame_stride = row_stride * H if frame_stride is None else frame_stride
        _check(self._lib.orbx_extract_match_batch(self._h, C.c_void_p(imgs_ptr), nframes, W, H, row_stride, frame_stride,
                                                  C.c_void_p(kps_ptr), C.c_void_p(desc_ptr), cap, C.c_void_p(counts_ptr),
                                                  th, ratio, C.c_void_p(midx_ptr), C.c_void_p(md1_ptr),
                                                  C.c_void_p(md2_ptr), C.c_void_p(mok_ptr)))

    def profile_stages(self, d_imgs, nframes, W, H, row_stride, frame_stride, d_kps, d_desc, cap, d_counts, stream=0):
        """Per-stage device ms (pyramid, fast, quadtree, blur, orient+desc, total) of one resident batch."""
        ms = np.zeros(6, np.float32)
        _check(self._lib.orbx_profile_stages(self._h, C.c_void_p(d_imgs), nframes, W, H, row_stride, frame_stride,
                                             C.c_void_p(d_kps), C.c_void_p(d_desc), cap, C.c_void_p(d_counts),
                                             C.c_void_p(stream), _p(ms)))
        return ms

    def launch_count(self):
        n = C.c_longlong(0)
        _check(self._lib.orbx_launch_count(self._h, C.byref(n)))
        return n.value

    # -- stage taps ------------------------------------------------------------------------------
    def debug_level(self, frame, level, blurred=False):
        w = C.c_int(); h = C.c_int()
        _check(self._lib.orbx_debug_level(self._h, frame, level, int(blurred), None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        _check(self._lib.orbx_debug_level(self._h, frame, level, int(blurred), _p(out), C.byref(w), C.byref(h)))
        return out

    def debug_candidates(self, frame, level, cap=1 << 20):
        out = np.empty((cap, 3), np.int32); n = C.c_int()
        _check(self._lib.orbx_debug_candidates(self._h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_selected(self, frame, level, cap=1 << 16):
        out = np.empty((cap, 3), np.int32); n = C.c_int()
        _check(self._lib.orbx_debug_selected(self._h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()


def knn2_pairs_device(d_desc, d_counts, cap, d_qf, d_tf, npairs, th, ratio, d_idx, d_d1, d_d2, d_ok, stream=0):
    _check(load_library().hamm_knn2_pairs_device(C.c_void_p(d_desc), C.c_void_p(d_counts), cap, C.c_void_p(d_qf),
                                                 C.c_void_p(d_tf), npairs, th, ratio, C.c_void_p(d_idx), C.c_void_p(d_d1),
                                                 C.c_void_p(d_d2), C.c_void_p(d_ok), C.c_void_p(stream)))


def knn2_merge_device(d_idx_in, d_d1_in, d_d2_in, nshards, nq, th, ratio, d_idx, d_d1, d_d2, d_ok, stream=0):
    _check(load_library().hamm_knn2_merge_device(C.c_void_p(d_idx_in), C.c_void_p(d_d1_in), C.c_void_p(d_d2_in), nshards, nq,
                                                 th, ratio, C.c_void_p(d_idx), C.c_void_p(d_d1), C.c_void_p(d_d2),
                                                 C.c_void_p(d_ok), C.c_void_p(stream)))


def exchange_alloc(device, world, max_queries):
    """This rank's exchange buffer for the fused sharded top-2 -> (device pointer, 64-byte CUDA IPC handle)."""
    buf = C.c_void_p()
    hdl = (C.c_ubyte * 64)()
    _check(load_library().hamm_exchange_alloc(device, world, max_queries, C.byref(buf), hdl))
    return buf.value, bytes(hdl)


def exchange_open(device, handle):
    """Map a peer's exchange buffer from its IPC handle -> device pointer valid in this process."""
    buf = C.c_void_p()
    hdl = (C.c_ubyte * 64).from_buffer_copy(handle)
    _check(load_library().hamm_exchange_open(device, hdl, C.byref(buf)))
    return buf.value


def exchange_close(ptr):
    _check(load_library().hamm_exchange_close(C.c_void_p(ptr)))


def exchange_free(ptr):
    _check(load_library().hamm_exchange_free(C.c_void_p(ptr)))


def knn2_sharded_device(d_q, nq, d_t, nt, shard_lo, th, ratio, rank, world, bufs, max_queries, epoch, d_idx, d_d1, d_d2, d_ok,
                        d_status, d_ws=0, ws_bytes=0, stream=0, phases=3):
    """Local shard top-2 + peer-memory scatter + merge (hamm_knn2_sharded[_phases]_device); bufs = list of `world` device
    pointers; phases: 1 = scan + scatter only, 2 = flag wait + merge only, 3 = both."""
    arr = (C.c_void_p * world)(*bufs)
    _check(load_library().hamm_knn2_sharded_phases_device(C.c_void_p(d_q), nq, C.c_void_p(d_t), nt, shard_lo, th, ratio, rank, world,
                                                          arr, max_queries, epoch, C.c_void_p(d_idx), C.c_void_p(d_d1),
                                                          C.c_void_p(d_d2), C.c_void_p(d_ok), C.c_void_p(d_status), C.c_void_p(d_ws),
                                                          ws_bytes, C.c_void_p(stream), phases))


def set_hamming_dynamic(mode):
    """Map-scale scans: 0 = fixed row range per CTA, 1 = automatic (default), 2 = dynamic block hand-out whenever the scan is split
    (same results); returns the previous mode."""
    return int(load_library().hamm_set_dynamic(int(mode)))


def set_hamming_variant(variant):
    """0 = POPC kernels (default), 1 = integer-tensor-pipe variant (same results); returns the previous value."""
    return int(load_library().hamm_set_variant(int(variant)))


def knn2_workspace_bytes(nq, nt):
    return int(load_library().hamm_knn2_workspace_bytes(nq, nt))


def knn2_device(d_q, nq, d_t, nt, th, ratio, d_idx, d_d1, d_d2, d_ok, d_ws=0, ws_bytes=0, stream=0):
    _check(load_library().hamm_knn2_device(C.c_void_p(d_q), nq, C.c_void_p(d_t), nt, th, ratio, C.c_void_p(d_idx),
                                           C.c_void_p(d_d1), C.c_void_p(d_d2), C.c_void_p(d_ok), C.c_void_p(d_ws),
                                           ws_bytes, C.c_void_p(stream)))


def medoid_descriptors(desc, start, device=0):
    """MapPoint::computeDescriptor (mappoint.cpp:118-179) for many map points: desc (total,32) uint8, start CSR (P+1)."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    start = np.ascontiguousarray(start, np.int32)
    P = len(start) - 1
    best = np.zeros(max(P, 1), np.int32)
    _check(load_library().orbx_medoid_descriptors(_p(desc), _p(start), P, _p(best), device))
    return best[:P].copy()


def grid_build(kps, bounds, device=0):
    """Frame::assignFeaturesToGrid (frame.cpp:72-89) -> CSR (cell_start[64*48+1], ids); cell = ix*48+iy."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    start = np.zeros(64 * 48 + 1, np.int32)
    ids = np.zeros(max(len(kps), 1), np.int32)
    _check(load_library().orbx_grid_build(_p(kps), len(kps), bounds[0], bounds[1], bounds[2], bounds[3], _p(start), _p(ids),
                                          device))
    return start, ids[:start[-1]].copy()


def camera(fx, fy, cx, cy, dist=(), bf=0.0, bounds=(0.0, 640.0, 0.0, 480.0)):
    """Camera constants of camera.cpp:10-48 as the C struct orbx_camera."""
    c = _Camera()
    c.fx, c.fy, c.cx, c.cy, c.bf = fx, fy, cx, cy, bf
    dist = list(dist)
    if len(dist) > 8:
        raise OrbError("at most 8 distortion coefficients (k1 k2 p1 p2 k3 k4 k5 k6)")
    for i, v in enumerate(dist):
        c.dist[i] = v
    c.ndist = len(dist)
    c.xmin, c.xmax, c.ymin, c.ymax = bounds
    return c


def frame_finish(kps, counts, cam, depth=None, device=0):
    """Frame::undistortKeyPoints + findDepth + assignFeaturesToGrid (frame.cpp:36-133) for a batch of frames.

    kps [F, cap] (KP_DTYPE) and counts [F] as returned by the batch extractor; depth [F, H, W] float32 or None.
    Returns (unkps [F, cap], uright [F, cap], depth [F, cap], cell_start [F, 64*48+1], ids [F, cap])."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    if kps.ndim == 1:
        kps = kps[None]
    F, cap = kps.shape
    counts = np.ascontiguousarray(counts, np.int32).reshape(F)
    un = np.zeros((F, cap), KP_DTYPE); ur = np.zeros((F, cap), np.float32); dp = np.zeros((F, cap), np.float32)
    start = np.zeros((F, 64 * 48 + 1), np.int32); ids = np.zeros((F, cap), np.int32)
    W = H = 0; rs = fs = 0; dptr = None
    if depth is not None:
        depth = np.ascontiguousarray(depth, np.float32).reshape(F, *depth.shape[-2:])
        H, W = depth.shape[1:]; rs, fs = depth.strides[1], depth.strides[0]; dptr = _p(depth)
    _check(load_library().orbx_frame_finish(C.byref(cam), _p(kps), _p(counts), F, cap, dptr, W, H, rs, fs, _p(un), _p(ur), _p(dp),
                                            _p(start), _p(ids), device))
    return un, ur, dp, start, ids


class Frame:
    """Device-resident frame (orbx_frame_t): Frame::Frame (frame.cpp:22-32) = extraction + undistortKeyPoints + findDepth +
    assignFeaturesToGrid in one call; keypoints / descriptors / grid stay in HBM for the searches (`*_h` entry points).
    Host copies of the members: kps, desc, unkps, uright, depth."""

    def __init__(self, extractor, cam, image, depth=None):
        img = np.ascontiguousarray(image, np.uint8)
        H, W = img.shape
        self._lib = load_library()
        self._ex = extractor             # the frame lives in its extractor's pool: keep the extractor alive
        self._h = C.c_void_p()
        n = C.c_int(0)
        dptr, ds = None, 0
        if depth is not None:
            depth = np.ascontiguousarray(depth, np.float32)
            assert depth.shape == (H, W)
            dptr, ds = _p(depth), depth.strides[0]
        _check(self._lib.orbx_frame_create(extractor._h, C.byref(cam), _p(img), W, H, img.strides[0], dptr, ds, C.byref(self._h),
                                           C.byref(n)))
        self.n = n.value
        self.bounds = (cam.xmin, cam.xmax, cam.ymin, cam.ymax)
        m = max(self.n, 1)
        self.kps = np.zeros(m, KP_DTYPE); self.desc = np.zeros((m, 32), np.uint8); self.unkps = np.zeros(m, KP_DTYPE)
        self.uright = np.zeros(m, np.float32); self.depth = np.zeros(m, np.float32)
        _check(self._lib.orbx_frame_get(self._h, _p(self.kps), _p(self.desc), _p(self.unkps), _p(self.uright), _p(self.depth), m))
        self.kps, self.desc, self.unkps = self.kps[:self.n], self.desc[:self.n], self.unkps[:self.n]
        self.uright, self.depth = self.uright[:self.n], self.depth[:self.n]

    def grid(self):
        """(cell_start[64*48+1], ids) like grid_build()."""
        start = np.zeros(64 * 48 + 1, np.int32); ids = np.zeros(max(self.n, 1), np.int32)
        _check(self._lib.orbx_frame_grid(self._h, _p(start), _p(ids), max(self.n, 1)))
        return start, ids[:start[-1]].copy()

    def close(self):
        if self._h:
            self._lib.orbx_frame_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class UploadedFrame:
    """orbx_frame_upload: an existing host-side frame dict (kps = unKeypoints_, desc, uright, bounds, scale_factors) made
    resident once; usable wherever a Frame handle is (the *_h searches)."""

    def __init__(self, frame, device=0):
        keep = []
        v = Matcher._frame_view(dict(frame, occupied0=np.zeros(max(len(frame["kps"]), 1), np.uint8)), keep)
        self._lib = load_library()
        self._h = C.c_void_p()
        _check(self._lib.orbx_frame_upload(C.byref(v), device, C.byref(self._h)))
        self.n = v.n

    def close(self):
        if self._h:
            self._lib.orbx_frame_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Matcher:
    """Mirror of myslam::Matcher (matcher.h:9-45) for the hot-path overloads; map points arrive already projected."""

    def __init__(self, ratio=0.7, device=0):
        self.ratio_ = float(ratio)
        self.device = device
        self._lib = load_library()

    @staticmethod
    def computeDistance(desp1, desp2, device=0):
        """Matcher::computeDistance (matcher.cpp:1240-1256) of two 32-byte descriptors."""
        a = np.ascontiguousarray(desp1, np.uint8).reshape(1, 32)
        b = np.ascontiguousarray(desp2, np.uint8).reshape(1, 32)
        idx, d1, d2, ok = Matcher(1.0, device).knn2(a, b, th=256)
        return int(d1[0])

    def knn2(self, queries, train, th=TH_LOW):
        """Best/second-best + ratio loop of matcher.cpp:481-507 over all pairs -> (idx, d1, d2, ok)."""
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        nq = len(q)
        idx = np.zeros(nq, np.int32); d1 = np.zeros(nq, np.int32); d2 = np.zeros(nq, np.int32); ok = np.zeros(nq, np.uint8)
        _check(self._lib.hamm_knn2(_p(q), nq, _p(t), len(t), th, self.ratio_, _p(idx), _p(d1), _p(d2), _p(ok), self.device))
        return idx, d1, d2, ok

    @staticmethod
    def _frame_view(frame, keep):
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        v = _FrameView()
        v.kps = a(frame["kps"], KP_DTYPE); v.desc = a(frame["desc"], np.uint8); v.uright = a(frame["uright"], np.float32)
        v.n = len(frame["kps"])
        v.xmin, v.xmax, v.ymin, v.ymax = [float(b) for b in frame["bounds"]]
        v.scale_factors = a(frame["scale_factors"], np.float32); v.nlevels = len(frame["scale_factors"])
        v.occupied0 = a(frame["occupied0"], np.uint8)
        return v

    def searchByProjection(self, frame_curr, frame_last_points, radius, checkRot=True, bf=40.0, forward=False,
                           backward=False):
        """Matcher::searchByProjection(Frame*, Frame*, radius, checkRot) (matcher.cpp:18-148).
        Returns (assign, match_cnt): assign[i] = map-point index written to frame_curr.mappoints_[i], -1 none,
        -2 cleared by the rotation check."""
        keep = []
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        fv = self._frame_view(frame_curr, keep)
        p = frame_last_points
        s = _SbpFramePoints()
        s.m = len(p["u"]); s.valid = a(p["valid"], np.uint8); s.u = a(p["u"], np.float32); s.v = a(p["v"], np.float32)
        s.invz = a(p["invz"], np.float32); s.octave = a(p["octave"], np.int32); s.angle = a(p["angle"], np.float32)
        s.desc = a(p["desc"], np.uint8); s.has_obs = a(p["has_obs"], np.uint8)
        assign = np.zeros(max(fv.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_frame(C.byref(fv), C.byref(s), radius, bf, int(forward), int(backward),
                                                         int(checkRot), _p(assign), C.byref(cnt), self.device))
        return assign[:fv.n].copy(), cnt.value

    def _frame_points(self, p, keep):
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        s = _SbpFramePoints()
        s.m = len(p["u"]); s.valid = a(p["valid"], np.uint8); s.u = a(p["u"], np.float32); s.v = a(p["v"], np.float32)
        s.invz = a(p["invz"], np.float32); s.octave = a(p["octave"], np.int32); s.angle = a(p["angle"], np.float32)
        s.desc = a(p["desc"], np.uint8); s.has_obs = a(p["has_obs"], np.uint8)
        return s

    def searchByProjectionKeyFrame(self, frame_curr, keyframe_points, radius, distThreshold, checkRot=True):
        """Matcher::searchByProjection(Frame*, KeyFrame*, radius, distThreshold, found, checkRot) (matcher.cpp:150-272)."""
        keep = []
        fv = self._frame_view(frame_curr, keep); s = self._frame_points(keyframe_points, keep)
        assign = np.zeros(max(fv.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_reloc(C.byref(fv), C.byref(s), radius, distThreshold, int(checkRot), _p(assign),
                                                         C.byref(cnt), self.device))
        return assign[:fv.n].copy(), cnt.value

    def searchByProjectionSim3(self, keyframe, loop_points, th):
        """Matcher::searchByProjection(KeyFrame*, Sim3&, loopMapPoints, matchMapPoints, th) (matcher.cpp:356-447)."""
        keep = []
        fv = self._frame_view(keyframe, keep); s = self._frame_points(loop_points, keep)
        assign = np.zeros(max(fv.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_sim3(C.byref(fv), C.byref(s), int(th), _p(assign), C.byref(cnt), self.device))
        return assign[:fv.n].copy(), cnt.value

    def windowArgmin(self, keyframe, points, thRadius, distThreshold, chi2=False):
        """Search core of searchBySim3 / fuseMapPoints (chi2=True) / fuseByPose: best keyframe feature per projected point
        (levels [l-1, l], strict '<' in window order), -1 when the best distance exceeds distThreshold."""
        keep = []
        fv = self._frame_view(keyframe, keep); s = self._frame_points(points, keep)
        best = np.zeros(max(s.m, 1), np.int32)
        _check(self._lib.orbx_window_argmin(C.byref(fv), C.byref(s), thRadius, distThreshold, int(chi2), _p(best), self.device))
        return best[:s.m].copy()

    def searchBySim3(self, kf1, pts12, kf2, pts21, th):
        """Matcher::searchBySim3 (matcher.cpp:679-865): (match12, found)."""
        keep = []
        v1 = self._frame_view(kf1, keep); v2 = self._frame_view(kf2, keep)
        s12 = self._frame_points(pts12, keep); s21 = self._frame_points(pts21, keep)
        match = np.zeros(max(s12.m, 1), np.int32); found = C.c_int(0)
        _check(self._lib.orbx_search_by_sim3(C.byref(v1), C.byref(s12), C.byref(v2), C.byref(s21), th, _p(match), C.byref(found),
                                             self.device))
        return match[:s12.m].copy(), found.value

    def searchForTriangulation(self, side1, side2, F12, epipole, scale_factors2, checkRot=True, th_low=TH_LOW):
        """Matcher::searchForTriangulation (matcher.cpp:867-1010).  side dicts = BoW side + 'kps' + 'uright'; returns
        (match12, match_cnt) with match12[i] = kf2 feature of kf1 feature i (-1 none, -2 cleared by the rotation check)."""
        keep = []
        def tri(d):
            def a(x, dt):
                y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
            t = _TriSide()
            t.side.n = len(d["desc"]); t.side.desc = a(d["desc"], np.uint8); t.side.angle = a(d["angle"], np.float32)
            t.side.valid = a(d["valid"], np.uint8); t.side.ngroups = len(d["node_ids"]); t.side.node_ids = a(d["node_ids"], np.uint32)
            t.side.group_start = a(d["group_start"], np.int32); t.side.feat_idx = a(d["feat_idx"], np.int32)
            t.kps = a(d["kps"], KP_DTYPE); t.uright = a(d["uright"], np.float32)
            return t
        ta, tb = tri(side1), tri(side2)
        F = np.ascontiguousarray(F12, np.float64).reshape(9); sc = np.ascontiguousarray(scale_factors2, np.float32)
        match = np.zeros(max(ta.side.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_for_triangulation(C.byref(ta), C.byref(tb), _p(F), epipole[0], epipole[1], _p(sc), len(sc),
                                                       th_low, int(checkRot), _p(match), C.byref(cnt), self.device))
        return match[:ta.side.n].copy(), cnt.value

    def searchByBoW(self, side_a, side_b, mode=0, checkRot=True, th_low=TH_LOW):
        """Matcher::searchByBoW: mode 0 = (KeyFrame*, Frame*) (matcher.cpp:449-559), mode 1 = (KeyFrame*, KeyFrame*)
        (matcher.cpp:561-677).  Each side: dict(desc, angle, valid, node_ids, group_start, feat_idx) with the DBoW3
        FeatureVector as a CSR sorted by node id.  Returns (match, match_cnt); see include/orb_b200.h."""
        keep = []
        def side(d):
            def a(x, dt):
                y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
            s = _BowSide()
            s.n = len(d["desc"]); s.desc = a(d["desc"], np.uint8); s.angle = a(d["angle"], np.float32)
            s.valid = a(d["valid"], np.uint8); s.ngroups = len(d["node_ids"]); s.node_ids = a(d["node_ids"], np.uint32)
            s.group_start = a(d["group_start"], np.int32); s.feat_idx = a(d["feat_idx"], np.int32)
            return s
        sa, sb = side(side_a), side(side_b)
        n_out = sb.n if mode == 0 else sa.n
        match = np.zeros(max(n_out, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_bow(C.byref(sa), C.byref(sb), mode, self.ratio_, th_low, int(checkRot), _p(match),
                                            C.byref(cnt), self.device))
        return match[:n_out].copy(), cnt.value

    # ---- the tracking-thread searches against a device-resident Frame (only the points and occupied0 go up) -----------------
    def searchByProjectionH(self, frame, occupied0, frame_last_points, radius, checkRot=True, bf=40.0, forward=False, backward=False):
        keep = []
        s = self._frame_points(frame_last_points, keep)
        occ = np.ascontiguousarray(occupied0, np.uint8)
        assign = np.zeros(max(frame.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_frame_h(frame._h, _p(occ), C.byref(s), radius, bf, int(forward), int(backward),
                                                           int(checkRot), _p(assign), C.byref(cnt)))
        return assign[:frame.n].copy(), cnt.value

    def searchByProjectionKeyFrameH(self, frame, occupied0, keyframe_points, radius, distThreshold, checkRot=True):
        keep = []
        s = self._frame_points(keyframe_points, keep)
        occ = np.ascontiguousarray(occupied0, np.uint8)
        assign = np.zeros(max(frame.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_reloc_h(frame._h, _p(occ), C.byref(s), radius, distThreshold, int(checkRot),
                                                           _p(assign), C.byref(cnt)))
        return assign[:frame.n].copy(), cnt.value

    def searchByProjectionLocalH(self, frame, occupied0, mappoints, thRadius):
        keep = []
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        p = mappoints
        s = _SbpLocalPoints()
        s.m = len(p["u"]); s.valid = a(p["valid"], np.uint8); s.u = a(p["u"], np.float32); s.v = a(p["v"], np.float32)
        s.ur = a(p["ur"], np.float32); s.level = a(p["level"], np.int32); s.view_cos = a(p["view_cos"], np.float32)
        s.desc = a(p["desc"], np.uint8); s.has_obs = a(p["has_obs"], np.uint8)
        occ = np.ascontiguousarray(occupied0, np.uint8)
        assign = np.zeros(max(frame.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_local_h(frame._h, _p(occ), C.byref(s), thRadius, self.ratio_, _p(assign), C.byref(cnt)))
        return assign[:frame.n].copy(), cnt.value

    def searchByBoWH(self, side_a, frame, frame_groups, checkRot=True, th_low=TH_LOW):
        """searchByBoW(KeyFrame*, Frame*) with the frame resident: frame_groups = dict(valid, node_ids, group_start, feat_idx)."""
        keep = []
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        sa = _BowSide()
        sa.n = len(side_a["desc"]); sa.desc = a(side_a["desc"], np.uint8); sa.angle = a(side_a["angle"], np.float32)
        sa.valid = a(side_a["valid"], np.uint8); sa.ngroups = len(side_a["node_ids"]); sa.node_ids = a(side_a["node_ids"], np.uint32)
        sa.group_start = a(side_a["group_start"], np.int32); sa.feat_idx = a(side_a["feat_idx"], np.int32)
        sb = _BowSide()
        sb.n = frame.n; sb.desc = None; sb.angle = None; sb.valid = a(frame_groups["valid"], np.uint8)
        sb.ngroups = len(frame_groups["node_ids"]); sb.node_ids = a(frame_groups["node_ids"], np.uint32)
        sb.group_start = a(frame_groups["group_start"], np.int32); sb.feat_idx = a(frame_groups["feat_idx"], np.int32)
        match = np.zeros(max(frame.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_bow_h(C.byref(sa), frame._h, C.byref(sb), self.ratio_, th_low, int(checkRot), _p(match), C.byref(cnt)))
        return match[:frame.n].copy(), cnt.value

    def searchByProjectionLocal(self, frame, mappoints, thRadius):
        """Matcher::searchByProjection(Frame*, const vector<MapPoint*>&, thRadius) (matcher.cpp:274-353)."""
        keep = []
        def a(x, dt):
            y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
        fv = self._frame_view(frame, keep)
        p = mappoints
        s = _SbpLocalPoints()
        s.m = len(p["u"]); s.valid = a(p["valid"], np.uint8); s.u = a(p["u"], np.float32); s.v = a(p["v"], np.float32)
        s.ur = a(p["ur"], np.float32); s.level = a(p["level"], np.int32); s.view_cos = a(p["view_cos"], np.float32)
        s.desc = a(p["desc"], np.uint8); s.has_obs = a(p["has_obs"], np.uint8)
        assign = np.zeros(max(fv.n, 1), np.int32); cnt = C.c_int(0)
        _check(self._lib.orbx_search_by_projection_local(C.byref(fv), C.byref(s), thRadius, self.ratio_, _p(assign),
                                                         C.byref(cnt), self.device))
        return assign[:fv.n].copy(), cnt.value
