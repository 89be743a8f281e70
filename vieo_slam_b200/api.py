"""ctypes binding of libvieo_b200.so (the C ABI in include/vieo_b200.h) with thin classes named after the
reference's (ORBextractor, ORBmatcher, IMUPreintegrator, Optimizer).  Used by tests and bench.py; a C++
caller uses vieo_slam_b200/host/*.h or the C ABI directly.  There is no fallback: if the library is missing
or no B200 is present every call raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvieo_b200.so")
if os.environ.get("VIEO_B200_LIB"):  # alternative build of the same library (profiling / kernel-variant experiments)
    LIB_PATH = os.environ["VIEO_B200_LIB"]

from .layouts import FRUSTUM_FRAME_DTYPE, PROJ_SEARCH_FRAME_DTYPE  # noqa: E402

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])


class VieoError(RuntimeError):
    pass


class VieoOrbConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("nfeatures", C.c_int32), ("scale_factor", C.c_float),
                ("nlevels", C.c_int32), ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32),
                ("max_batch", C.c_int32)]


class VieoSbpQueries(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("Xw", "level", "angle", "proj", "viewcos", "depth", "desc", "flags")]


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VieoError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.vieo_last_error.restype = C.c_char_p
        L.vieo_version.restype = C.c_char_p
        L.vieo_orb_create.argtypes = [C.POINTER(VieoOrbConfig), i32, C.POINTER(vp)]
        L.vieo_orb_destroy.argtypes = [vp]
        L.vieo_orb_destroy.restype = None
        L.vieo_orb_max_keypoints.argtypes = [vp]
        L.vieo_orb_get_tables.argtypes = [vp] * 8
        L.vieo_orb_extract.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp, vp]
        L.vieo_orb_extract_batch.argtypes = [vp, i32, vp, sz, i32, vp, vp, i32, vp]
        L.vieo_orb_extract_batch_dev.argtypes = [vp, i32, vp, sz, i32, vp, vp, i32, vp, vp]
        L.vieo_orb_debug_level.argtypes = [vp, i32, i32, vp]
        L.vieo_orb_debug_candidates.argtypes = [vp, i32, i32, vp, i32]
        L.vieo_orb_last_launches.argtypes = [vp]
        L.vieo_orb_profile.argtypes = [vp, i32]
        L.vieo_orb_profile_read.argtypes = [vp, vp, vp]
        L.vieo_hamming_knn2.argtypes = [vp, i32, vp, i32, vp, vp, i32]
        L.vieo_hamming_knn2_batch_dev.argtypes = [vp, sz, vp, i32, vp, sz, vp, i32, i32, i32, vp, vp, vp]
        L.vieo_sm_partition_create.argtypes = [i32, i32, C.POINTER(vp)]
        L.vieo_sm_partition_destroy.argtypes = [vp]
        L.vieo_sm_partition_destroy.restype = None
        L.vieo_sm_partition_sms.argtypes = [vp, i32]
        L.vieo_sm_partition_bind_thread.argtypes = [vp, i32]
        L.vieo_sm_partition_stream.argtypes = [vp, i32, i32]
        L.vieo_sm_partition_stream.restype = vp
        L.vieo_sbp_scratch_bytes.argtypes = [i32]
        L.vieo_sbp_scratch_bytes.restype = sz
        L.vieo_sbp_batch.argtypes = [i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]
        L.vieo_sbp_batch_dev.argtypes = [i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
        L.vieo_frustum_level_table.argtypes = [C.c_float, i32, vp]
        L.vieo_frustum_batch.argtypes = [vp, i32] + [vp] * 11 + [i32]
        L.vieo_frustum_batch_dev.argtypes = [vp, i32] + [vp] * 12
        L.vieo_search_local_points.argtypes = [vp, vp, i32] + [vp] * 21 + [i32]
        L.vieo_proj_search_batch.argtypes = [vp, i32] + [vp] * 12 + [i32]
        L.vieo_proj_search_batch_dev.argtypes = [vp, i32] + [vp] * 13
        L.vieo_distinctive_descriptors.argtypes = [vp, i32, vp, vp, i32, vp, vp, i32]
        L.vieo_distinctive_descriptors_dev.argtypes = [vp, vp, vp, i32, vp, vp, vp]
        L.vieo_imu_init_gyro_bias.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32]
        L.vieo_gyro_bias_init_dev.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.vieo_frontend_create.argtypes = [C.POINTER(VieoOrbConfig), i32, i32, C.POINTER(vp)]
        L.vieo_frontend_destroy.argtypes = [vp]
        L.vieo_frontend_destroy.restype = None
        L.vieo_frontend_max_keypoints.argtypes = [vp]
        L.vieo_frontend_last_launches.argtypes = [vp]
        L.vieo_frontend_process.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp]
        L.vieo_frontend_stereo_rectified.argtypes = [vp, i32, C.c_float, C.c_float, vp, vp, vp]
        L.vieo_orb_stereo_match_dev.argtypes = [vp, i32, vp, vp, vp, i32, C.c_float, C.c_float, vp, vp, vp, vp]
        L.vieo_hamming_csr.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, i32]
        L.vieo_imu_set_param.argtypes = [vp, vp, i32, C.c_double]
        L.vieo_imu_set_param.restype = None
        L.vieo_imu_preint_batch.argtypes = [vp, vp, vp, vp, vp, i32, vp, i32]
        L.vieo_imu_preint_batch_dev.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp]
        L.vieo_pose_opt_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32]
        L.vieo_pose_opt_batch_dev.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.vieo_multicam_frames.argtypes = [vp, i32, i32, vp, C.c_size_t, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        L.vieo_lapping_split_dev.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.vieo_fisheye_knn_dev.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.vieo_sft_scratch_bytes.argtypes = [i32, i32]
        L.vieo_sft_scratch_bytes.restype = sz
        L.vieo_search_for_triangulation_dev.argtypes = [vp, i32] + [vp] * 7 + [i32, i32, vp, vp, vp, vp, sz, vp]
        L.vieo_search_for_triangulation.argtypes = [vp, i32] + [vp] * 7 + [i32] * 6 + [vp, vp, vp, i32]
        L.vieo_search_by_bow_dev.argtypes = [vp, i32] + [vp] * 6 + [i32, vp, vp, vp, sz, vp]
        L.vieo_search_by_bow.argtypes = [vp, i32] + [vp] * 6 + [i32] * 5 + [vp, vp, i32]
        L.vieo_ba_create.argtypes = [i32, i32, i32, i32, i32, C.POINTER(vp)]
        L.vieo_ba_create_global.argtypes = [i32, i32, i32, i32, i32, C.POINTER(vp)]
        L.vieo_ba_stream_priority.argtypes = [i32]
        L.vieo_ba_stream_priority.restype = None
        L.vieo_ba_destroy.argtypes = [vp]
        L.vieo_ba_destroy.restype = None
        L.vieo_ba_set_sharding.argtypes = [vp, i32, i32, vp, vp]
        L.vieo_comm_unique_id.argtypes = [vp]
        L.vieo_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
        L.vieo_comm_destroy.argtypes = [vp]
        L.vieo_comm_destroy.restype = None
        L.vieo_comm_allreduce_f64.argtypes = [vp, vp, C.c_size_t, vp]
        L.vieo_ba_set_comm.argtypes = [vp, vp]
        L.vieo_ba_stream.argtypes = [vp]
        L.vieo_ba_stream.restype = vp
        L.vieo_local_ba_prv.argtypes = [vp] * 9
        L.vieo_local_ba_prv_begin.argtypes = [vp] * 4
        L.vieo_local_ba_prv_poll.argtypes = [vp]
        L.vieo_local_ba_prv_end.argtypes = [vp] * 6
        L.vieo_local_ba_prv_batch.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        L.vieo_global_ba_prv.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.vieo_global_ba_prv_ex.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        L.vieo_ba_get_border.argtypes = [vp, vp, vp]
        L.vieo_ba_set_problem.argtypes = [vp, vp, vp]
        L.vieo_ba_chi2_large_set_level.argtypes = [vp, C.c_float]
        L.vieo_ba_active_robust_chi2.argtypes = [vp, i32, vp]
        L.vieo_ba_optimize.argtypes = [vp, i32, C.c_double, vp]
        L.vieo_ba_reclassify.argtypes = [vp, i32, vp]
        L.vieo_ba_get.argtypes = [vp, vp, vp, vp]
        L.vieo_ba_debug_step.argtypes = [vp, C.c_double, vp, vp, vp, vp]
        L.vieo_ba_last_launches.argtypes = [vp]
        L.vieo_ba_last_ms.argtypes = [vp]
        L.vieo_ba_last_ms.restype = C.c_double
        L.vieo_ba_last_trials.argtypes = [vp]
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise VieoError(f"vieo_b200 error {rc}: {lib().vieo_last_error().decode()}")
    return rc


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


SM_FRONTEND, SM_BA = 0, 1


class SmPartition:
    """vieo_sm_partition_*: `ba_sms` SMs of the device for the bundle-adjustment streams, the rest for the front-end and
    tracking streams (CUDA green contexts).  Handles and the calling thread's staging stream are created inside the
    partition the thread is bound to: `with part.bound(SM_BA): ba = BundleAdjuster(...)`."""

    def __init__(self, ba_sms=16, device=0):
        self._h = C.c_void_p()
        _check(lib().vieo_sm_partition_create(device, ba_sms, C.byref(self._h)))

    def sms(self, which):
        return lib().vieo_sm_partition_sms(self._h, which)

    def bind(self, which):
        _check(lib().vieo_sm_partition_bind_thread(self._h, which))

    @staticmethod
    def unbind():
        _check(lib().vieo_sm_partition_bind_thread(None, 0))

    def bound(self, which):
        part = self

        class _Ctx:
            def __enter__(self):
                part.bind(which)

            def __exit__(self, *a):
                SmPartition.unbind()
        return _Ctx()

    def stream(self, which, high_priority=False):
        s = lib().vieo_sm_partition_stream(self._h, which, int(high_priority))
        if not s:
            raise VieoError("vieo_sm_partition_stream failed")
        return s


class ORBextractor:
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) (include/ORBextractor.h:31) for a
    fixed image size; `max_batch` images (cameras x frames) can be extracted per call."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height, max_batch=2, device=0):
        self.cfg = VieoOrbConfig(width, height, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_batch)
        self._h = C.c_void_p()
        _check(lib().vieo_orb_create(C.byref(self.cfg), device, C.byref(self._h)))
        self.nlevels = nlevels
        self.width, self.height = width, height
        self.cap = lib().vieo_orb_max_keypoints(self._h)
        t = [np.empty(nlevels, np.float32) for _ in range(4)] + [np.empty(nlevels, np.int32) for _ in range(3)]
        _check(lib().vieo_orb_get_tables(self._h, *[_p(x) for x in t]))
        (self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2, self.quota,
         self.level_w, self.level_h) = t
        self.mvImagePyramid = []

    def close(self):
        if getattr(self, "_h", None) and self._h and _lib is not None:  # _lib is gone at interpreter shutdown
            _lib.vieo_orb_destroy(self._h)
            self._h = None

    __del__ = close

    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(self.cfg.scale_factor)

    def GetScaleFactors(self):
        return self.mvScaleFactor

    def GetInverseScaleFactors(self):
        return self.mvInvScaleFactor

    def GetScaleSigmaSquares(self):
        return self.mvLevelSigma2

    def GetInverseScaleSigmaSquares(self):
        return self.mvInvLevelSigma2

    def __call__(self, image, mask=None, pvLappingArea=None, want_pyramid=False):
        """operator(): returns (ret, keypoints[KP_DTYPE], descriptors[n,32]); ret = monoIndex or -1 (empty image)."""
        if image is None or image.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 only (src/ORBextractor.cc:973)"
        assert image.shape == (self.height, self.width)
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        kps = np.empty(self.cap, KP_DTYPE)
        desc = np.empty((self.cap, 32), np.uint8)
        mono = C.c_int32(0)
        lap = None if pvLappingArea is None else np.asarray(pvLappingArea, np.int32)
        pyr = None
        if want_pyramid:
            self.mvImagePyramid = [np.empty((int(h), int(w)), np.uint8) for w, h in zip(self.level_w, self.level_h)]
            pyr = (C.c_void_p * self.nlevels)(*[x.ctypes.data for x in self.mvImagePyramid])
        n = _check(lib().vieo_orb_extract(self._h, _p(image), image.strides[0], _p(lap), _p(kps), _p(desc), self.cap,
                                          C.byref(mono), pyr))
        return mono.value, kps[:n].copy(), desc[:n].copy()

    def extract_batch(self, images):
        """images: (n, H, W) u8 host array -> (kps[n,cap], desc[n,cap,32], n_kp[n])."""
        images = np.ascontiguousarray(images, np.uint8)
        n = images.shape[0]
        kps = np.empty((n, self.cap), KP_DTYPE)
        desc = np.empty((n, self.cap, 32), np.uint8)
        nk = np.empty(n, np.int32)
        _check(lib().vieo_orb_extract_batch(self._h, n, _p(images), images.strides[0], images.strides[1], _p(kps),
                                            _p(desc), self.cap, _p(nk)))
        return kps, desc, nk

    def multicam_frames(self, images, n_cams, lapping=None):
        """Multi-camera frames (KB8 rigs): images (n_frames * n_cams, H, W) u8 ordered [frame][camera]; lapping =
        [n_cams][2] or None.  Per-camera ORBextractor::operator() with the lapping area (Frame::Frame, src/Frame.cc:259-278)
        + the brute-force half of Frame::ComputeStereoFishEyeMatches (:613-663) for every camera pair.
        -> dict(kps [n_img, cap], desc, n_kp, n_mono, pair_idx [n_frames, n_pairs, cap, 2], pair_dist, pair_good)."""
        images = np.ascontiguousarray(images, np.uint8)
        n_img = images.shape[0]
        assert n_img % n_cams == 0
        n_frames, n_pairs = n_img // n_cams, n_cams * (n_cams - 1) // 2
        kps = np.empty((n_img, self.cap), KP_DTYPE); desc = np.empty((n_img, self.cap, 32), np.uint8)
        nk = np.empty(n_img, np.int32); nm = np.empty(n_img, np.int32)
        pi = np.empty((n_frames, n_pairs, self.cap, 2), np.int32); pd = np.empty((n_frames, n_pairs, self.cap, 2), np.int32)
        pg = np.empty((n_frames, n_pairs, self.cap), np.uint8)
        lap = None if lapping is None else np.ascontiguousarray(lapping, np.int32).reshape(n_cams, 2)
        _check(lib().vieo_multicam_frames(self._h, n_frames, n_cams, _p(images), images.strides[0], images.strides[1],
                                          _p(lap), _p(kps), _p(desc), _p(nk), _p(nm), _p(pi), _p(pd), _p(pg)))
        return dict(kps=kps, desc=desc, n_kp=nk, n_mono=nm, pair_idx=pi, pair_dist=pd, pair_good=pg)

    def extract_batch_dev(self, imgs_ptr, n, img_stride, row_stride, kps_ptr, desc_ptr, cap, nkp_ptr, stream=0):
        """All pointers are device addresses (ints); asynchronous on `stream`."""
        _check(lib().vieo_orb_extract_batch_dev(self._h, n, imgs_ptr, img_stride, row_stride, kps_ptr, desc_ptr, cap,
                                                nkp_ptr, stream))

    def stereo_match_dev(self, n_frames, kps_ptr, desc_ptr, nkp_ptr, cap, bf, min_z, ur_ptr, depth_ptr, sad_ptr, stream=0):
        """Device-resident Frame::ComputeStereoMatches over the last extract_batch_dev call (images 2f / 2f+1)."""
        _check(lib().vieo_orb_stereo_match_dev(self._h, n_frames, kps_ptr, desc_ptr, nkp_ptr, cap, bf, min_z, ur_ptr,
                                               depth_ptr, sad_ptr, stream))

    def last_launches(self):
        return lib().vieo_orb_last_launches(self._h)

    def profile(self, enable=True):
        _check(lib().vieo_orb_profile(self._h, int(enable)))

    def profile_read(self):
        """-> (dict stage -> summed ms, number of profiled calls)"""
        ms = np.zeros(4, np.float32)
        n = C.c_int32(0)
        _check(lib().vieo_orb_profile_read(self._h, _p(ms), C.byref(n)))
        return dict(zip(("pyramid", "fast_cells", "quadtree", "orient_desc"), ms.tolist())), n.value

    def debug_level(self, img_index, level):
        out = np.empty((int(self.level_h[level]), int(self.level_w[level])), np.uint8)
        _check(lib().vieo_orb_debug_level(self._h, img_index, level, _p(out)))
        return out

    def debug_candidates(self, img_index, level):
        cap = 1 << 17
        out = np.empty((cap, 3), np.int32)
        n = _check(lib().vieo_orb_debug_candidates(self._h, img_index, level, _p(out), cap))
        return out[:n].copy()


def sbp_batch_dev(mode, frames_ptr, n_frames, kps_ptr, ur_ptr, desc_ptr, queries, blocked_ptr, kp_match_ptr, q_match_ptr,
                  q_dist_ptr, n_matches_ptr, scratch_ptr, scratch_bytes, stream=0):
    """vieo_sbp_batch_dev: device-resident guided searches; `queries` is a VieoSbpQueries of device pointers."""
    _check(lib().vieo_sbp_batch_dev(mode, frames_ptr, n_frames, kps_ptr, ur_ptr, desc_ptr, C.byref(queries), blocked_ptr,
                                    kp_match_ptr, q_match_ptr, q_dist_ptr, n_matches_ptr, scratch_ptr, scratch_bytes, stream))


def frustum_batch_dev(frames_ptr, n_frames, wP_ptr, normal_ptr, max_dist_ptr, min_dist_ptr, skip_ptr, inview_ptr, proj_ptr,
                      level_ptr, viewcos_ptr, depth_ptr, n_inview_ptr, stream=0):
    """vieo_frustum_batch_dev: device-resident Frame::isInFrustum; the frames' level_ratio tables must be filled
    (frustum_level_table)."""
    _check(lib().vieo_frustum_batch_dev(frames_ptr, n_frames, wP_ptr, normal_ptr, max_dist_ptr, min_dist_ptr, skip_ptr,
                                        inview_ptr, proj_ptr, level_ptr, viewcos_ptr, depth_ptr, n_inview_ptr, stream))


def _frustum_outputs(nq, nf):
    return dict(inview=np.zeros(nq, np.uint8), proj=np.zeros((nq, 3), np.float32), level=np.full(nq, -1, np.int32),
                viewcos=np.zeros(nq, np.float32), depth=np.zeros(nq, np.float32), n_inview=np.zeros(nf, np.int32))


def frustum_level_table(log_scale_factor, n_levels):
    """vieo_frustum_level_table: smallest float ratio reaching each pyramid level under MapPoint::PredictScale."""
    t = np.zeros(16, np.float32)
    _check(lib().vieo_frustum_level_table(C.c_float(log_scale_factor), int(n_levels), _p(t)))
    return t


def isInFrustum(pb, device=0):
    """Frame::isInFrustum (src/Frame.cc:335-416) + MapPoint::PredictScale for every candidate point of every frame of a
    synth.make_frustum_problem dict -> dict(inview, proj, level, viewcos, depth, n_inview)."""
    ff = np.ascontiguousarray(pb["frustum"], FRUSTUM_FRAME_DTYPE)
    wP, Pn = np.ascontiguousarray(pb["p_wP"], np.float32), np.ascontiguousarray(pb["p_normal"], np.float32)
    mx, mn = np.ascontiguousarray(pb["p_max_dist"], np.float32), np.ascontiguousarray(pb["p_min_dist"], np.float32)
    skip = None if pb.get("p_skip") is None else np.ascontiguousarray(pb["p_skip"], np.uint8)
    out = _frustum_outputs(len(mx), len(ff))
    _check(lib().vieo_frustum_batch(_p(ff), len(ff), _p(wP), _p(Pn), _p(mx), _p(mn), _p(skip), _p(out["inview"]),
                                    _p(out["proj"]), _p(out["level"]), _p(out["viewcos"]), _p(out["depth"]),
                                    _p(out["n_inview"]), device))
    return out


class ORBmatcher:
    """The Hamming kernels behind ORBmatcher / Frame stereo association (include/ORBmatcher.h:18-113)."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30  # src/ORBmatcher.cc:20-22

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio, self.mbCheckOrientation, self.device = nnratio, checkOri, device

    def SearchByProjection(self, pb):
        """Batched ORBmatcher::SearchByProjection (src/ORBmatcher.cc:1303-1467 when pb["mode"] == SBP_LAST_FRAME, :230-335
        when SBP_LOCAL_MAP) over the frames of a problem dict (see synth.make_sbp_problem for the arrays).  The matcher's
        mfNNratio / mbCheckOrientation are written into every frame record.
        Returns (kp_match, q_match, q_dist, n_matches)."""
        fr = np.ascontiguousarray(pb["frames"]).copy()
        fr["nn_ratio"] = self.mfNNratio
        fr["check_orientation"] = int(self.mbCheckOrientation)
        keep = []

        def arr(k, dt):
            a = pb.get(k)
            if a is None:
                return None
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a
        q = VieoSbpQueries()
        for name, key, dt in (("Xw", "q_Xw", np.float64), ("level", "q_level", np.int32), ("angle", "q_angle", np.float32),
                              ("proj", "q_proj", np.float32), ("viewcos", "q_viewcos", np.float32),
                              ("depth", "q_depth", np.float32), ("desc", "q_desc", np.uint8), ("flags", "q_flags", np.uint8)):
            a = arr(key, dt)
            setattr(q, name, a.ctypes.data if a is not None and a.size else None)
        kps, ur, desc = arr("kps", KP_DTYPE), arr("uright", np.float32), arr("desc", np.uint8)
        blk = arr("kp_blocked", np.uint8)
        nq = len(pb["q_level"])
        kp_match = np.full(len(kps), -1, np.int32)   # entries outside every frame's range stay -1
        q_match = np.full(nq, -1, np.int32); q_dist = np.full(nq, -1, np.int32)
        nm = np.zeros(len(fr), np.int32)
        _check(lib().vieo_sbp_batch(int(pb["mode"]), _p(fr), len(fr), _p(kps), _p(ur), _p(desc), C.byref(q),
                                    _p(blk) if blk is not None else None, _p(kp_match), _p(q_match), _p(q_dist), _p(nm),
                                    self.device))
        return kp_match, q_match, q_dist, nm

    def SearchByProjectionReloc(self, pb):
        """Batched ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist, th_far_pts)
        (src/ORBmatcher.cc:1471-1606) over the frames of a synth.make_reloc_problem dict.
        Returns (kp_match, q_match, q_dist, q_level, n_matches)."""
        from .layouts import SBP_RELOC_DTYPE
        fr = np.ascontiguousarray(pb["frames"]).copy()
        fr["check_orientation"] = int(self.mbCheckOrientation)
        rl = np.ascontiguousarray(pb["reloc"], SBP_RELOC_DTYPE)
        a = {k: np.ascontiguousarray(pb[k], dt) for k, dt in (("kps", KP_DTYPE), ("desc", np.uint8), ("q_Xw", np.float64),
                                                               ("q_angle", np.float32), ("q_max_dist", np.float32),
                                                               ("q_min_dist", np.float32), ("q_desc", np.uint8))}
        blk = None if pb.get("kp_blocked") is None else np.ascontiguousarray(pb["kp_blocked"], np.uint8)
        nq = len(a["q_angle"])
        kp_match = np.full(len(a["kps"]), -1, np.int32)
        q_match = np.full(nq, -1, np.int32); q_dist = np.full(nq, -1, np.int32); q_level = np.full(nq, -1, np.int32)
        nm = np.zeros(len(fr), np.int32)
        _check(lib().vieo_sbp_reloc_batch(_p(fr), _p(rl), len(fr), _p(a["kps"]), _p(a["desc"]), _p(a["q_Xw"]), _p(a["q_angle"]),
                                          _p(a["q_max_dist"]), _p(a["q_min_dist"]), _p(a["q_desc"]),
                                          _p(blk) if blk is not None else None, _p(kp_match), _p(q_match), _p(q_dist),
                                          _p(q_level), _p(nm), self.device))
        return kp_match, q_match, q_dist, q_level, nm

    def knnMatch2(self, q, t):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) -> (idx[nq,2], dist[nq,2])."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.empty((len(q), 2), np.int32)
        dist = np.empty((len(q), 2), np.int32)
        _check(lib().vieo_hamming_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist), self.device))
        return idx, dist

    def search_candidates(self, q, t, row_ptr, cand):
        """Best / second-best over per-row candidate lists -> (best_dist, best_idx, second_dist, second_idx)."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        cand = np.ascontiguousarray(cand, np.int32)
        n = len(row_ptr) - 1
        o = [np.empty(n, np.int32) for _ in range(4)]
        _check(lib().vieo_hamming_csr(_p(q), _p(t), len(t), _p(row_ptr), _p(cand), n, *[_p(x) for x in o], self.device))
        return o

    def SearchLocalPoints(self, pb, want_tracking_info=True):
        """Tracking::SearchLocalPoints (src/Tracking.cc:2308-2368): Frame::isInFrustum over every candidate local map point
        followed by SearchByProjection(F, vpMapPoints, th, th_far) on the same stream (vieo_search_local_points).  `pb` is
        a synth.make_frustum_problem dict that also carries the guided-search arrays.
        Returns (frustum outputs dict, kp_match, q_match, q_dist, n_matches); want_tracking_info=False leaves the
        projections / levels / cosines / depths on the device (only inview and n_inview are filled)."""
        fr = np.ascontiguousarray(pb["frames"]).copy()
        fr["nn_ratio"] = self.mfNNratio
        fr["check_orientation"] = int(self.mbCheckOrientation)
        ff = np.ascontiguousarray(pb["frustum"], FRUSTUM_FRAME_DTYPE)
        a = {k: np.ascontiguousarray(pb[k], dt) for k, dt in (("p_wP", np.float32), ("p_normal", np.float32),
                                                               ("p_max_dist", np.float32), ("p_min_dist", np.float32),
                                                               ("q_desc", np.uint8), ("q_flags", np.uint8),
                                                               ("kps", KP_DTYPE), ("uright", np.float32), ("desc", np.uint8))}
        skip = None if pb.get("p_skip") is None else np.ascontiguousarray(pb["p_skip"], np.uint8)
        blk = None if pb.get("kp_blocked") is None else np.ascontiguousarray(pb["kp_blocked"], np.uint8)
        nq, nk = len(a["p_max_dist"]), len(a["kps"])
        out = _frustum_outputs(nq, len(ff))
        kp_match = np.full(nk, -1, np.int32)
        q_match = np.full(nq, -1, np.int32); q_dist = np.full(nq, -1, np.int32)
        nm = np.zeros(len(fr), np.int32)
        _check(lib().vieo_search_local_points(_p(ff), _p(fr), len(fr), _p(a["p_wP"]), _p(a["p_normal"]), _p(a["p_max_dist"]),
                                              _p(a["p_min_dist"]), _p(skip), _p(a["q_desc"]), _p(a["q_flags"]), _p(a["kps"]),
                                              _p(a["uright"]), _p(a["desc"]), _p(blk), _p(out["inview"]),
                                              *[_p(out[k]) if want_tracking_info else None for k in ("proj", "level", "viewcos", "depth")],
                                              _p(out["n_inview"]),
                                              _p(kp_match), _p(q_match), _p(q_dist), _p(nm), self.device))
        return out, kp_match, q_match, q_dist, nm

    def SearchByProjectionBase(self, pb):
        """The search half of ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227; Fuse, Sim3 and keyframe
        projection searches) over the keyframes of a synth.make_fuse_problem dict.
        -> (best_idx, best_dist, level) per map point; the caller applies th_bestdist and the FuseMP policy."""
        fr = np.ascontiguousarray(pb["frames"], PROJ_SEARCH_FRAME_DTYPE)
        a = {k: np.ascontiguousarray(pb[k], dt) for k, dt in (("kps", KP_DTYPE), ("uright", np.float32), ("desc", np.uint8),
                                                               ("p_wP", np.float32), ("p_normal", np.float32),
                                                               ("p_max_dist", np.float32), ("p_min_dist", np.float32),
                                                               ("q_desc", np.uint8))}
        skip = None if pb.get("p_skip") is None else np.ascontiguousarray(pb["p_skip"], np.uint8)
        nq = len(a["p_max_dist"])
        best = np.full(nq, -1, np.int32); dist = np.full(nq, -1, np.int32); lvl = np.full(nq, -1, np.int32)
        _check(lib().vieo_proj_search_batch(_p(fr), len(fr), _p(a["kps"]), _p(a["uright"]), _p(a["desc"]), _p(a["p_wP"]),
                                            _p(a["p_normal"]), _p(a["p_max_dist"]), _p(a["p_min_dist"]), _p(a["q_desc"]),
                                            _p(skip), _p(best), _p(dist), _p(lvl), self.device))
        return best, dist, lvl

    def Fuse(self, pb, th_bestdist=None):
        """ORBmatcher::Fuse(pKF, vpMapPoints, th) (:1152-1165) up to the map-point bookkeeping: per map point the keypoint
        it would be fused into (bestDist <= TH_LOW), else -1; -> (keypoint index per point, nFused per keyframe)."""
        th = self.TH_LOW if th_bestdist is None else th_bestdist
        best, dist, _ = self.SearchByProjectionBase(pb)
        hit = np.where((best >= 0) & (dist <= th), best, -1).astype(np.int32)
        fr = pb["frames"]
        return hit, np.array([int((hit[int(f["q_begin"]):int(f["q_begin"]) + int(f["n_q"])] >= 0).sum()) for f in fr], np.int32)

    @staticmethod
    def sim3_agreement(best, dist, has_mp, n_kp2, n_kp1, prior12=None, th_bestdist=100):
        """The host half of ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1222-1302) over the output of its two searches run as ONE
        two-frame SearchByProjectionBase batch — frame 0: keyframe 1's n_kp1 map points into keyframe 2 (keypoints 0 .. n_kp2), frame
        1: keyframe 2's n_kp2 map points into keyframe 1 (keypoints n_kp2 ..).  The searches run in the one-match mode (mode =
        ~SBPMatchMultiCam): the arg-min keypoint counts only if it holds a map point of its keyframe (has_mp, keypoint order of the
        batch) and bestDist <= TH_HIGH (:205-222); a pair is accepted when both directions agree (:1281-1297).  prior12: matches
        known before the call (their points were skipped by the searches, :1245-1258) — kept, not counted.
        -> (match12 [n_kp1]: keyframe-2 keypoint whose map point vpMatches12[i1] holds, -1 none; nFound)."""
        best = np.asarray(best); dist = np.asarray(dist); has_mp = np.asarray(has_mp).astype(bool)
        bA, dA, bB, dB = best[:n_kp1], dist[:n_kp1], best[n_kp1:n_kp1 + n_kp2], dist[n_kp1:n_kp1 + n_kp2]
        hasA, hasB = has_mp[:n_kp2], has_mp[n_kp2:n_kp2 + n_kp1]
        mA = np.where((bA >= 0) & (dA <= th_bestdist) & hasA[np.maximum(bA, 0)], bA, -1)
        mB = np.where((bB >= 0) & (dB <= th_bestdist) & hasB[np.maximum(bB, 0)], bB, -1)
        match12 = np.full(n_kp1, -1, np.int32) if prior12 is None else np.array(prior12, np.int32)
        agree = (mA >= 0) & (mB[np.maximum(mA, 0)] == np.arange(n_kp1))
        match12[agree] = mA[agree]
        return match12, int(agree.sum())

    def SearchBySim3(self, pb):
        """ORBmatcher::SearchBySim3 (LoopClosing::ComputeSim3): pb = the two-frame batch described in sim3_agreement (the caller forms
        the frames' poses sR21 R1w, sR21 t1w + t21 and sR12 R2w, sR12 t2w + t12, th_radius = th, no bf gate, no viewing-cone test; skips
        points without a map point or already matched) + 'has_mp' per keypoint and optionally 'prior12'."""
        fr = pb["frames"]
        if len(fr) != 2 or int(fr[0]["n_q"]) != int(fr[1]["n_kp"]) or int(fr[1]["n_q"]) != int(fr[0]["n_kp"]):
            raise VieoError("SearchBySim3 needs the two mutual searches as frames 0 and 1")
        best, dist, _ = self.SearchByProjectionBase(pb)
        return self.sim3_agreement(best, dist, pb["has_mp"], int(fr[0]["n_kp"]), int(fr[1]["n_kp"]), pb.get("prior12"), self.TH_HIGH)

    def ComputeDistinctiveDescriptors(self, desc_pool, ptr, rows=None):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:314-378) for a CSR batch of map points.
        -> (best index into each point's observation list or -1, that row's median distance)."""
        pool = np.ascontiguousarray(desc_pool, np.uint8).reshape(-1, 32)
        ptr = np.ascontiguousarray(ptr, np.int32)
        rows = None if rows is None else np.ascontiguousarray(rows, np.int32)
        n = len(ptr) - 1
        best = np.empty(n, np.int32); med = np.empty(n, np.int32)
        _check(lib().vieo_distinctive_descriptors(_p(pool), len(pool), _p(rows), _p(ptr), n, _p(best), _p(med), self.device))
        return best, med

    def DescriptorDistance(self, a, b):
        """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1645) evaluated on the device."""
        bd, _, _, _ = self.search_candidates(np.asarray(a).reshape(1, 32), np.asarray(b).reshape(1, 32), [0, 1], [0])
        return int(bd[0])


def hamming_knn2_batch_dev(q_ptr, q_stride, nq_ptr, max_nq, t_ptr, t_stride, nt_ptr, max_nt, count_stride, n_pairs,
                           idx_ptr, dist_ptr, stream=0):
    """Device-resident batched knnMatch(k=2); all pointers are device addresses."""
    _check(lib().vieo_hamming_knn2_batch_dev(q_ptr, q_stride, nq_ptr, max_nq, t_ptr, t_stride, nt_ptr, max_nt,
                                             count_stride, n_pairs, idx_ptr, dist_ptr, stream))


class StereoFrontend:
    """Frame::Frame stereo constructor hot path (src/Frame.cc:218-316) for a batch of frames with host buffers."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height, max_frames, device=0):
        cfg = VieoOrbConfig(width, height, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, 2)
        self._h = C.c_void_p()
        _check(lib().vieo_frontend_create(C.byref(cfg), max_frames, device, C.byref(self._h)))
        self.cap = lib().vieo_frontend_max_keypoints(self._h)
        self.width, self.height, self.max_frames = width, height, max_frames

    def close(self):
        if getattr(self, "_h", None) and self._h and _lib is not None:
            _lib.vieo_frontend_destroy(self._h)
            self._h = None

    __del__ = close

    def alloc_outputs(self, n_frames, pinned=False):
        import torch
        mk = (lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()) if pinned else \
             (lambda shape, dt: torch.empty(shape, dtype=dt).numpy())
        import torch as T
        kps = mk((2 * n_frames, self.cap, 6), T.float32)
        desc = mk((2 * n_frames, self.cap, 32), T.uint8)
        nkp = mk((2 * n_frames,), T.int32)
        midx = mk((n_frames, self.cap, 2), T.int32)
        mdist = mk((n_frames, self.cap, 2), T.int32)
        return kps, desc, nkp, midx, mdist

    def process(self, imgs, outs):
        """imgs: (n_frames, 2, H, W) u8 host (pinned for speed); outs from alloc_outputs()."""
        n = imgs.shape[0]
        kps, desc, nkp, midx, mdist = outs
        _check(lib().vieo_frontend_process(self._h, n, _p(imgs), imgs.strides[2], _p(kps), _p(desc), _p(nkp), _p(midx),
                                           _p(mdist)))
        return kps.view(np.uint8).reshape(2 * n, self.cap, 24).view(KP_DTYPE).reshape(2 * n, self.cap), desc, nkp, midx, mdist

    def stereo_rectified(self, n_frames, bf, min_z):
        """Frame::ComputeStereoMatches for the frames of the last process() call.
        -> (uright f32[n,cap], depth f32[n,cap], sad i32[n,cap]) indexed by left keypoint (-1: no match)."""
        ur = np.empty((n_frames, self.cap), np.float32)
        dp = np.empty((n_frames, self.cap), np.float32)
        sad = np.empty((n_frames, self.cap), np.int32)
        _check(lib().vieo_frontend_stereo_rectified(self._h, n_frames, bf, min_z, _p(ur), _p(dp), _p(sad)))
        return ur, dp, sad

    def last_launches(self):
        return lib().vieo_frontend_last_launches(self._h)


# ---------------------------------------------------------------- IMU pre-integration
class VieoImuNoise(C.Structure):
    _fields_ = [("sigma_g", C.c_double), ("sigma_a", C.c_double), ("sigma_bg", C.c_double), ("sigma_ba", C.c_double),
                ("freq_ref", C.c_double), ("dt_cov_noise_fixed", C.c_int32), ("pad_", C.c_int32)]


PREINT_DTYPE = np.dtype([("Rij", "f8", (3, 3)), ("vij", "f8", 3), ("pij", "f8", 3), ("SigmaPRV", "f8", (9, 9)),
                         ("SigmaPVR", "f8", (9, 9)), ("Jgp", "f8", (3, 3)), ("Jap", "f8", (3, 3)), ("Jgv", "f8", (3, 3)),
                         ("Jav", "f8", (3, 3)), ("JgR", "f8", (3, 3)), ("dt", "f8"), ("status", "i4"), ("pad_", "i4")])

EUROC_IMU_SIGMA = (1.6968e-4, 2.0e-3, 1.9393e-5, 3.0e-3)  # Examples/Stereo/EuRoC/EuRoC_VIO.yaml:13-18


class IMUPreintegrator:
    """IMUPreIntegratorBase (src/Odom/OdomPreIntegrator.h:108-223) for batches of intervals."""

    def __init__(self, sigma=EUROC_IMU_SIGMA, dt_cov_noise_fixed=1, freq_ref=200.0, device=0):
        """sigma = IMU.sigma {gyro, acc, gyro-bias, acc-bias}; squared here as src/Tracking.cc:744-745 does."""
        self.noise = VieoImuNoise()
        s2 = (C.c_double * 4)(*[x * x for x in sigma])
        lib().vieo_imu_set_param(C.byref(self.noise), s2, dt_cov_noise_fixed, freq_ref)
        self.device = device

    def PreIntegration(self, samples, ti, tj, bg, ba):
        """One interval: samples (n,7) rows {t, a, w}.  -> PREINT_DTYPE record (status mirrors the return value)."""
        return self.preintegrate_batch(samples, [0, len(samples)], [[ti, tj]], [np.r_[bg, ba]])[0]

    def preintegrate_batch(self, samples, seg_ptr, ti_tj, bg_ba):
        samples = np.ascontiguousarray(samples, np.float64).reshape(-1, 7)
        seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
        ti_tj = np.ascontiguousarray(ti_tj, np.float64).reshape(-1, 2)
        bg_ba = np.ascontiguousarray(bg_ba, np.float64).reshape(-1, 6)
        n = len(seg_ptr) - 1
        out = np.zeros(n, PREINT_DTYPE)
        _check(lib().vieo_imu_preint_batch(_p(samples), _p(seg_ptr), _p(ti_tj), _p(bg_ba), C.byref(self.noise), n,
                                           _p(out), self.device))
        return out

    def OptimizeInitialGyroBias(self, pre, Rwb, bg, bInfo=True, samples=None, seg_ptr=None, ti_tj=None, ba=None):
        """Optimizer::OptimizeInitialGyroBias (include/Optimizer.h:819-892) on keyframe pre-integrations `pre`
        (PREINT_DTYPE[n_kf], entry 0 ignored) and body rotations Rwb (n_kf, 3, 3); with the sample lists given, every
        interval is re-integrated with the new bias on the same stream (IMUInitialization.cpp:640-648).
        -> (num_equations, bg + estimate, re-integrated PREINT_DTYPE[n_kf] or None)"""
        pre = np.ascontiguousarray(pre, PREINT_DTYPE)
        Rwb = np.ascontiguousarray(Rwb, np.float64).reshape(-1, 9)
        n = len(pre)
        assert len(Rwb) == n
        bg = np.array(bg, np.float64).reshape(3).copy()
        neq = C.c_int32(0)
        out = None
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.float64).reshape(-1, 7)
            seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
            ti_tj = np.ascontiguousarray(ti_tj, np.float64).reshape(-1, 2)
            ba = None if ba is None else np.ascontiguousarray(ba, np.float64).reshape(-1, 3)
            assert len(seg_ptr) == n + 1 and len(ti_tj) == n
            out = np.zeros(n, PREINT_DTYPE)
        _check(lib().vieo_imu_init_gyro_bias(_p(pre), _p(Rwb), n, int(bInfo), _p(bg), C.byref(neq), _p(samples), _p(seg_ptr),
                                             _p(ti_tj), _p(ba), C.byref(self.noise), _p(out), self.device))
        return neq.value, bg, out

    def preintegrate_batch_dev(self, samples_ptr, seg_ptr, ti_tj_ptr, bg_ba_ptr, n, out_ptr, stream=0):
        """Device-resident form; all pointers are device addresses."""
        _check(lib().vieo_imu_preint_batch_dev(samples_ptr, seg_ptr, ti_tj_ptr, bg_ba_ptr, C.byref(self.noise), n,
                                               out_ptr, stream))


# ---------------------------------------------------------------- bundle adjustment
from .layouts import (BA_RESULT_DTYPE, CAMERA_DTYPE, NAVSTATE_DTYPE, POSEOPT_PROBLEM_DTYPE,  # noqa: E402,F401
                      POSEOPT_RESULT_DTYPE)


def frustum_rig_batch(pb, device=0):
    """Frame::isInFrustum with a camera rig (vieo_frustum_rig_batch) over a synth.make_frustum_rig_problem dict ->
    dict(inview, cam_mask, proj [n][4][3], level [n][4], viewcos [n][4], depth, n_inview)"""
    from .layouts import FRUSTUM_RIG_FRAME_DTYPE
    rig = np.ascontiguousarray(pb["rig"], FRUSTUM_RIG_FRAME_DTYPE)
    wP = np.ascontiguousarray(pb["p_wP"], np.float32); Pn = np.ascontiguousarray(pb["p_normal"], np.float32)
    mx = np.ascontiguousarray(pb["p_max_dist"], np.float32); mn = np.ascontiguousarray(pb["p_min_dist"], np.float32)
    skip = None if pb.get("p_skip") is None else np.ascontiguousarray(pb["p_skip"], np.uint8)
    n = len(mx)
    out = dict(inview=np.zeros(n, np.uint8), cam_mask=np.zeros(n, np.uint8), proj=np.zeros((n, 4, 3), np.float32),
               level=np.full((n, 4), -1, np.int32), viewcos=np.zeros((n, 4), np.float32), depth=np.zeros(n, np.float32),
               n_inview=np.zeros(len(rig), np.int32))
    L = lib()
    L.vieo_frustum_rig_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 12 + [C.c_int]
    _check(L.vieo_frustum_rig_batch(_p(rig), len(rig), _p(wP), _p(Pn), _p(mx), _p(mn), None if skip is None else _p(skip),
                                    _p(out["inview"]), _p(out["cam_mask"]), _p(out["proj"]), _p(out["level"]), _p(out["viewcos"]),
                                    _p(out["depth"]), _p(out["n_inview"]), device))
    return out


def set_host_sync(mode, device=0):
    """vieo_set_host_sync: 0 auto, 1 spin, 2 yield, 3 blocking (how host threads wait for the device)"""
    _check(lib().vieo_set_host_sync(int(device), int(mode)))


class Optimizer:
    """Static surface of the reference's Optimizer (include/Optimizer.h:46-121) over flattened problems."""

    @staticmethod
    def PoseOptimizationBatch(pbs, cam, Xw, obs, inv_sigma2, flags, device=0):
        """A batch of Optimizer::PoseOptimization calls (visual or IMU/PVR, per problem `mode`).
        -> (results POSEOPT_RESULT_DTYPE[n], outlier u8[E], chi2 f64[E])"""
        pbs = np.ascontiguousarray(pbs, POSEOPT_PROBLEM_DTYPE)
        cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
        Xw = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 3)
        inv_sigma2 = np.ascontiguousarray(inv_sigma2, np.float32)
        flags = np.ascontiguousarray(flags, np.uint8)
        E = len(flags)
        assert len(Xw) == E and len(obs) == E and len(inv_sigma2) == E
        res = np.zeros(len(pbs), POSEOPT_RESULT_DTYPE)
        outlier = np.zeros(E, np.uint8)
        chi2 = np.zeros(E, np.float64)
        _check(lib().vieo_pose_opt_batch(_p(pbs), len(pbs), _p(cam), _p(Xw), _p(obs), _p(inv_sigma2), _p(flags), E,
                                         _p(res), _p(outlier), _p(chi2), device))
        return res, outlier, chi2

    @staticmethod
    def OptimizeSim3Batch(pbs, cam, Xc1, Xc2, obs1, obs2, inv_sigma2_1, inv_sigma2_2, device=0):
        """A batch of Optimizer::OptimizeSim3 calls (src/Optimizer.cc:2689-2920), one candidate keyframe pair per problem.
        -> (results SIM3_RESULT_DTYPE[n], keep u8[M] (0: vpMatches1[i] = nullptr), chi2_12 f64[M], chi2_21 f64[M])"""
        from .layouts import SIM3_PROBLEM_DTYPE, SIM3_RESULT_DTYPE
        pbs = np.ascontiguousarray(pbs, SIM3_PROBLEM_DTYPE)
        cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
        Xc1 = np.ascontiguousarray(Xc1, np.float64).reshape(-1, 3)
        Xc2 = np.ascontiguousarray(Xc2, np.float64).reshape(-1, 3)
        obs1 = np.ascontiguousarray(obs1, np.float32).reshape(-1, 2)
        obs2 = np.ascontiguousarray(obs2, np.float32).reshape(-1, 2)
        w1 = np.ascontiguousarray(inv_sigma2_1, np.float32)
        w2 = np.ascontiguousarray(inv_sigma2_2, np.float32)
        M = len(Xc1)
        assert len(Xc2) == M and len(obs1) == M and len(obs2) == M and len(w1) == M and len(w2) == M
        res = np.zeros(len(pbs), SIM3_RESULT_DTYPE)
        keep = np.zeros(M, np.uint8)
        c12 = np.zeros(M, np.float64)
        c21 = np.zeros(M, np.float64)
        _check(lib().vieo_optimize_sim3_batch(_p(pbs), len(pbs), _p(cam), _p(Xc1), _p(Xc2), _p(obs1), _p(obs2), _p(w1), _p(w2), M,
                                              _p(res), _p(keep), _p(c12), _p(c21), device))
        return res, keep, c12, c21

    @staticmethod
    def OptimizeEssentialGraph(pb, iterations=20, lambda_init=1e-16, device=0, want_Tcw=True):
        """The g2o part of Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:2309-2688) on a flattened graph (dict of
        synth.make_essential_graph: Scw, fixed, fix_scale, ei, ej, meas, info) -> (Scw_out SIM3_DTYPE[K], Tcw [K][3][4] or None,
        stats POSEGRAPH_STATS_DTYPE record).  One cooperative kernel launch runs the whole optimize(20)."""
        from .layouts import SIM3_DTYPE, POSEGRAPH_STATS_DTYPE
        S = np.ascontiguousarray(pb["Scw"], SIM3_DTYPE); fixed = np.ascontiguousarray(pb["fixed"], np.uint8)
        ei = np.ascontiguousarray(pb["ei"], np.int32); ej = np.ascontiguousarray(pb["ej"], np.int32)
        meas = np.ascontiguousarray(pb["meas"], SIM3_DTYPE)
        info = None if pb.get("info") is None else np.ascontiguousarray(pb["info"], np.float64).reshape(len(ei), 49)
        assert len(fixed) == len(S) and len(ej) == len(ei) and len(meas) == len(ei)
        out = np.zeros(len(S), SIM3_DTYPE); T = np.zeros((len(S), 3, 4)) if want_Tcw else None
        st = np.zeros(1, POSEGRAPH_STATS_DTYPE)
        L = lib()
        L.vieo_essential_graph_optimize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _check(L.vieo_essential_graph_optimize(len(S), _p(S), _p(fixed), int(pb["fix_scale"]), len(ei), _p(ei), _p(ej), _p(meas),
                                               None if info is None else _p(info), int(iterations), float(lambda_init), _p(out),
                                               None if T is None else _p(T), _p(st), device))
        return out, T, st[0]

    @staticmethod
    def essential_graph_debug_step(pb, lam, device=0):
        """vieo_essential_graph_debug_step -> (Scw_out, stats, H [n][n], b [n]) with n = 7 * stats["n_free"]"""
        from .layouts import SIM3_DTYPE, POSEGRAPH_STATS_DTYPE
        S = np.ascontiguousarray(pb["Scw"], SIM3_DTYPE); fixed = np.ascontiguousarray(pb["fixed"], np.uint8)
        ei = np.ascontiguousarray(pb["ei"], np.int32); ej = np.ascontiguousarray(pb["ej"], np.int32)
        meas = np.ascontiguousarray(pb["meas"], SIM3_DTYPE)
        info = None if pb.get("info") is None else np.ascontiguousarray(pb["info"], np.float64).reshape(len(ei), 49)
        nmax = 7 * len(S)
        out = np.zeros(len(S), SIM3_DTYPE); st = np.zeros(1, POSEGRAPH_STATS_DTYPE); H = np.zeros(nmax * nmax); b = np.zeros(nmax)
        L = lib()
        L.vieo_essential_graph_debug_step.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _check(L.vieo_essential_graph_debug_step(len(S), _p(S), _p(fixed), int(pb["fix_scale"]), len(ei), _p(ei), _p(ej), _p(meas),
                                                 None if info is None else _p(info), float(lam), _p(out), _p(st), _p(H), _p(b), device))
        n = 7 * int(st[0]["n_free"])
        return out, st[0], H[:n * n].reshape(n, n).copy(), b[:n].copy()

    @staticmethod
    def essential_graph_correct_points(Pw, ref, Scw_before, Scw_after, device=0):
        """The "Correct points" loop of Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:2645-2676) -> float32 [n][3]"""
        from .layouts import SIM3_DTYPE
        Pw = np.ascontiguousarray(Pw, np.float32).reshape(-1, 3); ref = np.ascontiguousarray(ref, np.int32)
        a = np.ascontiguousarray(Scw_before, SIM3_DTYPE); b = np.ascontiguousarray(Scw_after, SIM3_DTYPE)
        assert len(a) == len(b) and len(ref) == len(Pw)
        out = np.zeros_like(Pw)
        L = lib()
        L.vieo_essential_graph_correct_points.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _check(L.vieo_essential_graph_correct_points(len(Pw), _p(Pw), _p(ref), len(a), _p(a), _p(b), _p(out), device))
        return out

    @staticmethod
    def pose_opt_batch_dev(pbs_ptr, n, cam_ptr, Xw_ptr, obs_ptr, w_ptr, flags_ptr, res_ptr, outlier_ptr, chi2_ptr, stream=0):
        _check(lib().vieo_pose_opt_batch_dev(pbs_ptr, n, cam_ptr, Xw_ptr, obs_ptr, w_ptr, flags_ptr, res_ptr, outlier_ptr,
                                             chi2_ptr, stream))


class VieoBaProblem(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_imu", C.c_int32),
                ("states", C.c_void_p), ("state_flags", C.c_void_p), ("points", C.c_void_p), ("edge_state", C.c_void_p),
                ("edge_point", C.c_void_p), ("obs", C.c_void_p), ("inv_sigma2", C.c_void_p), ("edge_flags", C.c_void_p),
                ("imu_i", C.c_void_p), ("imu_j", C.c_void_p), ("preint", C.c_void_p), ("imu_dt_kf", C.c_void_p),
                ("gw", C.c_double * 3), ("inv_sigma_bg2", C.c_double), ("inv_sigma_ba2", C.c_double),
                ("large", C.c_int32), ("rec_init", C.c_int32), ("visual_only", C.c_int32), ("global_ba", C.c_int32),
                ("scale_init", C.c_double)]


class VieoGbaExtra(C.Structure):
    _fields_ = [("scale_opt", C.c_int32), ("imu_init", C.c_int32), ("scale", C.c_double), ("gw", C.c_double * 3)]


_BA_ARRAYS = (("states", NAVSTATE_DTYPE), ("state_flags", np.uint8), ("points", np.float64), ("edge_state", np.int32),
              ("edge_point", np.int32), ("obs", np.float32), ("inv_sigma2", np.float32), ("edge_flags", np.uint8),
              ("imu_i", np.int32), ("imu_j", np.int32), ("preint", PREINT_DTYPE), ("imu_dt_kf", np.float64))


def ba_problem(d, large=False, rec_init=False, visual_only=False):
    """dict of arrays (synth.make_lba_problem layout) -> (VieoBaProblem, keepalive dict)"""
    keep = {k: np.ascontiguousarray(d[k], dt) for k, dt in _BA_ARRAYS}
    pb = VieoBaProblem()
    pb.n_states, pb.n_points, pb.n_edges, pb.n_imu = len(keep["states"]), len(keep["points"]), len(keep["edge_state"]), len(keep["imu_i"])
    for k, a in keep.items():
        setattr(pb, k, a.ctypes.data)
    pb.gw = (C.c_double * 3)(*d["gw"])
    pb.inv_sigma_bg2, pb.inv_sigma_ba2 = d["inv_sigma_bg2"], d["inv_sigma_ba2"]
    pb.large, pb.rec_init, pb.visual_only = int(large), int(rec_init), int(visual_only)
    return pb, keep


def search_for_triangulation(pb, device=0):
    """ORBmatcher::SearchForTriangulation for a batch of keyframe pairs (synth.make_sft_problem layout, host arrays).
    -> (match12 [n_out_total], pairs_out [n_out_total, 2], n_matches [n_pairs])."""
    from .layouts import SFT_PAIR_DTYPE
    pairs = np.ascontiguousarray(pb["pairs"], SFT_PAIR_DTYPE)
    arrs = [np.ascontiguousarray(pb["kps"], KP_DTYPE), np.ascontiguousarray(pb["uright"], np.float32),
            np.ascontiguousarray(pb["desc"], np.uint8), np.ascontiguousarray(pb["has_mp"], np.uint8),
            np.ascontiguousarray(pb["fv_node"], np.int32), np.ascontiguousarray(pb["fv_ptr"], np.int32),
            np.ascontiguousarray(pb["fv_idx"], np.int32)]
    n_out = int(pb["n_out_total"])
    m12 = np.empty(max(n_out, 1), np.int32); po = np.empty((max(n_out, 1), 2), np.int32); nm = np.empty(max(len(pairs), 1), np.int32)
    _check(lib().vieo_search_for_triangulation(_p(pairs), len(pairs), *[_p(a) for a in arrs], len(arrs[0]), len(arrs[4]),
                                               len(arrs[5]), len(arrs[6]), n_out, int(pb["n_nodes1_total"]), _p(m12), _p(po),
                                               _p(nm), device))
    return m12[:n_out], po[:n_out], nm[:len(pairs)]


def search_by_bow(pb, device=0):
    """ORBmatcher::SearchByBoW(KeyFrame, Frame) for a batch of pairs (synth.make_bow_problem layout).
    -> (match_f [n_out_total], n_matches [n_pairs])."""
    from .layouts import BOW_PAIR_DTYPE
    pairs = np.ascontiguousarray(pb["pairs"], BOW_PAIR_DTYPE)
    arrs = [np.ascontiguousarray(pb["kps"], KP_DTYPE), np.ascontiguousarray(pb["desc"], np.uint8),
            np.ascontiguousarray(pb["mp_ok"], np.uint8), np.ascontiguousarray(pb["fv_node"], np.int32),
            np.ascontiguousarray(pb["fv_ptr"], np.int32), np.ascontiguousarray(pb["fv_idx"], np.int32)]
    n_out = int(pb["n_out_total"])
    mf = np.empty(max(n_out, 1), np.int32); nm = np.empty(max(len(pairs), 1), np.int32)
    _check(lib().vieo_search_by_bow(_p(pairs), len(pairs), *[_p(a) for a in arrs], len(arrs[0]), len(arrs[3]), len(arrs[4]),
                                    len(arrs[5]), n_out, _p(mf), _p(nm), device))
    return mf[:n_out], nm[:len(pairs)]


def local_ba_prv_batch(bas, problems, cam, **kw):
    """n LocalBA windows on n handles from ONE host thread: every window is enqueued before any is awaited."""
    for ba, d in zip(bas, problems):
        ba.begin(d, cam, **kw)
    return [ba.end() for ba in bas[:len(problems)]]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
COMM_ID_BYTES = 128


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it; the launcher hands it to the other ranks)."""
    buf = np.zeros(COMM_ID_BYTES, np.uint8)
    _check(lib().vieo_comm_unique_id(_p(buf)))
    return buf


class Comm:
    """Library-owned NCCL communicator (vieo_comm_*): the sharded bundle adjustment's all-reduce issued from C."""

    def __init__(self, unique_id, rank, world, device=0):
        self._h = C.c_void_p()
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == COMM_ID_BYTES
        _check(lib().vieo_comm_create(_p(uid), rank, world, device, C.byref(self._h)))
        self.rank, self.world = rank, world

    def allreduce_f64(self, ptr, count, stream=0):
        _check(lib().vieo_comm_allreduce_f64(self._h, ptr, count, stream))

    def close(self):
        if getattr(self, "_h", None) and self._h and _lib is not None:
            _lib.vieo_comm_destroy(self._h)
            self._h = None

    __del__ = close


class BundleAdjuster:
    """Device engine behind Optimizer::LocalBundleAdjustmentNavStatePRV / LocalBundleAdjustment (vieo_ba_* ABI)."""

    def __init__(self, max_states=256, max_points=8192, max_edges=65536, max_imu=64, device=0, global_ba=False):
        """global_ba: a handle for map-sized problems (GlobalBundleAdjustmentNavStatePRV; dense multi-CTA solver)."""
        self._h = C.c_void_p()
        create = lib().vieo_ba_create_global if global_ba else lib().vieo_ba_create
        _check(create(max_states, max_points, max_edges, max_imu, device, C.byref(self._h)))
        self._cb = None

    def close(self):
        if getattr(self, "_h", None) and self._h and _lib is not None:
            _lib.vieo_ba_destroy(self._h)
            self._h = None

    __del__ = close

    def stream(self):
        return lib().vieo_ba_stream(self._h)

    def set_sharding(self, rank, world, allreduce):
        """allreduce(dev_ptr:int, count:int, stream:int) sums `count` fp64 in place over the ranks (e.g. NCCL)."""
        def cb(ctx, buf, count, stream):
            try:
                allreduce(buf, count, stream)
                return 0
            except Exception:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1
        self._cb = ALLREDUCE_FN(cb) if allreduce else None
        _check(lib().vieo_ba_set_sharding(self._h, rank, world, C.cast(self._cb, C.c_void_p) if self._cb else None, None))

    def set_comm(self, comm):
        """Sharding with the library's own NCCL communicator (no host callback); None: single GPU."""
        self._comm = comm
        _check(lib().vieo_ba_set_comm(self._h, comm._h if comm is not None else None))

    def LocalBundleAdjustmentNavStatePRV(self, d, cam, large=False, rec_init=False, visual_only=False, stop=None):
        pb, keep = ba_problem(d, large, rec_init, visual_only)
        cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
        st = np.zeros(pb.n_states, NAVSTATE_DTYPE); pts = np.zeros((pb.n_points, 3)); chi2 = np.zeros(pb.n_edges)
        erase = np.zeros(pb.n_edges, np.uint8); res = np.zeros(1, BA_RESULT_DTYPE)
        _check(lib().vieo_local_ba_prv(self._h, C.byref(pb), _p(cam), _p(stop), _p(st), _p(pts), _p(chi2), _p(erase), _p(res)))
        return dict(states=st, points=pts, edge_chi2=chi2, erase=erase, res=res[0])

    def begin(self, d, cam, large=False, rec_init=False, visual_only=False, stop=None):
        """vieo_local_ba_prv_begin: enqueue the whole LocalBA routine on the handle's stream and return."""
        pb, keep = ba_problem(d, large, rec_init, visual_only)
        cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
        self._inflight = (pb.n_states, pb.n_points, pb.n_edges)
        self._stop_keep = stop
        _check(lib().vieo_local_ba_prv_begin(self._h, C.byref(pb), _p(cam), _p(stop)))

    def poll(self):
        return bool(lib().vieo_local_ba_prv_poll(self._h))

    def end(self):
        K, P, E = self._inflight
        st = np.zeros(K, NAVSTATE_DTYPE); pts = np.zeros((P, 3)); chi2 = np.zeros(E)
        erase = np.zeros(E, np.uint8); res = np.zeros(1, BA_RESULT_DTYPE)
        _check(lib().vieo_local_ba_prv_end(self._h, _p(st), _p(pts), _p(chi2), _p(erase), _p(res)))
        return dict(states=st, points=pts, edge_chi2=chi2, erase=erase, res=res[0])

    def GlobalBundleAdjustmentNavStatePRV(self, d, cam, nIterations=5, bRobust=True, stop=None, bScaleOpt=False,
                                          imu_init_gw=None):
        """Optimizer::GlobalBundleAdjustmentNavStatePRV (src/Optimizer.cc:771-1342) on the flattened map; the handle must
        have been created with global_ba=True.  bScaleOpt: the scale vertex of System::FinalGBA (result `scale`, points
        returned scaled).  imu_init_gw: the IMU initialiser's call (pimu_initiator) with its gravity estimate (result
        `gw` = refined gravity)."""
        pb, keep = ba_problem(d)
        cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
        st = np.zeros(pb.n_states, NAVSTATE_DTYPE); pts = np.zeros((pb.n_points, 3)); chi2 = np.zeros(pb.n_edges)
        res = np.zeros(1, BA_RESULT_DTYPE)
        if not bScaleOpt and imu_init_gw is None:
            it = _check(lib().vieo_global_ba_prv(self._h, C.byref(pb), _p(cam), int(nIterations), int(bRobust), _p(stop),
                                                 _p(st), _p(pts), _p(chi2), _p(res)))
            return dict(states=st, points=pts, edge_chi2=chi2, res=res[0], iterations=it)
        ex = VieoGbaExtra()
        ex.scale_opt, ex.imu_init = int(bool(bScaleOpt)), int(imu_init_gw is not None)
        if imu_init_gw is not None:
            ex.gw = (C.c_double * 3)(*[float(v) for v in imu_init_gw])
        it = _check(lib().vieo_global_ba_prv_ex(self._h, C.byref(pb), _p(cam), int(nIterations), int(bRobust), C.byref(ex),
                                                _p(stop), _p(st), _p(pts), _p(chi2), _p(res)))
        return dict(states=st, points=pts, edge_chi2=chi2, res=res[0], iterations=it, scale=float(ex.scale),
                    gw=np.array(list(ex.gw)))

    def set_problem(self, d, cam, global_ba=0, scale_init=0.0, **kw):
        """global_ba: VieoBaProblem.global_ba bits (1 global graph, 2 robust, 4 scale vertex, 8 gravity-direction vertex)."""
        pb, keep = ba_problem(d, **kw)
        pb.global_ba, pb.scale_init = int(global_ba), float(scale_init)
        cam = np.ascontiguousarray(cam, CAMERA_DTYPE).reshape(1)
        _check(lib().vieo_ba_set_problem(self._h, C.byref(pb), _p(cam)))
        self._n = (pb.n_states, pb.n_points, pb.n_edges)

    def get_border(self):
        sc = C.c_double(0)
        gw = np.zeros(3)
        _check(lib().vieo_ba_get_border(self._h, C.byref(sc), _p(gw)))
        return sc.value, gw

    def debug_step(self, lam):
        K, P, E = self._n
        n_ = 15 * K + 3
        xp = np.zeros(n_); xl = np.zeros((P, 3)); H = np.zeros(n_ ** 2); b = np.zeros(n_)
        n = _check(lib().vieo_ba_debug_step(self._h, lam, _p(xp), _p(xl), _p(H), _p(b)))
        return xp[:n], xl, H[:n * n].reshape(n, n), b[:n]

    def optimize(self, iterations, lambda_init=0.0, stop=None):
        return _check(lib().vieo_ba_optimize(self._h, iterations, lambda_init, _p(stop)))

    def active_robust_chi2(self, recompute=True):
        c = C.c_double(0)
        _check(lib().vieo_ba_active_robust_chi2(self._h, int(recompute), C.byref(c)))
        return c.value

    def get(self):
        K, P, E = self._n
        st = np.zeros(K, NAVSTATE_DTYPE); pts = np.zeros((P, 3)); chi2 = np.zeros(E)
        _check(lib().vieo_ba_get(self._h, _p(st), _p(pts), _p(chi2)))
        return st, pts, chi2

    def last_launches(self):
        return lib().vieo_ba_last_launches(self._h)

    def last_ms(self):
        return lib().vieo_ba_last_ms(self._h)
