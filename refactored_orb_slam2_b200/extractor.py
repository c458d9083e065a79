"""ORBextractor — host-side mirror of ORB_SLAM2::ORBextractor over the C ABI.

Same constructor arguments, call signature, getters and public ``mvImagePyramid`` as the reference
class (Source/Libraries/ORB_SLAM2/include/ORBextractor.h:47-75); the work happens in
liborbgpu.so's sm_100a kernels.  Keypoints come back as a numpy record array with cv::KeyPoint's
layout (pt.x, pt.y, size, angle, response, octave, class_id), descriptors as [N, 32] uint8.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, ExtractorParams, check, lib, ptr


class ORBextractor:
    HARRIS_SCORE = 0   # ORBextractor.h:49
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, *, max_width=0, max_height=0,
                 max_batch=1, device=-1):
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(scaleFactor), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        self._max = (int(max_width), int(max_height), int(max_batch))
        self._device = int(device)
        self._h = None
        self._last = None          # (batch, width, height) of the last call
        if max_width > 0 and max_height > 0:
            self._create(max_width, max_height, max_batch)

    # -- handle management ----------------------------------------------------------------
    def _create(self, w, h, batch):
        self.close()
        p = ExtractorParams(self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST, self.minThFAST,
                            int(w), int(h), int(batch), self._device)
        out = C.c_void_p()
        check(lib().orbgpu_extractor_create(C.byref(p), C.byref(out)))
        self._h = out
        self._max = (int(w), int(h), int(batch))
        self._load_tables()

    def _ensure(self, w, h, batch):
        mw, mh, mb = self._max
        if self._h is None or w > mw or h > mh or batch > mb:
            self._create(max(w, mw), max(h, mh), max(batch, mb))

    def _load_tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        check(lib().orbgpu_extractor_tables(self._h, ptr(sc), ptr(isc), ptr(s2), ptr(is2), ptr(q)))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = sc, isc, s2, is2
        self.mnFeaturesPerLevel = q

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().orbgpu_extractor_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def capacity(self):
        """Keypoint rows per image in the output buffers."""
        if self._h is None:
            self._tables_only()
        return check(lib().orbgpu_extractor_capacity(self._h))

    def _tables_only(self):
        # getters must work before the first image, as in the reference: size a minimal handle
        self._create(max(self._max[0], 128), max(self._max[1], 128), max(self._max[2], 1))

    # -- ORBextractor.h:59-73 --------------------------------------------------------------
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return np.float32(self.scaleFactor)

    def _table(self, name):
        if self._h is None:
            self._tables_only()
        return getattr(self, name).copy()

    def GetScaleFactors(self):
        return self._table("mvScaleFactor")

    def GetInverseScaleFactors(self):
        return self._table("mvInvScaleFactor")

    def GetScaleSigmaSquares(self):
        return self._table("mvLevelSigma2")

    def GetInverseScaleSigmaSquares(self):
        return self._table("mvInvLevelSigma2")

    # -- operator() ------------------------------------------------------------------------
    def __call__(self, image, mask=None):
        """ORBextractor::operator()(image, mask, keypoints, descriptors); the mask is ignored, as in
        the reference (ORBextractor.cc:978-1039).  Returns (keypoints, descriptors)."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)     # ORBextractor.cc:981-982
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("image must be a 2-D uint8 array (CV_8UC1)")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        self._ensure(w, h, 1)
        cap = self.capacity
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        check(lib().orbgpu_extract(self._h, ptr(image), w, h, image.strides[0], ptr(kps), ptr(desc), cap, C.byref(n)))
        self._last = (1, w, h)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """operator() over [B, H, W] uint8 host images in one call.  Returns (keypoints [B, cap],
        descriptors [B, cap, 32], counts [B]); rows beyond counts[i] are unspecified."""
        images = np.ascontiguousarray(images, np.uint8)
        b, h, w = images.shape
        self._ensure(w, h, b)
        cap = self.capacity
        kps = np.empty((b, cap), KP_DTYPE)
        desc = np.empty((b, cap, 32), np.uint8)
        counts = np.zeros(b, np.int32)
        check(lib().orbgpu_extract_batch(self._h, ptr(images), b, w, h, w, h * w, ptr(kps), ptr(desc), cap, ptr(counts)))
        self._last = (b, w, h)
        return kps, desc, counts

    def extract_batch_device(self, d_images, count, width, height, row_step=None, image_step=None):
        """Asynchronous extraction of images already resident in HBM (`d_images` is a device address
        or an object with .data_ptr()).  Results stay on the device: see results_device()."""
        addr = d_images.data_ptr() if hasattr(d_images, "data_ptr") else int(d_images)
        self._ensure(width, height, count)
        row_step = width if row_step is None else row_step
        image_step = row_step * height if image_step is None else image_step
        check(lib().orbgpu_extract_batch_device(self._h, C.c_void_p(addr), count, width, height, row_step, image_step))
        self._last = (count, width, height)

    def results_device(self):
        k, d, c, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        check(lib().orbgpu_results_device(self._h, C.byref(k), C.byref(d), C.byref(c), C.byref(cap)))
        return k.value, d.value, c.value, cap.value

    def download_results(self):
        b = self._last[0]
        cap = self.capacity
        kps = np.empty((b, cap), KP_DTYPE)
        desc = np.empty((b, cap, 32), np.uint8)
        counts = np.zeros(b, np.int32)
        check(lib().orbgpu_download_results(self._h, ptr(kps), ptr(desc), cap, ptr(counts)))
        return kps, desc, counts

    def synchronize(self):
        check(lib().orbgpu_synchronize(self._h))

    def set_profiling(self, enable=True):
        check(lib().orbgpu_set_profiling(self._h, int(enable)))

    def stage_times(self):
        """{stage name: milliseconds} of the last batch (needs set_profiling(True)); synchronises."""
        ms = np.zeros(6, np.float32)
        check(lib().orbgpu_stage_times(self._h, ptr(ms), 6))
        return {lib().orbgpu_stage_name(i).decode(): float(ms[i]) for i in range(6)}

    @property
    def stream(self):
        return lib().orbgpu_extractor_stream(self._h)

    # -- mvImagePyramid (ORBextractor.h:75) --------------------------------------------------
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(lib().orbgpu_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level, image=0, with_border=False):
        w, h = self.level_size(level)
        b = 19 if with_border else 0
        out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
        check(lib().orbgpu_get_pyramid_level(self._h, image, level, int(with_border), ptr(out), out.strides[0]))
        return out

    def blurred_level(self, level, image=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        check(lib().orbgpu_get_blurred_level(self._h, image, level, ptr(out), out.strides[0]))
        return out

    @property
    def mvImagePyramid(self):
        """Pyramid of image 0 of the last call, fetched from HBM on demand."""
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    def stage(self, level, stage, image=0):
        """Parity-test accessor: stage 0 = FAST candidates (reference order), 1 = octree selection
        (list order); rows (x, y, response) in octree coordinates."""
        n = check(lib().orbgpu_get_stage(self._h, image, level, stage, None, 0))
        out = np.zeros((max(n, 1), 3), np.int32)
        check(lib().orbgpu_get_stage(self._h, image, level, stage, ptr(out), n))
        return out[:n]
