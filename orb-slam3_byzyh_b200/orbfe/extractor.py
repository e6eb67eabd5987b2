"""ORBextractor: mirror of ORB_SLAM3::ORBextractor (/root/reference/include/ORBextractor.h:46-112,
src/ORBextractor.cc:468-571, 1557-1682) over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, lib, ptr


class _Pyramid:
    """mvImagePyramid (ORBextractor.h:83): level images of frame `frame` of the last call, fetched
    from the device on demand."""

    def __init__(self, ex):
        self._ex = ex

    def __len__(self):
        return self._ex.nlevels

    def __getitem__(self, level):
        return self._ex.pyramid_level(level)


class ORBextractor:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0):
        self.nfeatures, self.nlevels, self.device = nfeatures, nlevels, device
        self._h = C.c_void_p()
        check(lib().orbfe_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device,
                                           C.byref(self._h)))
        self.capacity = lib().orbfe_max_keypoints(self._h)
        self.mvImagePyramid = _Pyramid(self)
        self._shape = None
        self._rect = None       # (rows, cols) of the rectified image while rectification maps are set

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().orbfe_extractor_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    # ---- accessors, ORBextractor.h:61-81 ----
    def GetLevels(self):
        return lib().orbfe_get_levels(self._h)

    def GetScaleFactor(self):
        return lib().orbfe_get_scale_factor(self._h)

    def _tables(self):
        t = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        check(lib().orbfe_scale_tables(self._h, *[ptr(a) for a in t]))
        return t

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        n = np.empty(self.nlevels, np.int32)
        check(lib().orbfe_features_per_level(self._h, ptr(n)))
        return n

    # ---- operator(), ORBextractor.cc:1557-1682 ----
    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints[KP_DTYPE], descriptors[n,32]); monoIndex == -1 and empty
        outputs for an empty image, as the reference."""
        image = np.asarray(image)
        if image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected (ORBextractor.cc:1567)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.capacity_for(*image.shape)
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = check(lib().orbfe_extract(self._h, ptr(image), image.shape[0], image.shape[1], image.strides[0],
                                         int(vLappingArea[0]), int(vLappingArea[1]), ptr(kps), ptr(desc), cap,
                                         C.byref(n)))
        self._shape = self._rect or image.shape
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def capacity_for(self, rows, cols):
        """Keypoint capacity a rows x cols frame needs: `capacity`, or more for frames wider than 8.5 : 1."""
        return max(self.capacity, lib().orbfe_max_keypoints_for(self._h, int(rows), int(cols)))

    def extract_batch(self, images, vLappingArea=(0, 0), out=None):
        """images: [B, rows, cols] uint8 host array (numpy, or a pinned torch CPU tensor).  Returns
        (n[B], mono[B], keypoints[B, cap], descriptors[B, cap, 32]) host arrays."""
        B, rows, cols = images.shape
        cap = self.capacity_for(rows, cols)
        if out is None:
            out = (np.empty(B, np.int32), np.empty(B, np.int32), np.empty((B, cap), KP_DTYPE),
                   np.empty((B, cap, 32), np.uint8))
        n, mono, kps, desc = out
        if isinstance(images, np.ndarray):
            step, fstride = images.strides[1], images.strides[0]
        else:
            step, fstride = images.stride(1), images.stride(0)
        check(lib().orbfe_extract_batch(self._h, ptr(images), B, rows, cols, step, fstride, int(vLappingArea[0]),
                                        int(vLappingArea[1]), ptr(kps), ptr(desc), cap, ptr(n), ptr(mono)))
        self._shape = self._rect or (rows, cols)
        return n, mono, kps, desc

    def extract_batch_submit(self, images, vLappingArea, out):
        """Streaming form: enqueue the batch and return; `out` = (n, mono, kps, desc) PINNED host arrays that, like
        `images`, must stay alive and untouched until the matching extract_batch_wait()."""
        B, rows, cols = images.shape
        n, mono, kps, desc = out
        if isinstance(images, np.ndarray):
            step, fstride = images.strides[1], images.strides[0]
        else:
            step, fstride = images.stride(1), images.stride(0)
        check(lib().orbfe_extract_batch_submit(self._h, ptr(images), B, rows, cols, step, fstride, int(vLappingArea[0]),
                                               int(vLappingArea[1]), ptr(kps), ptr(desc), self.capacity, ptr(n), ptr(mono)))
        self._shape = self._rect or (rows, cols)

    def extract_batch_wait(self):
        """Block until the oldest submitted batch has its results in its host arrays."""
        check(lib().orbfe_extract_batch_wait(self._h))

    def extract_batch_device(self, d_images, vLappingArea, d_kps, d_desc, d_n, d_mono, stream=None):
        """Device-resident form: torch CUDA tensors in, torch CUDA tensors out, enqueued on `stream`
        (a torch.cuda.Stream or None for the extractor's own stream); no synchronisation."""
        B, rows, cols = d_images.shape
        cap = d_kps.shape[1]
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        check(lib().orbfe_extract_batch_device(self._h, ptr(d_images), B, rows, cols, d_images.stride(1),
                                               d_images.stride(0), int(vLappingArea[0]), int(vLappingArea[1]),
                                               ptr(d_kps), ptr(d_desc), cap, ptr(d_n), ptr(d_mono), st))
        self._shape = self._rect or (rows, cols)

    # ---- mvImagePyramid and stage taps ----
    def level_size(self, level, shape=None):
        rows, cols = shape or self._shape
        w, h = C.c_int(), C.c_int()
        check(lib().orbfe_level_size(self._h, rows, cols, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level, frame=0, with_border=False):
        w, h = self.level_size(level)
        pad = 38 if with_border else 0
        out = np.empty((h + pad, w + pad), np.uint8)
        check(lib().orbfe_pyramid_level(self._h, frame, level, int(with_border), ptr(out), out.strides[0]))
        return out

    def _img_tap(self, fn, level, frame):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        check(fn(self._h, frame, level, ptr(out), out.strides[0]))
        return out

    def debug_blurred(self, level, frame=0):
        return self._img_tap(lib().orbfe_debug_blurred, level, frame)


    def _list_tap(self, fn, level, frame):
        w, h = self.level_size(level)
        cap = w * h // 4 + 16
        out = np.empty((cap, 3), np.int32)
        n = C.c_int(0)
        check(fn(self._h, frame, level, ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_candidates(self, level, frame=0):
        return self._list_tap(lib().orbfe_debug_candidates, level, frame)

    def debug_level_keypoints(self, level, frame=0):
        return self._list_tap(lib().orbfe_debug_level_keypoints, level, frame)

    def debug_octree(self, xys, minX, maxX, minY, maxY, N):
        xys = np.ascontiguousarray(xys, np.int32)
        keep = np.empty(N + 64, np.int32)
        n = C.c_int(0)
        check(lib().orbfe_debug_octree(self._h, ptr(xys), len(xys), minX, maxX, minY, maxY, N, ptr(keep), len(keep),
                                       C.byref(n)))
        return keep[:n.value].copy()

    def set_profiling(self, on=True):
        check(lib().orbfe_set_profiling(self._h, int(on)))

    def stage_ms(self):
        ms = np.zeros(_lib.NUM_STAGES, np.float32)
        check(lib().orbfe_stage_ms(self._h, ptr(ms)))
        return dict(zip(_lib.STAGE_NAMES, ms.tolist()))

    def launch_count(self):
        return lib().orbfe_launch_count(self._h)

    def set_rectification(self, map_x=None, map_y=None):
        """Fuse cv::remap(frame, M1, M2, INTER_LINEAR) (System.cc:286-293) into pyramid level 0: later calls take the RAW
        frames.  map_x / map_y: CV_32FC1 maps of the rectified size; None switches rectification off."""
        if map_x is None:
            check(lib().orbfe_extractor_set_rectification(self._h, None, None, 0, 0))
            self._rect = None
            return
        mx = np.ascontiguousarray(map_x, np.float32)
        my = np.ascontiguousarray(map_y, np.float32)
        assert mx.shape == my.shape and mx.ndim == 2
        check(lib().orbfe_extractor_set_rectification(self._h, ptr(mx), ptr(my), mx.shape[0], mx.shape[1]))
        self._rect = tuple(mx.shape)

    def set_max_bytes(self, nbytes):
        check(lib().orbfe_set_max_bytes(self._h, int(nbytes)))

    def frame_geometry(self):
        cells, slots, kpcap = C.c_int(), C.c_int(), C.c_int()
        stride, per = C.c_ulonglong(), C.c_ulonglong()
        check(lib().orbfe_frame_geometry(self._h, C.byref(cells), C.byref(slots), C.byref(kpcap), C.byref(stride),
                                         C.byref(per)))
        return dict(cells=cells.value, slots=slots.value, kpcap=kpcap.value, pyr_stride=stride.value,
                    per_frame_bytes=per.value)
