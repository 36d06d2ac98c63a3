"""ORACLE (test infrastructure): ctypes front-ends for the two CPU
DetectionMatching implementations.

  detection_matching      oracle/det_matching_oracle.cc, our C++ restatement of
                          nms_net/matching_module/det_matching.cc:95-159
  ref_detection_matching  the reference's own det_matching.cc compiled
                          unmodified against oracle/tf_shim into oracle/_ref/
                          (present when `make -C oracle` ran where
                          /root/reference exists; the built .so travels)

Signature mirrors the reference op (matching_module/__init__.py:13):
(iou[N,G] f32, score[N] f32, ignore[G] bool) -> (labels f32, weights f32,
assignment i32).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, '_build', 'libdet_matching_oracle.so')
_REF = os.path.join(_HERE, '_ref', 'libdet_matching_ref.so')
_cache = {}


def build():
    subprocess.check_call(['make', '-C', _HERE, '-s'])


def _load(path):
    if path not in _cache:
        if not os.path.exists(path):
            build()
        _cache[path] = ctypes.CDLL(path)
    return _cache[path]


def have_reference_build():
    return os.path.exists(_REF)


def _prep(iou, score, ignore):
    iou = np.ascontiguousarray(iou, dtype=np.float32)
    score = np.ascontiguousarray(score, dtype=np.float32)
    ignore = np.ascontiguousarray(ignore, dtype=np.uint8)
    if iou.ndim != 2 or score.ndim != 1 or ignore.ndim != 1:
        raise ValueError('DetectionMatching expects a matrix and two vectors')
    if iou.shape[0] != score.shape[0] or iou.shape[1] != ignore.shape[0]:
        raise ValueError('DetectionMatching: inconsistent shapes')
    n, g = iou.shape
    return iou, score, ignore, n, g, (np.empty(n, np.float32),
                                      np.empty(n, np.float32),
                                      np.empty(n, np.int32))


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def detection_matching(iou, score, ignore, return_order=False):
    lib = _load(_LIB)
    iou, score, ignore, n, g, (lab, w, asg) = _prep(iou, score, ignore)
    order = np.empty(n, np.int64)
    rc = lib.oracle_detection_matching(
        _p(iou, ctypes.c_float), _p(score, ctypes.c_float), _p(ignore, ctypes.c_uint8),
        ctypes.c_int(n), ctypes.c_int(g), _p(lab, ctypes.c_float),
        _p(w, ctypes.c_float), _p(asg, ctypes.c_int32), _p(order, ctypes.c_int64))
    if rc != 0:
        raise RuntimeError('oracle_detection_matching failed: %d' % rc)
    return (lab, w, asg, order) if return_order else (lab, w, asg)


def ref_detection_matching(iou, score, ignore):
    lib = _load(_REF)
    iou, score, ignore, n, g, (lab, w, asg) = _prep(iou, score, ignore)
    err = ctypes.create_string_buffer(512)
    rc = lib.ref_detection_matching(
        _p(iou, ctypes.c_float), _p(score, ctypes.c_float), _p(ignore, ctypes.c_uint8),
        ctypes.c_int(n), ctypes.c_int(g), _p(lab, ctypes.c_float),
        _p(w, ctypes.c_float), _p(asg, ctypes.c_int32), err, ctypes.c_int(512))
    if rc != 0:
        raise RuntimeError('reference DetectionMatching failed (%d): %s'
                           % (rc, err.value.decode()))
    return lab, w, asg
