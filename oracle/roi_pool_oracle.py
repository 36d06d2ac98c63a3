"""ORACLE (test infrastructure): ctypes front-ends for the CPU ROI max pooling.

  roi_pool / roi_pool_grad          oracle/roi_pool_oracle.c (C restatement of
                                    roi_pooling_op.cc:128-187 / :374-449)
  ref_roi_pool / ref_roi_pool_grad  the reference's own roi_pooling_op.cc compiled
                                    unmodified against oracle/tf_shim (oracle/_ref)
Signatures mirror the reference ops (roi_pooling_op.cc:35-54):
roi_pool(data[B,H,W,C], rois[R,5], ph, pw, scale) -> (top[R,ph,pw,C], argmax i32).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, '_build', 'libroi_pool_oracle.so')
_REF = os.path.join(_HERE, '_ref', 'libroi_pool_ref.so')
_cache = {}
F, I = ctypes.c_float, ctypes.c_int


def _load(path):
    if path not in _cache:
        if not os.path.exists(path):
            subprocess.check_call(['make', '-C', _HERE, '-s'])
        _cache[path] = ctypes.CDLL(path)
    return _cache[path]


def have_reference_build():
    return os.path.exists(_REF)


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _prep(data, rois):
    data = np.ascontiguousarray(data, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    return data, rois


def roi_pool(data, rois, ph, pw, scale):
    data, rois = _prep(data, rois)
    b, h, w, c = data.shape
    r = rois.shape[0]
    top = np.empty((r, ph, pw, c), np.float32)
    arg = np.empty((r, ph, pw, c), np.int32)
    _load(_LIB).oracle_roi_pool_fwd(_p(data, F), I(b), I(h), I(w), I(c), _p(rois, F), I(r), I(ph),
                                    I(pw), F(scale), _p(top, F), _p(arg, ctypes.c_int32))
    return top, arg


def roi_pool_grad(data, rois, argmax, grad, ph, pw, scale):
    data, rois = _prep(data, rois)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    grad = np.ascontiguousarray(grad, dtype=np.float32)
    b, h, w, c = data.shape
    out = np.empty_like(data)
    _load(_LIB).oracle_roi_pool_bwd(I(b), I(h), I(w), I(c), _p(rois, F), I(rois.shape[0]),
                                    _p(argmax, ctypes.c_int32), _p(grad, F), I(ph), I(pw), F(scale),
                                    _p(out, F))
    return out


def ref_roi_pool(data, rois, ph, pw, scale, threads=None):
    data, rois = _prep(data, rois)
    if data.ndim != 4:
        data4 = data.reshape((1,) * (4 - data.ndim) + data.shape) if data.ndim < 4 else data
    b, h, w, c = data.shape
    r = rois.shape[0]
    top = np.empty((r, ph, pw, c), np.float32)
    arg = np.empty((r, ph, pw, c), np.int32)
    err = ctypes.create_string_buffer(256)
    rc = _load(_REF).ref_roi_pool_fwd(_p(data, F), I(b), I(h), I(w), I(c), _p(rois, F), I(r),
                                      I(1 if rois.ndim == 2 else 0), I(ph), I(pw), F(scale),
                                      _p(top, F), _p(arg, ctypes.c_int32),
                                      I(threads or os.cpu_count()), err, I(256))
    if rc:
        raise ValueError('reference RoiPool failed (%d): %s' % (rc, err.value.decode()))
    return top, arg


def ref_roi_pool_grad(data, rois, argmax, grad, ph, pw, scale, threads=None):
    data, rois = _prep(data, rois)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    grad = np.ascontiguousarray(grad, dtype=np.float32)
    b, h, w, c = data.shape
    out = np.empty_like(data)
    err = ctypes.create_string_buffer(256)
    rc = _load(_REF).ref_roi_pool_bwd(_p(data, F), I(b), I(h), I(w), I(c), _p(rois, F),
                                      I(rois.shape[0]), _p(argmax, ctypes.c_int32), _p(grad, F),
                                      I(ph), I(pw), F(scale), _p(out, F),
                                      I(threads or os.cpu_count()), err, I(256))
    if rc:
        raise ValueError('reference RoiPoolGrad failed (%d): %s' % (rc, err.value.decode()))
    return out
