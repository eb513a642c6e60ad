"""ctypes/numpy doors onto oracle/liboracle.so (the plain-C restatement) plus the
pure-torch-CPU restatements of the small host-side pieces (anchors, focal loss).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liboracle.so")], check=True)
        L = ctypes.CDLL(path)
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.oracle_decode.restype = ctypes.c_int
        L.oracle_decode.argtypes = [ctypes.c_int, f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_int, ctypes.c_int, f32p, f32p, f32p]
        L.oracle_nms.restype = ctypes.c_int
        L.oracle_nms.argtypes = [ctypes.c_int, f32p, f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                 ctypes.c_int, ctypes.c_int, f32p, f32p, f32p, i32p]
        L.oracle_rotated_overlap.restype = ctypes.c_float
        L.oracle_rotated_overlap.argtypes = [f32p, f32p, ctypes.c_int]
        L.oracle_aligned_overlap.restype = ctypes.c_float
        L.oracle_aligned_overlap.argtypes = [f32p, f32p]
        L.oracle_focal_loss.restype = ctypes.c_double
        L.oracle_focal_loss.argtypes = [f32p, f32p, f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_float, f32p, f32p]
        L.oracle_smooth_l1.restype = ctypes.c_double
        L.oracle_smooth_l1.argtypes = [f32p, f32p, f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_float, f32p, f32p]
        L.oracle_preprocess_u8.restype = None
        L.oracle_preprocess_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, f32p, f32p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def decode(cls_head, box_head, anchors, scale, thresh, top_n, rotated=False):
    """cls_head [B,A*C,H,W], box_head [B,A*nbox,H,W] fp32; anchors flat list (4A).
    Mirrors odtk._C.decode (csrc/extensions.cpp:69-115)."""
    cls_head, box_head = _f32(cls_head), _f32(box_head)
    anchors = _f32(anchors).reshape(-1)
    nbox = 6 if rotated else 4
    B, AC, H, W = cls_head.shape
    A = anchors.size // 4 if anchors.size else box_head.shape[1] // nbox
    C = AC // A
    os_ = np.zeros((B, top_n), np.float32)
    ob = np.zeros((B, top_n, nbox), np.float32)
    oc = np.zeros((B, top_n), np.float32)
    rc = lib().oracle_decode(B, _p(cls_head), _p(box_head), H, W, int(scale), A, C, _p(anchors),
                             int(anchors.size), float(thresh), int(top_n), nbox, _p(os_), _p(ob), _p(oc))
    assert rc == 0
    return os_, ob, oc


def nms(scores, boxes, classes, nms_thresh, detections, rotated=False, fixed_angle=False, return_index=False):
    """Mirrors odtk._C.nms (csrc/extensions.cpp:117-158)."""
    scores, boxes, classes = _f32(scores), _f32(boxes), _f32(classes)
    nbox = 6 if rotated else 4
    B, N = scores.shape
    os_ = np.zeros((B, detections), np.float32)
    ob = np.zeros((B, detections, nbox), np.float32)
    oc = np.zeros((B, detections), np.float32)
    oi = np.zeros((B, detections), np.int32)
    rc = lib().oracle_nms(B, _p(scores), _p(boxes), _p(classes), N, int(detections), float(nms_thresh), nbox,
                          int(bool(fixed_angle)), _p(os_), _p(ob), _p(oc),
                          oi.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert rc == 0
    return (os_, ob, oc, oi) if return_index else (os_, ob, oc)


def rotated_overlap(ibox, mbox, fixed_angle=False):
    a, b = _f32(ibox), _f32(mbox)
    return float(lib().oracle_rotated_overlap(_p(a), _p(b), int(bool(fixed_angle))))


def aligned_overlap(ibox, mbox):
    a, b = _f32(ibox), _f32(mbox)
    return float(lib().oracle_aligned_overlap(_p(a), _p(b)))


def focal_loss(logits, target, mask=None, alpha=0.25, gamma=2.0, grad_scale=1.0):
    """Returns (masked sum, per-element masked loss, grad of the sum * grad_scale)."""
    x, t = _f32(logits).reshape(-1), _f32(target).reshape(-1)
    m = None if mask is None else _f32(mask).reshape(-1)
    lo = np.empty_like(x)
    g = np.empty_like(x)
    tot = lib().oracle_focal_loss(_p(x), _p(t), _p(m) if m is not None else None, x.size, alpha, gamma,
                                  grad_scale, _p(lo), _p(g))
    return tot, lo, g


# ---------------------------------------------------------------------------------------------
def generate_anchors(stride, ratio_vals, scales_vals):
    """odtk/box.py:8-20 restated in float32 numpy.  Returns [A,4]."""
    f = np.float32
    scales = np.repeat(np.asarray(scales_vals, dtype=f), len(ratio_vals)).reshape(-1, 1)
    ratios = np.asarray(list(ratio_vals) * len(scales_vals), dtype=f)
    wh = np.full((len(ratios), 2), stride, dtype=f)
    ws = np.sqrt(wh[:, 0] * wh[:, 1] / ratios).astype(f)
    dwh = np.stack([ws, ws * ratios], axis=1).astype(f)
    xy1 = f(0.5) * (wh - dwh * scales)
    xy2 = f(0.5) * (wh + dwh * scales)
    return np.concatenate([xy1, xy2], axis=1).astype(f)


def generate_anchors_rotated_axis(stride, ratio_vals, scales_vals, angles_vals):
    """The [A*len(angles),4] axis-aligned table of odtk/box.py:23-64 (anchors_axis), which is
    the only part decode uses (odtk/box.py:258-259, decode_rotate.cu:139)."""
    f = np.float32
    scales = np.repeat(np.asarray(scales_vals, dtype=f), len(ratio_vals)).reshape(-1, 1)
    ratios = np.asarray(list(ratio_vals) * len(scales_vals), dtype=f)
    wh = np.full((len(ratios), 2), stride, dtype=f)
    ws = np.round(np.sqrt(wh[:, 0] * wh[:, 1] / ratios).astype(f))
    dwh = np.stack([ws, np.round(ws * ratios)], axis=1).astype(f)
    xy0 = f(0.5) * (wh - dwh * scales)
    xy2 = f(0.5) * (wh + dwh * scales) - f(1)
    na = len(angles_vals)
    return np.concatenate([np.tile(xy0, (na, 1)), np.tile(xy2, (na, 1))], axis=1).astype(f)


DEFAULT_RATIOS = [1.0, 2.0, 0.5]
DEFAULT_SCALES = [4 * 2 ** (i / 3) for i in range(3)]
DEFAULT_ANGLES = [-math.pi / 6, 0, math.pi / 6]


def smooth_l1(pred, target, mask=None, beta=0.11, grad_scale=1.0):
    """odtk/loss.py:27-31.  Returns (masked sum, per-element loss, grad of the sum * grad_scale)."""
    x, t = _f32(pred).reshape(-1), _f32(target).reshape(-1)
    m = None if mask is None else _f32(mask).reshape(-1)
    lo, g = np.empty_like(x), np.empty_like(x)
    tot = lib().oracle_smooth_l1(_p(x), _p(t), _p(m) if m is not None else None, x.size, beta, grad_scale, _p(lo), _p(g))
    return tot, lo, g


def preprocess_u8(image, stride=128, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """odtk/data.py:113-123: uint8 [H, W, 3] -> float32 [3, Hs, Ws] normalised and zero-padded."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    h, w, _ = img.shape
    hs, ws = (h + stride - 1) // stride * stride, (w + stride - 1) // stride * stride
    out = np.empty((3, hs, ws), np.float32)
    lib().oracle_preprocess_u8(img.ctypes.data_as(ctypes.c_void_p), h, w, hs, ws, _p(_f32(mean)), _p(_f32(std)), _p(out))
    return out
