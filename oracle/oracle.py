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
        L.oracle_iou.restype = None
        L.oracle_iou.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, f32p]
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


def iou(boxes, anchors):
    """odtk._C.iou restated (oracle_iou): boxes [nb, 8], anchors [na, 8] corner lists -> [na, nb]."""
    b, a = _f32(boxes).reshape(-1, 8), _f32(anchors).reshape(-1, 8)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().oracle_iou(_p(b), _p(a), b.shape[0], a.shape[0], _p(out))
    return out


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


def snap_to_anchors(boxes, size, stride, anchors, num_classes, anchor_ious):
    """Restatement of odtk/box.py:134-186 (snap_to_anchors) + :67-78 (box2delta) in numpy fp32, same operation
    order.  boxes [G, 5] (x, y, w, h, class) of one image, size = [W*stride, H*stride], anchors [A, 4].
    Returns (cls_target [A,C,H,W], box_target [A,4,H,W], depth [A,1,H,W], cls_index [A,H,W] int32), the last
    being the class-index form (class, -1 background, -2 ignored) of the same assignment."""
    f = np.float32
    anchors = np.asarray(anchors, dtype=f).reshape(-1, 4)
    A = anchors.shape[0]
    width, height = int(size[0] / stride), int(size[1] / stride)
    boxes = np.asarray(boxes, dtype=f).reshape(-1, 5)
    boxes = boxes[boxes[:, 4] > -1]                                   # odtk/model.py:174
    if boxes.size == 0:                                               # :140-143
        return (np.zeros((A, num_classes, height, width), f), np.zeros((A, 4, height, width), f),
                np.zeros((A, 1, height, width), f), np.full((A, height, width), -1, np.int32))
    classes = boxes[:, 4]
    xs, ys = np.arange(0, size[0], stride, dtype=f), np.arange(0, size[1], stride, dtype=f)
    x, y = np.meshgrid(xs, ys, indexing="ij")                         # [W, H] (torch.meshgrid default)
    xyxy = np.stack((x, y, x, y), 2)[None]
    anc = (xyxy + anchors.reshape(-1, 1, 1, 4)).astype(f).reshape(-1, 4)
    b = np.concatenate([boxes[:, :2], boxes[:, :2] + boxes[:, 2:4] - f(1)], 1).astype(f)
    d = np.clip(np.minimum(anc[:, None, 2:], b[:, 2:]) - np.maximum(anc[:, None, :2], b[:, :2]) + f(1), 0, None)
    inter = d[..., 0] * d[..., 1]
    b_area = (b[:, 2] - b[:, 0] + f(1)) * (b[:, 3] - b[:, 1] + f(1))
    a_area = (anc[:, 2] - anc[:, 0] + f(1)) * (anc[:, 3] - anc[:, 1] + f(1))
    overlap = inter / (a_area[:, None] + b_area - inter)
    idx = overlap.argmax(1)                                           # first maximum, like torch.max
    ov = overlap[np.arange(overlap.shape[0]), idx]
    bb = b[idx]
    a_wh = anc[:, 2:] - anc[:, :2] + f(1)                             # box2delta
    a_ctr = anc[:, :2] + f(0.5) * a_wh
    b_wh = bb[:, 2:] - bb[:, :2] + f(1)
    b_ctr = bb[:, :2] + f(0.5) * b_wh
    delta = np.concatenate([(b_ctr - a_ctr) / a_wh, np.log(b_wh / a_wh)], 1).astype(f)
    box_target = delta.reshape(A, width, height, 4).transpose(0, 3, 2, 1)
    depth = np.full(ov.shape, -1, f)
    depth[ov < f(anchor_ious[0])] = 0
    fg = ov >= f(anchor_ious[1])
    depth[fg] = classes[idx][fg] + f(1)
    cls = classes[idx].astype(np.int64)
    onehot = np.where(ov < f(anchor_ious[0]), num_classes, cls)
    cls_target = np.zeros((anc.shape[0], num_classes + 1), f)
    cls_target[np.arange(anc.shape[0]), onehot] = 1
    cls_target = cls_target[:, :num_classes].reshape(A, width, height, num_classes).transpose(0, 3, 2, 1)
    cls_index = np.where(ov < f(anchor_ious[0]), -1, np.where(fg, cls, -2)).astype(np.int32)
    return (np.ascontiguousarray(cls_target), np.ascontiguousarray(box_target),
            np.ascontiguousarray(depth.reshape(A, width, height).transpose(0, 2, 1))[:, None],
            np.ascontiguousarray(cls_index.reshape(A, width, height).transpose(0, 2, 1)))
