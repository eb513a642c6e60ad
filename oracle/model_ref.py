"""oracle/model_ref.py -- TEST INFRASTRUCTURE ONLY.

Pure-PyTorch fp32 CPU restatement of the reference's Model.forward for ResNet-FPN backbones, written
against the reference's state_dict key layout.  It is the CPU "port" used (a) to check the CUDA
convolution stack (the reference itself runs nn.Conv2d: torch is the arithmetic oracle here), and
(b) as bench.py's cpu_baseline / --impl reference leg on the GPU box, where /root/reference does
not exist.  Pinned against the unmodified reference Model by tests/golden/model_*.npz
(oracle/gen_golden_model.py) and, when /root/reference is mounted, by a live comparison.

Follows: odtk/model.py:125-165 (forward), odtk/backbones/fpn.py:45-61 (FPN),
odtk/backbones/resnet.py:24-39 (feature extractor), torchvision/models/resnet.py BasicBlock :59-105,
Bottleneck :108-163 (stride on the 3x3: "v1.5"), stem/maxpool :197-200, eps = 1e-5."""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle

LAYERS = {"ResNet18FPN": ("basic", [2, 2, 2, 2]), "ResNet34FPN": ("basic", [3, 4, 6, 3]),
          "ResNet50FPN": ("bottleneck", [3, 4, 6, 3]), "ResNet101FPN": ("bottleneck", [3, 4, 23, 3]),
          "ResNet152FPN": ("bottleneck", [3, 8, 36, 3])}


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def _conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def features(sd, backbone, x):
    block, layers = LAYERS[backbone]
    f = "backbones.%s.features." % backbone
    x = F.relu(_bn(sd, f + "bn1", _conv(sd, f + "conv1", x, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nblocks in enumerate(layers):
        for b in range(nblocks):
            stride = 2 if (b == 0 and li > 0) else 1
            p = f + "layer%d.%d." % (li + 1, b)
            identity = x
            if block == "bottleneck":
                out = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x)))
                out = F.relu(_bn(sd, p + "bn2", _conv(sd, p + "conv2", out, stride, 1)))
                out = _bn(sd, p + "bn3", _conv(sd, p + "conv3", out))
            else:
                out = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x, stride, 1)))
                out = _bn(sd, p + "bn2", _conv(sd, p + "conv2", out, 1, 1))
            if (p + "downsample.0.weight") in sd:
                identity = _bn(sd, p + "downsample.1", _conv(sd, p + "downsample.0", x, stride))
            x = F.relu(out + identity)
        if li >= 1:
            outs.append(x)
    c3, c4, c5 = outs
    n = "backbones.%s." % backbone
    p5 = _conv(sd, n + "lateral5", c5)
    p4 = F.interpolate(p5, scale_factor=2) + _conv(sd, n + "lateral4", c4)
    p3 = F.interpolate(p4, scale_factor=2) + _conv(sd, n + "lateral3", c3)
    p6 = _conv(sd, n + "pyramid6", c5, 2, 1)
    p7 = _conv(sd, n + "pyramid7", F.relu(p6), 2, 1)
    return [_conv(sd, n + "smooth3", p3, 1, 1), _conv(sd, n + "smooth4", p4, 1, 1), _conv(sd, n + "smooth5", p5, 1, 1), p6, p7]


def head(sd, name, t):
    for i in (0, 2, 4, 6):
        t = F.relu(_conv(sd, "%s.%d" % (name, i), t, 1, 1))
    return _conv(sd, "%s.8" % name, t, 1, 1)


def forward_heads(sd, backbone, x, sigmoid=True):
    """== reference Model.forward with exporting=True (odtk/model.py:130-144)."""
    with torch.no_grad():
        sd = {k: v.float() for k, v in sd.items()}
        feats = features(sd, backbone, x.float())
        cls = [head(sd, "cls_head", t) for t in feats]
        box = [head(sd, "box_head", t) for t in feats]
        if sigmoid:
            cls = [c.sigmoid() for c in cls]
    return cls, box


def postprocess(cls_heads, box_heads, width, ratios=None, scales=None, angles=None, rotated=False,
                threshold=0.05, top_n=1000, nms=0.5, detections=100, return_index=False):
    """odtk/model.py:146-165 with the CUDA semantics of decode / nms (oracle/odtk_oracle.c)."""
    ratios = ratios or oracle.DEFAULT_RATIOS
    scales = scales or oracle.DEFAULT_SCALES
    outs = []
    for c, b in zip(cls_heads, box_heads):
        c = c.numpy() if torch.is_tensor(c) else c
        b = b.numpy() if torch.is_tensor(b) else b
        stride = width // c.shape[-1]
        anchors = (oracle.generate_anchors_rotated_axis(stride, ratios, scales, angles or oracle.DEFAULT_ANGLES)
                   if rotated else oracle.generate_anchors(stride, ratios, scales))
        outs.append(oracle.decode(c, b, anchors.reshape(-1), stride, threshold, top_n, rotated))
    cat = [np.concatenate(t, 1) for t in zip(*outs)]
    return oracle.nms(cat[0], cat[1], cat[2], nms, detections, rotated=rotated, return_index=return_index), cat


def forward(sd, backbone, x, **kw):
    cls, box = forward_heads(sd, backbone, x)
    return postprocess(cls, box, x.shape[-1], **kw)[0]
