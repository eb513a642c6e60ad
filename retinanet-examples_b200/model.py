"""RetinaNet inference model -- host-side mirror of the reference's odtk/model.py `Model`
(constructor arguments, `config` keys, state_dict key names, `forward(x) -> (scores, boxes,
classes)`), executing on the sm_100a kernels behind the C ABI instead of nn.Conv2d/cuDNN +
odtk._C.  `_compute_loss` (training-mode forward) runs target assignment + ONE fused loss kernel; `save` / `load` /
`initialize(pre_trained)` read and write the reference's checkpoint dict; ONNX/TensorRT export is outside the
hot path; `load_state_dict` ingests exactly the reference's key layout
(odtk/model.py:217-258): `backbones.<Name>.features.*`, `backbones.<Name>.{lateral,pyramid,smooth}*`,
`cls_head.{0,2,4,6,8}.*`, `box_head.{0,2,4,6,8}.*`.

Data flow of forward() (reference: odtk/model.py:125-165, odtk/backbones/fpn.py:45-61,
odtk/backbones/resnet.py:24-39, torchvision BasicBlock/Bottleneck):
  NHWC fp16 activations; every conv is one launch of the tcgen05 kernel with BatchNorm folded
  into weights/bias, ReLU / residual add / FPN upsample-add fused in its epilogue; the last conv
  of each head writes fp32 NCHW (sigmoid fused for the class head) which feeds the all-levels
  decode (3 launches) and the batched NMS (1 launch)."""
import math
import os

import numpy as np
import torch

from . import _C, box, engine

RESNET_LAYERS = {"ResNet18FPN": ("basic", [2, 2, 2, 2]), "ResNet34FPN": ("basic", [3, 4, 6, 3]),
                 "ResNet50FPN": ("bottleneck", [3, 4, 6, 3]), "ResNet101FPN": ("bottleneck", [3, 4, 23, 3]),
                 "ResNet152FPN": ("bottleneck", [3, 8, 36, 3]),
                 "ResNeXt50_32x4dFPN": ("bottleneck", [3, 4, 6, 3]), "ResNeXt101_32x8dFPN": ("bottleneck", [3, 4, 23, 3])}
# grouped bottlenecks (odtk/backbones/fpn.py:85-91): (groups, width_per_group); conv2 has planes * width_per_group / 64 * groups channels
RESNEXT = {"ResNeXt50_32x4dFPN": (32, 4), "ResNeXt101_32x8dFPN": (32, 8)}
# MobileNetV2 (odtk/backbones/mobilenet.py:5-25, fpn.py:92-93; torchvision mobilenetv2.py inverted_residual_setting):
# (expansion t, output channels c, repeats n, first stride s); FPN taps features[6 / 13 / 17] = 32 / 96 / 320 channels
MOBILENET = {"MobileNetV2FPN": [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]}
MOBILENET_TAPS = (6, 13, 17)
BACKBONES = tuple(RESNET_LAYERS) + tuple(MOBILENET)


def mobilenet_blocks(backbone):
    """[(features index, cin, cout, stride, t)] of the inverted residual blocks (features[1 .. 17])."""
    out, cin, idx = [], 32, 1
    for t, c, n, s in MOBILENET[backbone]:
        for i in range(n):
            out.append((idx, cin, c, s if i == 0 else 1, t))
            cin, idx = c, idx + 1
    return out


def _head_specs(classes, num_anchors, rotated):
    specs = []
    nbox = 6 if rotated else 4
    for head, out in (("cls_head", classes * num_anchors), ("box_head", nbox * num_anchors)):
        for i in (0, 2, 4, 6):
            specs.append(("%s.%d" % (head, i), "convb", (256, 256, 3, 3)))
        specs.append(("%s.8" % head, "convb_final", (out, 256, 3, 3)))
    return specs


def _fpn_specs(n, ch):
    return [(n + "lateral3", "convb", (256, ch[0], 1, 1)), (n + "lateral4", "convb", (256, ch[1], 1, 1)),
            (n + "lateral5", "convb", (256, ch[2], 1, 1)), (n + "pyramid6", "convb", (256, ch[2], 3, 3)),
            (n + "pyramid7", "convb", (256, 256, 3, 3)), (n + "smooth3", "convb", (256, 256, 3, 3)),
            (n + "smooth4", "convb", (256, 256, 3, 3)), (n + "smooth5", "convb", (256, 256, 3, 3))]


def mobilenet_specs(backbone, classes, num_anchors, rotated):
    f = "backbones.%s.features.features." % backbone       # FPN.features = MobileNet module, MobileNet.features = Sequential
    specs = [(f + "0.0", "conv", (32, 3, 3, 3)), (f + "0.1", "bn", 32)]
    for idx, cin, cout, stride, t in mobilenet_blocks(backbone):
        hidden, k = cin * t, 0
        p = f + "%d.conv." % idx
        if t != 1:
            specs += [(p + "0.0", "conv", (hidden, cin, 1, 1)), (p + "0.1", "bn", hidden)]
            k = 1
        specs += [(p + "%d.0" % k, "conv", (hidden, 1, 3, 3)), (p + "%d.1" % k, "bn", hidden),
                  (p + "%d" % (k + 1), "conv", (cout, hidden, 1, 1)), (p + "%d" % (k + 2), "bn_last", cout)]
    return specs + _fpn_specs("backbones.%s." % backbone, [32, 96, 320]) + _head_specs(classes, num_anchors, rotated)


def conv_specs(backbone, classes=80, num_anchors=9, rotated=False):
    """Every convolution / batch-norm of the model as (state_dict prefix, kind, shape info), in
    forward order.  kind: 'conv' (weight only, followed by 'bn') or 'convb' (weight + bias)."""
    if backbone in MOBILENET:
        return mobilenet_specs(backbone, classes, num_anchors, rotated)
    block, layers = RESNET_LAYERS[backbone]
    f = "backbones.%s.features." % backbone
    specs = [(f + "conv1", "conv", (64, 3, 7, 7)), (f + "bn1", "bn", 64)]
    inplanes = 64
    exp = 4 if block == "bottleneck" else 1
    for li, (planes, nblocks) in enumerate(zip([64, 128, 256, 512], layers)):
        for b in range(nblocks):
            stride = 2 if (b == 0 and li > 0) else 1
            p = f + "layer%d.%d." % (li + 1, b)
            if block == "bottleneck":
                groups, wpg = RESNEXT.get(backbone, (1, 64))
                width = planes * wpg // 64 * groups
                specs += [(p + "conv1", "conv", (width, inplanes, 1, 1)), (p + "bn1", "bn", width),
                          (p + "conv2", "conv", (width, width // groups, 3, 3)), (p + "bn2", "bn", width),
                          (p + "conv3", "conv", (planes * 4, width, 1, 1)), (p + "bn3", "bn_last", planes * 4)]
            else:
                specs += [(p + "conv1", "conv", (planes, inplanes, 3, 3)), (p + "bn1", "bn", planes),
                          (p + "conv2", "conv", (planes, planes, 3, 3)), (p + "bn2", "bn_last", planes)]
            if b == 0 and (stride != 1 or inplanes != planes * exp):
                specs += [(p + "downsample.0", "conv", (planes * exp, inplanes, 1, 1)),
                          (p + "downsample.1", "bn", planes * exp)]
            inplanes = planes * exp
    ch = [128, 256, 512] if block == "basic" else [512, 1024, 2048]
    n = "backbones.%s." % backbone
    specs += [(n + "lateral3", "convb", (256, ch[0], 1, 1)), (n + "lateral4", "convb", (256, ch[1], 1, 1)),
              (n + "lateral5", "convb", (256, ch[2], 1, 1)), (n + "pyramid6", "convb", (256, ch[2], 3, 3)),
              (n + "pyramid7", "convb", (256, 256, 3, 3)), (n + "smooth3", "convb", (256, 256, 3, 3)),
              (n + "smooth4", "convb", (256, 256, 3, 3)), (n + "smooth5", "convb", (256, 256, 3, 3))]
    nbox = 6 if rotated else 4
    for head, out in (("cls_head", classes * num_anchors), ("box_head", nbox * num_anchors)):
        for i in (0, 2, 4, 6):
            specs.append(("%s.%d" % (head, i), "convb", (256, 256, 3, 3)))
        specs.append(("%s.8" % head, "convb_final", (out, 256, 3, 3)))
    return specs


def make_state_dict(backbone="ResNet50FPN", classes=80, num_anchors=9, rotated=False, seed=0, cls_prior=0.01):
    """Deterministic random-init weights in the reference's state_dict layout (no checkpoint or
    pretrained weights exist offline).  BatchNorm statistics are randomised so that folding is
    exercised; the last BN of each residual block is damped so
    that activations stay well inside fp16 range through 50+ layers."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, kind, shp in conv_specs(backbone, classes, num_anchors, rotated):
        if kind.startswith("conv"):
            cout, cin, kh, kw = shp
            std = 0.01 if name.startswith(("cls_head", "box_head")) else math.sqrt(2.0 / (cin * kh * kw))   # cin = per-group fan-in
            sd[name + ".weight"] = torch.randn(shp, generator=g) * std
            if kind != "conv":
                sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.01
                if kind == "convb_final" and name.startswith("cls_head"):
                    sd[name + ".bias"] = torch.full((cout,), -math.log((1 - cls_prior) / cls_prior))
        else:
            c = shp
            lo, hi = (0.1, 0.3) if kind == "bn_last" else (0.5, 1.5)
            sd[name + ".weight"] = torch.rand(c, generator=g) * (hi - lo) + lo
            sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
            sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
            sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5
    return sd


class _Conv:
    """One packed convolution: fp16 weights in the kernel's K order, fp32 bias, geometry."""

    def __init__(self, weight, bias, stride=1, device="cuda", groups=1):
        cout, cin, kh, kw = weight.shape
        self.groups = groups
        if groups > 1:                       # grouped 3x3 (ResNeXt): block-diagonal packing, see engine.pack_weight_grouped
            self.cout, self.cin, self.ks, self.stride = cout, cin * groups, kh, stride
            self.direct, self.stem, self.w_stem = stride == 1, False, None
            self.kpad = kh * kw * 64
            self.w = engine.pack_weight_grouped(weight, groups).to(device)
            self.b = bias.float().contiguous().to(device) if bias is not None else None
            self.bop = engine.pack_bias(self.b) if self.b is not None else None
            return
        self.cout, self.cin, self.ks, self.stride = cout, cin, kh, stride
        self.direct = stride == 1 and kh in (1, 3) and cin % 64 == 0
        self.stem = (cin == 3 and kh == 7 and stride == 2 and cout % 16 == 0 and cout <= 256)
        self.w_stem = engine.pack_stem_weight(weight).to(device) if self.stem else None
        kreal = kh * kw * cin
        self.kpad = kreal if self.direct else (kreal + 63) // 64 * 64
        self.w = engine.pack_weight(weight, self.kpad).to(device)
        self.b = bias.float().contiguous().to(device) if bias is not None else None
        self.bop = engine.pack_bias(self.b) if self.b is not None else None     # bias as a tensor-core K block

    def __call__(self, x, relu=False, residual=None, upsample=None, out_mode=engine.OUT_NHWC_F16, in_relu=False,
                 sink=None, out=None, tile_tab=None, valid_px=None):
        oh, ow = (x.shape[1] - 1) // self.stride + 1, (x.shape[2] - 1) // self.stride + 1
        px = x.shape[0] * oh * ow if valid_px is None else valid_px          # atlas launches: only the pixels of the levels count
        engine.STATS["conv_flops"] += 2 * px * self.cout * self.ks * self.ks * self.cin // self.groups
        if self.stem and x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0 and residual is None and upsample is None:
            return engine.stem_conv(x, self.w_stem, self.b, self.cout, relu)
        if self.direct:
            assert not in_relu
            return engine.conv2d(x, self.w, self.b, self.cout, self.ks, relu, residual, upsample, out_mode,
                                 bias_op=self.bop, sink=sink, groups=self.groups, out=out, tile_tab=tile_tab)
        assert sink is None
        if self.stride == 2 and self.ks in (1, 3) and self.cin % 64 == 0 and upsample is None:
            # stride-2 convolution, any size: strided TMA view (even sizes: parity split, odd sizes: element-strided
            # boxes), no gather pre-pass; ReLU on the input (FPN pyramid7) is one tiny elementwise launch
            if in_relu:
                x = engine.relu_from_view(x) if isinstance(x, engine.AtlasView) else engine.relu(x)
            return engine.conv2d(x, self.w, self.b, self.cout, self.ks, relu, residual, None, out_mode, stride=2,
                                 bias_op=self.bop, groups=self.groups, out=out)
        low = engine.lower_conv(x, self.ks, self.stride, self.ks // 2, self.kpad if self.cin % 8 else None, in_relu)
        return engine.conv2d(low, self.w, self.b, self.cout, 1, relu, residual, upsample, out_mode, bias_op=self.bop)


class Model:
    'RetinaNet - https://arxiv.org/abs/1708.02002 (reference: odtk/model.py:15-72)'

    def __init__(self, backbones='ResNet50FPN', classes=80, ratios=[1.0, 2.0, 0.5],
                 scales=[4 * 2 ** (i / 3) for i in range(3)], angles=None, rotated_bbox=False,
                 anchor_ious=[0.4, 0.5], config={}):
        if isinstance(backbones, (list, tuple)):
            if len(backbones) != 1:
                raise ValueError("one backbone per model on the B200 path")
            backbones = backbones[0]
        if backbones not in BACKBONES:
            raise ValueError("unsupported backbone %r (hot path: %s)" % (backbones, ", ".join(BACKBONES)))
        self.backbone = backbones
        self.name = 'RetinaNet'
        self.exporting = False
        self.training = False
        self.rotated_bbox = rotated_bbox
        self.anchor_ious = anchor_ious
        self.ratios, self.scales = ratios, scales
        self.angles = angles if angles is not None else [-np.pi / 6, 0, np.pi / 6] if rotated_bbox else None
        self.anchors = {}
        self.classes = classes
        self.threshold = config.get('threshold', 0.05)
        self.top_n = config.get('top_n', 1000)
        self.nms = config.get('nms', 0.5)
        self.detections = config.get('detections', 100)
        self.stride = 128
        if self.top_n > 4096 or self.detections > 1024 or 5 * self.top_n > 6144:
            raise ValueError("config outside the sm_100a kernels' limits (include/odtk_b200.h): top_n <= 4096, "
                             "5 * top_n <= 6144 candidates per image into NMS, detections <= 1024; got top_n=%d detections=%d"
                             % (self.top_n, self.detections))
        self.num_anchors = len(ratios) * len(scales) * (len(self.angles) if rotated_bbox else 1)
        self._sd = None
        self._packed = None
        self.device = None
        self.parallel_heads = True
        self.fused_candidates = os.environ.get("ODTK_FUSED_CANDIDATES", "1") != "0"
        self.fused_stem = os.environ.get("ODTK_FUSED_STEM", "1") != "0"
        self.merged_heads = os.environ.get("ODTK_MERGED_HEADS", "1") != "0"
        self.fused_bneck = os.environ.get("ODTK_FUSED_BNECK", "1") != "0"
        self.fused_conv1 = os.environ.get("ODTK_FUSED_CONV1", "1") != "0"
        self._atlas = {}
        self._fused = {}
        self._head_streams = None
        self.gather = None            # peer.PeerGather: image-wise sharding, detections pushed to all ranks by the NMS kernel
        self._packed_out = {}

    def __repr__(self):
        return '\n'.join(['     model: {}'.format(self.name), '  backbone: {}'.format(self.backbone),
                          '   classes: {}, anchors: {}'.format(self.classes, self.num_anchors)])

    # ---- weights ---------------------------------------------------------------------------------
    def initialize(self, pre_trained=None, seed=0):
        """Reference odtk/model.py:79-123.  `pre_trained`: a checkpoint file written by `save` (or by the reference):
        every weight except the class head's last layer (and, rotated, the box head's) is taken from it -- fine-tuning, as
        the reference does; those layers get the prior initialisation.  Otherwise seeded random init (no ImageNet
        backbone checkpoints exist offline)."""
        sd = make_state_dict(self.backbone, self.classes, self.num_anchors, self.rotated_bbox, seed)
        if pre_trained:
            if not os.path.isfile(pre_trained):
                raise ValueError('No checkpoint {}'.format(pre_trained))
            chk = torch.load(pre_trained, map_location="cpu", weights_only=False)
            ignored = ['cls_head.8.bias', 'cls_head.8.weight']
            if self.rotated_bbox:
                ignored += ['box_head.8.bias', 'box_head.8.weight']
            sd.update({k: v for k, v in chk['state_dict'].items() if k not in ignored and k in sd})
        self.load_state_dict(sd)
        return self

    # ---- checkpoint files (reference odtk/model.py:217-258: same dict layout, readable by either side) ---------
    def save(self, state):
        checkpoint = {'backbone': [self.backbone], 'classes': self.classes, 'state_dict': self.state_dict(),
                      'ratios': self.ratios, 'scales': self.scales}
        if self.rotated_bbox and self.angles:
            checkpoint['angles'] = self.angles
        for key in ('iteration', 'optimizer', 'scheduler'):
            if key in state:
                checkpoint[key] = state[key]
        torch.save(checkpoint, state['path'])

    @classmethod
    def load(cls, filename, rotated_bbox=False):
        if not os.path.isfile(filename):
            raise ValueError('No checkpoint {}'.format(filename))
        checkpoint = torch.load(filename, map_location="cpu", weights_only=False)
        kwargs = {k: checkpoint[k] for k in ('ratios', 'scales', 'angles') if k in checkpoint}
        if ('angles' in checkpoint) or rotated_bbox:
            kwargs['rotated_bbox'] = True
        model = cls(backbones=checkpoint['backbone'], classes=checkpoint['classes'], **kwargs)
        model.load_state_dict(checkpoint['state_dict'])
        state = {key: checkpoint[key] for key in ('iteration', 'optimizer', 'scheduler') if key in checkpoint}
        return model, state

    def state_dict(self):
        return self._sd

    def load_state_dict(self, sd):
        need = [n + ".weight" for n, _, _ in conv_specs(self.backbone, self.classes, self.num_anchors, self.rotated_bbox)]
        missing = [k for k in need if k not in sd]
        if missing:
            raise RuntimeError("missing keys in state_dict: %s" % missing[:4])
        self._sd = {k: v.detach().float().cpu() for k, v in sd.items() if torch.is_tensor(v)}
        self._packed = None
        return self

    def cuda(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._pack()
        return self

    def to(self, *args, **kwargs):   # memory_format / dtype arguments of the reference call sites are accepted
        for a in args:
            if isinstance(a, (str, torch.device)) and str(a).startswith("cuda"):
                return self.cuda(torch.device(a).index)
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        """Training-mode forward (odtk/model.py:130-138): `model([images, targets])` returns (cls_loss, box_loss).  The
        convolutions have no backward kernels here (the training loop is outside the hot path): the
        losses and -- `_compute_loss(..., with_grad=True)` -- their gradients w.r.t. the head outputs are what this
        provides."""
        self.training = bool(mode)
        return self

    def share_memory(self):
        return self

    def _pack(self):
        sd, dev = self._sd, self.device
        P = {}

        def conv_bn(prefix_conv, prefix_bn, stride=1, groups=1):
            w, b = engine.fold_bn(sd[prefix_conv + ".weight"], sd[prefix_bn + ".weight"], sd[prefix_bn + ".bias"],
                                  sd[prefix_bn + ".running_mean"], sd[prefix_bn + ".running_var"])
            return _Conv(w, b, stride, dev, groups)

        def conv_b(prefix, stride=1):
            return _Conv(sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride, dev)

        if self.backbone in MOBILENET:
            self._pack_mobilenet(P)
            self._packed = P
            return
        block, layers = RESNET_LAYERS[self.backbone]
        f = "backbones.%s.features." % self.backbone
        P["stem"] = conv_bn(f + "conv1", f + "bn1", 2)
        blocks = []
        for li, nblocks in enumerate(layers):
            for b in range(nblocks):
                stride = 2 if (b == 0 and li > 0) else 1
                p = f + "layer%d.%d." % (li + 1, b)
                blk = {"level": li + 2, "last": b == nblocks - 1}
                if block == "bottleneck":
                    blk["convs"] = [conv_bn(p + "conv1", p + "bn1"),
                                    conv_bn(p + "conv2", p + "bn2", stride, RESNEXT.get(self.backbone, (1, 64))[0]),
                                    conv_bn(p + "conv3", p + "bn3")]
                else:
                    blk["convs"] = [conv_bn(p + "conv1", p + "bn1", stride), conv_bn(p + "conv2", p + "bn2")]
                blk["down"] = conv_bn(p + "downsample.0", p + "downsample.1", stride) if (p + "downsample.0.weight") in sd else None
                blocks.append(blk)
        P["blocks"] = blocks
        n = "backbones.%s." % self.backbone
        for k in ("lateral3", "lateral4", "lateral5", "smooth3", "smooth4", "smooth5"):
            P[k] = conv_b(n + k)
        P["pyramid6"] = conv_b(n + "pyramid6", 2)
        P["pyramid7"] = conv_b(n + "pyramid7", 2)
        for head in ("cls_head", "box_head"):
            P[head] = [conv_b("%s.%d" % (head, i)) for i in (0, 2, 4, 6, 8)]
        self._packed = P

    def _pack_mobilenet(self, P):
        """MobileNetV2FPN (odtk/backbones/mobilenet.py, fpn.py:92-93).  The tensor-core kernels work on 64-channel K
        blocks: every activation is carried with its channel count padded to a multiple of 64 (zero weights / zero bias in
        the padding, so the padding stays exactly 0 through ReLU6 and the residual adds)."""
        sd, dev = self._sd, self.device
        f = "backbones.%s.features.features." % self.backbone

        def up64(c):
            return (c + 63) // 64 * 64

        def fold(pc, pb):
            return engine.fold_bn(sd[pc + ".weight"], sd[pb + ".weight"], sd[pb + ".bias"], sd[pb + ".running_mean"], sd[pb + ".running_var"])

        def padded_conv(w, b, stride=1, pad_in=True):
            cout, cin = w.shape[0], w.shape[1]
            wp = w.new_zeros((up64(cout), up64(cin) if pad_in else cin, w.shape[2], w.shape[3]))
            wp[:cout, :cin] = w
            bp = b.new_zeros(up64(cout))
            bp[:cout] = b
            return _Conv(wp, bp, stride, dev)

        P["stem"] = padded_conv(*fold(f + "0.0", f + "0.1"), stride=2, pad_in=False)       # 3x3 s2, 3 -> 32: receptive-field gather + GEMM
        blocks = []
        for idx, cin, cout, stride, t in mobilenet_blocks(self.backbone):
            p, k = f + "%d.conv." % idx, 0
            blk = {"idx": idx, "stride": stride, "res": stride == 1 and cin == cout, "expand": None}
            if t != 1:
                blk["expand"] = padded_conv(*fold(p + "0.0", p + "0.1"))
                k = 1
            wd, bd = fold(p + "%d.0" % k, p + "%d.1" % k)                                 # depthwise [hidden, 1, 3, 3]
            hidden = wd.shape[0]
            wdp = wd.new_zeros((9, up64(hidden)))
            wdp[:, :hidden] = wd.reshape(hidden, 9).t()
            bdp = bd.new_zeros(up64(hidden))
            bdp[:hidden] = bd
            blk["dw_w"], blk["dw_b"] = wdp.to(torch.float16).contiguous().to(dev), bdp.float().contiguous().to(dev)
            blk["hidden"] = hidden
            blk["project"] = padded_conv(*fold(p + "%d" % (k + 1), p + "%d" % (k + 2)))
            blocks.append(blk)
        P["blocks"] = blocks
        n = "backbones.%s." % self.backbone
        for kname in ("lateral3", "lateral4", "lateral5"):
            P[kname] = padded_conv(sd[n + kname + ".weight"], sd[n + kname + ".bias"])
        for kname in ("smooth3", "smooth4", "smooth5"):
            P[kname] = _Conv(sd[n + kname + ".weight"], sd.get(n + kname + ".bias"), 1, dev)
        P["pyramid6"] = _Conv(sd[n + "pyramid6.weight"], sd.get(n + "pyramid6.bias"), 2, dev)
        P["pyramid7"] = _Conv(sd[n + "pyramid7.weight"], sd.get(n + "pyramid7.bias"), 2, dev)
        for head in ("cls_head", "box_head"):
            P[head] = [_Conv(sd["%s.%d.weight" % (head, i)], sd.get("%s.%d.bias" % (head, i)), 1, dev) for i in (0, 2, 4, 6, 8)]

    def _features_mobilenet(self, x):
        """`x`: NHWC fp16 image [N, H, W, 3].  torchvision MobileNetV2.features[0 .. 17] with ReLU6, taps after blocks
        6 / 13 / 17 (odtk/backbones/mobilenet.py:18-25), then the shared FPN."""
        P = self._packed
        x = P["stem"](x, relu=2)
        outs = {}
        for blk in P["blocks"]:
            inp = x
            hcur = blk["expand"](x, relu=2) if blk["expand"] is not None else x
            engine.STATS["conv_flops"] += 2 * hcur.shape[0] * ((hcur.shape[1] - 1) // blk["stride"] + 1) * ((hcur.shape[2] - 1) // blk["stride"] + 1) * blk["hidden"] * 9
            hcur = engine.depthwise3x3(hcur, blk["dw_w"], blk["dw_b"], blk["stride"], act=2)
            x = blk["project"](hcur, relu=False, residual=inp if blk["res"] else None)
            if blk["idx"] in MOBILENET_TAPS:
                outs[blk["idx"]] = x
        return self._fpn(outs[6], outs[13], outs[17])

    # ---- forward ---------------------------------------------------------------------------------
    def _stem(self, x=None, padded=None):
        """conv1 + bn1 + relu + maxpool (odtk/backbones/resnet.py:25-28).  One fused kernel when the stem has the
        standard 64 output channels and the image size is even; `padded` = (xp, h, w) from engine.preprocess_u8."""
        if self.backbone in MOBILENET:
            if padded is not None:
                raise RuntimeError("the fused uint8 input path feeds the ResNet stem; pass a float / half image to MobileNetV2FPN")
            return x                                        # the 3x3 stem runs inside _features_mobilenet
        stem = self._packed["stem"]
        if padded is not None:
            xp, h, w = padded
            n = xp.shape[0]
        else:
            n, h, w = x.shape[0], x.shape[1], x.shape[2]
        engine.STATS["conv_flops"] += 2 * n * (h // 2) * (w // 2) * stem.cout * 49 * 3
        fused = self.fused_stem and stem.stem and stem.cout == 64 and h % 2 == 0 and w % 2 == 0
        if padded is not None:
            if fused:
                return engine.stem_pool_padded(xp, h, w, stem.w_stem, stem.b, stem.cout, relu=True)
            return engine.maxpool3x3s2(engine.stem_conv_padded(xp, h, w, stem.w_stem, stem.b, stem.cout, relu=True))
        if fused:
            return engine.stem_pool(x, stem.w_stem, stem.b, stem.cout, relu=True)
        engine.STATS["conv_flops"] -= 2 * n * (h // 2) * (w // 2) * stem.cout * 49 * 3      # _Conv.__call__ counts it
        return engine.maxpool3x3s2(stem(x, relu=True))

    def _features(self, x):
        """`x`: the POOLED stem output [N, H/4, W/4, 64] (from _stem)."""
        if self.backbone in MOBILENET:
            return self._features_mobilenet(x)
        P = self._packed
        outs = {}
        blocks = P["blocks"]
        z_next = None                 # conv1 of THIS block, already computed by the previous block's fused tail (GEMM3)
        for bi, blk in enumerate(blocks):
            cs = blk["convs"]
            dn = blk["down"]
            z_in, z_next = z_next, None
            if (self.fused_bneck and dn is not None and len(cs) == 3 and dn.ks == 1 and dn.stride == 1 and dn.cin == 64 and
                    dn.groups == 1 and cs[1].stride == 1 and cs[1].groups == 1 and cs[1].ks == 3 and cs[1].cin == 64 and
                    cs[1].cout == 64 and cs[2].cin == 64 and cs[2].cout == dn.cout and dn.cout % 128 == 0 and dn.cout <= 512):
                # first block of layer1: conv2 + conv3 + the 1x1 PROJECTION of the block input (torchvision `downsample`)
                # + ReLU in one kernel: neither the 3x3's output nor the projected identity touches HBM
                if "b3d" not in blk:
                    blk["b3d"] = (cs[2].b + dn.b).contiguous()
                out = cs[0](x, relu=True)
                px = out.shape[0] * out.shape[1] * out.shape[2]
                engine.STATS["conv_flops"] += 2 * px * (64 * (9 * 64 + cs[2].cout) + 64 * dn.cout)
                x = engine.bottleneck_tail(out, cs[1].w, cs[1].b, cs[2].w, blk["b3d"], None, relu=True, xproj=x, wproj=dn.w)
                if blk["last"]:
                    outs[blk["level"]] = x
                continue
            identity = x if dn is None else dn(x)
            if (self.fused_bneck and len(cs) == 3 and cs[1].stride == 1 and cs[1].groups == 1 and cs[1].ks == 3 and
                    cs[1].cin in (64, 128) and cs[1].cout == cs[1].cin and cs[2].cin == cs[1].cin and cs[2].cout % 128 == 0 and
                    cs[2].cout <= 512):
                # conv2 (3x3) + conv3 (1x1) + identity + ReLU in one kernel: the 3x3's output never leaves the SM
                out = z_in if z_in is not None else cs[0](x, relu=True)
                px = out.shape[0] * out.shape[1] * out.shape[2]
                engine.STATS["conv_flops"] += 2 * px * cs[1].cin * (9 * cs[1].cin + cs[2].cout)
                # GEMM3: the NEXT block's conv1 (1x1, C2 -> 64 / 128, + ReLU) from this block's output while it is still in
                # shared memory -- removes a full read of the widest tensor of the level
                nxt = blocks[bi + 1]["convs"] if bi + 1 < len(blocks) else None
                if (self.fused_conv1 and cs[1].cin == 64 and nxt is not None and len(nxt) == 3 and nxt[0].ks == 1 and nxt[0].stride == 1 and
                        nxt[0].groups == 1 and nxt[0].cin == cs[2].cout and nxt[0].cout in (64, 128) and nxt[0].b is not None):
                    engine.STATS["conv_flops"] += 2 * px * nxt[0].cin * nxt[0].cout
                    x, z_next = engine.bottleneck_tail(out, cs[1].w, cs[1].b, cs[2].w, cs[2].b, identity, relu=True,
                                                       w_next=nxt[0].w, b_next=nxt[0].b)
                else:
                    x = engine.bottleneck_tail(out, cs[1].w, cs[1].b, cs[2].w, cs[2].b, identity, relu=True)
            else:
                out = z_in if z_in is not None else x
                for c in (cs[1:-1] if z_in is not None else cs[:-1]):
                    out = c(out, relu=True)
                x = cs[-1](out, relu=True, residual=identity)
            if blk["last"]:
                outs[blk["level"]] = x
        return self._fpn(outs[3], outs[4], outs[5])

    def _fpn(self, c3, c4, c5):
        """FPN (odtk/backbones/fpn.py:45-61) on the three backbone taps."""
        P = self._packed
        p5 = P["lateral5"](c5)
        p4 = P["lateral4"](c4, upsample=p5)
        p3 = P["lateral3"](c3, upsample=p4)
        if self.merged_heads and c5.shape[3] % 64 == 0:
            # pyramid atlas: the five levels stacked in ONE zero-separated NHWC buffer, so that every head-tower layer is
            # one launch over all levels instead of five
            n = c5.shape[0]
            h6, w6 = (c5.shape[1] - 1) // 2 + 1, (c5.shape[2] - 1) // 2 + 1
            sizes = ((p3.shape[1], p3.shape[2]), (p4.shape[1], p4.shape[2]), (p5.shape[1], p5.shape[2]), (h6, w6),
                     ((h6 - 1) // 2 + 1, (w6 - 1) // 2 + 1))
            at = self._atlas_for(n, sizes, c5.device)
            v = [engine.AtlasView(at["buf"][0], r, h, w) for r, (h, w) in zip(at["rows"], sizes)]
            P["smooth3"](p3, out=v[0])
            P["smooth4"](p4, out=v[1])
            P["smooth5"](p5, out=v[2])
            P["pyramid6"](c5, out=v[3])
            P["pyramid7"](v[3], in_relu=True, out=v[4])
            return v
        p6 = P["pyramid6"](c5)
        p7 = P["pyramid7"](p6, in_relu=True)
        return [P["smooth3"](p3), P["smooth4"](p4), P["smooth5"](p5), p6, p7]

    def _atlas_for(self, n, sizes, device):
        """Geometry + persistent, zero-initialised buffers of the pyramid atlas for one batch size / level geometry.
        Level l occupies rows [rows[l], rows[l] + h_l) x columns [0, w_l); level starts are multiples of 8 (the tile
        height) with at least one gap row in between; gap rows / columns are never written and stay zero (they ARE the
        convolutions' zero padding between levels).  tab: every 8 x 16 tile of one image as (row0, col0, row_limit,
        col_limit)."""
        key = (n, sizes, str(device))
        at = self._atlas.get(key)
        if at is None:
            rows, r = [], 0
            for (h, w) in sizes:
                rows.append(r)
                r = (r + h + 1 + 7) // 8 * 8
            ha, wa = rows[-1] + sizes[-1][0], max(w for _, w in sizes)
            tiles = [(r0 + 8 * ty, 16 * tx, r0 + h, w) for (h, w), r0 in zip(sizes, rows)
                     for ty in range((h + 7) // 8) for tx in range((w + 15) // 16)]
            at = {"rows": rows, "ha": ha, "wa": wa, "px": n * sum(h * w for h, w in sizes),
                  "tab": torch.tensor(tiles, dtype=torch.int32, device=device),
                  "buf": [torch.zeros((n, ha, wa, 256), dtype=torch.float16, device=device) for _ in range(5)]}
            self._atlas[key] = at
        return at

    def _heads(self, features, sigmoid=True, sinks=None):
        """Class / box heads on the five levels (odtk/model.py:134-135).  The ten conv chains are
        independent, so each runs on its own CUDA stream (fork after the FPN, join before decode): the
        small levels' launches (P5-P7 use 1-32 CTAs) overlap each other and the tail of the big
        ones; inside a CUDA graph the chains become parallel branches."""
        P = self._packed
        nl = len(features)
        cls_heads, box_heads = [None] * nl, [None] * nl
        cls_mode = engine.OUT_NCHW_F32_SIGMOID if sigmoid else engine.OUT_NCHW_F32
        if isinstance(features[0], engine.AtlasView):
            return self._heads_atlas(features, cls_mode, sinks)

        def chain(head, t, final_mode, sink=None):
            for conv in P[head][:-1]:
                t = conv(t, relu=True)
            if sink is not None:      # class scores go straight to the decode workspace (no dense map)
                return P[head][-1](t, out_mode=engine.OUT_CANDIDATES, sink=sink)
            return P[head][-1](t, out_mode=final_mode)

        if not self.parallel_heads:
            for i, t in enumerate(features):
                cls_heads[i] = chain("cls_head", t, cls_mode, sinks[i] if sinks is not None else None)
                box_heads[i] = chain("box_head", t, engine.OUT_NCHW_F32)
            return cls_heads, box_heads
        main = torch.cuda.current_stream()
        if self._head_streams is None or len(self._head_streams) < 2 * nl:
            self._head_streams = [torch.cuda.Stream(device=self.device) for _ in range(2 * nl)]
        fork = torch.cuda.Event()
        fork.record(main)
        joins = []
        for i, t in enumerate(features):
            for j, (head, mode) in enumerate((("cls_head", cls_mode), ("box_head", engine.OUT_NCHW_F32))):
                sink = sinks[i] if (sinks is not None and j == 0) else None
                if i == 0 and j == 0:          # the biggest chain stays on the main stream
                    cls_heads[0] = chain(head, t, mode, sink)
                    continue
                st = self._head_streams[2 * i + j]
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    out = chain(head, t, mode, sink)
                    ev = torch.cuda.Event()
                    ev.record(st)
                joins.append(ev)
                (cls_heads if j == 0 else box_heads)[i] = out
        for ev in joins:
            main.wait_event(ev)
        return cls_heads, box_heads

    def _heads_atlas(self, views, cls_mode, sinks):
        """Heads over the pyramid atlas: each of the 4 + 4 tower layers is ONE launch over all five levels (a tile table
        lists the 8 x 16 tiles of every level; weights are shared between levels anyway, odtk/model.py:134-135); the
        two final layers run per level on views of the last tower output (their outputs are per-level tensors).  The class
        and the box tower are independent: two streams / two graph branches."""
        P = self._packed
        atlas0 = views[0].atlas
        sizes = tuple((v.h, v.w) for v in views)
        at = self._atlas_for(atlas0.shape[0], sizes, atlas0.device)
        nl = len(views)
        cls_heads, box_heads = [None] * nl, [None] * nl

        def tower(head, bufs, final_mode, use_sinks):
            t = atlas0
            for i, conv in enumerate(P[head][:-1]):
                nxt = bufs[i & 1]
                conv(t, relu=True, out=nxt, tile_tab=at["tab"], valid_px=at["px"])
                t = nxt
            outs = []
            for l, v in enumerate(views):
                lv = engine.AtlasView(t, at["rows"][l], v.h, v.w)
                if use_sinks is not None:
                    outs.append(P[head][-1](lv, out_mode=engine.OUT_CANDIDATES, sink=use_sinks[l]))
                else:
                    outs.append(P[head][-1](lv, out_mode=final_mode))
            return outs

        if not self.parallel_heads:
            cls_heads = tower("cls_head", at["buf"][1:3], cls_mode, sinks)
            box_heads = tower("box_head", at["buf"][3:5], engine.OUT_NCHW_F32, None)
            return cls_heads, box_heads
        main = torch.cuda.current_stream()
        if self._head_streams is None or len(self._head_streams) < 1:
            self._head_streams = [torch.cuda.Stream(device=self.device)]
        side = self._head_streams[0]
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            box_heads = tower("box_head", at["buf"][3:5], engine.OUT_NCHW_F32, None)
            join = torch.cuda.Event()
            join.record(side)
        cls_heads = tower("cls_head", at["buf"][1:3], cls_mode, sinks)
        main.wait_event(join)
        return cls_heads, box_heads

    @staticmethod
    def _to_nhwc_half(x):
        if x.dim() != 4:
            raise ValueError("expected a [B, 3, H, W] batch")
        return x.permute(0, 2, 3, 1).contiguous().to(torch.float16)   # zero-copy for channels_last fp16

    def forward_heads_u8(self, images, sigmoid=True):
        """Input side of `odtk infer` fused in: `images` is a uint8 HWC batch
        [B, H, W, 3]; normalisation, stride padding (odtk/data.py:113-123) and the stem's zero border are one
        kernel that writes the stem's input buffer directly."""
        if self._packed is None:
            raise RuntimeError("call .cuda() after loading weights: there is no CPU path")
        P = self._packed
        xp, hs, ws = engine.preprocess_u8(images, self.stride)
        return self._heads(self._features(self._stem(padded=(xp, hs, ws))), sigmoid), (hs, ws)

    def forward_heads(self, x, sigmoid=True):
        """The `exporting=True` view of the reference (odtk/model.py:142-144): per-level
        (sigmoid) class maps [B, A*C, H, W] and box maps [B, A*4|6, H, W], fp32 NCHW."""
        if self._packed is None:
            raise RuntimeError("call .cuda() after loading weights: there is no CPU path")
        return self._heads(self._features(self._stem(self._to_nhwc_half(x))), sigmoid)

    def _level_anchors(self, widths, width):
        strides, anchors = [], []
        for w in widths:
            stride = width // w                            # width only (odtk/model.py:155)
            if stride not in self.anchors:
                self.anchors[stride] = (box.generate_anchors_rotated(stride, self.ratios, self.scales, self.angles)
                                        if self.rotated_bbox else box.generate_anchors(stride, self.ratios, self.scales))
            a = self.anchors[stride][0] if self.rotated_bbox else self.anchors[stride]
            strides.append(stride)
            anchors.append(a.reshape(-1).tolist())
        return strides, anchors

    def _forward_fused(self, x):
        """Inference branch with the class head's last convolution appending its above-threshold scores
        directly to the decode workspace (odtk_decode_fused_begin/_finish): the dense [B, A*C, H, W] score maps
        of odtk/model.py:140 are never written or re-read.  Same detections as the dense route."""
        if self._packed is None:
            raise RuntimeError("call .cuda() after loading weights: there is no CPU path")
        if x.dtype == torch.uint8:
            xp, hs, width = engine.preprocess_u8(x, self.stride)
            features = self._features(self._stem(padded=(xp, hs, width)))
        else:
            width = x.shape[-1]
            features = self._features(self._stem(self._to_nhwc_half(x)))
        sizes = tuple((f.shape[1], f.shape[2]) for f in features)
        batch = features[0].shape[0]
        key = (batch, sizes, width, self.threshold, self.top_n, self.rotated_bbox)
        fd = self._fused.get(key)
        if fd is None:
            strides, anchors = self._level_anchors([s[1] for s in sizes], width)
            num_anchors = len(anchors[0]) // 4
            fd = self._fused[key] = _C.FusedDecode(batch, sizes, num_anchors, self.classes, anchors, strides,
                                                   self.threshold, self.top_n, self.rotated_bbox, features[0].device)
        sinks = fd.begin()
        _, box_heads = self._heads(features, True, sinks)
        return self._nms(fd.finish(box_heads))

    def _nms(self, decoded):
        """box.nms / nms_rotated (odtk/model.py:162-165) + the packed [B, D, 2 + nbox] rows (self.last_packed) and, when a
        PeerGather is attached, the exchange with the other ranks -- all in the one NMS launch."""
        b = decoded[0].shape[0]
        key = (b, self.detections, self.rotated_bbox, decoded[0].device)
        packed = self._packed_out.get(key)
        if packed is None:
            packed = self._packed_out[key] = torch.empty((b, self.detections, 2 + (6 if self.rotated_bbox else 4)),
                                                         dtype=torch.float32, device=decoded[0].device)
        self.last_packed = packed
        return tuple(_C.nms(*decoded, self.nms, self.detections, self.rotated_bbox, packed=packed, gather=self.gather))

    def attach_gather(self, gather):
        """Image-wise sharded inference: every forward also delivers this rank's detections to all ranks
        (peer.PeerGather.gathered()).  Attach BEFORE the CUDA graph of a shape is captured."""
        self.gather = gather
        if getattr(self, "_graphs", None):
            self._graphs = {}
        return self

    # ---- training-side losses (reference odtk/model.py:167-210) ---------------------------------------------------
    def _extract_targets(self, targets, stride, size):
        """Per level: class-index targets [B, A, H, W] int32, box targets [B, A, nbox, H, W], depth [B, A, 1, H, W].
        Axis-aligned: the whole batch in one launch (odtk_snap_to_anchors); rotated: per image, polygon IoUs on
        odtk_iou (odtk/model.py:167-184)."""
        if stride not in self.anchors:
            self.anchors[stride] = (box.generate_anchors_rotated(stride, self.ratios, self.scales, self.angles)
                                    if self.rotated_bbox else box.generate_anchors(stride, self.ratios, self.scales))
        anchors = self.anchors[stride]
        h, w = int(size[0]), int(size[1])
        if not self.rotated_bbox:
            _, box_target, depth, cls_index = box.snap_to_anchors_batch(targets, (h, w), stride, anchors, self.classes,
                                                                        self.anchor_ious, dense=False)
            return cls_index, box_target, depth
        bts, dps = [], []
        for target in targets:
            target = target[target[:, -1] > -1]
            _, bt, dp = box.snap_to_anchors_rotated(target, [w * stride, h * stride], stride, anchors, self.classes,
                                                    targets.device, self.anchor_ious)
            bts.append(bt)
            dps.append(dp)
        box_target, depth = torch.stack(bts), torch.stack(dps)
        d = depth[:, :, 0]
        cls_index = torch.where(d > 0, d - 1, torch.where(d == 0, torch.full_like(d, -1), torch.full_like(d, -2))).int()
        return cls_index.contiguous(), box_target, depth

    def _compute_loss(self, x, cls_heads, box_heads, targets, with_grad=False):
        """odtk/model.py:186-210: target assignment per level, then ONE fused launch for the focal + smooth-L1 losses of
        all levels, the masks, the foreground counts and the normalisation (loss.retina_loss -> odtk_retina_loss)."""
        from . import loss as loss_mod
        width = x.shape[-1] if torch.is_tensor(x) else int(x)
        cls_idx, box_tgt = [], []
        for cls_head in cls_heads:
            size = cls_head.shape[-2:]
            stride = width // cls_head.shape[-1]
            ci, bt, _ = self._extract_targets(targets, stride, size)
            cls_idx.append(ci)
            box_tgt.append(bt)
        return loss_mod.retina_loss(cls_heads, box_heads, cls_idx, box_tgt, self.classes, with_grad=with_grad)

    def forward(self, x, rotated_bbox=None):
        if self.training:
            x, targets = x
            cls_heads, box_heads = self.forward_heads(x, sigmoid=False)
            return self._compute_loss(x, cls_heads, box_heads, targets.float().to(cls_heads[0].device))
        if self.fused_candidates and not self.exporting:
            return self._forward_fused(x)
        if x.dtype == torch.uint8:                       # raw HWC images: fused input side
            (cls_heads, box_heads), (_, width) = self.forward_heads_u8(x)
        else:
            cls_heads, box_heads = self.forward_heads(x)
            width = x.shape[-1]
        if self.exporting:
            self.strides = [width // c.shape[-1] for c in cls_heads]
            return cls_heads, box_heads
        strides, anchors = self._level_anchors([c.shape[-1] for c in cls_heads], width)
        decoded = _C.decode_levels(cls_heads, box_heads, anchors, strides, self.threshold, self.top_n, self.rotated_bbox)
        return self._nms(decoded)

    # ---- CUDA graph replay ---------------------------------------------------------------------
    def enable_cuda_graph(self, enabled=True):
        """Capture the whole forward (≈125 kernel launches, no host sync anywhere on the path) into a
        CUDA graph per input shape and replay it: removes the per-launch host cost, which dominates at
        small batch.  The reference has no equivalent (its decode/nms block the host B*5+B times)."""
        self._graphs = {} if enabled else None
        return self

    def _forward_graphed(self, x, rotated_bbox=None, static_input=False):
        """static_input=True: the caller promises to reuse THIS tensor's storage for every call (e.g. the
        device side of a double-buffered upload): the graph reads it in place, no staging copy."""
        key = (tuple(x.shape), x.dtype, bool(x.is_contiguous(memory_format=torch.channels_last)),
               x.data_ptr() if static_input else 0)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = x if static_input else x.clone()
            for _ in range(2):                      # warm-up: lazy attribute set-up, allocator, anchors
                self.forward(static_in, rotated_bbox)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.forward(static_in, rotated_bbox)
            entry = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        if not static_input:
            static_in.copy_(x, non_blocking=True)
        graph.replay()
        if self.gather is not None:
            self.gather.steps += 1
        return static_out

    def __call__(self, x, rotated_bbox=None, static_input=False):
        if getattr(self, "_graphs", None) is not None and not self.exporting and not self.training:
            return self._forward_graphed(x, rotated_bbox, static_input)
        return self.forward(x, rotated_bbox)
