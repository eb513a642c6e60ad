"""Debug helper: MobileNetV2FPN block by block, CUDA path vs the fp16-emulating oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from oracle import model_ref
from retinanet_examples_b200 import engine
from retinanet_examples_b200.model import Model, make_state_dict
bb = "MobileNetV2FPN"
sd = make_state_dict(bb, 5, 9, False, 3)
x = torch.randn((1, 3, 128, 128), generator=torch.Generator().manual_seed(1))
m = Model(bb, classes=5).load_state_dict(sd).cuda()
P = m._packed
f = "backbones.%s.features.features." % bb
q = lambda t: t.half().float()
# oracle intermediates
xo = q(x)
xo = q(F.relu6(model_ref._cb(sd, f + "0.0", f + "0.1", xo, 2, 1, True)))
xg = P["stem"](m._to_nhwc_half(x.cuda()), relu=2)
def cmp(name, g, o):
    c = o.shape[1]
    gg = g.float().cpu().permute(0, 3, 1, 2)
    err = float((gg[:, :c] - o).abs().max()); pad = float(gg[:, c:].abs().max()) if gg.shape[1] > c else 0.0
    print("%-22s max|ref| %.4f err %.5f pad %.5f" % (name, float(o.abs().max()), err, pad), flush=True)
cmp("stem", xg, xo)
cin, idx = 32, 1
for blk, (t, c, n, s) in zip([None] * 0, []):
    pass
from retinanet_examples_b200.model import mobilenet_blocks
for blk, (idx, cin, cout, stride, t) in zip(P["blocks"], mobilenet_blocks(bb)):
    p, k = f + "%d.conv." % idx, 0
    ho, hg = xo, xg
    if t != 1:
        ho = q(F.relu6(model_ref._cb(sd, p + "0.0", p + "0.1", ho, 1, 0, True)))
        hg = blk["expand"](hg, relu=2)
        cmp("b%d expand" % idx, hg, ho)
        k = 1
    ho = q(F.relu6(model_ref._cb(sd, p + "%d.0" % k, p + "%d.1" % k, ho, stride, 1, True, groups=ho.shape[1])))
    hg = engine.depthwise3x3(hg, blk["dw_w"], blk["dw_b"], blk["stride"], act=2)
    cmp("b%d dw s%d" % (idx, stride), hg, ho)
    ho = model_ref._cb(sd, p + "%d" % (k + 1), p + "%d" % (k + 2), ho, 1, 0, True)
    xo_new = q(ho + xo if (stride == 1 and cin == cout) else ho)
    xg = blk["project"](hg, relu=False, residual=xg if blk["res"] else None)
    xo = xo_new
    cmp("b%d out res=%s" % (idx, blk["res"]), xg, xo)
