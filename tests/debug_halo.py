"""Manual diagnostic (not collected by pytest): 3x3 convolutions through the halo mode (patch loaded once,
nine shifted shared-memory views) against torch, error summary per shape.  Run with
ODTK_CONV_HALO=0/1 and ODTK_CONV_HALO_BOFF=0/1 to compare."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from retinanet_examples_b200 import engine

DEV = "cuda:0"


def run(name, n, h, w, cin, cout, mode=0, bias_op=False, residual=False):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn((n, h, w, cin), generator=g)).half()
    wt = (torch.randn((cout, cin, 3, 3), generator=g) * 0.05).half()
    b = torch.randn(cout, generator=g)
    res = torch.randn((n, h, w, cout), generator=g).half() if residual else None
    bop = engine.pack_bias(b.to(DEV)) if bias_op else None
    y = engine.conv2d(x.to(DEV), engine.pack_weight(wt.float()).to(DEV), b.to(DEV), cout, 3, out_mode=mode, bias_op=bop,
                      residual=res.to(DEV) if residual else None)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1)
    if residual:
        ref = ref + res.float().permute(0, 3, 1, 2)
    got = y.float().cpu() if mode else y.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    bad = err > 0.02 + 0.01 * ref.abs()
    print("%-34s max err %.4g  ref max %.4g  bad %d / %d" % (name, err.max().item(), ref.abs().max().item(),
                                                             int(bad.sum()), err.numel()), flush=True)
    if bad.any():
        badh = bad.any(dim=0).any(dim=0).any(dim=1).nonzero().flatten().tolist()
        badw = bad.any(dim=0).any(dim=0).any(dim=0).nonzero().flatten().tolist()
        print("   bad h", badh[:24], "bad w", badw[:24], flush=True)


print("HALO=%s BOFF=%s" % (os.environ.get("ODTK_CONV_HALO", "1"), os.environ.get("ODTK_CONV_HALO_BOFF", "0")))
run("64->64 16x8 one tile", 1, 16, 8, 64, 64)
run("64->64 32x24", 1, 32, 24, 64, 64)
run("64->64 40x64 (transposed)", 1, 40, 64, 64, 64)
run("128->128 33x47 n2", 2, 33, 47, 128, 128)
run("256->256 48x64 n4 bias_op (pairs)", 4, 48, 64, 256, 256, bias_op=True)
run("256->256 100x160 n2 bias_op", 2, 100, 160, 256, 256, bias_op=True)
run("256->36 nchw 48x40", 1, 48, 40, 256, 36, mode=1)
run("256->720 nchw 32x40 n2", 2, 32, 40, 256, 720, mode=1, bias_op=True)
run("64->64 +res 32x32", 1, 32, 32, 64, 64, residual=True)
print("done")
