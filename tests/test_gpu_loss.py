"""GPU parity tests of the fused focal-loss kernel (loss.cu) through the C ABI against the golden
vectors of the reference's own FocalLoss + autograd (tests/golden/focal.npz), the CPU oracle, and
torch autograd semantics.  Tolerance: fp32 transcendental maths, rtol 2e-5 / atol 1e-7 (stated)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from retinanet_examples_b200 import loss as loss_mod

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_focal_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "focal.npz"))
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g["t"]).to(DEV)
    elem = loss_mod.FocalLoss()(x.detach(), t)
    np.testing.assert_allclose(elem.cpu().numpy(), g["loss"], rtol=2e-5, atol=1e-7)
    total = loss_mod.focal_loss_sum(x, t)
    total.backward()
    np.testing.assert_allclose(float(total), g["loss"].astype(np.float64).sum(), rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["grad"], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("shape,gamma", [((2, 9, 80, 25, 40), 2.0), ((1, 9, 3, 7, 10), 2.0), ((3, 1001), 1.5), ((5, 77), 0.0)])
def test_focal_matches_oracle_with_mask(shape, gamma):
    rng = np.random.default_rng(len(shape) * 100 + int(gamma * 10))
    x = rng.normal(-2, 3, size=shape).astype(np.float32)
    t = (rng.uniform(size=shape) < 0.02).astype(np.float32)
    m = (rng.uniform(size=shape) < 0.9).astype(np.float32)
    tot, lo, gr = oracle.focal_loss(x, t, m, 0.25, gamma, 1.0)
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    total = loss_mod.focal_loss_sum(xt, torch.from_numpy(t).to(DEV), torch.from_numpy(m).to(DEV), gamma=gamma)
    (total * 0.5).backward()
    np.testing.assert_allclose(float(total), tot, rtol=2e-5)
    np.testing.assert_allclose(xt.grad.cpu().numpy().reshape(-1), 0.5 * gr, rtol=5e-5, atol=1e-7)


def test_focal_class_index_targets_equal_dense_one_hot():
    rng = np.random.default_rng(8)
    groups, C, hw = 2 * 9, 80, 13 * 20
    x = rng.normal(-3, 2, size=(groups, C, hw)).astype(np.float32)
    idx = rng.integers(-2, C, size=(groups, hw)).astype(np.int32)        # -2 ignored, -1 background
    t = np.zeros_like(x)
    gi, pi = np.nonzero(idx >= 0)
    t[gi, idx[gi, pi], pi] = 1.0
    m = np.broadcast_to((idx != -2)[:, None, :], x.shape).astype(np.float32)
    xa = torch.from_numpy(x).to(DEV).requires_grad_(True)
    xb = torch.from_numpy(x).to(DEV).requires_grad_(True)
    la = loss_mod.focal_loss_sum(xa, torch.from_numpy(t).to(DEV), torch.from_numpy(m).to(DEV))
    lb = loss_mod.focal_loss_sum(xb, cls_index=torch.from_numpy(idx).to(DEV))
    la.backward(); lb.backward()
    assert float(la) == float(lb)
    assert torch.equal(xa.grad, xb.grad)
    tot, _, _ = oracle.focal_loss(x, t, m)
    np.testing.assert_allclose(float(lb), tot, rtol=2e-5)


def test_smooth_l1_matches_reference_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "l1_preproc.npz"))
    p = torch.from_numpy(g["p"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g["t"]).to(DEV)
    elem = loss_mod.SmoothL1Loss(beta=0.11)(p.detach(), t)
    np.testing.assert_allclose(elem.cpu().numpy(), g["loss"], rtol=1e-5, atol=1e-7)
    total = loss_mod.smooth_l1_loss_sum(p, t)
    total.backward()
    np.testing.assert_allclose(p.grad.cpu().numpy(), g["grad"], rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(4)
    m = (rng.uniform(size=g["p"].shape) < 0.5).astype(np.float32)
    tot, _, gr = oracle.smooth_l1(g["p"], g["t"], m)
    p2 = torch.from_numpy(g["p"]).to(DEV).requires_grad_(True)
    l2 = loss_mod.smooth_l1_loss_sum(p2, t, torch.from_numpy(m).to(DEV))
    l2.backward()
    np.testing.assert_allclose(float(l2), tot, rtol=1e-5)
    np.testing.assert_allclose(p2.grad.cpu().numpy().reshape(-1), gr, rtol=1e-5, atol=1e-6)


def test_compute_loss_one_launch_matches_reference_golden(golden_dir):
    """Model._compute_loss (odtk/model.py:186-210) of the UNMODIFIED reference on seeded heads / targets
    (tests/golden/compute_loss.npz, oracle/gen_golden_loss.py) == target assignment + ONE fused launch here:
    both normalised losses (rtol 2e-5) and the gradients autograd produced for every head tensor."""
    import os
    from retinanet_examples_b200.model import Model
    g = np.load(os.path.join(golden_dir, "compute_loss.npz"))
    classes = int(g["classes"])
    model = Model("ResNet18FPN", classes=classes)
    cls_heads = [torch.from_numpy(g["cls%d" % i]).to(DEV) for i in range(5)]
    box_heads = [torch.from_numpy(g["box%d" % i]).to(DEV) for i in range(5)]
    targets = torch.from_numpy(g["targets"]).to(DEV)
    cls_loss, box_loss, cg, bg = model._compute_loss(int(g["width"]), cls_heads, box_heads, targets, with_grad=True)
    np.testing.assert_allclose(float(cls_loss), float(g["cls_loss"]), rtol=2e-5)
    np.testing.assert_allclose(float(box_loss), float(g["box_loss"]), rtol=2e-5)
    for i in range(5):
        np.testing.assert_allclose(cg[i].cpu().numpy(), g["cls_grad%d" % i], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(bg[i].cpu().numpy(), g["box_grad%d" % i], rtol=2e-4, atol=1e-7)
    # deterministic: bit-identical on a second run
    c2, b2 = model._compute_loss(int(g["width"]), cls_heads, box_heads, targets)
    assert float(c2) == float(cls_loss) and float(b2) == float(box_loss)


def test_training_mode_forward_returns_losses():
    """model.train(); model([images, targets]) -> (cls_loss, box_loss), as odtk/model.py:130-138."""
    from retinanet_examples_b200.model import Model, make_state_dict
    classes = 4
    model = Model("ResNet18FPN", classes=classes).load_state_dict(make_state_dict("ResNet18FPN", classes, 9, False, seed=2)).cuda(0)
    x = torch.randn((2, 3, 128, 256), generator=torch.Generator().manual_seed(1)).to(DEV)
    t = torch.tensor([[[20., 30., 60., 40., 1.], [100., 10., 50., 90., 3.]], [[5., 5., 100., 100., 0.], [-1., -1., -1., -1., -1.]]]).to(DEV)
    cls_loss, box_loss = model.train()([x, t])
    assert torch.isfinite(cls_loss) and torch.isfinite(box_loss) and float(cls_loss) > 0 and float(box_loss) > 0
    model.eval()
    s, b, c = model(x)
    assert s.shape == (2, 100)
