"""CPU test of the N>1 path: image-wise sharding + the single packed all-gather, world_size 2, gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from retinanet_examples_b200 import infer as infer_mod


class _FakeModel:
    """Deterministic stand-in for Model.forward: detections are a function of the image content."""

    def __call__(self, data, rotated_bbox=None):
        b = data.shape[0]
        key = data.reshape(b, -1).sum(dim=1)
        scores = key[:, None] + torch.arange(4)[None, :].float()
        boxes = scores[..., None].repeat(1, 1, 4)
        classes = torch.floor(scores) % 3
        return scores, boxes, classes


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    images = torch.arange(8 * 3 * 4 * 4, dtype=torch.float32).reshape(8, 3, 4, 4)
    mine = list(infer_mod.shard_batch(8, world, rank))
    batches = [images[mine[:2]], images[mine[2:]]]
    s, b, c = infer_mod.infer(_FakeModel(), batches, world=world)
    ret[rank] = (s.clone(), b.clone(), c.clone())
    dist.destroy_process_group()


def test_sharded_infer_gathers_all_ranks_in_order():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    images = torch.arange(8 * 3 * 4 * 4, dtype=torch.float32).reshape(8, 3, 4, 4)
    es, eb, ec = _FakeModel()(images)
    for rank in range(2):
        s, b, c = ret[rank]
        assert torch.equal(s, es) and torch.equal(b, eb) and torch.equal(c, ec)


def test_pack_roundtrip_and_shard_errors():
    s, b, c = torch.rand(3, 5), torch.rand(3, 5, 6), torch.rand(3, 5)
    p = infer_mod.pack_detections(s, b, c)
    assert p.shape == (3, 5, 8)
    s2, b2, c2 = infer_mod.unpack_detections(p)
    assert torch.equal(s, s2) and torch.equal(b, b2) and torch.equal(c, c2)
    try:
        infer_mod.shard_batch(10, 4, 0)
        assert False
    except RuntimeError:
        pass
