"""World-size-2 CPU (gloo) test of the sharding + gather plumbing used for N > 1 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from encodec_b200 import dist as ebdist


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class FakeModel:
    """Deterministic stand-in with EncodecModel.forward's return signature (CPU tensors) and the geometry attributes
    dist.output_shapes reads (so that ranks with an empty shard can size their part of the collective)."""
    quantizer = _NS(bins=1024)
    encoder = _NS(hop_length=10)

    def _segments(self, length):
        return [(0, length)], length

    def _n_q(self):
        return 3

    def __call__(self, x):
        b, c, t = x.shape
        audio = x * 2.0 + 1.0
        codes = (x[:, 0, :: max(1, t // 5)][:, :5].abs() * 1000).long().unsqueeze(1).repeat(1, 3, 1) % 1024
        return audio, codes, torch.zeros(3, 1), torch.zeros(3, 1)


def _worker(rank, world, port, n_clips, out, dst=0, micro_batch=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_clips, 2, 50, generator=g)
        audio, codes = ebdist.forward_sharded(FakeModel(), x, dst=dst, micro_batch=micro_batch)
        if rank == dst:
            ref_audio, ref_codes, _, _ = FakeModel()(x)
            assert torch.equal(audio, ref_audio)
            assert torch.equal(codes, ref_codes) and codes.dtype == torch.int64
            out.put("ok")
        else:
            assert audio is None and codes is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips,dst,micro_batch", [(8, 0, None), (5, 0, None), (1, 0, None), (1, 1, None), (5, 1, 2), (7, 0, 3)])
def test_forward_sharded_gathers_in_order(n_clips, dst, micro_batch):
    """Ragged shards, an EMPTY shard on the destination rank (n_clips < world, dst = 1) and micro-batched gathers."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, out, dst, micro_batch)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
