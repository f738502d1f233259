"""World-size-2 CPU (gloo) test of the sharding + gather plumbing used for N > 1 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from encodec_b200 import dist as ebdist


class FakeModel:
    """Deterministic stand-in with EncodecModel.forward's return signature (CPU tensors)."""

    def __call__(self, x):
        b, c, t = x.shape
        audio = x * 2.0 + 1.0
        codes = (x[:, 0, :: max(1, t // 5)][:, :5].abs() * 1000).long().unsqueeze(1).repeat(1, 3, 1) % 1024
        return audio, codes, torch.zeros(3, 1), torch.zeros(3, 1)


def _worker(rank, world, port, n_clips, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_clips, 2, 50, generator=g)
        audio, codes = ebdist.forward_sharded(FakeModel(), x, dst=0)
        if rank == 0:
            ref_audio, ref_codes, _, _ = FakeModel()(x)
            assert torch.equal(audio, ref_audio)
            assert torch.equal(codes, ref_codes) and codes.dtype == torch.int64
            out.put("ok")
        else:
            assert audio is None and codes is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5, 1])
def test_forward_sharded_gathers_in_order(n_clips):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
