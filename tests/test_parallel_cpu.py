"""world_size-2 gloo tests (CPU) for the N>1 path: shard bounds + the final gather, with a stand-in vocoder."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from e2e_tts_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 2, 5, 16, 255, 256):
        for w in (1, 2, 3, 4, 8):
            b = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


class FakeVocoder:
    """Deterministic stand-in with the HifiGan contract ([b,80,T] -> [b,1,hop*T]); utterances independent."""
    hop = 4

    def __call__(self, mel):
        return mel.sum(1, keepdim=True).repeat_interleave(self.hop, dim=2) + 1.0


class FakePcmVocoder(FakeVocoder):
    """Stand-in for HifiGan.forward_pcm16: int16 [b, hop*T] (exercises the byte-view gather, NCCL has no int16)."""

    def __call__(self, mel):
        return (super().__call__(mel).squeeze(1) * 100).to(torch.int16)


def _worker(rank, world, port, B, T, q, pcm=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        mel = torch.randn(B, 80, T, generator=g)
        voc = FakePcmVocoder() if pcm else FakeVocoder()
        out = parallel.synthesize_sharded(voc, mel, dst=0)
        if rank == 0:
            want = voc(mel) if pcm else voc(mel).squeeze(1)
            q.put(bool(out is not None and out.shape == want.shape and torch.equal(out, want)))
        else:
            q.put(out is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,pcm", [(6, False), (5, False), (1, False), (5, True), (1, True)])
def test_sharded_synthesis_gathers_on_rank0_gloo_world2(B, pcm):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 7, q, pcm)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)
