"""Multi-GPU synthesis: utterances are independent (the generator has no cross-batch op — generator.py:37-53 has
no BatchNorm), so a batch shards contiguously across ranks with NO data-path collective; the only exchange is the
final gather of waveforms (SURVEY.md §8 e).  One process per GPU, torch.distributed (NCCL over NVLink on the GPU
box; the same code runs under gloo on CPU for the host-logic tests, with any callable standing in for the
vocoder).  The reference itself is single-process (no DP/DDP anywhere): this is new surface."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n utterances owned by `rank`; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def synthesize_shard(vocoder: Callable[[torch.Tensor], torch.Tensor], mel_local: torch.Tensor) -> torch.Tensor:
    """Local part: mel [b, 80, T] -> wav [b, hop*T] (the `.squeeze(1)` of utils.py:144)."""
    if mel_local.shape[0] == 0:  # this rank owns no utterance (B < world size): one dummy call fixes shape and dtype
        with torch.no_grad():
            probe = vocoder(mel_local.new_zeros((1,) + tuple(mel_local.shape[1:])))
        probe = probe.squeeze(1) if probe.dim() == 3 else probe
        return probe.new_zeros((0, probe.shape[-1]))
    with torch.no_grad():
        wav = vocoder(mel_local)
    return wav.squeeze(1) if wav.dim() == 3 else wav  # [b,1,S] fp32 (forward) or [b,S] int16 (forward_pcm16)


def gather_waveforms(wav_local: torch.Tensor, total: int, dst: int = 0,
                     group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Gather the per-rank [b_r, S] waveforms (b_r from shard_bounds(total, world, r)) on rank `dst`.
    Equal shards use one all_gather_into_tensor-style collective into a preallocated [total, S] buffer;
    ragged shards are padded to the largest shard and trimmed.  Returns the [total, S] tensor on `dst`, None
    elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    S = wav_local.shape[1]
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    assert wav_local.shape[0] == counts[rank], "local shard does not match shard_bounds"
    bmax = max(counts)
    if bmax == 0:
        return wav_local.new_zeros((0, S)) if rank == dst else None
    send = wav_local
    if send.shape[0] != bmax:
        send = torch.cat([wav_local, wav_local.new_zeros((bmax - wav_local.shape[0], S))], 0)
    send = send.contiguous()
    dtype = send.dtype
    if dtype == torch.int16:  # NCCL has no int16: int16 PCM (HifiGan.forward_pcm16) travels as bytes
        send = send.view(torch.uint8)
    if rank == dst:
        recv: List[torch.Tensor] = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, recv, dst=dst, group=group)
        return torch.cat([recv[r].view(dtype)[:counts[r]] for r in range(world)], 0)
    dist.gather(send, None, dst=dst, group=group)
    return None


def synthesize_sharded(vocoder: Callable[[torch.Tensor], torch.Tensor], mel: torch.Tensor, dst: int = 0,
                       group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Every rank passes the same global mel [B, 80, T] (or at least its own rows valid); rank r synthesises
    rows shard_bounds(B, world, r) and rank `dst` receives all B waveforms [B, hop*T]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(mel.shape[0], world, rank)
    wav = synthesize_shard(vocoder, mel[lo:hi])
    return gather_waveforms(wav, mel.shape[0], dst, group)
