"""Host-to-host synthesis loop for serving: pinned host mel batches in, pinned host waveforms out, with the
host->device copy of batch i+1 and the device->host copy of batch i-1 overlapped with the synthesis of batch i
(copy-in stream, the caller's compute stream, copy-out stream; double-buffered device inputs).

The reference's loop is strictly serial per batch (e2e_tts/src/api/utils.py:130-149: acoustic model -> vocoder ->
`.detach().cpu().numpy()`); this is new surface around the same `HifiGan.forward` / `forward_pcm16` call."""
from __future__ import annotations

import inspect
from typing import Callable, List, Optional

import torch


def _accepts_out(fn) -> bool:
    """Does the callable take an `out=` keyword (HifiGan.forward / forward_pcm16 do; a plain callable may not)?"""
    target = fn.forward if isinstance(fn, torch.nn.Module) else fn
    try:
        params = inspect.signature(target).parameters
    except (TypeError, ValueError):
        return False
    return "out" in params or any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())


class HostPipeline:
    """pipe = HostPipeline(vocoder)                       # or HostPipeline(vocoder.forward_pcm16)
    for i, mel in enumerate(batches):                     # mel: pinned host [B, 80, T] fp32
        pipe.submit(mel, wav_host[i % 2])                 # wav_host[*]: pinned host [B, hop*T] (fp32 or int16)
    pipe.drain()                                          # all submitted waveforms are in their host buffers

    submit() never blocks the host on the GPU as long as at most `depth` batches are in flight; reuse of a host output
    buffer is the caller's contract (wait(i) / drain() before reading or overwriting it)."""

    def __init__(self, vocoder: Callable[[torch.Tensor], torch.Tensor], device: Optional[torch.device] = None,
                 depth: int = 2) -> None:
        self.vocoder = vocoder
        owner = getattr(vocoder, "__self__", vocoder)
        if device is None:
            device = next(owner.parameters()).device
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the vocoder on a CUDA device (there is no CPU path)")
        self.depth = int(depth)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._mel_dev: List[Optional[torch.Tensor]] = [None] * self.depth
        self._wav_dev: List[Optional[torch.Tensor]] = [None] * self.depth   # preallocated results (no allocator traffic)
        self._ev_copied: List[Optional[torch.cuda.Event]] = [None] * self.depth
        self._takes_out = _accepts_out(vocoder)
        self._ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        self._ev_comp = [torch.cuda.Event() for _ in range(self.depth)]
        self._ev_out: List[torch.cuda.Event] = []   # copy-out events of batches _base .. _n - 1
        self._base = 0
        self._n = 0

    def submit(self, mel_host: torch.Tensor, wav_host: torch.Tensor) -> int:
        """Enqueue one batch; returns its index for wait()."""
        if mel_host.is_cuda or wav_host.is_cuda:
            raise ValueError("HostPipeline takes host tensors (pin them for asynchronous copies)")
        i, slot = self._n, self._n % self.depth
        compute = torch.cuda.current_stream(self.device)
        buf = self._mel_dev[slot]
        fresh = buf is None or buf.shape != mel_host.shape or buf.dtype != mel_host.dtype
        if fresh:
            buf = torch.empty(mel_host.shape, dtype=mel_host.dtype, device=self.device)
            self._mel_dev[slot] = buf
            if self._wav_dev[slot] is not None:                # the cached result has the old batch's shape: drop it,
                self._wav_dev[slot].record_stream(self.s_out)  # but not before its copy-out has finished reading it
                self._wav_dev[slot] = None
        with torch.cuda.stream(self.s_in):
            if i >= self.depth:
                self.s_in.wait_event(self._ev_comp[slot])      # the forward that read this slot's input buffer is done
            if fresh:
                # a block the caching allocator hands out may have been freed on the compute stream only a moment ago
                # (e.g. a vocoder workspace still in use by the forward in flight): order the copy after that stream
                self.s_in.wait_stream(compute)
            buf.copy_(mel_host, non_blocking=True)
            self._ev_in[slot].record(self.s_in)
        compute.wait_event(self._ev_in[slot])
        if self._ev_copied[slot] is not None:
            compute.wait_event(self._ev_copied[slot])          # the previous result in this slot has left the device
        with torch.no_grad():
            if self._takes_out:
                wav = self.vocoder(buf, out=self._wav_dev[slot])
            else:
                wav = self.vocoder(buf)
        self._wav_dev[slot] = wav if self._takes_out else None
        flat = wav.squeeze(1) if wav.dim() == 3 else wav
        self._ev_comp[slot].record(compute)
        ev_out = torch.cuda.Event()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self._ev_comp[slot])
            wav_host.copy_(flat, non_blocking=True)
            if not self._takes_out:
                flat.record_stream(self.s_out)                 # keep the allocator from reusing it before the copy ends
            ev_out.record(self.s_out)
        self._ev_copied[slot] = ev_out
        self._ev_out.append(ev_out)
        self._n += 1
        # a serving loop that never calls drain() must not accumulate one event per batch: forget the oldest events once
        # they have completed (the copy-out stream is in order, so everything older has completed too)
        while len(self._ev_out) > 4 * self.depth + 32 and self._ev_out[0].query():
            self._ev_out.pop(0)
            self._base += 1
        return i

    def wait(self, index: int) -> None:
        """Blocks the host until batch `index` is in its host buffer."""
        if index >= self._n:
            raise IndexError("batch %d has not been submitted" % index)
        if index >= self._base:                                # (older batches completed before their events were dropped)
            self._ev_out[index - self._base].synchronize()

    def join(self) -> None:
        """Makes the caller's current stream wait for everything submitted so far (no host blocking)."""
        if self._ev_out:
            torch.cuda.current_stream(self.device).wait_event(self._ev_out[-1])
            torch.cuda.current_stream(self.device).wait_stream(self.s_in)

    def drain(self) -> None:
        """Blocks the host until every submitted batch is in its host buffer."""
        if self._ev_out:
            self._ev_out[-1].synchronize()
        self._ev_out.clear()
        self._base = 0
        self._n = 0
