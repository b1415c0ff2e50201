#!/usr/bin/env python
"""Benchmark of the e2e-tts synthesis hot path on B200 (BASELINE.json metric: audio-seconds synthesised per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the vocoder hot path (HifiGan.forward: mel -> waveform) over one batch of synthetic
log-mel-like input with random-init weights of the reference's default HiFi-GAN V1 config.  Workload at every N:
BASELINE.json configs[1] per GPU — 16 utterances x 5 s (T = 431 mel frames -> 110 336 samples each), bf16 tensor-core
operands with fp32 accumulation, bf16 activation+residual stream, fp32 resblock sum.  N > 1 is weak scaling: every rank synthesises its own
16 utterances (no data-path collective) and the step ends with the NCCL gather of all waveforms on rank 0
(SURVEY.md §8 e).  Prints ONE JSON line on rank 0.

--impl reference times the reference's CPU path: the oracle port of generator.py:37-53 in eager torch fp32 on the
host cores (the reference is pure Python and /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR = 22050
HOP = 256
T_FRAMES = 431           # 5 s utterance (110 336 samples = 5.004 s)
B_PER_GPU = 16
FLOP_PER_FRAME = 614105088          # 2*MACs of all 78 convs per mel frame
CONV_TC_FLOP_PER_FRAME = FLOP_PER_FRAME - 114688   # everything but conv_post runs in the tcgen05 kernel
METRIC = "audio_seconds_synthesized_per_second"
UNIT = "audio-s/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the tcgen05 launches of ONE step, from the committed
    ncu capture of this same command (scripts/gpu_final_profile.sh -> profiles/roofline_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_step_tcgen05"]), str(t.get("source", ""))
    except Exception:
        return None, ""


def mel_like(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.active = False      # only samples taken while the timed region runs are kept
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            self.ready.set()
            while not self._stop_evt.is_set():
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # NVML missing: report that, never fake numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)
            self.ready.set()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_oracle_throughput(seconds_budget: float, utterances: int):
    """Oracle port of the reference generator on the host cores: audio-s/s on `utterances` x 5 s, repeated until
    about `seconds_budget` s of CPU time were spent (at least once)."""
    from oracle import hifigan_oracle as ho
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ho.DEFAULT_CONFIG
    sd = ho.make_state_dict(cfg, 1, "strong")
    mel = mel_like(utterances, T_FRAMES, 0)
    with torch.no_grad():
        ho.hifigan_forward(sd, cfg, mel[:1])     # warm-up (thread pool, oneDNN primitives)
        times = []
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            ho.hifigan_forward(sd, cfg, mel)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > seconds_budget:
                break
    best = min(times)
    audio_s = utterances * T_FRAMES * HOP / SR
    return audio_s / best, cores, "%d x 5 s utterances (T=431) per pass, best of %d passes" % (utterances, len(times))


def run_reference(args):
    """The reference arm: CPU oracle port, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    utter = 2
    from oracle import hifigan_oracle as ho
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ho.DEFAULT_CONFIG
    sd = ho.make_state_dict(cfg, 1, "strong")
    mel = mel_like(utter, T_FRAMES, 0)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            ho.hifigan_forward(sd, cfg, mel[:1])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ho.hifigan_forward(sd, cfg, mel)
        dt = time.perf_counter() - t0
    audio_s = utter * T_FRAMES * HOP / SR
    value = audio_s * args.steps / dt
    sample = "each step = %d of the 16 utterances x 5 s (T=431), oracle port of generator.py:37-53, fp32 eager torch" % utter
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 16 utterances x 5 s (T=431), HiFi-GAN V1 default config, random-init weights; "
                               "CPU sample: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import e2e_tts_b200 as pkg
    from e2e_tts_b200 import parallel
    from oracle import hifigan_oracle as ho   # only for the synthetic checkpoint generator and cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = ho.DEFAULT_CONFIG
    voc = pkg.HifiGan(cfg)
    voc.load_state_dict(ho.make_state_dict(cfg, 1, "strong"))
    voc = voc.eval().to(dev)
    B, T = B_PER_GPU, T_FRAMES
    S = HOP * T
    n_in = 4   # rotate inputs; the per-step working set (~1.1 GB of activations) is far larger than the 126 MB L2
    mels_host = [mel_like(B, T, 100 * rank + i).pin_memory() for i in range(n_in)]
    mels_dev = [m.to(dev) for m in mels_host]
    wav_host = torch.empty((B, S), dtype=torch.float32).pin_memory()
    gathered = torch.empty((world * B, S), dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None

    out_dev = torch.empty((B, 1, S), dtype=torch.float32, device=dev)

    def step(i):
        with torch.no_grad():
            wav = voc(mels_dev[i % n_in], out=out_dev).squeeze(1)
            if world > 1:   # final gather of waveforms on rank 0 over NVLink (part of the step)
                dist.gather(wav, list(gathered.split(B)) if rank == 0 else None, dst=0)
        return wav

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.ready.wait(timeout=10)

    # ---- timed region: device-resident inputs -------------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.active = True
    e0.record()
    for i in range(args.steps):
        ev[i][0].record()   # materialise the lazily created cudaEvent_t handles; the library re-records them
        ev[i][1].record()   # around its tensor-core convolution launches (e2e_voc_set_profile_events)
        voc._profile_events = (ev[i][0].cuda_event, ev[i][1].cuda_event)
        step(i)
    e1.record()
    barrier()
    sampler.active = False
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    conv_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    audio_s_step = world * B * S / SR
    value = audio_s_step / (ms_step * 1e-3)

    # ---- end-to-end: pinned host mel -> H2D -> forward -> D2H waveform, every step ---------------------------
    # through the package's host-to-host serving loop (e2e_tts_b200.serving.HostPipeline): the copies of neighbouring
    # steps overlap the synthesis of the current one; every step's input comes from pinned host memory and every step's
    # waveform lands in pinned host memory inside the timed region.
    from e2e_tts_b200.serving import HostPipeline
    wav_hosts = [wav_host, torch.empty((B, S), dtype=torch.float32).pin_memory()]
    pipe = HostPipeline(voc, dev)
    for i in range(3):
        pipe.submit(mels_host[i % n_in], wav_hosts[i % 2])
    pipe.drain()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for i in range(args.steps):
        pipe.submit(mels_host[i % n_in], wav_hosts[i % 2])
    pipe.join()
    h1.record()
    pipe.drain()
    barrier()
    t = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = audio_s_step / (t.item() / args.steps * 1e-3)

    peak, peak_src = peaks()
    traffic, traffic_src = ncu_traffic()
    conv_tflops = B * T * CONV_TC_FLOP_PER_FRAME / (conv_ms * 1e-3) * 1e-12
    launches = voc.launches_per_forward()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "cfg2 per GPU: 16 utterances x 5 s (T=431 -> 110336 samples), HiFi-GAN V1 default config "
                               "(model_config.yaml:75-82), random-init weights (fan-in-scaled 'strong' regime), bf16 "
                               "operands / fp32 accumulate / bf16 activation+residual stream / fp32 resblock sum",
                   "global_batch": world * B, "mel_frames": T, "parallelism": "batch-sharded x%d + gather" % world,
                   "l2": "no flush: per-step working set ~1.1 GB >> 126 MB L2; inputs rotate over %d buffers" % n_in},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 80 * T * 4, "d2h_bytes_per_step": B * S * 4},
        "gpu_launches": launches * args.steps,
        "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peak, "unit": "TFLOP/s",
                     "frac": conv_tflops / peak, "traffic": traffic, "kernel": "pair_tc_kernel+conv_tc_kernel",
                     "peak_source": peak_src,
                     "traffic_note": ("DRAM bytes per step over the same launches, " + traffic_src) if traffic else "",
                     "note": "%d tcgen05 launches per step (pair_tc_kernel + conv_tc_kernel: the same implicit-GEMM "
                             "pipeline, fused and unfused), %.3f ms of the %.3f ms step" %
                             (launches - 2, conv_ms, ms_step)},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_oracle_throughput(12.0, 2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
