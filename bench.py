#!/usr/bin/env python
"""Benchmark of the e2e-tts synthesis hot path on B200 (BASELINE.json metric: audio-seconds synthesised per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--passes R] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (one JSON line on rank 0):
  N = 1   BASELINE.json configs[1]: 16 utterances x 5 s (T = 431 mel frames -> 110 336 samples each) through
          HifiGan.forward, random-init weights of the reference's default HiFi-GAN V1 config.
  N > 1   BASELINE.json configs[3]: 256 utterances x 5 s sharded contiguously, 256 / N per rank, no data-path
          collective, one NCCL gather of all waveforms on rank 0 per pass (strong scaling).  The
          gather of pass i runs on NCCL's stream while pass i + 1 is synthesised; it is also timed alone with its own
          CUDA events (`gather_ms`) and the gathered rows are verified against every rank's own result after the timed
          region.  A 16-per-rank weak-scaling sub-record is kept (`weak16`).
A "pass" is one forward over one batch; a "step" is R back-to-back passes (R chosen so the K timed steps last >= 2 s:
the sustained-clock denominator of the roofline is only honest for a seconds-long region; --passes overrides).
`value` = audio-seconds / time over the K steps with inputs resident in HBM; `e2e` = the same passes through
e2e_tts_b200.HostPipeline (pinned host mel -> H2D -> forward -> D2H waveform, every pass).

Extra objects on the N = 1 line: `workloads` (cfg3 8 x 30 s; cfg4 on one GPU = the strong-scaling base; cfg5 mel
front-end 1024 x 10 s with its own HBM roofline; single-utterance / small-batch latency), `gpu_eager_baseline` (the
reference generator in eager PyTorch on the same B200: fp32 as shipped, fp32 after remove_weight_norm(), bf16 — the
bar to beat), `cpu_baseline` (the reference class on the host cores, bounded sample).

--impl reference times the reference's own CPU path on the host cores: the UNMODIFIED reference generator from
baseline/_ref (staged by oracle/build_ref.py; `kind: "reference"`), or the oracle port when those files are absent
(`kind: "port"`), same config, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR = 22050
HOP = 256
T_FRAMES = 431           # 5 s utterance (110 336 samples = 5.004 s)
T_LONG = 2584            # 30 s utterance
B_CFG2 = 16
B_CFG4 = 256
FLOP_PER_FRAME = 614105088          # 2*MACs of all 78 convs per mel frame
CONV_TC_FLOP_PER_FRAME = FLOP_PER_FRAME - 114688   # everything but conv_post runs on the tensor cores
MEL_L = 220500           # 10 s clip
MEL_CLIPS = 1024
MEL_DRAM_BYTES_NCU = 904063232 + 265385472   # dram__bytes_read.sum + dram__bytes_write.sum of one 1024-clip launch, same file
MEL_WARP_INSTR_PER_FRAME = 1043   # ncu smsp__inst_executed.sum / 881 664 frames, profiles/r02_ncu_full_mel_h.txt
METRIC = "audio_seconds_synthesized_per_second"
UNIT = "audio-s/s"
NUMERICS = ("bf16 operands / fp32 accumulate; activations stored once as bf16 leaky_relu(x); residual adds and the "
            "resblock sum are fp32 adds in the epilogue (see DESIGN.md §2 for which tensors are kept in fp32)")


def workload_config(world: int) -> dict:
    """Identical for both arms (the driver compares the two `config` objects)."""
    if world == 1:
        return {"workload": "cfg2: 16 utterances x 5 s (T=431 -> 110336 samples), HiFi-GAN V1 default config "
                            "(model_config.yaml:75-82), random-init weights (fan-in-scaled 'strong' regime)",
                "global_batch": B_CFG2, "mel_frames": T_FRAMES, "parallelism": "single GPU"}
    return {"workload": "cfg4: 256 utterances x 5 s (T=431) sharded %d per rank over %d GPUs + NCCL gather of the "
                        "waveforms on rank 0, HiFi-GAN V1 default config, random-init weights ('strong' regime)"
                        % (B_CFG4 // world, world),
            "global_batch": B_CFG4, "mel_frames": T_FRAMES, "parallelism": "batch-sharded x%d + gather" % world}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"sustained": float(p["bf16_tflops_sustained"]), "burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"sustained": 1400.0, "burst": 1590.0, "hbm": 6650.0,
                "source": "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained, 6.65 TB/s)"}


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the tcgen05 launches of ONE cfg2 pass, from the
    committed ncu capture of this same command (profiles/roofline_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_step_tcgen05"]), str(t.get("source", "")), t.get("tensor_pipe_active_pct_tcgen05")
    except Exception:
        return None, "", None


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
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report that, never fake numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)
            self.ready.set()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------
# Baseline legs (the only places that touch oracle/ and baseline/_ref)
# ---------------------------------------------------------------------------------------------------------------
def reference_forward_fn(device: str, variant: str = "fp32"):
    """Returns (fn(mel) -> wav, kind).  variant: "fp32" (as shipped: weight-norm hooks on, e2e_tts/src/api/utils.py:53-56),
    "fp32_nowm" (after remove_weight_norm(), generator.py:55-62), "bf16" (.to(bfloat16) after remove_weight_norm())."""
    from e2e_tts_b200 import synthetic as sy
    from oracle import ref_loader
    cfg = sy.DEFAULT_CONFIG
    sd = sy.make_state_dict(cfg, 1, "strong")
    m = ref_loader.build_reference_hifigan(cfg, sd)
    if m is not None:
        if variant != "fp32":
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                m.remove_weight_norm()   # prints "Removing weight norm..."
        m = m.to(device)
        if variant == "bf16":
            m = m.to(torch.bfloat16)
            return (lambda mel: m(mel.to(torch.bfloat16)).float()), "reference"
        return (lambda mel: m(mel)), "reference"
    # the staged reference files are absent: oracle port (a functional restatement of the same eager ops)
    from oracle import hifigan_oracle as ho
    dt = torch.bfloat16 if variant == "bf16" else torch.float32
    sdd = {k: v.to(device) for k, v in sd.items()}
    if variant != "fp32":
        folded = {}
        for name, *_ in sy.layer_names(cfg):
            folded[name + ".weight"] = ho.fold_weight_norm(sdd[name + ".weight_g"], sdd[name + ".weight_v"])
            folded[name + ".bias"] = sdd[name + ".bias"]
        sdd = folded
    return (lambda mel: ho.hifigan_forward(sdd, cfg, mel, dtype=dt).float()), "port"


def cpu_reference_throughput(seconds_budget: float, utterances: int):
    """The reference generator on the host cores: audio-s/s on `utterances` x 5 s, repeated until about
    `seconds_budget` s were spent (at least one pass)."""
    from e2e_tts_b200 import synthetic as sy
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind = reference_forward_fn("cpu", "fp32")
    mel = sy.mel_like(utterances, T_FRAMES, 0)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fn(mel[:1])     # warm-up (thread pool, oneDNN primitives)
        times = []
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            fn(mel)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > seconds_budget:
                break
    best = min(times)
    audio_s = utterances * T_FRAMES * HOP / SR
    what = ("unmodified reference HifiGan (baseline/_ref, generator.py:13-62, weight-norm hooks on)" if kind == "reference"
            else "oracle port of generator.py:37-53")
    return audio_s / best, cores, kind, "%d x 5 s utterances (T=431) per pass, best of %d passes, fp32 eager torch, %s" % (
        utterances, len(times), what)


def run_reference(args):
    """The reference arm: the reference's own CPU implementation on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from e2e_tts_b200 import synthetic as sy
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind = reference_forward_fn("cpu", "fp32")
    full = B_CFG2 if world == 1 else B_CFG4
    n_pass = args.steps + min(args.warmup, 2)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = sy.mel_like(1, T_FRAMES, 0)
        fn(one)
        t0 = time.perf_counter()
        fn(one)
        t1 = time.perf_counter() - t0          # one utterance: sizes the bounded sample (~150 s for the whole run)
        utter = int(max(1, min(B_CFG2, 150.0 / (n_pass * t1 * 0.8))))
        mel = sy.mel_like(utter, T_FRAMES, 0)
        for _ in range(min(args.warmup, 2)):
            fn(mel)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(mel)
        dt = time.perf_counter() - t0
    audio_s = utter * T_FRAMES * HOP / SR
    value = audio_s * args.steps / dt
    what = ("unmodified reference HifiGan from baseline/_ref (generator.py:13-62, weight-norm hooks on as "
            "src/api/utils.py:53-56 runs it)" if kind == "reference" else "oracle port of generator.py:37-53 (baseline/_ref absent)")
    sample = "each step = one forward over %d of the %d utterances x 5 s (T=431), fp32 eager torch on %d host threads, %s" % (
        utter, full, cores, what)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cfg1_cpu_reference():
    """SURVEY.md §8 d2 / BASELINE config 1: single-utterance inference on the CPU exactly as the reference runs it -
    the reference class with its constructor's default init (torch.manual_seed(1234)), weight-norm hooks left on
    (e2e_tts/src/api/utils.py:53-56), input torch.randn(1, 80, 431), fp32, all host threads."""
    from e2e_tts_b200 import synthetic as sy
    from oracle import ref_loader
    cls = ref_loader.reference_hifigan_class()
    if cls is None:
        return {"unavailable": "baseline/_ref not staged"}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cls(sy.DEFAULT_CONFIG).eval()
        x = torch.randn(1, 80, T_FRAMES)
        times = []
        with torch.no_grad():
            for i in range(8):
                t0 = time.perf_counter()
                m(x)
                if i >= 3:
                    times.append(time.perf_counter() - t0)
    times.sort()
    audio_s = T_FRAMES * HOP / SR
    return {"workload": "1 utterance x 5 s on the CPU, reference HifiGan, default init, weight-norm hooks on",
            "cores": cores, "kind": "reference", "best_s": times[0], "median_s": times[len(times) // 2],
            "value": audio_s / times[0], "unit": UNIT}


def mel_reference_baselines(dev, quick: bool):
    """the reference TorchSTFT (baseline/_ref; librosa's filterbank = the restated algorithm) on the CPU
    - per clip as e2e_tts/src/tools/tools_for_data.py:143-178 loops, and batched x 64 - and in eager PyTorch on the GPU."""
    from oracle import ref_loader
    cls = ref_loader.reference_stft_class()
    if cls is None:
        return {"unavailable": "baseline/_ref not staged"}
    out = {"kind": "reference", "unit": UNIT}
    g = torch.Generator().manual_seed(0)
    clips = torch.rand(64, MEL_L, generator=g) * 2 - 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cpu = cls()
        torch.set_num_threads(os.cpu_count() or 1)
        cpu.mel_spectrogram(clips[:1], return_energy=True)
        n = 4 if quick else 16
        t0 = time.perf_counter()
        for i in range(n):
            cpu.mel_spectrogram(clips[i:i + 1], return_energy=True)
        dt = time.perf_counter() - t0
        out["cpu_per_clip"] = {"value": n * MEL_L / SR / dt, "ms_per_clip": dt / n * 1e3, "cores": os.cpu_count() or 1}
        t0 = time.perf_counter()
        cpu.mel_spectrogram(clips, return_energy=True)
        dt = time.perf_counter() - t0
        out["cpu_batched_64"] = {"value": 64 * MEL_L / SR / dt, "ms_per_clip": dt / 64 * 1e3}
        try:
            gpu = cls(device=str(dev)).to(dev)
            xb = clips.to(dev)
            for _ in range(2):
                gpu.mel_spectrogram(xb, return_energy=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gpu.mel_spectrogram(xb, return_energy=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            out["gpu_eager_batched_64"] = {"value": 64 * MEL_L / SR / (ms * 1e-3), "ms_per_64_clips": ms,
                                           "note": "includes the reference's host-synchronising range assert"}
        except Exception as e:
            out["gpu_eager_batched_64"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    return out


def gpu_eager_baseline(dev, quick: bool):
    """the reference modules in eager PyTorch on this same B200 (cuDNN, cudnn.benchmark on)."""
    from e2e_tts_b200 import synthetic as sy
    out = {"workload": "cfg2: 16 x 5 s (T=431)", "unit": UNIT, "cudnn_benchmark": True, "torch": torch.__version__}
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    mel = sy.mel_like(B_CFG2, T_FRAMES, 0).to(dev)
    audio_s = B_CFG2 * T_FRAMES * HOP / SR
    try:
        for variant in ("fp32", "fp32_nowm", "bf16"):
            try:
                fn, kind = reference_forward_fn(str(dev), variant)
                out["kind"] = kind
                with torch.no_grad(), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    for _ in range(3):
                        fn(mel)
                    torch.cuda.synchronize()
                    n = 3 if quick else 10
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n):
                        fn(mel)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out[variant] = {"value": audio_s / (ms * 1e-3), "ms_per_pass": ms}
            except Exception as e:   # a baseline leg must never take the bench line down
                out[variant] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = old
    return out


# ---------------------------------------------------------------------------------------------------------------
# Side workloads (N = 1)
# ---------------------------------------------------------------------------------------------------------------
def time_forward(voc, mels, iters, warm=3):
    """Device time per forward over `iters` back-to-back passes (rotating inputs, one output buffer), ms.  The warm-up
    walks the input buffers `warm` times with the same output buffer, so plan building and CUDA-graph capture (second
    sighting of a buffer combination) happen before the timed loop."""
    with torch.no_grad():
        out = voc(mels[0])
        for i in range(warm * len(mels)):
            voc(mels[i % len(mels)], out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            voc(mels[i % len(mels)], out=out)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def side_workloads(voc, dev, peaks, quick: bool):
    import e2e_tts_b200 as pkg
    from e2e_tts_b200 import synthetic as sy
    w = {}
    # cfg3: long-form, 8 x 30 s
    mels = [sy.mel_like(8, T_LONG, 300 + i).to(dev) for i in range(2)]
    ms = time_forward(voc, mels, 4 if quick else 40)
    tf = 8 * T_LONG * FLOP_PER_FRAME / (ms * 1e-3) * 1e-12
    w["cfg3"] = {"workload": "8 utterances x 30 s (T=2584 -> 661504 samples)", "ms_per_pass": ms,
                 "value": 8 * T_LONG * HOP / SR / (ms * 1e-3), "unit": UNIT, "tflops": tf,
                 "frac_sustained": tf / peaks["sustained"], "frac_burst": tf / peaks["burst"]}
    del mels
    # cfg4 on ONE GPU: the strong-scaling base of the N > 1 lines
    mels = [sy.mel_like(B_CFG4, T_FRAMES, 400 + i).to(dev) for i in range(2)]
    ms = time_forward(voc, mels, 2 if quick else 12, warm=3)
    w["cfg4_n1"] = {"workload": "256 utterances x 5 s on one GPU (no gather)", "ms_per_pass": ms,
                    "value": B_CFG4 * T_FRAMES * HOP / SR / (ms * 1e-3), "unit": UNIT}
    del mels
    voc._workspaces.clear()
    torch.cuda.empty_cache()
    # small batches: what src/api/utils.py:131-145 does per request (one bucket at a time)
    for b in (1, 4):
        mels = [sy.mel_like(b, T_FRAMES, 500 + i).to(dev) for i in range(4)]
        ms_dev = time_forward(voc, mels, 10 if quick else 100)
        lat = []
        with torch.no_grad():
            out = voc(mels[0])
            for i in range(12):                      # plans and graphs for every input buffer exist before timing
                voc(mels[i % 4], out=out)
            for i in range(5 if quick else 30):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                voc(mels[i % 4], out=out)
                torch.cuda.synchronize()
                lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        w["latency_b%d" % b] = {"workload": "%d utterance(s) x 5 s (T=431)" % b, "device_ms_back_to_back": ms_dev,
                                "host_latency_ms_median": lat[len(lat) // 2], "host_latency_ms_min": lat[0],
                                "value": b * T_FRAMES * HOP / SR / (ms_dev * 1e-3), "unit": UNIT,
                                "launches": voc.launches_per_forward()}
    # cfg5: mel front-end on 1024 x 10 s clips (HBM roofline)
    clips = 64 if quick else MEL_CLIPS
    wav = torch.rand(clips, MEL_L, device=dev) * 2 - 1
    stft = pkg.TorchSTFT()
    for _ in range(3):
        mel, en = stft.mel_spectrogram(wav, return_energy=True, check_range=False)
    torch.cuda.synchronize()
    n = 3 if quick else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        mel, en = stft.mel_spectrogram(wav, return_energy=True, check_range=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    Tm = mel.shape[-1]
    alg = clips * (4 * MEL_L + 4 * 81 * Tm)          # fp32 audio in, fp32 mel (80) + energy (1) out
    gbs = alg / (ms * 1e-3) * 1e-9
    # What bounds the kernel is instruction issue / latency, not HBM: 1 043 warp instructions per frame
    # (ncu smsp__inst_executed.sum / frames, profiles/r02_ncu_full_mel_h.txt) at 4 issue slots per cycle and SM
    issue_floor_ms = clips * Tm * MEL_WARP_INSTR_PER_FRAME / (148 * 4 * 1.965e9) * 1e3
    w["cfg5_mel"] = {"workload": "%d clips x 10 s (L=220500 -> T=%d frames), TorchSTFT.mel_spectrogram(return_energy=True), "
                                 "range check off (no host sync)" % (clips, Tm),
                     "ms_per_pass": ms, "value": clips * MEL_L / SR / (ms * 1e-3), "unit": UNIT,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                                  "frac": gbs / peaks["hbm"], "traffic": MEL_DRAM_BYTES_NCU if clips == 1024 else None,
                                  "kernel": "mel_kernel",
                                  "algorithmic_bytes": alg, "hbm_floor_ms": alg / (peaks["hbm"] * 1e9) * 1e3,
                                  "issue_slot_floor_ms": issue_floor_ms, "issue_slot_frac": issue_floor_ms / ms,
                                  "note": "SURVEY.md §8 d7 names HBM as this path's roofline and `frac` is reported against it; "
                                          "the kernel is bound by dependent-instruction latency at about half of the issue slots "
                                          "(%d warp instructions per frame, ncu): " % MEL_WARP_INSTR_PER_FRAME +
                                          "issue_slot_frac = time at 100 % issue-slot use / measured time",
                                  "peak_source": peaks["source"]}}
    del wav, mel, en
    torch.cuda.empty_cache()
    return w


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--passes", type=int, default=0, help="passes per step (0 = auto: timed region >= 2 s)")
    ap.add_argument("--quick", action="store_true", help="short side workloads, no CPU / eager baselines (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="headline only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import e2e_tts_b200 as pkg
    from e2e_tts_b200 import synthetic as sy

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

    cfg = sy.DEFAULT_CONFIG
    voc = pkg.HifiGan(cfg)
    voc.load_state_dict(sy.make_state_dict(cfg, 1, "strong"))
    voc = voc.eval().to(dev)
    peaks = load_peaks()
    T = T_FRAMES
    S = HOP * T
    if world > 1 and B_CFG4 % world:
        raise SystemExit("cfg4 shards 256 utterances: --gpus must divide 256")
    B = B_CFG2 if world == 1 else B_CFG4 // world
    n_in = 4   # rotating inputs; the per-pass working set (>= 1.1 GB of activations) is far larger than the 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    class Job:
        """One sharded workload: this rank's `b` utterances per pass (+ the gather on rank 0 when world > 1).  Results
        alternate between two device buffers so that the NCCL gather of pass i (its own stream) overlaps the synthesis
        of pass i + 1; a buffer is reused only after its gather has completed (stream-ordered wait, no host block)."""

        def __init__(self, b):
            self.b = b
            self.mels_host = [sy.mel_like(b, T, 1000 * rank + i).pin_memory() for i in range(n_in)]
            self.mels_dev = [m.to(dev) for m in self.mels_host]
            self.out_dev = [torch.empty((b, 1, S), dtype=torch.float32, device=dev) for _ in range(2)]
            self.gathered = ([torch.empty((world * b, S), dtype=torch.float32, device=dev) for _ in range(2)]
                             if (world > 1 and rank == 0) else None)
            self.views = [list(g.split(b)) for g in self.gathered] if self.gathered else None
            self.works = [None, None]
            self.gather_events = []
            self.last_slot = 0

        def one_pass(self, i, serial_gather=False):
            slot = i & 1
            self.last_slot = slot
            with torch.no_grad():
                if self.works[slot] is not None:
                    self.works[slot].wait()       # the gather that read this buffer two passes ago has completed
                    self.works[slot] = None
                wav = voc(self.mels_dev[i % n_in], out=self.out_dev[slot]).squeeze(1)
                if world > 1:   # final gather of waveforms on rank 0 over NVLink (part of every pass)
                    recv = self.views[slot] if rank == 0 else None
                    if serial_gather:   # measurement of the collective alone: CUDA events on the compute stream
                        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        g0.record()
                        dist.gather(wav, recv, dst=0)
                        g1.record()
                        self.gather_events.append((g0, g1))
                    else:
                        self.works[slot] = dist.gather(wav, recv, dst=0, async_op=True)
            return wav

        def finish(self):
            for k in range(2):
                if self.works[k] is not None:
                    self.works[k].wait()
                    self.works[k] = None

        def verify_gather(self):
            """After a pass: rank 0 checks that gathered[r*b:(r+1)*b] is bit-identical to rank r's own result
            (wrap-around int64 sum of the fp32 bit patterns of every row + a position-weighted one)."""
            if world == 1:
                return None
            bits = self.out_dev[self.last_slot].view(torch.int32).reshape(self.b, S).to(torch.int64)
            wts = (torch.arange(S, device=dev, dtype=torch.int64) % 8191) + 1
            mine = torch.stack([bits.sum(dim=1), (bits * wts).sum(dim=1)], dim=1)       # [b, 2]
            allc = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine)
            if rank != 0:
                return None
            gb = self.gathered[self.last_slot].view(torch.int32).to(torch.int64)
            got = torch.stack([gb.sum(dim=1), (gb * wts).sum(dim=1)], dim=1)
            want = torch.cat(allc, dim=0)
            bad = int((got != want).any(dim=1).sum().item())
            return {"rows_checked": int(got.shape[0]), "rows_mismatched": bad, "ok": bad == 0}

    job = Job(B)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        job.one_pass(i)
    job.finish()
    barrier()
    # size the step: R passes so that the K timed steps last >= 2 s (same R on every rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        job.one_pass(i)
    job.finish()
    e1.record()
    barrier()
    pass_ms_est = max_over_ranks(e0.elapsed_time(e1) / 3)
    R = args.passes if args.passes > 0 else int(min(64, max(1, -(-2000.0 // (args.steps * pass_ms_est)))))
    sampler.ready.wait(timeout=10)

    # ---- timed region: device-resident inputs -------------------------------------------------------------------
    n_pass = args.steps * R
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_pass)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.active = True
    e0.record()
    for i in range(n_pass):
        ev[i][0].record()   # materialise the lazily created cudaEvent_t handles; the library re-records them
        ev[i][1].record()   # around its tensor-core convolution launches (e2e_voc_set_profile_events)
        voc._profile_events = (ev[i][0].cuda_event, ev[i][1].cuda_event)
        job.one_pass(i)
    job.finish()
    e1.record()
    barrier()
    sampler.active = False
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    conv_ms = sum(a.elapsed_time(b) for a, b in ev) / n_pass
    ms_step = ms_total / args.steps
    audio_s_pass = world * B * S / SR
    value = audio_s_pass * R / (ms_step * 1e-3)
    gather = None
    if world > 1:
        for i in range(12):     # the collective alone (not overlapped), after the timed region
            job.one_pass(i, serial_gather=True)
        barrier()
        gms = sorted(a.elapsed_time(b) for a, b in job.gather_events[2:])
        gather = {"gather_ms_median": max_over_ranks(gms[len(gms) // 2]), "gather_ms_mean": max_over_ranks(sum(gms) / len(gms)),
                  "bytes_to_rank0_per_pass": (world - 1) * B * S * 4,
                  "share_of_pass_if_serial": max_over_ranks(sum(gms) / len(gms)) / (ms_step / R),
                  "overlapped": True,
                  "note": "gather_ms: CUDA events on the compute stream around a blocking dist.gather, measured over 10 "
                          "passes after the timed region (rank 0 waits for every rank's result, so the figure includes "
                          "skew between ranks); max over ranks.  Inside the timed region the gather of pass i runs on "
                          "NCCL's stream while pass i + 1 is synthesised (two result buffers)"}
        job.one_pass(0, serial_gather=True)
        gather["verified"] = job.verify_gather()

    # ---- end-to-end: pinned host mel -> H2D -> forward -> D2H waveform, every pass --------------------------------
    from e2e_tts_b200.serving import HostPipeline
    wav_hosts = [torch.empty((B, S), dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = HostPipeline(voc, dev)
    for i in range(3):
        pipe.submit(job.mels_host[i % n_in], wav_hosts[i % 2])
    pipe.drain()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for i in range(n_pass):
        pipe.submit(job.mels_host[i % n_in], wav_hosts[i % 2])
    pipe.join()
    h1.record()
    pipe.drain()
    barrier()
    e2e_ms = max_over_ranks(h0.elapsed_time(h1))
    e2e_value = audio_s_pass * n_pass / (e2e_ms * 1e-3)
    del pipe

    traffic, traffic_src, tensor_pipe_pct = ncu_traffic()
    conv_tflops = B * T * CONV_TC_FLOP_PER_FRAME / (conv_ms * 1e-3) * 1e-12
    launches = voc.launches_per_forward()
    config = workload_config(world)
    config.update({"passes_per_step": R, "utterances_per_pass_per_gpu": B,
                   "step": "%d back-to-back forward passes (timed region %.2f s)" % (R, ms_total * 1e-3),
                   "numerics": NUMERICS,
                   "l2": "no flush: per-pass working set >= 1.1 GB >> 126 MB L2; inputs rotate over %d buffers" % n_in})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "ms_per_pass": ms_step / R, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": R * B * 80 * T * 4,
                "d2h_bytes_per_step": R * B * S * 4},
        "gpu_launches": launches * n_pass,
        "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": conv_tflops / peaks["sustained"], "frac_burst": conv_tflops / peaks["burst"],
                     "traffic": traffic if (world == 1) else None, "kernel": "stage_tc/pair_tc/conv_tc (tcgen05 implicit-GEMM convolutions)",
                     "peak_source": peaks["source"] + ": bf16_tflops_sustained (the kernels are timed inside a >= 2 s region); "
                                    "frac_burst uses bf16_tflops",
                     "traffic_note": ("DRAM bytes per cfg2 pass over the same launches, " + traffic_src) if traffic else "",
                     # BASELINE metric, second half ("vocoder tensor-pipe util %"): ncu sm__pipe_tensor_cycles_active of the
                     # same launches, weighted by launch time, from the committed launch list (not measured in this run)
                     "tensor_pipe_active_pct_ncu": tensor_pipe_pct if (world == 1) else None,
                     "note": "%d tensor-core launches per pass, %.3f ms of the %.3f ms pass (CUDA events around them, "
                             "mean over %d passes, rank %d)" % (launches - 2, conv_ms, ms_step / R, n_pass, rank)},
    }
    if gather is not None:
        line["gather"] = gather
        # weak-scaling sub-record: 16 utterances per rank (the N = 1 headline workload on every GPU) + gather
        del job
        torch.cuda.empty_cache()
        wjob = Job(B_CFG2)
        for i in range(3):
            wjob.one_pass(i)
        wjob.finish()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nw = 100
        w0.record()
        for i in range(nw):
            wjob.one_pass(i)
        wjob.finish()
        w1.record()
        barrier()
        wms = max_over_ranks(w0.elapsed_time(w1)) / nw
        line["weak16"] = {"workload": "16 utterances x 5 s per rank + gather", "ms_per_pass": wms,
                          "value": world * B_CFG2 * S / SR / (wms * 1e-3), "unit": UNIT, "passes": nw}
    if rank == 0 and world == 1 and not args.no_side:
        try:
            line["workloads"] = side_workloads(voc, dev, peaks, args.quick)
        except Exception as e:
            line["workloads"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        if not args.quick:
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev, args.quick)
            for key, fn in (("cfg1_cpu_reference", cfg1_cpu_reference),
                            ("cfg5_mel_reference", lambda: mel_reference_baselines(dev, args.quick))):
                try:
                    line["workloads"][key] = fn()
                except Exception as e:   # a baseline leg must never take the bench line down
                    line["workloads"][key] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        v, cores, kind, sample = cpu_reference_throughput(12.0, 2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
