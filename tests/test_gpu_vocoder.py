"""GPU parity tests for the vocoder path, through the C ABI (e2e_voc_forward) via the drop-in module.

Tolerance (bf16 tensor-core operands, fp32 accumulation; activations are stored as bf16, the residual / resblock-sum
adds run in fp32 - DESIGN.md §2; vs the fp32 reference/oracle):
    max |wav - ref| <= 2e-2 * max|ref|     and     mean |wav - ref| <= 3e-3 * max|ref|
also enforced separately on the first / last 4096 samples, where per-layer zero padding matters."""
import os

import numpy as np
import pytest
import torch

import e2e_tts_b200 as pkg
from oracle import hifigan_oracle as ho
import margins

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAX_TOL, MEAN_TOL = 2e-2, 3e-3


def mel_like(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


def check(got, want, what=""):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    assert torch.isfinite(got).all()
    scale = want.abs().max().item()
    def one(a, b, tag):
        d = (a - b).abs()
        margins.record("%s %s" % (what, tag), max_rel=d.max().item() / scale, mean_rel=d.mean().item() / scale,
                       bound_max=MAX_TOL, bound_mean=MEAN_TOL, scale=scale)
        assert d.max().item() <= MAX_TOL * scale, "%s %s: max err %.3g vs scale %.3g" % (what, tag, d.max().item(), scale)
        assert d.mean().item() <= MEAN_TOL * scale, "%s %s: mean err %.3g vs scale %.3g" % (what, tag, d.mean().item(), scale)
    one(got, want, "all")
    n = min(4096, got.shape[-1])
    one(got[..., :n], want[..., :n], "head")
    one(got[..., -n:], want[..., -n:], "tail")


def build(cfg, seed, regime, operand_dtype=None):
    sd = ho.make_state_dict(cfg, seed, regime)
    voc = pkg.HifiGan(cfg, operand_dtype=operand_dtype)
    voc.load_state_dict(sd)
    return voc.eval().to("cuda"), sd


@pytest.mark.parametrize("name", ["voc_default_init", "voc_strong_init", "voc_strong_resblock2"])
def test_against_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = dict(ho.DEFAULT_CONFIG)
    cfg["resblock"] = int(g["resblock"])
    voc, _ = build(cfg, int(g["seed"]), str(g["regime"]))
    with torch.no_grad():
        wav = voc(torch.from_numpy(g["mel"]).cuda())
    assert wav.is_cuda and wav.dtype == torch.float32
    check(wav, torch.from_numpy(g["wav"]), name)


@pytest.mark.parametrize("B,T,regime", [(1, 1, "strong"), (3, 7, "strong"), (2, 130, "strong"), (1, 431, "default"),
                                        (2, 431, "strong")])
def test_against_oracle(B, T, regime):
    cfg = ho.DEFAULT_CONFIG
    voc, sd = build(cfg, 40 + T, regime)
    mel = mel_like(B, T, 7 * T + B)
    with torch.no_grad():
        want = ho.hifigan_forward(sd, cfg, mel)
        got = voc(mel.cuda())
    assert tuple(got.shape) == (B, 1, 256 * T)
    check(got, want, "B%d T%d %s" % (B, T, regime))


def test_transposed_view_input_and_call_site_contract():
    """utils.py:144-145: vocoder(mel_predicted.transpose(1, 2)).squeeze(1).detach().cpu().numpy()"""
    cfg = ho.DEFAULT_CONFIG
    voc, sd = build(cfg, 5, "strong")
    mel_btc = mel_like(2, 40, 99).transpose(1, 2).contiguous()          # [B, T, 80] as the acoustic model emits
    with torch.no_grad():
        audio = voc(mel_btc.cuda().transpose(1, 2)).squeeze(1).detach().cpu().numpy()
        want = ho.hifigan_forward(sd, cfg, mel_btc.transpose(1, 2))
    assert audio.shape == (2, 40 * 256)
    check(torch.from_numpy(audio)[:, None], want, "transposed view")


def test_batch_items_are_independent_and_deterministic():
    cfg = ho.DEFAULT_CONFIG
    voc, _ = build(cfg, 6, "strong")
    mel = mel_like(3, 50, 1).cuda()
    with torch.no_grad():
        a = voc(mel)
        b = voc(mel[1:2])
        c = voc(mel)
    assert torch.equal(a, c)
    assert torch.equal(a[1:2], b)


def test_long_form_tile_seams():
    """30 s utterance (config 3 shape, one item): compare a window around every 128*mt-row tile seam."""
    cfg = ho.DEFAULT_CONFIG
    voc, sd = build(cfg, 8, "strong")
    T = 2584
    mel = mel_like(1, T, 2)
    with torch.no_grad():
        got = voc(mel.cuda()).cpu()
        want = ho.hifigan_forward(sd, cfg, mel)
    check(got, want, "long")
    d = (got - want).abs()[0, 0]
    scale = want.abs().max().item()
    for seam in range(512, d.numel(), 512):                              # every possible output-tile boundary
        assert d[seam - 8: seam + 8].max().item() <= MAX_TOL * scale


def test_weights_reload_after_in_place_update():
    cfg = ho.DEFAULT_CONFIG
    voc, _ = build(cfg, 9, "strong")
    mel = mel_like(1, 12, 3).cuda()
    with torch.no_grad():
        a = voc(mel)
        sd2 = ho.make_state_dict(cfg, 10, "strong")
        voc.load_state_dict(sd2)
        b = voc(mel)
        want = ho.hifigan_forward(sd2, cfg, mel.cpu())
    assert not torch.equal(a, b)
    check(b, want, "reloaded")


def test_forward_pcm16_is_the_callers_post_processing_fused():
    """HifiGan.forward_pcm16 == combine_audio's trim / * max_wav_value / astype(int16) (src/api/utils.py:108-117)
    applied to forward(): same kernel arithmetic, so the int16 stream must be bit-identical."""
    from oracle import postprocess_oracle as po
    voc, _ = build(ho.DEFAULT_CONFIG, 11, "strong")
    mel = mel_like(3, 40, 21).cuda()
    lengths = [40, 17, 0]
    with torch.no_grad():
        wav = voc(mel).squeeze(1)
        pcm = voc.forward_pcm16(mel, lengths)
        pcm_full = voc.forward_pcm16(mel)
    assert pcm.dtype == torch.int16 and pcm.shape == wav.shape
    want_full = torch.trunc(wav * 32768.0).clamp(-32768, 32767).to(torch.int16)
    assert torch.equal(pcm_full, want_full)
    for b, n in enumerate(lengths):
        assert torch.equal(pcm[b, : n * 256], want_full[b, : n * 256])
        assert int(pcm[b, n * 256:].abs().max().item()) == 0 if n < 40 else True
    ref_stream = po.combine_audio(list(wav.cpu().numpy()), lengths, 100)
    assert np.array_equal(pkg.combine_audio(list(pcm.cpu()), lengths, 100), ref_stream)
    with pytest.raises(ValueError):
        voc.forward_pcm16(mel, [1, 2])


def test_empty_inputs_and_size_guards():
    """Empty batches / zero frames return empty waveforms (the reference's convs do the same for B = 0); utterances
    whose waveform index would not fit 31 bits are refused with a Python exception, not a crash."""
    voc, _ = build(ho.DEFAULT_CONFIG, 12, "strong")
    with torch.no_grad():
        assert voc(torch.zeros(0, 80, 5).cuda()).shape == (0, 1, 1280)
        assert voc(torch.zeros(2, 80, 0).cuda()).shape == (2, 1, 0)
        assert voc.forward_pcm16(torch.zeros(0, 80, 5).cuda()).shape == (0, 1280)
        voc(torch.zeros(1, 80, 3).cuda())                               # creates the native handle
    from e2e_tts_b200 import _native
    lib = _native.lib()
    rc = lib.e2e_voc_forward(voc._handle, 8, 1, 1, 1, 1, 1 << 23, 8, 1024, 1 << 40, None)   # T * 256 >= 2^31
    assert rc != 0 and b"too long" in lib.e2e_last_error_string()
    with pytest.raises(ValueError):
        voc(torch.zeros(1, 81, 4).cuda())
    with pytest.raises(ValueError):
        voc(torch.zeros(1, 80, 4, dtype=torch.float64).cuda())


def test_host_pipeline_matches_direct_calls():
    """e2e_tts_b200.HostPipeline (pinned host in / out, copies overlapped with synthesis) returns exactly what the
    direct forward() / forward_pcm16() calls return, for more batches than it has buffers."""
    voc, _ = build(ho.DEFAULT_CONFIG, 14, "strong")
    mels = [mel_like(2, 30, 50 + i).pin_memory() for i in range(5)]
    outs = [torch.empty(2, 30 * 256).pin_memory() for _ in range(5)]
    pipe = pkg.HostPipeline(voc)
    for m, o in zip(mels, outs):
        pipe.submit(m, o)
    pipe.drain()
    with torch.no_grad():
        for m, o in zip(mels, outs):
            assert torch.equal(o, voc(m.cuda()).squeeze(1).cpu())
    pcm = [torch.empty(2, 30 * 256, dtype=torch.int16).pin_memory() for _ in range(3)]
    pipe16 = pkg.HostPipeline(voc.forward_pcm16)
    idx = [pipe16.submit(m, o) for m, o in zip(mels[:3], pcm)]
    pipe16.wait(idx[-1])
    with torch.no_grad():
        for m, o in zip(mels[:3], pcm):
            assert torch.equal(o, voc.forward_pcm16(m.cuda()).cpu())


FP16_MAX_TOL, FP16_MEAN_TOL = 4e-3, 6e-4   # 5x tighter than the bf16 bounds: fp16 operands carry 3 more significand bits


@pytest.mark.parametrize("name", ["voc_default_init", "voc_strong_init", "voc_strong_resblock2"])
def test_fp16_operand_mode_against_reference_golden(name):
    """operand_dtype="fp16": same kernels, fp16 instead of bf16 operands / activations (e2e_voc_set_operand_dtype)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = dict(ho.DEFAULT_CONFIG)
    cfg["resblock"] = int(g["resblock"])
    voc, _ = build(cfg, int(g["seed"]), str(g["regime"]), operand_dtype="fp16")
    with torch.no_grad():
        wav = voc(torch.from_numpy(g["mel"]).cuda()).cpu()
    ref = torch.from_numpy(g["wav"])
    scale = ref.abs().max().item()
    d = (wav - ref).abs()
    margins.record("fp16 " + name, max_rel=d.max().item() / scale, mean_rel=d.mean().item() / scale,
                   bound_max=FP16_MAX_TOL, bound_mean=FP16_MEAN_TOL, scale=scale)
    assert d.max().item() <= FP16_MAX_TOL * scale and d.mean().item() <= FP16_MEAN_TOL * scale, (d.max().item(), scale)


def test_fp16_operand_mode_full_utterance_and_pcm():
    voc, sd = build(ho.DEFAULT_CONFIG, 21, "strong", operand_dtype="fp16")
    mel = mel_like(2, 431, 31)
    with torch.no_grad():
        wav = voc(mel.cuda()).cpu()
        pcm = voc.forward_pcm16(mel.cuda()).cpu()
        ref = ho.hifigan_forward(sd, ho.DEFAULT_CONFIG, mel)
    scale = ref.abs().max().item()
    d = (wav - ref).abs()
    margins.record("fp16 2 x 5 s", max_rel=d.max().item() / scale, mean_rel=d.mean().item() / scale,
                   bound_max=FP16_MAX_TOL, bound_mean=FP16_MEAN_TOL, scale=scale)
    assert d.max().item() <= FP16_MAX_TOL * scale and d.mean().item() <= FP16_MEAN_TOL * scale
    assert torch.equal(pcm, (wav.squeeze(1) * 32768.0).clamp(-32768, 32767).to(torch.int16))
    with pytest.raises(ValueError):
        pkg.HifiGan(ho.DEFAULT_CONFIG, operand_dtype="fp8")


@pytest.mark.skipif(os.environ.get("E2E_NO_GRAPH") is not None, reason="E2E_NO_GRAPH switches CUDA-graph replay off")
def test_cuda_graph_replay_is_bit_identical_and_tracks_weight_reloads():
    """A forward on buffers seen before is captured into a CUDA graph (2nd call) and replayed (3rd on): same bits as
    the eager launches, a changed input buffer content is picked up (the graph reads the buffer, not a snapshot),
    and new weights invalidate the captured graphs."""
    voc, sd = build(ho.DEFAULT_CONFIG, 16, "strong")
    mel_a, mel_b = mel_like(1, 60, 80).cuda(), mel_like(1, 60, 81).cuda()
    buf = mel_a.clone()
    out = torch.empty(1, 1, 60 * 256, device="cuda")
    with torch.no_grad():
        eager = voc(mel_a).clone()                       # fresh output buffer every call: never a graph
        assert not voc.last_forward_was_graph()
        for i in range(4):
            voc(buf, out=out)
            assert voc.last_forward_was_graph() == (i >= 1)
            assert torch.equal(out, eager)
        buf.copy_(mel_b)
        voc(buf, out=out)
        assert voc.last_forward_was_graph() and torch.equal(out, voc(mel_b))
        sd2 = ho.make_state_dict(ho.DEFAULT_CONFIG, 17, "strong")
        voc.load_state_dict(sd2)
        voc(buf, out=out)
        assert not voc.last_forward_was_graph()          # plans and graphs were dropped with the old weights
        want = ho.hifigan_forward(sd2, ho.DEFAULT_CONFIG, mel_b.cpu())
    check(out, want, "after reload")


@pytest.mark.parametrize("kind,C,k,dil", [(1, 128, 3, (1, 3, 5)), (1, 32, 7, (1, 3, 5)), (2, 64, 5, (2, 6)), (1, 256, 11, (1, 3, 5))])
def test_resblock_modules_are_callable(kind, C, k, dil):
    """ResBlock1.forward / ResBlock2.forward (layers.py:33-40, 60-65) as standalone modules, against the same eager ops
    the oracle uses; also on a non-contiguous input view."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    rb = (pkg.ResBlock1 if kind == 1 else pkg.ResBlock2)(C, k, dil).cuda()
    for p in rb.parameters():                       # O(1) activations through the block
        if p.dim() == 3 and p.shape[1] > 1:
            p.data.mul_(2.0)
    x = torch.randn(2, 300, C).transpose(1, 2)      # [B, C, T] view, like the generator's caller passes
    with torch.no_grad():
        got = rb(x.cuda()).cpu()
        y = x.clone()
        if kind == 1:
            for c1, c2, d in zip(rb.convs1, rb.convs2, rb.dilation):
                xt = F.conv1d(F.leaky_relu(y, 0.1), c1.folded_weight().cpu(), c1.bias.cpu(), dilation=d, padding=ho.get_padding(k, d))
                xt = F.conv1d(F.leaky_relu(xt, 0.1), c2.folded_weight().cpu(), c2.bias.cpu(), padding=ho.get_padding(k, 1))
                y = xt + y
        else:
            for c1, d in zip(rb.convs, rb.dilation):
                xt = F.conv1d(F.leaky_relu(y, 0.1), c1.folded_weight().cpu(), c1.bias.cpu(), dilation=d, padding=ho.get_padding(k, d))
                y = xt + y
    assert got.shape == y.shape
    check(got, y, "resblock%d C=%d k=%d" % (kind, C, k))
    with pytest.raises(RuntimeError):
        rb.cpu()(x)


def test_host_pipeline_variable_batch_shapes():
    """Serving batches differ in B and T from one submit to the next (one <=300-symbol bucket per request,
    e2e_tts/src/api/utils.py:131-145): the pipeline must re-size its device buffers, also with a plain callable that has
    no `out=` keyword."""
    voc, _ = build(ho.DEFAULT_CONFIG, 15, "strong")
    shapes = [(2, 30), (1, 47), (3, 30), (2, 30), (1, 9), (1, 47), (4, 12)]
    mels = [mel_like(b, t, 70 + i).pin_memory() for i, (b, t) in enumerate(shapes)]
    with torch.no_grad():
        want = [voc(m.cuda()).squeeze(1).cpu() for m in mels]
    for fn in (voc, lambda x: voc(x)):
        outs = [torch.empty(b, t * 256).pin_memory() for b, t in shapes]
        pipe = pkg.HostPipeline(fn, device="cuda")
        for m, o in zip(mels, outs):
            pipe.submit(m, o)
        pipe.drain()
        for o, w in zip(outs, want):
            assert torch.equal(o, w)


def test_host_pipeline_long_loop_keeps_a_bounded_event_list():
    """A serving loop never calls drain(): the pipeline forgets completed copy-out events (wait() on a forgotten batch
    returns at once), so its bookkeeping stays bounded, and the results stay those of direct calls."""
    voc, _ = build(ho.DEFAULT_CONFIG, 16, "strong")
    mels = [mel_like(1, 12, 90 + i).pin_memory() for i in range(3)]
    outs = [torch.empty(1, 12 * 256).pin_memory() for _ in range(4)]
    with torch.no_grad():
        want = [voc(m.cuda()).squeeze(1).cpu() for m in mels]
    pipe = pkg.HostPipeline(voc)
    n = 300
    for i in range(n):
        idx = pipe.submit(mels[i % 3], outs[i % 4])
        if i % 4 == 3:
            pipe.wait(idx - 2)      # the caller's contract: a host buffer is reused only after its batch has landed
        assert idx == i
    pipe.wait(n - 1)
    pipe.wait(0)                    # long forgotten: returns immediately
    assert len(pipe._ev_out) <= 4 * pipe.depth + 33
    assert torch.equal(outs[(n - 1) % 4], want[(n - 1) % 3])
    with pytest.raises(IndexError):
        pipe.wait(n)
    pipe.drain()


ALT_CONFIGS = {
    # four x4 stages (hop 256), ResBlock1
    "x4x4x4x4": {"resblock": 1, "upsample_rates": [4, 4, 4, 4], "upsample_kernel_sizes": [8, 8, 8, 8],
                 "upsample_initial_channel": 512, "resblock_kernel_sizes": [3, 7, 11],
                 "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]]},
    # HiFi-GAN V3-like: three stages, ResBlock2, other kernel sizes / dilations
    "v3like": {"resblock": 2, "upsample_rates": [8, 8, 4], "upsample_kernel_sizes": [16, 16, 8],
               "upsample_initial_channel": 256, "resblock_kernel_sizes": [3, 5, 7],
               "resblock_dilation_sizes": [[1, 2], [2, 6], [3, 12]]},
    # two kernels per stage, wide dilations
    "two_kernels": {"resblock": 1, "upsample_rates": [8, 8, 2, 2], "upsample_kernel_sizes": [16, 16, 4, 4],
                    "upsample_initial_channel": 512, "resblock_kernel_sizes": [5, 9],
                    "resblock_dilation_sizes": [[1, 2, 4], [1, 4, 9]]},
}


@pytest.mark.parametrize("name", sorted(ALT_CONFIGS))
def test_other_generator_configs_against_oracle(name):
    """The C ABI takes any `hifigan:` mapping within its stated limits (channel counts of 32 or multiples of 64,
    kernel = 2 x stride): other stage counts, kernel sizes, dilations and ResBlock2 against the oracle."""
    cfg = ALT_CONFIGS[name]
    voc, sd = build(cfg, 60, "strong")
    hop = int(np.prod(cfg["upsample_rates"]))
    mel = mel_like(2, 37, 61)
    with torch.no_grad():
        got = voc(mel.cuda())
        want = ho.hifigan_forward(sd, cfg, mel)
    assert got.shape == (2, 1, hop * 37)
    check(got, want, name)


def test_unsupported_config_is_refused_cleanly():
    cfg = dict(ho.DEFAULT_CONFIG, upsample_initial_channel=256)     # last stage would have 16 channels
    voc, _ = build(cfg, 61, "strong")
    with pytest.raises(Exception) as ei:
        with torch.no_grad():
            voc(mel_like(1, 8, 1).cuda())
    assert "channel" in str(ei.value)
