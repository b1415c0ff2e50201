"""Drop-in HiFi-GAN generator for the e2e-tts synthesis path, running on hand-written sm_100a kernels.

Mirrors the reference's public contract (e2e_tts/models/vocoder/generator.py:13-62, layers.py:10-69,
function.py:4-17) as seen by its callers (e2e_tts/src/api/utils.py:53-56,144-145 and
e2e_tts/src/tools/tools_for_model.py:45-54,94-140):

    voc = HifiGan(config["models"]["hifigan"]); voc.load_state_dict(ckpt["state_dict"]); voc.eval().to("cuda")
    wav = voc(mel.transpose(1, 2))            # [B, 80, T] fp32 (any strides) -> [B, 1, 256*T] fp32

The module owns parameters with the reference's state-dict names (`*.weight_g`, `*.weight_v`, `*.bias`; or
`*.weight` after remove_weight_norm()), so checkpoints load unchanged.  forward() is inference-only: it folds
weight-norm once, packs bf16 GEMM-layout weights through the C ABI, and enqueues the CUDA kernels on the current
torch stream without synchronising.  There is no CPU path and no fallback: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List

import torch
import torch.nn as nn

from . import _native

LRELU_SLOPE = 0.1  # generator.py:10, layers.py:7


def _operand_dtype(name) -> str:
    import os
    name = name or os.environ.get("E2E_OPERAND_DTYPE") or "bf16"
    name = {"bfloat16": "bf16", "float16": "fp16", "half": "fp16", "f16": "fp16"}.get(str(name).lower(), str(name).lower())
    if name not in ("bf16", "fp16"):
        raise ValueError("operand_dtype must be 'bf16' or 'fp16', got %r" % (name,))
    return name


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """function.py:16-17."""
    return int((kernel_size * dilation - dilation) / 2)


def init_weights(m, mean: float = 0.0, std: float = 0.01) -> None:
    """function.py:4-7.  (On a weight-normed conv this writes the derived `.weight`, which the reference's
    pre-forward hook overwrites, so it has no effect there either; kept for API parity.)"""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1 and hasattr(m, "weight") and isinstance(m.weight, torch.Tensor):
        m.weight.data.normal_(mean, std)


def apply_weight_norm(m) -> None:
    """function.py:10-13: `torch.nn.utils.weight_norm` on every Conv* module (used with `module.apply`)."""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.utils.weight_norm(m)


class _WNConv(nn.Module):
    """Parameter holder with the reference's weight-norm naming.  `transposed` selects ConvTranspose1d's
    [C_in, C_out, k] weight layout (weight-norm dim 0 is then C_in)."""

    def __init__(self, cin: int, cout: int, k: int, transposed: bool = False) -> None:
        super().__init__()
        self.cin, self.cout, self.k, self.transposed = cin, cout, k, transposed
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        v = torch.empty(shape)
        # PyTorch's default Conv init (kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))): the reference's
        # effective random init.  fan_in follows torch: weight.size(1) * k.
        fan_in = shape[1] * k
        bound = 1.0 / math.sqrt(fan_in)
        nn.init.uniform_(v, -bound, bound)
        self.weight_v = nn.Parameter(v)
        self.weight_g = nn.Parameter(v.detach().reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1).clone())
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))

    def folded_weight(self) -> torch.Tensor:
        """w = g * v / ||v||, norm over all dims but 0 (torch._weight_norm, SURVEY.md §8 a'4), in fp32."""
        if "weight" in self._parameters:
            return self.weight.detach().float()
        v = self.weight_v.detach().float()
        g = self.weight_g.detach().float()
        n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
        return v * (g / n)

    def remove_weight_norm(self) -> None:
        if "weight" in self._parameters:
            raise ValueError("weight_norm of this layer was already removed")
        w = self.folded_weight()
        dev = self.weight_v.device
        del self._parameters["weight_g"]
        del self._parameters["weight_v"]
        self.weight = nn.Parameter(w.to(dev))


class _ResBlockBase(nn.Module):
    """Shared native plumbing of the standalone ResBlock1 / ResBlock2 modules (inside HifiGan the same convolutions run
    through the generator's fused launches; these classes are then parameter holders)."""

    _kind = 1

    def _rb_layers(self) -> List[tuple]:
        raise NotImplementedError

    def _rb_init(self) -> None:
        self._rb_handle = None
        self._rb_device = None
        self._rb_version = None
        self._rb_ws: Dict[tuple, torch.Tensor] = {}

    def _rb_sync(self, device: torch.device) -> None:
        L = _native.lib()
        if self._rb_handle is not None and self._rb_device != device:
            L.e2e_resblock_destroy(self._rb_handle)
            self._rb_handle = None
            self._rb_ws.clear()
        if self._rb_handle is None:
            h = ctypes.c_void_p()
            dil = (ctypes.c_int32 * len(self.dilation))(*[int(d) for d in self.dilation])
            _native.check(L.e2e_resblock_create(self._kind, int(self.channels), int(self.kernel_size), dil,
                                                len(self.dilation), ctypes.byref(h)), "e2e_resblock_create")
            self._rb_handle, self._rb_device, self._rb_version = h, device, None
        fp = tuple((id(p), p._version) for p in self.parameters())
        if self._rb_version == fp:
            return
        for name, layer in self._rb_layers():
            w = layer.folded_weight().cpu().contiguous()
            b = layer.bias.detach().float().cpu().contiguous()
            _native.check(L.e2e_resblock_load_layer(self._rb_handle, name.encode(),
                                                    ctypes.cast(w.data_ptr(), ctypes.POINTER(ctypes.c_float)), w.numel(),
                                                    ctypes.cast(b.data_ptr(), ctypes.POINTER(ctypes.c_float)), b.numel()),
                          "e2e_resblock_load_layer(%s)" % name)
        self._rb_version = fp

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, channels, T] float32 CUDA tensor (any strides) -> [B, channels, T] float32.  Inference only."""
        if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[1] != self.channels:
            raise ValueError("expected a [B, %d, T] tensor, got %s" % (self.channels, tuple(getattr(x, "shape", ()))))
        if not x.is_cuda:
            raise RuntimeError("e2e_tts_b200 residual blocks run on CUDA (sm_100a) only; there is no CPU path")
        if x.dtype != torch.float32:
            raise ValueError("expected float32 input, got %s" % x.dtype)
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("e2e_tts_b200 residual blocks are inference-only; call them under torch.no_grad()")
        B, C, T = x.shape
        if B == 0 or T == 0:
            return x.new_zeros((B, C, T))
        with torch.cuda.device(x.device):
            self._rb_sync(x.device)
            L = _native.lib()
            key = (B, T, str(x.device))
            ws = self._rb_ws.get(key)
            if ws is None:
                nbytes = int(L.e2e_resblock_workspace_bytes(self._rb_handle, B, T))
                if len(self._rb_ws) >= 4:
                    self._rb_ws.clear()
                ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
                self._rb_ws[key] = ws
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            out = torch.empty((B, C, T), dtype=torch.float32, device=x.device)
            rc = L.e2e_resblock_forward(self._rb_handle, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), B, T,
                                        out.data_ptr(), ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()),
                                        torch.cuda.current_stream(x.device).cuda_stream)
            _native.check(rc, "e2e_resblock_forward")
        return out

    def __getstate__(self):
        # native handles belong to this instance: copies / pickles start without one and rebuild it on first use
        d = self.__dict__.copy()
        d.update(_rb_handle=None, _rb_device=None, _rb_version=None, _rb_ws={})
        return d

    def __del__(self):
        try:
            if getattr(self, "_rb_handle", None) is not None:
                _native.lib().e2e_resblock_destroy(self._rb_handle)
                self._rb_handle = None
        except Exception:
            pass


class ResBlock1(_ResBlockBase):
    """layers.py:10-46 - 3 x [lrelu -> conv(k, d_i) -> lrelu -> conv(k, 1) -> + x]."""

    _kind = 1

    def __init__(self, channels: int, kernel_size: int = 3, dilation=(1, 3, 5)) -> None:
        super().__init__()
        self.channels, self.kernel_size, self.dilation = channels, kernel_size, tuple(dilation)
        self.convs1 = nn.ModuleList([_WNConv(channels, channels, kernel_size) for _ in self.dilation])
        self.convs2 = nn.ModuleList([_WNConv(channels, channels, kernel_size) for _ in self.dilation])
        self._rb_init()

    def _rb_layers(self) -> List[tuple]:
        return ([("convs1.%d" % m, l) for m, l in enumerate(self.convs1)] +
                [("convs2.%d" % m, l) for m, l in enumerate(self.convs2)])

    def remove_weight_norm(self) -> None:
        for layer in self.convs1:
            layer.remove_weight_norm()
        for layer in self.convs2:
            layer.remove_weight_norm()


class ResBlock2(_ResBlockBase):
    """layers.py:49-69 - 2 x [lrelu -> conv(k, d_i) -> + x]."""

    _kind = 2

    def __init__(self, channels: int, kernel_size: int = 3, dilation=(1, 3)) -> None:
        super().__init__()
        # the reference builds exactly two convs from dilation[0] and dilation[1] (layers.py:52-57)
        self.channels, self.kernel_size, self.dilation = channels, kernel_size, tuple(dilation)[:2]
        self.convs = nn.ModuleList([_WNConv(channels, channels, kernel_size) for _ in self.dilation])
        self._rb_init()

    def _rb_layers(self) -> List[tuple]:
        return [("convs.%d" % m, l) for m, l in enumerate(self.convs)]

    def remove_weight_norm(self) -> None:
        for layer in self.convs:
            layer.remove_weight_norm()


class HifiGan(nn.Module):
    """generator.py:13-62 on B200.  `config` is the `hifigan:` mapping of model_config.yaml:75-82."""

    def __init__(self, config: dict, operand_dtype: str = None) -> None:
        """`operand_dtype` (not in the reference): "bf16" (default; what BASELINE names) or "fp16" - the format of the
        tensor-core operands and of the 16-bit activations between layers.  fp16 runs at the same speed with 11
        instead of 8 significand bits (about 8x smaller waveform error); conversions saturate at +-65504.  The
        environment variable E2E_OPERAND_DTYPE sets the default."""
        super().__init__()
        self.config = dict(config)
        self.operand_dtype = _operand_dtype(operand_dtype)
        self.num_kernels = len(config["resblock_kernel_sizes"])
        self.num_upsamples = len(config["upsample_rates"])
        c0 = int(config["upsample_initial_channel"])
        self.in_channels = 80  # hard-coded at generator.py:18
        self.conv_pre = _WNConv(self.in_channels, c0, 7)
        self._resblock_type = self._select_resblock(config)
        resblock = ResBlock1 if self._resblock_type == 1 else ResBlock2
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(config["upsample_rates"], config["upsample_kernel_sizes"])):
            self.ups.append(_WNConv(c0 // (2 ** i), c0 // (2 ** (i + 1)), int(k), transposed=True))
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(config["resblock_kernel_sizes"], config["resblock_dilation_sizes"]):
                self.resblocks.append(resblock(ch, int(k), tuple(int(x) for x in d)))
        self.post_n_fft = self._post_n_fft(config)   # 0: HiFi-GAN head; n: iSTFTNet head (conv_post -> n + 2)
        self.conv_post = _WNConv(ch, self.post_n_fft + 2 if self.post_n_fft else 1, 7)
        self.hop = 1
        for u in config["upsample_rates"]:
            self.hop *= int(u)
        self._handle = None          # e2e_voc* (created lazily on the first CUDA forward)
        self._handle_device = None
        self._loaded_version = None  # parameter-version fingerprint the packed weights correspond to
        self._param_cache = None
        self._workspaces: Dict[tuple, torch.Tensor] = {}
        self._profile_events = None  # (cudaEvent_t, cudaEvent_t) handles for the next forward (measurement hook)

    @staticmethod
    def _select_resblock(config: dict) -> int:
        return 1 if config["resblock"] == 1 else 2  # generator.py:19 (int compare)

    @staticmethod
    def _post_n_fft(config: dict) -> int:
        return 0

    # ------------------------------------------------------------------ state-dict compatibility
    def _wn_layers(self) -> List[tuple]:
        out = [("conv_pre", self.conv_pre)]
        for i, l in enumerate(self.ups):
            out.append(("ups.%d" % i, l))
        for n, rb in enumerate(self.resblocks):
            if isinstance(rb, ResBlock1):
                for m, l in enumerate(rb.convs1):
                    out.append(("resblocks.%d.convs1.%d" % (n, m), l))
                for m, l in enumerate(rb.convs2):
                    out.append(("resblocks.%d.convs2.%d" % (n, m), l))
            else:
                for m, l in enumerate(rb.convs):
                    out.append(("resblocks.%d.convs.%d" % (n, m), l))
        out.append(("conv_post", self.conv_post))
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the reference's 234-key weight_g/weight_v/bias layout (utils.py:54-55) and also the folded
        `*.weight` layout a checkpoint saved after remove_weight_norm() has."""
        sd = dict(state_dict)
        for name, layer in self._wn_layers():
            wkey = name + ".weight"
            has_wn = "weight_v" in layer._parameters
            if wkey in sd and has_wn:
                w = sd.pop(wkey).detach().float()
                sd[name + ".weight_v"] = w
                sd[name + ".weight_g"] = w.reshape(w.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            elif (name + ".weight_v") in sd and not has_wn:
                v = sd.pop(name + ".weight_v").detach().float()
                g = sd.pop(name + ".weight_g").detach().float()
                sd[wkey] = v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1))
        res = super().load_state_dict(sd, strict=strict, **kw)
        self.invalidate()
        return res

    def remove_weight_norm(self) -> None:
        """generator.py:55-62 (output unchanged; the packed weights were folded anyway)."""
        print("Removing weight norm...")
        for layer in self.ups:
            layer.remove_weight_norm()
        for layer in self.resblocks:
            layer.remove_weight_norm()
        self.conv_pre.remove_weight_norm()
        self.conv_post.remove_weight_norm()
        self.invalidate()

    # ------------------------------------------------------------------ native side
    def _native_config(self) -> _native.VocConfig:
        c = self.config
        cfg = _native.VocConfig()
        cfg.in_channels = self.in_channels
        cfg.upsample_initial_channel = int(c["upsample_initial_channel"])
        cfg.resblock = self._resblock_type
        cfg.num_upsamples = self.num_upsamples
        if self.num_upsamples > _native.E2E_MAX_UPSAMPLES or self.num_kernels > _native.E2E_MAX_KERNELS:
            raise ValueError("too many upsample stages / resblock kernels")
        for i, (u, k) in enumerate(zip(c["upsample_rates"], c["upsample_kernel_sizes"])):
            cfg.upsample_rates[i] = int(u)
            cfg.upsample_kernel_sizes[i] = int(k)
        cfg.num_kernels = self.num_kernels
        cfg.istft_n_fft = self.post_n_fft
        for j, (k, d) in enumerate(zip(c["resblock_kernel_sizes"], c["resblock_dilation_sizes"])):
            d = list(d)
            if self._resblock_type != 1:
                d = d[:2] if len(d) >= 2 else d  # ResBlock2 is built from dilation[0], dilation[1] (layers.py:52-57)
            if len(d) > _native.E2E_MAX_DILATIONS:
                raise ValueError("too many dilations")
            cfg.resblock_kernel_sizes[j] = int(k)
            cfg.num_dilations[j] = len(d)
            for m, x in enumerate(d):
                cfg.resblock_dilation_sizes[j][m] = int(x)
        return cfg

    def _version_fingerprint(self, device) -> tuple:
        # (id, in-place version counter) of every parameter; the parameter list itself is cached (walking the module
        # tree costs more than a single-utterance forward) and rebuilt whenever the set of parameters can have changed
        if self._param_cache is None:
            self._param_cache = list(self.parameters())
        return (str(device),) + tuple((id(p), p._version) for p in self._param_cache)

    def invalidate(self) -> None:
        """Forces the next forward() to re-fold weight-norm and re-pack the device weights.  Needed only after writes
        the version counters cannot see: `p.data.copy_(...)` / `p.data.normal_()` (what `init_weights` does), or
        storage shared with an optimizer that updates through `.data`.  Ordinary in-place updates (`p.copy_()`,
        optimizer steps, `load_state_dict`, `.to()`) are detected automatically."""
        self._loaded_version = None
        self._param_cache = None

    def _apply(self, fn, *a, **kw):
        self._param_cache = None
        self._loaded_version = None
        return super()._apply(fn, *a, **kw)

    def _sync_native(self, device: torch.device) -> None:
        L = _native.lib()
        if self._handle is not None and self._handle_device != device:
            L.e2e_voc_destroy(self._handle)
            self._handle = None
            self._workspaces.clear()
        if self._handle is None:
            h = ctypes.c_void_p()
            cfg = self._native_config()
            _native.check(L.e2e_voc_create(ctypes.byref(cfg), ctypes.byref(h)), "e2e_voc_create")
            _native.check(L.e2e_voc_set_operand_dtype(
                h, _native.E2E_OPERAND_FP16 if self.operand_dtype == "fp16" else _native.E2E_OPERAND_BF16),
                "e2e_voc_set_operand_dtype")
            self._handle = h
            self._handle_device = device
            self._loaded_version = None
        fp = self._version_fingerprint(device)
        if self._loaded_version == fp:
            return
        for name, layer in self._wn_layers():
            w = layer.folded_weight().cpu().contiguous()
            b = layer.bias.detach().float().cpu().contiguous()
            _native.check(
                L.e2e_voc_load_layer(self._handle, name.encode(),
                                     ctypes.cast(w.data_ptr(), ctypes.POINTER(ctypes.c_float)), w.numel(),
                                     ctypes.cast(b.data_ptr(), ctypes.POINTER(ctypes.c_float)), b.numel()),
                "e2e_voc_load_layer(%s)" % name)
        if L.e2e_voc_missing_layers(self._handle) != 0:
            raise _native.NativeError("not every generator layer received weights")
        self._loaded_version = fp

    def launches_per_forward(self) -> int:
        if self._handle is None:
            raise RuntimeError("call forward() once first")
        return int(_native.lib().e2e_voc_launches_per_forward(self._handle))

    def last_forward_was_graph(self) -> bool:
        """True if the last forward replayed a captured CUDA graph (same input / output / workspace buffers as a
        previous call - what HostPipeline and `out=` loops do) instead of enqueueing its kernels one by one."""
        return self._handle is not None and int(_native.lib().e2e_voc_last_forward_was_graph(self._handle)) == 1

    def _workspace(self, B: int, T: int, device: torch.device) -> torch.Tensor:
        key = (B, T, str(device))
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = int(_native.lib().e2e_voc_workspace_bytes(self._handle, B, T))
            if nbytes == 0:
                raise _native.NativeError("e2e_voc_workspace_bytes returned 0")
            if len(self._workspaces) >= 4:
                self._workspaces.clear()
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._workspaces[key] = ws
        return ws

    def _check_input(self, x: torch.Tensor) -> None:
        if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[1] != self.in_channels:
            raise ValueError("expected a [B, %d, T] tensor, got %s" % (self.in_channels, tuple(getattr(x, "shape", ()))))
        if not x.is_cuda:
            raise RuntimeError("e2e_tts_b200.HifiGan runs on CUDA (sm_100a) only: move the module and its input "
                               "to a B200 device; there is no CPU path")
        if x.dtype != torch.float32:
            raise ValueError("expected float32 input, got %s" % x.dtype)
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("e2e_tts_b200.HifiGan is inference-only; call it under torch.no_grad()")
        p0 = next(self.parameters())
        if p0.device != x.device:
            raise RuntimeError("module parameters are on %s but the input is on %s" % (p0.device, x.device))

    def _check_out(self, out, shape, dtype, device):
        if out is None:
            return torch.empty(shape, dtype=dtype, device=device)
        if tuple(out.shape) != tuple(shape) or out.dtype != dtype or out.device != device or not out.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor of shape %s on %s" % (dtype, tuple(shape), device))
        return out

    def forward(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """generator.py:37-53.  x: [B, 80, T] float32 CUDA tensor (non-contiguous views are fine).  `out` (not in the
        reference): an optional preallocated [B, 1, hop*T] float32 result tensor (serving loops, HostPipeline)."""
        self._check_input(x)
        B, _, T = x.shape
        if B == 0 or T == 0:
            return x.new_zeros((B, 1, self.hop * T))
        with torch.cuda.device(x.device):
            self._sync_native(x.device)
            ws = self._workspace(B, T, x.device)
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            out = self._check_out(out, (B, 1, self.hop * T), torch.float32, x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            if self._profile_events is not None:   # bench.py: time the tensor-core segment of this call
                ev0, ev1 = self._profile_events
                self._profile_events = None
                _native.check(_native.lib().e2e_voc_set_profile_events(self._handle, ev0, ev1),
                              "e2e_voc_set_profile_events")
            rc = _native.lib().e2e_voc_forward(self._handle, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2),
                                               B, T, out.data_ptr(), ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()),
                                               stream)
            _native.check(rc, "e2e_voc_forward")
        return out

    def forward_pcm16(self, x: torch.Tensor, mel_lengths=None, max_wav_value: float = 32768.0,
                      out: torch.Tensor = None) -> torch.Tensor:
        """forward() fused with the caller's post-processing (combine_audio, e2e_tts/src/api/utils.py:108-117):
        returns int16 PCM [B, hop*T] = trunc(wav * max_wav_value), zero beyond mel_lengths[b] * hop.  Half the
        device->host (and multi-GPU gather) bytes of forward().  mel_lengths: None, a sequence, or an int tensor [B]."""
        self._check_input(x)
        B, _, T = x.shape
        if B == 0 or T == 0:
            return torch.zeros((B, self.hop * T), dtype=torch.int16, device=x.device)
        lens = None
        if mel_lengths is not None:
            lens = torch.as_tensor(mel_lengths).to(device=x.device, dtype=torch.int32).contiguous()
            if lens.shape != (B,):
                raise ValueError("mel_lengths must have shape [%d], got %s" % (B, tuple(lens.shape)))
        with torch.cuda.device(x.device):
            self._sync_native(x.device)
            ws = self._workspace(B, T, x.device)
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            out = self._check_out(out, (B, self.hop * T), torch.int16, x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            rc = _native.lib().e2e_voc_forward_pcm16(self._handle, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2),
                                                     B, T, lens.data_ptr() if lens is not None else None,
                                                     float(max_wav_value), out.data_ptr(), ws_ptr,
                                                     ws.numel() - (ws_ptr - ws.data_ptr()), stream)
            _native.check(rc, "e2e_voc_forward_pcm16")
        return out

    def __getstate__(self):
        # the native handle (packed weights, plans, graphs) belongs to this instance: copy.deepcopy / pickle start
        # without one and rebuild it on their first forward
        d = self.__dict__.copy()
        d.update(_handle=None, _handle_device=None, _loaded_version=None, _param_cache=None, _workspaces={},
                 _profile_events=None)
        return d

    def __del__(self):
        try:
            if self._handle is not None:
                _native.lib().e2e_voc_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


class iSTFT(HifiGan):
    """iSTFTNet generator (reference class iSTFT, generator.py:65-119; `istft:` mapping of model_config.yaml:83-92;
    selected by load_vocoder(use_complex=True), src/tools/tools_for_model.py:45-50).  Same trunk kernels as HifiGan
    with the configured stages, then ReflectionPad1d((1, 0)), conv_post C -> gen_istft_n_fft + 2, exp / sin:

        spec, phase = g(mel)                       # each [B, n_fft/2 + 1, hop*T + 1] fp32
        wav = inverse_stft(spec, phase, n_fft, hop_size, win_size)   # e2e_tts_b200.inverse_stft, stft.py:138-148

    The reference selects its residual block with `config['resblock'] == '1'` (a STRING compare, generator.py:71), so
    the shipped config (`resblock: 1`, an int) builds ResBlock2; reproduced here so checkpoints load."""

    @staticmethod
    def _select_resblock(config: dict) -> int:
        return 1 if config["resblock"] == "1" else 2  # generator.py:71

    @staticmethod
    def _post_n_fft(config: dict) -> int:
        return int(config["gen_istft_n_fft"])  # generator.py:85

    def forward(self, x: torch.Tensor):
        """generator.py:91-109."""
        self._check_input(x)
        B, _, T = x.shape
        nb = self.post_n_fft // 2 + 1
        if B == 0 or T == 0:
            z = x.new_zeros((B, nb, self.hop * T + (1 if T else 0)))
            return z, z.clone()
        with torch.cuda.device(x.device):
            self._sync_native(x.device)
            ws = self._workspace(B, T, x.device)
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            spec = torch.empty((B, nb, self.hop * T + 1), dtype=torch.float32, device=x.device)
            phase = torch.empty_like(spec)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            rc = _native.lib().e2e_voc_forward_spec(self._handle, x.data_ptr(), x.stride(0), x.stride(1), x.stride(2),
                                                    B, T, spec.data_ptr(), phase.data_ptr(), ws_ptr,
                                                    ws.numel() - (ws_ptr - ws.data_ptr()), stream)
            _native.check(rc, "e2e_voc_forward_spec")
        return spec, phase

    def forward_pcm16(self, *a, **kw):
        raise RuntimeError("iSTFT returns (spec, phase); synthesise with inverse_stft(spec, phase, ...)")

