"""Drop-in Postnet of the acoustic model — the step right before the vocoder (SURVEY.md §8 f, N2).

Mirrors e2e_tts/models/acoustic/unsupervised_fastspeech2/layers.py:507-563 (class Postnet) and its ConvNorm
(sublayers.py:72-103) as seen by the caller (model.py:60-63,137,188: `postnet_output = self.postnet(output) + output`):

    postnet = Postnet(n_channels=80, config=config["postnet"])     # embedding_dim 512, conv_layers 5, kernel_size 5
    postnet.load_state_dict(...)                                   # convolutions.<i>.0.conv.{weight,bias},
    y = postnet.eval().to("cuda")(x)                               # convolutions.<i>.1.{weight,bias,running_*}
                                                                   # x, y: [B, T, 80] fp32

Inference only (eval mode: BatchNorm1d uses its running statistics and is folded into the convolution once, dropout is
the identity).  The convolutions run in the same tcgen05 implicit-GEMM kernel as the vocoder, with a tanh epilogue; the
input is already channels-last, so the reference's two transposes disappear.  CUDA only, no fallback."""
from __future__ import annotations

import ctypes
from typing import Dict

import torch
import torch.nn as nn

from . import _native


class ConvNorm(nn.Module):
    """Parameter holder with the reference's naming (sublayers.py:72-103): `.conv.weight`, `.conv.bias`."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 1, stride: int = 1, padding=None,
                 dilation: int = 1, bias: bool = True, w_init_gain: str = "linear") -> None:
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))


class Postnet(nn.Module):
    """layers.py:507-563 on B200."""

    def __init__(self, n_channels: int, config: dict) -> None:
        super().__init__()
        self.n_channels = int(n_channels)
        self.embedding_dim = int(config["embedding_dim"])
        self.kernel_size = int(config["kernel_size"])
        self.n_layers = int(config["conv_layers"])
        pad = int((self.kernel_size - 1) / 2)
        self.convolutions = nn.ModuleList()
        for i in range(self.n_layers):
            cin = self.n_channels if i == 0 else self.embedding_dim
            cout = self.n_channels if i == self.n_layers - 1 else self.embedding_dim
            gain = "linear" if i == self.n_layers - 1 else "tanh"
            self.convolutions.append(nn.Sequential(
                ConvNorm(cin, cout, kernel_size=self.kernel_size, stride=1, padding=pad, dilation=1, w_init_gain=gain),
                nn.BatchNorm1d(cout)))
        self._handle = None
        self._handle_device = None
        self._loaded_version = None
        self._workspaces: Dict[tuple, torch.Tensor] = {}

    def folded(self, i: int):
        """Conv weight / bias of layer i with its eval-mode BatchNorm1d folded in (fp32)."""
        conv, bn = self.convolutions[i][0].conv, self.convolutions[i][1]
        w = conv.weight.detach().float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        g = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        return w * g[:, None, None], (b - bn.running_mean.detach().float()) * g + bn.bias.detach().float()

    def _fingerprint(self, device) -> tuple:
        ts = list(self.parameters()) + list(self.buffers())
        return (str(device),) + tuple((id(t), t._version) for t in ts)

    def _sync_native(self, device: torch.device) -> None:
        L = _native.lib()
        if self._handle is not None and self._handle_device != device:
            L.e2e_postnet_destroy(self._handle)
            self._handle = None
            self._workspaces.clear()
        if self._handle is None:
            h = ctypes.c_void_p()
            _native.check(L.e2e_postnet_create(self.n_channels, self.embedding_dim, self.n_layers, self.kernel_size,
                                               ctypes.byref(h)), "e2e_postnet_create")
            self._handle, self._handle_device, self._loaded_version = h, device, None
        fp = self._fingerprint(device)
        if self._loaded_version == fp:
            return
        for i in range(self.n_layers):
            w, b = self.folded(i)
            w, b = w.cpu().contiguous(), b.cpu().contiguous()
            _native.check(L.e2e_postnet_load_layer(self._handle, i,
                                                   ctypes.cast(w.data_ptr(), ctypes.POINTER(ctypes.c_float)), w.numel(),
                                                   ctypes.cast(b.data_ptr(), ctypes.POINTER(ctypes.c_float)), b.numel()),
                          "e2e_postnet_load_layer(%d)" % i)
        self._loaded_version = fp

    def forward(self, x: torch.Tensor, add_input: bool = False) -> torch.Tensor:
        """layers.py:556-563.  x: [B, T, n_channels] fp32 CUDA.  add_input=True also adds x (the caller's
        `postnet(output) + output`, model.py:188) in the last kernel."""
        if self.training:
            raise RuntimeError("e2e_tts_b200.Postnet is inference-only (eval-mode BatchNorm, no dropout): call .eval()")
        if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[2] != self.n_channels:
            raise ValueError("expected a [B, T, %d] tensor, got %s" % (self.n_channels, tuple(getattr(x, "shape", ()))))
        if not x.is_cuda:
            raise RuntimeError("e2e_tts_b200.Postnet runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dtype != torch.float32:
            raise ValueError("expected float32 input, got %s" % x.dtype)
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("e2e_tts_b200.Postnet is inference-only; call it under torch.no_grad()")
        B, T, _ = x.shape
        if B == 0 or T == 0:
            return x.new_zeros(x.shape)
        xc = x.contiguous()  # layers.py:557
        with torch.cuda.device(x.device):
            self._sync_native(x.device)
            key = (B, T, str(x.device))
            ws = self._workspaces.get(key)
            if ws is None:
                nbytes = int(_native.lib().e2e_postnet_workspace_bytes(self._handle, B, T))
                if len(self._workspaces) >= 4:
                    self._workspaces.clear()
                ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
                self._workspaces[key] = ws
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            out = torch.empty_like(xc)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            rc = _native.lib().e2e_postnet_forward(self._handle, xc.data_ptr(), B, T, 1 if add_input else 0,
                                                   out.data_ptr(), ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()), stream)
            _native.check(rc, "e2e_postnet_forward")
        return out

    def __getstate__(self):
        d = self.__dict__.copy()   # copies / pickles start without a native handle
        d.update(_handle=None, _handle_device=None, _loaded_version=None, _workspaces={})
        return d

    def __del__(self):
        try:
            if self._handle is not None:
                _native.lib().e2e_postnet_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
