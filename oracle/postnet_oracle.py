"""TEST INFRASTRUCTURE ONLY (never imported by e2e_tts_b200/): CPU restatement of the acoustic model's Postnet in
eval mode (e2e_tts/models/acoustic/unsupervised_fastspeech2/layers.py:507-563 + ConvNorm, sublayers.py:72-103),
pinned to the unmodified reference class by oracle/make_golden.py (tests/golden/postnet_default.npz)."""
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CONFIG = {"embedding_dim": 512, "conv_layers": 5, "kernel_size": 5}   # model_config.yaml:71-74


def make_state_dict(n_channels: int, config: dict, seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic checkpoint with the reference's key names; non-trivial BatchNorm statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    H, k, n = config["embedding_dim"], config["kernel_size"], config["conv_layers"]
    for i in range(n):
        cin = n_channels if i == 0 else H
        cout = n_channels if i == n - 1 else H
        p = "convolutions.%d." % i
        sd[p + "0.conv.weight"] = torch.randn(cout, cin, k, generator=g) * (1.3 / np.sqrt(cin * k))
        sd[p + "0.conv.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * 0.1
        sd[p + "1.weight"] = 0.7 + 0.6 * torch.rand(cout, generator=g)
        sd[p + "1.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * 0.2
        sd[p + "1.running_mean"] = (torch.rand(cout, generator=g) * 2 - 1) * 0.3
        sd[p + "1.running_var"] = 0.5 + torch.rand(cout, generator=g)
        sd[p + "1.num_batches_tracked"] = torch.tensor(100)
    return sd


def postnet_forward(sd: dict, config: dict, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """x [B, T, C] -> [B, T, C] (layers.py:556-563, self.training == False so F.dropout is the identity)."""
    n, k = config["conv_layers"], config["kernel_size"]
    x = x.to(dtype).contiguous().transpose(1, 2)                                        # :557
    for i in range(n):
        p = "convolutions.%d." % i
        y = F.conv1d(x, sd[p + "0.conv.weight"].to(dtype), sd[p + "0.conv.bias"].to(dtype), padding=int((k - 1) / 2))
        y = F.batch_norm(y, sd[p + "1.running_mean"].to(dtype), sd[p + "1.running_var"].to(dtype),
                         sd[p + "1.weight"].to(dtype), sd[p + "1.bias"].to(dtype), training=False, eps=1e-5)
        x = torch.tanh(y) if i < n - 1 else y                                           # :558-560
    return x.contiguous().transpose(1, 2)                                               # :561
