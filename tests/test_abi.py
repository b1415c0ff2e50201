"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares,
and the Python mirror keeps the reference's module contract.  No compute calls (no GPU here)."""
import os
import re

import pytest
import torch

import e2e_tts_b200 as pkg
from e2e_tts_b200 import _native
from oracle import hifigan_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "e2e_tts_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(e2e_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_native.SYMBOLS), "ctypes table and header disagree"
    lib = _native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.e2e_version_string()


def test_bad_arguments_are_reported_not_crashed():
    lib = _native.lib()
    assert lib.e2e_voc_create(None, None) != 0
    assert b"null" in lib.e2e_last_error_string()
    assert lib.e2e_voc_workspace_bytes(None, 1, 1) == 0
    assert lib.e2e_voc_missing_layers(None) == -1


def test_state_dict_contract_matches_reference_layout():
    voc = pkg.HifiGan(ho.DEFAULT_CONFIG)
    sd = voc.state_dict()
    want = ho.make_state_dict(ho.DEFAULT_CONFIG, 3, "strong")
    assert set(sd) == set(want) and len(sd) == 234
    for k in want:
        assert tuple(sd[k].shape) == tuple(want[k].shape), k
    voc.load_state_dict(want)                                  # the 234-key weight_g/weight_v/bias layout
    for name, layer in voc._wn_layers():
        w = ho.fold_weight_norm(want[name + ".weight_g"], want[name + ".weight_v"])
        assert torch.allclose(layer.folded_weight(), w, atol=1e-7), name
    assert len(list(voc.parameters())) == 234
    assert voc.eval() is voc


def test_remove_weight_norm_and_folded_checkpoints(capsys):
    cfg = ho.DEFAULT_CONFIG
    want = ho.make_state_dict(cfg, 4, "strong")
    a = pkg.HifiGan(cfg)
    a.load_state_dict(want)
    folded_before = {n: l.folded_weight().clone() for n, l in a._wn_layers()}
    a.remove_weight_norm()
    assert "Removing weight norm" in capsys.readouterr().out   # generator.py:56
    sd = a.state_dict()
    assert "conv_pre.weight" in sd and "conv_pre.weight_g" not in sd and len(sd) == 156
    for n, l in a._wn_layers():
        assert torch.equal(l.folded_weight(), folded_before[n])
    b = pkg.HifiGan(cfg)                                       # a folded checkpoint loads into a fresh module
    b.load_state_dict(sd)
    for n, l in b._wn_layers():
        assert torch.allclose(l.folded_weight(), folded_before[n], atol=1e-6), n
    a2 = pkg.HifiGan(cfg)
    a2.remove_weight_norm()
    a2.load_state_dict(want)                                   # and an unfolded one into a folded module
    for n, l in a2._wn_layers():
        assert torch.allclose(l.folded_weight(), folded_before[n], atol=1e-6), n
    with pytest.raises(ValueError):
        a.remove_weight_norm()


def test_resblock2_selection_like_the_reference():
    cfg = dict(ho.DEFAULT_CONFIG)
    cfg["resblock"] = "1"                                      # generator.py:19 compares with the int 1
    voc = pkg.HifiGan(cfg)
    assert isinstance(voc.resblocks[0], pkg.ResBlock2) and len(voc.resblocks[0].convs) == 2
    assert set(voc.state_dict()) == set(ho.make_state_dict({**cfg, "resblock": 2}, 0))


def test_no_cpu_fallback():
    voc = pkg.HifiGan(ho.DEFAULT_CONFIG)
    with pytest.raises(RuntimeError, match="CUDA"):
        voc(torch.zeros(1, 80, 8))
    with pytest.raises(ValueError):
        voc(torch.zeros(1, 81, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            pkg.TorchSTFT().mel_spectrogram(torch.zeros(1, 4096))


def test_stft_public_attributes_and_helpers():
    s = pkg.TorchSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0)   # positional, dataloader.py:82-86
    assert (s.sampling_rate, s.hop_length, s.n_mel_channels, s.filter_length, s.win_length) == (22050, 256, 80, 1024, 1024)
    assert s.stft_pad == (384, 384) and tuple(s.mel_basis.shape) == (80, 513) and s.window.shape == (1024,)
    assert pkg.get_padding(11, 5) == 25 and pkg.get_padding(7, 1) == 3
    x = torch.tensor([1e-7, 0.5, 2.0])
    assert torch.allclose(pkg.dynamic_range_compression(x), torch.log(torch.clamp(x, min=1e-5)))
    assert torch.allclose(pkg.dynamic_range_decompression(pkg.dynamic_range_compression(x[1:])), x[1:])
    with pytest.raises(NotImplementedError):
        s.mel_spectrogram(torch.zeros(1, 4096), center=True)


def test_istft_generator_contract():
    """class iSTFT (generator.py:65-119): state-dict names, ResBlock selection quirk, head shape, no CPU path."""
    from oracle import hifigan_oracle as ho
    g = pkg.iSTFT(ho.ISTFT_CONFIG)
    sd = ho.make_state_dict(ho.ISTFT_CONFIG, 1, "strong")
    assert set(g.state_dict().keys()) == set(sd.keys())
    g.load_state_dict(sd)
    assert type(g.resblocks[0]).__name__ == "ResBlock2" and len(g.resblocks[0].convs) == 2
    assert g.conv_post.weight_v.shape == (18, 128, 7) and g.hop == 64
    assert type(pkg.iSTFT(dict(ho.ISTFT_CONFIG, resblock="1")).resblocks[0]).__name__ == "ResBlock1"
    with pytest.raises(RuntimeError):
        g(torch.zeros(1, 80, 4))
    with pytest.raises(RuntimeError):
        pkg.inverse_stft(torch.zeros(1, 9, 5), torch.zeros(1, 9, 5), 16, 4, 16)


def test_modules_copy_and_pickle_without_their_native_handles():
    """copy.deepcopy / pickle of a module must not duplicate its native handle (double free): copies start without one."""
    import copy
    import pickle
    import e2e_tts_b200 as pkg
    from e2e_tts_b200 import synthetic as sy
    voc = pkg.HifiGan(sy.DEFAULT_CONFIG)
    voc._handle = 123                      # stand-in for a live e2e_voc*
    try:
        clone = copy.deepcopy(voc)
        assert clone._handle is None and clone._workspaces == {} and voc._handle == 123
        assert list(clone.state_dict()) == list(voc.state_dict())
    finally:
        voc._handle = None
    stft = pkg.TorchSTFT()
    stft._handles["cuda:0"] = object()
    assert copy.deepcopy(stft)._handles == {}
    stft._handles.clear()
    rb = pickle.loads(pickle.dumps(pkg.ResBlock1(64)))
    assert rb._rb_handle is None and len(rb.convs1) == 3
