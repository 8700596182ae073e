"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "facodec_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fac_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/facodec_b200.h but not exported"
    from facodec_b200 import _lib
    assert sorted(_lib.EXPORTED) == declared
    L = _lib.load()
    assert L.fac_abi_version() == 1


def test_encode_frames_matches_reference_rule(built_lib):
    """ceil(T / 300) for the causal strided stack (encodec.py:71-78 extra padding)."""
    from facodec_b200 import _lib
    L = _lib.load()
    for T in (300, 900, 1500, 7000, 7200, 96000, 96001, 12345):
        t = T
        for s in (2, 5, 5, 6):
            t = -(-t // s)
        assert L.fac_encode_frames(T) == t


def test_no_gpu_fails_loudly(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import facodec_b200 as fb
    model = fb.build_model()
    for m in model.values():
        m.eval()
    with pytest.raises(fb.FacError):
        model.encoder(torch.zeros(1, 1, 3000))


def test_state_dict_surface_matches_reference_keys():
    import facodec_b200 as fb
    from facodec_b200 import synth
    model = fb.build_model()
    sds = synth.synth_state_dicts(0)
    for name in ("encoder", "quantizer", "decoder"):
        sd = model[name].state_dict()
        assert list(sd.keys()) == list(sds[name].keys())
        for k in sd:
            assert tuple(sd[k].shape) == tuple(sds[name][k].shape)
        model[name].load_state_dict(sds[name])
        assert torch.equal(model[name].state_dict()["%s" % list(sd.keys())[3]], sds[name][list(sd.keys())[3]])
        bad = dict(sds[name])
        bad.pop(next(iter(bad)))
        with pytest.raises(RuntimeError):
            model[name].load_state_dict(bad)
    assert sum(p.numel() for p in model.encoder.parameters()) == 36283520
    assert sum(p.numel() for p in model.decoder.parameters()) == 85536866
    with pytest.raises(NotImplementedError):
        model.encoder.train()
        model.encoder(torch.zeros(1, 1, 3000))


def test_build_model_accepts_reference_config():
    import yaml
    import facodec_b200 as fb
    cfg = yaml.safe_load("""
model_params:
  causal: True
  lstm: 2
  separate_prosody_encoder: True
  n_c_codebooks: 2
  timbre_norm: True
  DAC: {encoder_dim: 64, encoder_rates: [2, 5, 5, 6], decoder_dim: 1536, decoder_rates: [6, 5, 5, 2], sr: 24000}
""")
    m = fb.build_model(cfg["model_params"])
    assert set(m.keys()) == {"encoder", "quantizer", "decoder"}
    assert m.encoder._engine is m.decoder._engine is m.quantizer._engine
    with pytest.raises(NotImplementedError):
        fb.build_model({"causal": False})


def test_mel_buffers_close_to_torchaudio():
    import torchaudio
    from facodec_b200 import synth
    fb = synth.melscale_fbanks_htk()
    ref = torchaudio.functional.melscale_fbanks(1025, 0.0, 12000.0, 80, 24000, norm=None, mel_scale="htk")
    assert (fb - ref).abs().max() < 2e-5
    assert (synth.hann_window_periodic(1200) - torch.hann_window(1200)).abs().max() < 5e-7


def test_shard_range_partitions():
    from facodec_b200.distributed import shard_range
    for n in (0, 1, 7, 32, 256, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
