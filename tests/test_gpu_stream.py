"""Streaming: the causal encoder / decoder fed in chunks through fac_stream_* must give the results
of ONE offline call (dac/model/dac.py:103-104, :164-165) - checked against the engine's own offline path and against the
committed golden fixture made from the imported reference (b1_t96000: 4 s, 320 frames).

Bars: latents within Z_RTOL of the offline latents (the chunked windows run the same kernels on the same values; only a
kernel-plan change with the window length could re-associate fp32 sums), waveform RMS <= 1e-4 (north_star), codes of the
chunked latents bit-exact against the golden codes.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, case_inputs, load_golden
from test_gpu_parity import RMS_TOL, Z_RTOL, model_for, rms

pytestmark = pytest.mark.gpu


def chunks_of(total, sizes):
    out, pos, i = [], 0, 0
    while pos < total:
        n = min(sizes[i % len(sizes)], total - pos)
        out.append((pos, n))
        pos += n
        i += 1
    return out


@pytest.mark.parametrize("sizes", [[3000, 300, 9000, 24000, 600], [30000], [4500, 1500]])
def test_stream_equals_offline(sizes, built_lib):
    import facodec_b200 as fb
    m = model_for(1)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(77)
    B, T = 3, 60000
    x = (torch.randn(B, 1, T, generator=g) * 0.1).to(dev)
    z_off = m.encoder(x)
    y_off = m.decoder(z_off)
    with fb.CodecStream(m, B) as s:
        zs = [s.encode(x[:, :, p:p + n].contiguous()) for p, n in chunks_of(T, sizes)]
        z_st = torch.cat(zs, dim=2)
        ys = [s.decode(z_off[:, :, p // 300:(p + n) // 300].contiguous()) for p, n in chunks_of(T, sizes)]
        y_st = torch.cat(ys, dim=2)
    torch.cuda.synchronize()
    assert z_st.shape == z_off.shape and y_st.shape == y_off.shape
    dz = float((z_st - z_off).abs().max()) / float(z_off.abs().max())
    assert dz <= Z_RTOL, dz
    assert rms(y_st, y_off) <= RMS_TOL
    # the engine runs the same kernels on the same values: report (not require) bit-equality
    print("stream vs offline: z max rel %.3g bit-equal %s | y rms %.3g bit-equal %s"
          % (dz, bool(torch.equal(z_st, z_off)), rms(y_st, y_off), bool(torch.equal(y_st, y_off))))


def test_stream_against_reference_golden(built_lib):
    """4 s utterance in 0.25 s chunks == the reference's offline z / codes / waveform (golden fixture)."""
    import facodec_b200 as fb
    name = "b1_t96000"
    c = GOLDEN_CASES[name]
    gold = load_golden(name)
    m = model_for(c["wseed"])
    x, kw = case_inputs(c)
    dev = torch.device("cuda:0")
    xd = x.to(dev)
    B, _, T = xd.shape
    with fb.CodecStream(m, B) as s:
        z_st = torch.cat([s.encode(xd[:, :, p:p + n].contiguous()) for p, n in chunks_of(T, [6000])], dim=2)
        zg = torch.as_tensor(gold["z"])
        assert float((z_st.cpu() - zg).abs().max()) <= Z_RTOL * float(zg.abs().max())
        # per-frame VQ on the streamed latents (timbre branch sees the whole wave, as offline)
        q = m.quantizer(z_st, xd, n_c=c["n_c"], return_codes=True, **{k: v.to(dev) for k, v in kw.items()})
        codes = q[5]
        zq = q[0]
        y_st = torch.cat([s.decode(zq[:, :, p:p + n].contiguous()) for p, n in chunks_of(zq.shape[2], [20])], dim=2)
    torch.cuda.synchronize()
    gy = torch.as_tensor(gold["y"])
    assert y_st.shape == gy.shape
    assert rms(y_st, gy) <= RMS_TOL
    for k, cg in zip(("codes_p", "codes_c", "codes_r"), codes):
        assert np.array_equal(cg.cpu().numpy(), gold[k]), k


def test_stream_error_paths(built_lib):
    import facodec_b200 as fb
    m = model_for(1)
    dev = torch.device("cuda:0")
    with fb.CodecStream(m, 1) as s:
        with pytest.raises(fb.FacError):
            s.encode(torch.zeros(1, 1, 2700, device=dev))      # first chunk < 3000
        with pytest.raises(fb.FacError):
            s.encode(torch.zeros(1, 1, 3100, device=dev))      # not a multiple of 300
        with pytest.raises(fb.FacError):
            s.decode(torch.zeros(1, 1024, 4, device=dev))      # first chunk < 10 frames
        with pytest.raises(fb.FacError):
            s.encode(torch.zeros(1, 1, 3000))                  # CPU tensor
    with pytest.raises(fb.FacError):
        fb.CodecStream(m, 33)
