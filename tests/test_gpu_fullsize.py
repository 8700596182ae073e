"""GPU parity at the BENCHMARKED sizes, every output compared with the oracle (the CPU restatement pinned
bit-for-bit to the imported reference, tests/test_oracle.py), run live on this box's host cores:

* BASELINE configs[1]: all 32 utterances x 4 s -- all 61 440 VQ indices equal, per-utterance waveform RMS <= 1e-4
  (round 1 compared one utterance of the batch with the reference and the other 31 only with the repo's own fp32 path);
* the reference's real inference shape, B = 1 x 30 s (reconstruct.py:52), which also takes the long-sequence
  attention kernel (T' = 2400 frames);
* n_c = 1 (the quantizer's default, modules/quantize.py:375) at 4 s;
* quantize/rvq.py ResidualVQ at 65 536 frames (BASELINE configs[3] geometry) vs the oracle, every index.
"""
import os

import numpy as np
import pytest
import torch

from conftest import state_dicts

pytestmark = pytest.mark.gpu

RMS_TOL = 1e-4


def _threads():
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        pass
    return n


def _model(seed=0):
    import facodec_b200 as fb
    m = fb.build_model()
    sds = state_dicts(seed)
    for k in ("encoder", "quantizer", "decoder"):
        m[k].load_state_dict(sds[k])
        m[k].eval()
    return m, sds


def _oracle_chunks(sds, x, n_c, chunk):
    from oracle import facodec_oracle as O
    torch.set_num_threads(_threads())
    codes, ys = [[], [], []], []
    for i in range(0, x.shape[0], chunk):
        _, q, y = O.codec_forward(sds, x[i:i + chunk], n_c=n_c)
        for k in range(3):
            codes[k].append(q[5][k])
        ys.append(y)
    return [torch.cat(c) for c in codes], torch.cat(ys)


def _rms_per_utt(a, b):
    d = (a.double().cpu() - b.double()) ** 2
    return d.flatten(1).mean(1).sqrt()


def test_all_32_utterances_of_configs1_vs_oracle(built_lib):
    import facodec_b200 as fb
    from facodec_b200 import synth
    m, sds = _model(0)
    codec = fb.Codec(m)
    x = synth.synth_waves(32, 96000)
    y, codes, _ = codec.forward(x.cuda(), n_c=2)
    torch.cuda.synchronize()
    ocodes, oy = _oracle_chunks(sds, x, 2, 4)
    total = 0
    for name, a, b in zip(("codes_p", "codes_c", "codes_r"), codes, ocodes):
        a = a.cpu()
        assert a.shape == b.shape
        nbad = int((a != b).sum())
        assert nbad == 0, f"{name}: {nbad} of {b.numel()} indices differ from the oracle"
        total += b.numel()
    assert total == 61440
    r = _rms_per_utt(y, oy)
    assert float(r.max()) <= RMS_TOL, f"per-utterance waveform RMS {r.tolist()}"
    print(f"FULLSIZE B=32: 61440/61440 indices equal, waveform RMS max {float(r.max()):.3e} mean {float(r.mean()):.3e}")


@pytest.mark.parametrize("seed", [3, 11])
def test_other_weight_sets_at_4s(seed, built_lib):
    """The benchmark geometry with OTHER synthetic checkpoints (weight seeds 3 and 11; every fixture and the B = 32 test use
    seed 0 / 1): 8 utterances x 4 s, all 15 360 indices against the oracle, per-utterance RMS <= 1e-4."""
    import facodec_b200 as fb
    from facodec_b200 import synth
    m, sds = _model(seed)
    codec = fb.Codec(m)
    x = synth.synth_waves(8, 96000, seed=500 + seed)
    y, codes, _ = codec.forward(x.cuda(), n_c=2)
    torch.cuda.synchronize()
    ocodes, oy = _oracle_chunks(sds, x, 2, 4)
    total = 0
    for name, a, b in zip(("codes_p", "codes_c", "codes_r"), codes, ocodes):
        nbad = int((a.cpu() != b).sum())
        assert nbad == 0, f"seed {seed} {name}: {nbad} of {b.numel()} indices differ from the oracle"
        total += b.numel()
    assert total == 15360
    r = _rms_per_utt(y, oy)
    assert float(r.max()) <= RMS_TOL, f"per-utterance waveform RMS {r.tolist()}"


def test_b1_30s_reference_inference_shape(built_lib):
    """reconstruct.py:52 crops to 30 s: B = 1 x 720 000 samples, 2400 frames (long-sequence attention path)."""
    import facodec_b200 as fb
    from facodec_b200 import synth
    m, sds = _model(0)
    codec = fb.Codec(m)
    x = synth.synth_waves(1, 720000, seed=31)
    y, codes, timbre = codec.forward(x.cuda(), n_c=2)
    torch.cuda.synchronize()
    ocodes, oy = _oracle_chunks(sds, x, 2, 1)
    for a, b in zip(codes, ocodes):
        assert torch.equal(a.cpu(), b), f"{int((a.cpu() != b).sum())} of {b.numel()} indices differ"
    r = _rms_per_utt(y, oy)
    assert float(r.max()) <= RMS_TOL, f"waveform RMS {float(r.max())}"
    # the three-call surface gives the same bits as the fused call at this length
    z = m.encoder(x.cuda())
    q = m.quantizer(z, x.cuda(), n_c=2, return_codes=True)
    assert torch.equal(m.decoder(q[0]), y)
    for a, b in zip(q[5], codes):
        assert torch.equal(a, b)


def test_attention_long_sequence_kernel_matches_stored_scores(built_lib):
    """The recomputing attention kernel (sequences beyond the shared-memory score block) against the stored-score kernel on
    the same 4 s batch: timbre within fp32 round-off, codes identical."""
    import facodec_b200 as fb
    from facodec_b200 import synth
    m, sds = _model(0)
    codec = fb.Codec(m)
    eng = codec.engine
    x = synth.synth_waves(3, 24000, seed=5).cuda()
    y0, c0, t0 = codec.forward(x, n_c=2)
    try:
        eng.set_option("attention_stream", 1, torch.device("cuda:0"))
        y1, c1, t1 = codec.forward(x, n_c=2)
    finally:
        eng.set_option("attention_stream", 0, torch.device("cuda:0"))
    torch.cuda.synchronize()
    assert float((t0 - t1).abs().max()) <= 2e-6 * max(1.0, float(t0.abs().max()))
    for a, b in zip(c0, c1):
        assert torch.equal(a, b)
    # the two attention kernels differ by fp32 round-off in `timbre`; downstream of the VQ the k = 7 convs run one fp16 pass
    # (decoder_conv7_fp16), where a 1e-6 input change re-rounds operands: the waveform moves by the class's own noise
    # (1.4e-5 RMS against the oracle), so the bound is that noise, not fp32 round-off
    assert float(((y0 - y1).double() ** 2).mean().sqrt()) <= 3e-5


def test_n_c_1_at_4s_vs_oracle(built_lib):
    import facodec_b200 as fb
    from facodec_b200 import synth
    m, sds = _model(0)
    codec = fb.Codec(m)
    x = synth.synth_waves(4, 96000, seed=77)
    y, codes, _ = codec.forward(x.cuda(), n_c=1)
    torch.cuda.synchronize()
    ocodes, oy = _oracle_chunks(sds, x, 1, 4)
    assert tuple(codes[1].shape) == (4, 1, 320)
    for a, b in zip(codes, ocodes):
        assert torch.equal(a.cpu(), b)
    assert float(_rms_per_utt(y, oy).max()) <= RMS_TOL


def test_rvq_65536_frames_vs_oracle(built_lib):
    """quantize/rvq.py ResidualVQ, BASELINE configs[3] geometry (4 quantizers x 1024 entries, 1024 -> 8) at 2^16 frames:
    every index equal to the oracle's, quantized output within 1e-5."""
    import facodec_b200 as fb
    from oracle import facodec_oracle as O
    torch.set_num_threads(_threads())
    rvq = fb.ResidualVQ(num_quantizers=4, codebook_size=10, dim=1024, codebook_dim=8, commitment=0.25).eval()
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(64, 1024, 1024, generator=g)
    layers = []
    for i in range(4):
        layers.append(dict(in_w=rvq._folded(i, "in_proj"), in_b=rvq._p[f"layers/{i}/in_proj/bias"].detach(),
                           out_w=rvq._folded(i, "out_proj"), out_b=rvq._p[f"layers/{i}/out_proj/bias"].detach(),
                           codebook=rvq._p[f"layers/{i}/_codebook/weight"].detach()))
    with torch.no_grad():
        qo, io, _, _ = O.fvq_residual_vq(layers, x)
    q, idx, _, _ = rvq(x.cuda(), return_all=False)
    nbad = int((idx.cpu() != io).sum())
    assert nbad == 0, f"{nbad} of {io.numel()} indices differ"
    assert float((q.cpu() - qo).abs().max()) <= 1e-5
