"""Training-side reconstruction loss of the reference, forward only (losses.py:65-89), on the B200 front-end kernels.

``reconstruction_loss(x, G_x)`` mirrors losses.py:65-89: ``100 * mse(x, G_x)`` plus, for s = 64 ... 2048, ``l1 + sqrt(s/2) * l2``
between 64-band torchaudio mel spectrograms (``sample_rate=16000, n_fft=max(s,512), win_length=s, hop_length=s//4``).  Per
scale the two signals' frames go through ONE tensor-core GEMM against the window-folded DFT basis, a per-frame kernel
forms both mel spectra and the frame's share of the two terms, and fp64 sums in a fixed order give the scalars
(``fac_reconstruction_loss``).  Forward only -- no autograd graph is attached to the result (SURVEY.md 8f rank 3 stops
at the loss value); a CPU tensor or a missing library raises ``FacError``.
"""
import torch

from . import _lib
from .modules import Engine, _ptr, _stream

LAMBDA_WAV = 100            # losses.py:58

_ENGINE = None


def _engine(device):
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine()
    _ENGINE._ensure(device)
    return _ENGINE


def _flat(t, name):
    if not torch.is_tensor(t) or t.device.type != "cuda":
        raise _lib.FacError(f"reconstruction_loss: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dim() == 3 and t.shape[1] == 1:
        t = t[:, 0]
    if t.dim() == 1:
        t = t[None]
    if t.dim() != 2:
        raise _lib.FacError(f"reconstruction_loss: {name} must be [B, T] or [B, 1, T], got {tuple(t.shape)}")
    return t.float().contiguous()


def reconstruction_loss(x, G_x, eps=1e-7, return_terms=False):
    """losses.py:65-89.  x, G_x: [B, 1, T] or [B, T] on the GPU, T > 1024.  Returns the 0-d loss (and, with
    ``return_terms``, the 13 components: mse, then (l1, l2) for s = 64, 128, ..., 2048)."""
    if eps != 1e-7:
        raise _lib.FacError("reconstruction_loss: eps is fixed at the reference's 1e-7")
    a, b = _flat(x, "x"), _flat(G_x, "G_x")
    if a.shape != b.shape or a.device != b.device:
        raise _lib.FacError(f"reconstruction_loss: shapes / devices differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    e = _engine(a.device)
    out = torch.empty(14, device=a.device)
    B, T = a.shape
    _lib.check(e.handle, e.L.fac_reconstruction_loss(e.handle, _ptr(a), _ptr(b), B, T, _ptr(out), _ptr(out[1:]), _stream(a.device)),
               "fac_reconstruction_loss")
    return (out[0], out[1:]) if return_terms else out[0]
