"""GAN training losses (reference: util/losses.py).  Same names and argument meaning.

multiscale_feat_loss runs on the tdvc reduction kernels (one accumulation buffer for all 30 feature maps).
multiscale_spec_loss and contrastive_loss are SURVEY.md 8(f) "next" rows: restated here on torch ops
(cuFFT / matmul / gather), not yet on hand-written kernels."""
import functools
import math

import torch
import torch.nn.functional as F

from tdvc import ops


def multiscale_feat_loss(feat_sig_list, feat_ref_list, norm_p=1):
    """sum over branches and maps of mean |f_sig - f_ref.detach()|  (reference util/losses.py:55-68)."""
    if norm_p != 1:
        # the reference's norm_p == 2 branch calls a torch function that does not exist (losses.py:65)
        raise AttributeError("module 'torch.nn.functional' has no attribute 'rms_loss'")
    sig, ref = [], []
    for feat_sig, feat_ref in zip(feat_sig_list, feat_ref_list):
        for map_sig, map_ref in zip(feat_sig, feat_ref):
            sig.append(map_sig)
            ref.append(map_ref)
    if not sig:
        return 0
    return ops.l1_mean_sum(sig, ref)


def multiscale_feat_loss_rows(feat_sig_list, row0, nrows, feat_ref_list):
    """multiscale_feat_loss(f_sig, f_ref) where every map of `feat_sig_list` holds several signals stacked along the
    batch and f_sig is its rows [row0, row0 + nrows)  (not in the reference: tdvc.train_step runs the discriminator once
    on the stacked generator outputs)."""
    sig, ref = [], []
    for feat_sig, feat_ref in zip(feat_sig_list, feat_ref_list):
        for map_sig, map_ref in zip(feat_sig, feat_ref):
            sig.append(map_sig)
            ref.append(map_ref)
    if not sig:
        return 0
    return ops.l1_mean_sum_rows(sig, row0, nrows, ref)


@functools.lru_cache(maxsize=None)
def _mel_operands(sr, n_fft, n_mels, device):
    """Hann window and slaney-normalised HTK mel filterbank, as torchaudio.transforms.MelSpectrogram(sr, n_fft,
    hop_length=n_fft//4, n_mels=n_mels, norm='slaney') builds them (reference util/losses.py:28-31)."""
    window = torch.hann_window(n_fft, periodic=True)
    n_freqs = n_fft // 2 + 1
    freqs = torch.linspace(0, sr // 2, n_freqs)
    mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
    pts = torch.linspace(mel(0.0), mel(sr / 2), n_mels + 2)
    hz = 700.0 * (10.0 ** (pts / 2595.0) - 1.0)
    width = hz[1:] - hz[:-1]
    slopes = hz.unsqueeze(0) - freqs.unsqueeze(1)
    fb = torch.clamp(torch.min(-slopes[:, :-2] / width[:-1], slopes[:, 2:] / width[1:]), min=0)
    fb = fb * (2.0 / (hz[2:n_mels + 2] - hz[:n_mels])).unsqueeze(0)
    return window.to(device), fb.to(device)


def _log_mel(signal, n_fft):
    window, fb = _mel_operands(16000, n_fft, 80, signal.device)
    window, fb = window.to(signal.dtype), fb.to(signal.dtype)
    lead = signal.shape[:-1]
    spec = torch.stft(signal.reshape(-1, signal.shape[-1]), n_fft, hop_length=n_fft // 4, win_length=n_fft,
                      window=window, center=True, pad_mode='reflect', normalized=False, onesided=True,
                      return_complex=True)
    power = spec.real.square() + spec.imag.square()
    melspec = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)
    melspec = melspec.reshape(lead + melspec.shape[-2:])
    return torch.log(torch.clamp(melspec, min=1e-5))


def multiscale_spec_loss(signal, ref, fft_sizes, spectype='both', return_separated=False, norm_p=1):
    """L1 between log-mel spectrograms.  As in the reference (util/losses.py:33-53) the function returns
    from inside the loop, so only fft_sizes[0] contributes."""
    losses = []
    for fft_size in fft_sizes:
        if norm_p != 1:
            raise AttributeError("module 'torch.nn.functional' has no attribute 'rms_loss'")
        loss = F.l1_loss(_log_mel(signal, fft_size), _log_mel(ref, fft_size).detach())
        losses.append(loss)
        if return_separated:
            return sum(losses), losses
        return sum(losses)


def contrastive_loss(sig_X, sig_Y, num_negatives=100, temp=1, _raw_draws=None):
    """InfoNCE over time frames with in-utterance negatives (reference util/losses.py:70-116).  `temp` is
    accepted and, as in the reference, not applied (its inner call drops it).  `_raw_draws` (not in the
    reference's signature) lets a test supply the two torch.randint draws so both sides share them.
    The gather + cosine + cross-entropy run as one tdvc kernel per direction; the random draws stay torch."""
    B, Cc, T = sig_X.shape
    if _raw_draws is not None:
        raw_X, raw_Y = (d.to(sig_X.device) for d in _raw_draws)
    else:
        raw_X = torch.randint(low=0, high=T - 1, size=(B, T, num_negatives), device=sig_X.device)
        raw_Y = torch.randint(low=0, high=T - 1, size=(B, T, num_negatives), device=sig_X.device)
    return ops.contrastive_loss(sig_X, sig_Y, raw_X, raw_Y)
