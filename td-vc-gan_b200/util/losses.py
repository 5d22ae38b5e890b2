"""GAN training losses (reference: util/losses.py).  Same names and argument meaning.

multiscale_feat_loss runs on the tdvc reduction kernels (one launch for all 30 feature maps).  multiscale_spec_loss is
the library's own as well: framing + window, the 2048-point real DFT of the 18 frames as a GEMM against a (cos | -sin)
basis on the convolution kernels (tcgen05 in bf16 mode, with the operands split into bf16 high and low parts so that the
spectrum keeps fp32-class accuracy), |X|^2, the mel projection as a second GEMM, log(clamp) and the L1 reduction -- no
cuFFT / cuBLAS.  contrastive_loss is one fused gather + cosine + cross-entropy kernel per direction."""
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


@functools.lru_cache(maxsize=None)
def _dft_basis(n_fft, device, split):
    """[rows, n_fft, 1] conv weight of the real DFT: rows [0, nfreq) = cos(2 pi f k / N), rows [nfreq, 2 nfreq) = -sin(...),
    zero rows up to a multiple of 128 (the tensor-core kernels' wide-N tiling).  split: three column blocks
    [B_hi | B_hi | B_lo] (bf16 high / low parts, stored as fp32) matching ops.stft_frames(split=True)."""
    nfreq = n_fft // 2 + 1
    rows = (2 * nfreq + 127) // 128 * 128
    k = torch.arange(n_fft, dtype=torch.float64)
    f = torch.arange(nfreq, dtype=torch.float64).unsqueeze(1)
    ang = 2.0 * math.pi * f * k / n_fft
    B = torch.zeros(rows, n_fft, dtype=torch.float64)
    B[:nfreq] = torch.cos(ang)
    B[nfreq:2 * nfreq] = -torch.sin(ang)
    B = B.float()
    if split:
        hi = B.to(torch.bfloat16).float()
        lo = (B - hi).to(torch.bfloat16).float()
        B = torch.cat([hi, hi, lo], dim=1)
    return ops.mark_constant(B.unsqueeze(2).contiguous().to(device)), nfreq, nfreq


@functools.lru_cache(maxsize=None)
def _mel_weight(n_fft, device, split):
    """the mel filterbank as a [n_mels, nfreq, 1] conv weight; split: column blocks [fb_hi | fb_hi | fb_lo] (power is
    handed over as [hi | lo | hi]) so that the bf16 tensor-core GEMM keeps fp32-class accuracy over |X|^2's 8 decades"""
    _, fb = _mel_operands(16000, n_fft, 80, device)
    W = fb.t().contiguous().float()
    if split:
        hi = W.to(torch.bfloat16).float()
        lo = (W - hi).to(torch.bfloat16).float()
        W = torch.cat([hi, hi, lo], dim=1)
    return ops.mark_constant(W.unsqueeze(2).contiguous())


def _log_mel(signal, n_fft):
    """log(clamp(MelSpectrogram(signal), 1e-5)) -> [..., n_mels, frames] on the tdvc kernels (CUDA fp32 tensors)."""
    window, fb = _mel_operands(16000, n_fft, 80, signal.device)
    lead = signal.shape[:-1]
    x = signal.reshape(-1, signal.shape[-1])
    Bn = x.shape[0]
    split = ops.get_precision() == "bf16"          # bf16 tensor-core GEMM: hi / lo operand split keeps fp32-class accuracy
    basis, nfreq, im_off = _dft_basis(n_fft, signal.device, split)
    frames = ops.stft_frames(x, window, n_fft, n_fft // 4, split=split)              # [(3*)n_fft, B*NF]
    spec = ops.conv1d(frames.unsqueeze(0), basis)                                    # [1, rows, B*NF]: re | im
    power = ops.power_spectrum(spec.squeeze(0), nfreq, im_off, split=split)          # [(3*)nfreq, B*NF]
    mel = ops.conv1d(power.unsqueeze(0), _mel_weight(n_fft, signal.device, split))   # [1, n_mels, B*NF]
    logmel = ops.log_clamp(mel.squeeze(0), 1e-5)                                     # [n_mels, B*NF]
    nf = logmel.shape[1] // Bn
    return logmel.view(logmel.shape[0], Bn, nf).permute(1, 0, 2).reshape(lead + (logmel.shape[0], nf))


def _log_mel_torch(signal, n_fft):
    """The same on torch ops (torch.stft + matmul): what a CPU tensor gets -- tests of the host logic only."""
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
        if signal.is_cuda and signal.dtype == torch.float32:
            with torch.no_grad():
                lm_ref = _log_mel(ref, fft_size)
            loss = ops.l1_mean_sum([_log_mel(signal, fft_size)], [lm_ref])
        else:
            loss = F.l1_loss(_log_mel_torch(signal, fft_size), _log_mel_torch(ref, fft_size).detach())
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
