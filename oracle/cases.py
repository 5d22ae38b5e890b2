"""Shared case table for the golden generator and the parity tests (TEST INFRASTRUCTURE)."""
import zlib

import torch

# model section of config/conv_enc-stage1.yaml:44-72 (identical in every shipped stage YAML)
FULL = dict(ratios=(10, 8, 2, 2), channels=(256, 128, 64, 32, 16), content_dim=128, cond_dim=128, nspk=100,
            num_disc=3, d_layers=4, d_base=16)
TINY = dict(ratios=(4, 4, 2, 2), channels=(32, 16, 16, 8, 8), content_dim=16, cond_dim=16, nspk=6,
            num_disc=3, d_layers=3, d_base=4)

CASES = {
    "g_tiny": dict(TINY, B=2, T=2048, seed=1),
    "d_tiny": dict(TINY, B=2, T=2048, seed=2),
    "g_full": dict(FULL, B=1, T=8960, seed=3),
    "d_full": dict(FULL, B=1, T=8960, seed=4),
    "step_tiny": dict(TINY, B=2, T=2048, seed=5),
    "step_full": dict(FULL, B=2, T=8960, seed=6),
}

# train sections of the shipped YAMLs (config/conv_enc-stage{1,2_1,2_2}.yaml:5-35); lambda_f0 forced to 0
# (torchcrepe unavailable offline)
_COMMON = dict(lambda_feat=2, lambda_spec=5, lambda_wave=0, lambda_latcls=0, lambda_cont_emb=10,
               lambda_corrupted=1, lambda_converted=0, lambda_f0=0, jitter_amp=0, grad_max_norm_D=None, grad_max_norm_G=None)
HP_STAGE1 = dict(_COMMON, no_conv=False, lambda_rec=0, lambda_idt=5)
HP_STAGE2_1 = dict(_COMMON, no_conv=True, lambda_rec=0, lambda_idt=20)
HP_STAGE2_2 = dict(_COMMON, no_conv=False, lambda_rec=10, lambda_idt=1)
# BASELINE.json config 2: conv_enc-stage2_1 with the latent classifier + gradient reversal switched on
HP_LATCLS = dict(HP_STAGE2_1, lambda_latcls=1)
# options no shipped YAML switches on, pinned so that they are not silently ignored: the waveform L1 term of
# train.py:358-361,382-385 and clip_grad_norm_ on D and G (train.py:288-289,488-489); stored gradients are post-clip
HP_WAVE_CLIP = dict(HP_STAGE2_2, lambda_wave=3.0, grad_max_norm_D=0.05, grad_max_norm_G=0.5)


def rand_like(t: torch.Tensor, tag: int, dtype=torch.float64) -> torch.Tensor:
    """Deterministic N(0,1) tensor of t's shape keyed by an integer tag (projection vectors for
    scalar test losses)."""
    g = torch.Generator(); g.manual_seed(1000003 * tag + 17)
    return torch.randn(tuple(t.shape), generator=g, dtype=torch.float64).to(dtype)
