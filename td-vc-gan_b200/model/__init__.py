"""Drop-in `model` package: same import names as the reference (train.py:22-25,
generate_with_target.py:12).  generator / discriminator / conditional_instance_norm / latent_classifier /
grad_rev / ssl_encoder (its WaveNet stack; WavLM itself is imported from a reference checkout) are re-implemented on the tdvc
CUDA kernels; the sub-module that is out of the hot-path scope (f0_estimator) resolves to the reference's own file when a
reference checkout is available (TDVC_REFERENCE or /root/reference) -- nothing is copied."""
import os as _os

_ref = _os.environ.get("TDVC_REFERENCE", "/root/reference")
_ref_model = _os.path.join(_ref, "model")
if _os.path.isdir(_ref_model) and _ref_model not in __path__:
    __path__.append(_ref_model)
