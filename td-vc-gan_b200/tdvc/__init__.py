"""tdvc -- B200-native kernels and host glue for td-vc-gan's Generator / CIN / Discriminator / loss hot path."""
from . import ops  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401
