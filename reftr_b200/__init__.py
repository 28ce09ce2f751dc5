"""reftr_b200: B200-native (sm_100a) implementation of the RefTR forward/backward hot path behind the reference's
``build_reftr(args)`` / nn.Module surface.  See DESIGN.md."""
from .api import build_reftr  # noqa: F401
