"""mirrorfusion_b200 — B200-native MirrorFusion denoising hot path (BrushNet + SD1.5 UNet + CFG + scheduler)."""
from .config import NetConfig, SD15, TINY, MICRO  # noqa: F401
