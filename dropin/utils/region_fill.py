"""Drop-in for tool/utils/region_fill.py: same `regionfill(I, mask, factor=1.0)`; `diffusion` is the batched
form of the driver's helper (tool/video_inpainting.py:44-52)."""
from fgt_b200.regionfill import diffusion, regionfill, regionfill_batch  # noqa: F401
