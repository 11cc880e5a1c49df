"""Drop-in for tool/utils/Poisson_blend_img.py: same `Poisson_blend_img(imgTrg, imgSrc_gx, imgSrc_gy, holeMask,
gradientMask=None, edge=None)` -> (imgBlend, UnfilledMask) and `getUnfilledMask(holeMask, gradientMask)`;
`poisson_blend_clip` / `poisson_blend_batch` are the batched forms (all frames of the driver's loop,
tool/video_inpainting.py:643-656, in one call)."""
from fgt_b200.poisson import Poisson_blend_img, getUnfilledMask, poisson_blend_batch, poisson_blend_clip  # noqa: F401
