"""Shadow of /root/reference/tool/get_flowNN_gradient.py (driver import: tool/video_inpainting.py:33)."""
from fgt_b200.propagation import get_flowNN_gradient  # noqa: F401
