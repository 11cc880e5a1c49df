"""Shadow of /root/reference/LAFC/models/lafc.py (driver: tool/video_inpainting.py:200-214)."""
from fgt_b200.lafc_model import Model, P3DNet  # noqa: F401
