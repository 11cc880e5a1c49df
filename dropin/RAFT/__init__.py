"""Shadow of /root/reference/RAFT/__init__.py: `from RAFT import RAFT` (tool/video_inpainting.py:26)
yields the B200-native model; other submodules (RAFT.utils, ...) still resolve from the reference's
RAFT/ directory further down sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
from fgt_b200.raft_model import RAFT  # noqa: E402,F401
