"""Shadow of /root/reference/tool/utils/__init__.py: `import utils.region_fill as rf`
(tool/video_inpainting.py:28) picks up the GPU region fill below; every other `utils.*` module still
resolves from the reference's tool/utils/ directory further down sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
