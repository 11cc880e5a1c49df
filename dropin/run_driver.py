"""Launcher that makes the reference driver pick up every fgt_b200 shim:

    python dropin/run_driver.py /path/to/FGT/tool/video_inpainting.py [driver arguments ...]

`python tool/video_inpainting.py` puts tool/ at sys.path[0], AHEAD of PYTHONPATH, so the three modules the driver
imports from its own directory (`utils.region_fill`, `utils.Poisson_blend_img`, `get_flowNN_gradient`,
tool/video_inpainting.py:28-33) would resolve to the reference's CPU code no matter what PYTHONPATH says. This
launcher builds the search path the other way round — dropin/, the fgt_b200 repository, then tool/ — and runs the
unchanged driver as __main__; the driver then appends the reference root, FGT/ and LAFC/ itself (:4-6), which is
where `RAFT.utils`, `utils.*` helpers the shims do not replace, and the config files come from.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def setup_path(driver_path):
    """sys.path = [dropin, repo, <driver's directory>, ...previous entries...]; returns the driver's directory."""
    tool_dir = os.path.dirname(os.path.abspath(driver_path))
    for p in (HERE, REPO, tool_dir):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[0:0] = [HERE, REPO, tool_dir]
    return tool_dir


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        sys.exit("usage: python dropin/run_driver.py /path/to/FGT/tool/video_inpainting.py [driver arguments ...]")
    driver = os.path.abspath(argv[0])
    setup_path(driver)
    sys.argv = [driver] + list(argv[1:])
    runpy.run_path(driver, run_name="__main__")


if __name__ == "__main__":
    main()
