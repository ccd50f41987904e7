"""A/B helper (development): run a script against another build of the library.
    PVR_AB_LIB=pvr_habitat_b200/lib/libpvr_b200_ab.so python tools/run_with_lib.py bench.py --no-extra ..."""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvr_habitat_b200 import _lib  # noqa: E402

if os.environ.get("PVR_AB_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["PVR_AB_LIB"])
sys.argv = sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
