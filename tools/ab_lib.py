"""Run a script of this repo against an A/B build of the library: python tools/ab_lib.py tools/ab/libe2b_<name>.so <script.py> [args...]
(same box, same call: the only way to compare two kernels under this pool's box-to-box clock spread of +-3 %)."""
import os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'video-to-audio-and-piano-rp_b200'), os.path.join(ROOT, 'tests')]
from e2_tts_pytorch import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
script = sys.argv[2]
sys.argv = sys.argv[2:]
runpy.run_path(script, run_name='__main__')
