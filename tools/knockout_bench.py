"""Critical-path sensitivity of the training step: run bench.py with some C-ABI entry points KNOCKED OUT (their launches are
skipped, so results are garbage -- timing only) and see how much the step shortens.  What a component costs on the
critical path is what its knock-out saves, which is NOT its summed kernel time when streams overlap.

    GS_KNOCKOUT=gs_conv2d_wgrad python tools/knockout_bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200  # noqa: E402,F401
from gaia_seg_b200 import _lib  # noqa: E402

skip = {s for s in os.environ.get('GS_KNOCKOUT', '').split(',') if s}
orig = _lib.call


def call(name, *args):
    if name in skip:
        return None
    return orig(name, *args)


for mod in list(sys.modules.values()):
    if getattr(mod, '__name__', '').startswith('gaia_seg_b200') and getattr(mod, 'call', None) is orig:
        mod.call = call
sys.argv = [os.path.join(ROOT, 'bench.py')] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, 'bench.py'), run_name='__main__')
