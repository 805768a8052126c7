"""Single-process target for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): ONE small launch set of every
kernel family of libgaiaseg_b200.so through the parity checks of tests/ (so results are also compared with the oracle).

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py [conv] [bn] [loss] [misc] [sgd]

One process = one `import torch` under the sanitizer (tools/gpu_diag.py spawns one process per group, too slow here)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import gaia_seg_b200 as gs  # noqa: E402
import gs_checks as C  # noqa: E402
import gs_checks_path as P  # noqa: E402

want = set(sys.argv[1:]) or {'conv', 'bn', 'loss', 'misc', 'sgd'}
res, t0 = [], time.time()
by_name = {c[0]: c for c in C.CONV_CASES}
if 'conv' in want:
    # single-CTA kernel (short K), strided dgrad (zero insertion), CTA-pair kernel with a phantom tile (long K, odd tile count)
    for name in ('1x1_k320_n80', '3x3_s2_ragged', '3x3_oddtiles'):
        res += C.conv_case_checks(by_name[name], gs, 'tc')
    res += C.conv_epilogue_checks(gs)
if 'bn' in want:
    res += C.bn_checks(gs)
    res += P.bn_bwd_one_pass_checks(gs)
if 'loss' in want:
    res += C.loss_checks(gs, cases=((2, 19, 7, 9, 33, 50), (1, 150, 8, 8, 64, 64)))
    res += C.argmax_checks(gs)
if 'misc' in want:
    res += C.maxpool_checks(gs) + C.image_conv_checks(gs) + P.dropout_checks(gs)[:5]
if 'sgd' in want:
    res += P.sgd_checks(gs, steps=2)
torch.cuda.synchronize()
bad = [r for r in res if not r['ok']]
print(json.dumps(dict(groups=sorted(want), checks=len(res), failed=[r['name'] for r in bad], launches=gs._lib.launch_count(),
                      seconds=round(time.time() - t0, 1))))
sys.exit(1 if bad else 0)
