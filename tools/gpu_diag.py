"""Run every GPU parity check group in its own subprocess (a trapped kernel kills the CUDA context of its
process only) and write gpurun_out/diag.json + a readable summary.  Usage: python tools/gpu_diag.py [groups...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')

GROUPS = {
    # name: (env, python expression over `C` (tests/gs_checks) and `gs`)
    # (the *_simt groups route the three conv entry points to the TEST-ONLY CUDA-core twins, tests/libgaiaseg_simt.so)
    'simt_conv': ({}, "sum((C.conv_case_checks(c, gs, 'simt') for c in C.CONV_CASES), [])"),
    'elementwise': ({}, "C.with_simt(gs, lambda: C.image_conv_checks(gs) + C.bn_checks(gs) + C.standalone_bn_checks(gs)) + C.maxpool_checks(gs) + C.loss_checks(gs) + C.argmax_checks(gs)"),
    'model_simt': ({}, "C.with_simt(gs, lambda: C.model_checks(gs))"),
    'tc_1x1': ({}, "sum((C.conv_case_checks(c, gs, 'tc') for c in C.CONV_CASES if c[0].startswith('1x1') and 's2' not in c[0]), [])"),
    'tc_3x3': ({}, "sum((C.conv_case_checks(c, gs, 'tc') for c in C.CONV_CASES if c[0].startswith('3x3') and 's2' not in c[0]), [])"),
    'tc_s2': ({}, "sum((C.conv_case_checks(c, gs, 'tc') for c in C.CONV_CASES if 's2' in c[0]), [])"),
    'tc_epilogue': ({}, "C.conv_epilogue_checks(gs) + C.image_conv_checks(gs)"),
    'tc_bn': ({}, "C.bn_checks(gs)"),
    'stage_tc': ({}, "C.stage_checks(gs)"),
    'psp_ops': ({}, "C.psp_op_checks(gs)"),
    'full_size': ({}, "C.full_size_checks(gs)"),
    'model_tc': ({}, "C.model_checks(gs)"),
    'sgd': ({}, "P.sgd_checks(gs)"),
    'bn_calib': ({}, "P.bn_calibration_checks(gs)"),
    'dropout': ({}, "P.dropout_checks(gs)"),
    'rescale': ({}, "P.rescale_checks(gs)"),
    'loss_full': ({}, "P.loss_full_size_checks(gs)"),
    'big_conv': ({}, "sum((P.big_conv_case_checks(c, gs) for c in P.BIG_CONV_CASES), [])"),
    'deep_stage': ({}, "P.deep_stage_checks(gs)"),
    'config3': ({}, "P.config3_full_size_checks(gs)"),
    'loss': ({}, "C.loss_checks(gs) + C.argmax_checks(gs)"),
    'small_ops': ({}, "C.maxpool_checks(gs) + C.standalone_bn_checks(gs) + P.dropout_checks(gs)[:5]"),
}

CHILD = r'''
import json, sys, os, traceback
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import torch
import gaia_seg_b200 as gs
import gs_checks as C
import gs_checks_path as P
res = []
try:
    res = {expr}
except Exception as e:
    res.append(dict(name={name!r} + '.EXCEPTION', ok=False, err=float('inf'), tol=0, why=f'{{type(e).__name__}}: {{e}}', tb=traceback.format_exc()[-3000:]))
json.dump(res, open({out!r}, 'w'), default=str)
'''


def main():
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(GROUPS)
    report, t_all = {}, time.time()
    for name in names:
        env_extra, expr = GROUPS[name]
        out = os.path.join(OUT, f'diag_{name}.json')
        if os.path.exists(out):
            os.remove(out)
        env = dict(os.environ, **env_extra)
        code = CHILD.format(root=ROOT, expr=expr, name=name, out=out)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=int(os.environ.get('GS_DIAG_TIMEOUT', '420')))
            tail = (r.stdout[-1500:] + r.stderr[-2500:])
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            tail, rc = f'TIMEOUT after 420 s: {str(e.stdout)[-1000:]} {str(e.stderr)[-1000:]}', -9
        res = json.load(open(out)) if os.path.exists(out) else []
        report[name] = dict(rc=rc, seconds=round(time.time() - t0, 1), results=res, tail=tail if (rc != 0 or not res) else '')
        nfail = sum(1 for x in res if not x.get('ok'))
        print(f'== {name}: rc={rc} {len(res)} checks, {nfail} failed, {report[name]["seconds"]} s', flush=True)
        for x in res:
            if not x.get('ok'):
                print('   FAIL', {k: v for k, v in x.items() if k != 'tb'}, flush=True)
                if 'tb' in x:
                    print(x['tb'], flush=True)
        if rc != 0 or not res:
            print(tail, flush=True)
    json.dump(report, open(os.path.join(OUT, 'diag.json'), 'w'), indent=1, default=str)
    print(f'total {time.time() - t_all:.0f} s')


if __name__ == '__main__':
    main()
