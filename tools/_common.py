"""Shared CLI helpers for the GAIA-seg entry points (tools/train_supernet.py, test_supernet.py, extract_subnet.py,
finetune_supernet.py keep the reference's argument names)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class DictAction(argparse.Action):
    """mmcv.DictAction: KEY=VALUE pairs with int / float / bool / list parsing."""

    @staticmethod
    def _parse(val):
        for cast in (int, float):
            try:
                return cast(val)
            except ValueError:
                pass
        if val.lower() in ('true', 'false'):
            return val.lower() == 'true'
        if val.startswith('[') and val.endswith(']'):
            return [DictAction._parse(v) for v in val[1:-1].split(',') if v]
        return val

    def __call__(self, parser, namespace, values, option_string=None):
        out = {}
        for kv in values:
            k, v = kv.split('=', maxsplit=1)
            out[k] = self._parse(v)
        setattr(namespace, self.dest, out)


def setup_dist(args, cfg):
    import gaia_seg_b200 as gs
    if 'LOCAL_RANK' not in os.environ:
        os.environ['LOCAL_RANK'] = str(getattr(args, 'local_rank', 0))
    if getattr(args, 'launcher', 'none') == 'none':
        return False
    gs.init_dist(args.launcher, **cfg.get('dist_params', dict(backend='nccl')))
    return True
