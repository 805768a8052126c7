"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of one
sandwich cycle: per-kernel launches / time / share / DRAM bytes per launch -> markdown + JSON under profiles/, plus the
DRAM-traffic file `bench.py` reads for `roofline.traffic` (stamped with the git revision of the profiled kernels).

    python tools/summarize_launches.py gpurun_out/launches.csv r02 [kineto_kernels.json]
"""
import csv, json, os, re, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r'\(.*$', '', name).replace('void ', '')
    name = name.replace('at::native::', 'at::')
    return name[:70]


def main():
    src, tag = sys.argv[1], sys.argv[2]
    kin = sys.argv[3] if len(sys.argv) > 3 else None
    rows = {}
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r['ID']), {'kernel': short(r['Kernel Name'])})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        m = r['Metric Name']
        if m.startswith('gpu__time_duration'):
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)      # -> us
            d['us'] = v
        else:
            v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
            d['rd' if 'read' in m else 'wr'] = v
    agg = OrderedDict()
    for d in rows.values():
        a = agg.setdefault(d['kernel'], dict(launches=0, us=0.0, rd=0.0, wr=0.0))
        a['launches'] += 1
        a['us'] += d.get('us', 0.0)
        a['rd'] += d.get('rd', 0.0)
        a['wr'] += d.get('wr', 0.0)
    tot = sum(a['us'] for a in agg.values())
    order = sorted(agg.items(), key=lambda kv: -kv[1]['us'])
    try:
        rev = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:   # noqa: BLE001
        rev = 'unknown'
    out = [dict(kernel=k, launches=a['launches'], ms=round(a['us'] / 1e3, 3), share=round(a['us'] / tot, 4),
                dram_read_MB_per_launch=round(a['rd'] / a['launches'] / 1e6, 2),
                dram_write_MB_per_launch=round(a['wr'] / a['launches'] / 1e6, 2)) for k, a in order]
    pdir = os.path.join(ROOT, 'profiles')
    json.dump(dict(git=rev, total_kernel_ms=round(tot / 1e3, 2), launches=len(rows), kernels=out),
              open(os.path.join(pdir, f'{tag}_launch_list_one_cycle.json'), 'w'), indent=1)
    ig = [(k, a) for k, a in agg.items() if 'igemm_kernel' in k]
    n = sum(a['launches'] for _, a in ig)
    if n:
        json.dump(dict(kernel='gs::igemm_kernel (conv fwd + dgrad), all instantiations', git=rev,
                       source=f'profiles/{tag}_launch_list_one_cycle.json: ncu --metrics gpu__time_duration.sum,'
                              'dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one sandwich cycle '
                              '(cold cache: ncu flushes between replays)',
                       launches=n, dram_bytes_per_launch=sum(a['rd'] + a['wr'] for _, a in ig) / n),
                  open(os.path.join(pdir, f'{tag}_igemm_dram_traffic.json'), 'w'), indent=1)
    live = {}
    if kin and os.path.exists(kin):
        for r in json.load(open(kin)):
            live[short(r['kernel'])] = r
    md = [f'# {tag} -- ncu launch list of one steady-state sandwich cycle (kernels of git {rev})', '',
          '    CMD="python bench.py --ncu-cycle --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-infer"   # exited 0 without ncu first',
          '    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\',
          '        --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD', '',
          f'One cycle [MAX, MIN, rand, rand] between cudaProfilerStart/Stop: {len(rows)} kernel launches, {tot / 1e3:.1f} ms of summed '
          'kernel time.  Under ncu every launch is serialised and starts from a flushed cache, so absolute times are upper bounds; '
          'the SHARES are what is compared with the live (CUPTI, no ncu) cycle in the last two columns.', '',
          '| kernel | launches | ms (ncu, serialised) | share | DRAM read MB / launch | DRAM write MB / launch | live ms (kineto) | live share |',
          '|---|---:|---:|---:|---:|---:|---:|---:|']
    ltot = sum(r['ms'] for r in live.values()) or 1.0
    for o in out:
        if o['share'] < 0.001:
            continue
        norm = lambda k: k.replace('true', '1').replace('false', '0').replace(' ', '')
        lv = None
        for lk, r in live.items():
            if norm(lk) == norm(o['kernel']):
                lv = r
        md.append(f"| `{o['kernel']}` | {o['launches']} | {o['ms']:.3f} | {100 * o['share']:.1f} % | {o['dram_read_MB_per_launch']:.2f} | "
                  f"{o['dram_write_MB_per_launch']:.2f} | " + (f"{lv['ms']:.2f} | {100 * lv['ms'] / ltot:.1f} % |" if lv else '| |'))
    open(os.path.join(pdir, f'{tag}_launch_list_summary.md'), 'w').write('\n'.join(md) + '\n')
    print('\n'.join(md[:24]))


if __name__ == '__main__':
    main()
