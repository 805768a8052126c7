"""Summarise `ncu --set full` reports (.ncu-rep) into one markdown table under profiles/: per captured launch the duration,
grid, registers, tensor-pipe utilisation, DRAM bytes and throughput, L2 hit rate, issue-slot utilisation.

    python tools/summarize_ncu_full.py profiles/r02_ncu_full_top_kernels.md gpurun_out/prof_conv_fwd.ncu-rep [more.ncu-rep ...]

Reads the reports with `ncu -i <rep> --page raw --csv` (no GPU needed)."""
import csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COLS = [  # (header, metric name, format)
    ('µs', 'gpu__time_duration.sum', '{:.1f}'),
    ('grid', 'launch__grid_size', '{:.0f}'),
    ('block', 'launch__block_size', '{:.0f}'),
    ('regs', 'launch__registers_per_thread', '{:.0f}'),
    ('cluster', 'launch__cluster_size', '{:.0f}'),
    ('tensor pipe % (active)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', '{:.1f}'),
    ('SM throughput %', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', '{:.1f}'),
    ('IPC (active)', 'sm__inst_executed.avg.per_cycle_active', '{:.2f}'),
    ('DRAM rd MB', 'dram__bytes_read.sum', '{:.2f}'),
    ('DRAM wr MB', 'dram__bytes_write.sum', '{:.2f}'),
    ('DRAM % peak', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '{:.1f}'),
    ('L2 hit %', 'lts__t_sector_hit_rate.pct', '{:.1f}'),
    ('L2 throughput %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', '{:.1f}'),
    ('achieved occupancy %', 'sm__warps_active.avg.pct_of_peak_sustained_active', '{:.1f}'),
]
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def short(name):
    name = re.sub(r'\(.*$', '', name).replace('void ', '')
    return name[:60]


def load(rep):
    r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units = rows[0], rows[1]

    def find(metric):
        for i, h in enumerate(hdr):
            if h == metric:
                return i
        return None
    idx = {c[0]: find(c[1]) for c in COLS}
    ik = hdr.index('Kernel Name')
    out = []
    for row in rows[2:]:
        d = {'kernel': short(row[ik])}
        for name, _, fmt in COLS:
            i = idx[name]
            if i is None or row[i] in ('', 'n/a'):
                d[name] = ''
                continue
            try:
                v = float(row[i].replace(',', ''))
            except ValueError:
                d[name] = ''
                continue
            u = units[i]
            if name in ('µs', 'DRAM rd MB', 'DRAM wr MB'):
                v *= SCALE.get(u, 1.0)
            d[name] = fmt.format(v)
        out.append(d)
    return out


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    try:
        rev = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:   # noqa: BLE001
        rev = 'unknown'
    lines = [f'# `ncu --set full --clock-control none --import-source on` captures mid-cycle (kernels of git {rev})', '',
             'Command profiled: `python bench.py --ncu-cycle --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-infer`'
             ' (exited 0 without ncu first; `tools/gpu_profile.sh`).  Under ncu every launch is serialised and replayed'
             ' from a flushed cache: absolute times are upper bounds of the live ones.', '']
    for rep in reps:
        rows = load(rep)
        lines += [f'## {os.path.basename(rep)}', '', '| # | kernel | ' + ' | '.join(c[0] for c in COLS) + ' |',
                  '|---|---|' + '---:|' * len(COLS)]
        for i, d in enumerate(rows):
            lines.append(f'| {i} | `{d["kernel"]}` | ' + ' | '.join(d[c[0]] for c in COLS) + ' |')
        lines.append('')
    open(dst, 'w').write('\n'.join(lines))
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
