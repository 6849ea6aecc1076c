"""Turn ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches <launches.csv> <out.md> [title]
  python scripts/summarize_profiles.py full <report.ncu-rep> <out.md> [title]
"""
import collections
import csv
import subprocess
import sys


def launches(path, out, title, last=0, drop_tail=0):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    if drop_tail:                              # launches after the step (bench bookkeeping)
        rows = rows[:-drop_tail]
    if last:                                   # only the last `last` launches (one step of a multi-step capture)
        rows = rows[-last:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r['Kernel Name'].split('(')[0].replace('void ', '')
        agg[n][0] += 1
        agg[n][1] += float(r['Metric Value'])
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nSource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: '
                f'compare SHARES, not absolutes).\n\n{len(rows)} launches, {tot / 1e3:.1f} us total.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n')
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'| `{n[:100]}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |\n')
        for kn in ('k_conv_tc', 'k_conv_chain'):
            seq = [(r['Grid Size'], float(r['Metric Value']) / 1e3) for r in rows if kn in r['Kernel Name']]
            if seq:
                f.write(f'\n`{kn}` launches in order (grid: us):\n\n```\n' + ' '.join(f"{g}:{t:.0f}" for g, t in seq) + '\n```\n')


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum']


def full(path, out, title):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nSource: `{path}` (ncu --set full --clock-control none --import-source on).\n')
        for d in data:
            f.write(f"\n## {d[hdr.index('Kernel Name')][:80]}  grid {d[hdr.index('Grid Size')]}  block {d[hdr.index('Block Size')]}\n\n| metric | value |\n|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f'| {k} | {d[hdr.index(k)]} {units[hdr.index(k)]} |\n')
            stalls = []
            for i, h in enumerate(hdr):
                if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                    try:
                        stalls.append((float(d[i]), h.split('warp_issue_stalled_')[1].replace('_per_warp_active.pct', '')))
                    except ValueError:
                        pass
            f.write('\nTop warp-issue stall reasons (% of active warps): ' + ', '.join(f'{n} {v:.0f}' for v, n in sorted(stalls, reverse=True)[:6]) + '\n')


if __name__ == '__main__':
    mode, src, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    if mode == 'launches':
        launches(src, out, title, int(sys.argv[5]) if len(sys.argv) > 5 else 0, int(sys.argv[6]) if len(sys.argv) > 6 else 0)
    else:
        full(src, out, title)
