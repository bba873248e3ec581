"""Digest of an .ncu-rep: per-launch headline metrics and the SASS regions by executed instructions (a reading aid)."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
  return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
  rep = sys.argv[1]
  rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
  hdr, units = rows[0], rows[1]
  for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')])
    for k in KEYS:
      if k in hdr:
        print(f'  {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}')
    for i, k in enumerate(hdr):
      if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(r[i] or 0) > 0.15:
        print(f'  stall {k.split("issue_stalled_")[1].split("_per_")[0]}: {float(r[i]):.2f}')
  src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
  secs, cur = [], None
  for r in src:
    if r and r[0] == 'Kernel Name':
      cur = {'name': r[1], 'rows': []}
      secs.append(cur)
    elif r and r[0] == 'Address':
      cur['hdr'] = r
    elif cur is not None and r:
      cur['rows'].append(r)
  seen = set()
  for s in secs:
    h = s['hdr']
    ia, isrc, ismp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    tot = sum(int(r[ia]) for r in s['rows'])
    if (s['name'], tot) in seen:
      continue
    seen.add((s['name'], tot))
    print('==== regions of', s['name'][:60], 'total inst', tot)
    byop = collections.Counter()
    seg = []
    for i, r in enumerate(s['rows']):
      n, sm = int(r[ia]), int(r[ismp] or 0)
      m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
      byop[m.group(2).split('.')[0] if m else '?'] += n
      if seg and abs(seg[-1]['n'] - n) <= 0.02 * max(n, 1):
        seg[-1]['cnt'] += 1; seg[-1]['sum'] += n; seg[-1]['smp'] += sm; seg[-1]['end'] = i
      else:
        seg.append({'n': n, 'cnt': 1, 'sum': n, 'smp': sm, 'start': i, 'end': i})
    tsm = max(1, sum(x['smp'] for x in seg))
    for x in seg:
      if x['sum'] > 0.005 * tot or x['smp'] > 0.02 * tsm:
        print(f"  rows {x['start']:5d}-{x['end']:5d} exec/row {x['n']:10d} rows {x['cnt']:4d} inst {100 * x['sum'] / tot:5.1f}%"
              f" samples {100 * x['smp'] / tsm:5.1f}%  {s['rows'][x['start']][isrc].strip()[:50]}")
    print('  ops:', ', '.join(f'{op} {100 * n / tot:.1f}%' for op, n in byop.most_common(12)))


if __name__ == '__main__':
  main()
