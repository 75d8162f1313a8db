"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: whole iterations are cut out of the
captured window (pack_weights_kernel starts one), then time per kernel name, share, and the in-order list of one
iteration (`--order`)."""
import collections
import csv
import re
import sys


def main(path, order, nsteps=0):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void |dsr::|\(anonymous namespace\)::|<unnamed>::', '', name)
        rows.append((name, v, row['Grid Size']))
    starts = [i for i, r in enumerate(rows) if r[0].startswith('pack_weights')]
    if nsteps:                      # --steps N: the capture holds exactly N whole steps (tools/gant_step.py)
        steps, rows_used = nsteps, rows
    elif len(starts) >= 2:
        steps = len(starts) - 1
        rows_used = rows[starts[0]:starts[-1]]
    else:
        steps, rows_used = 1, rows
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for name, v, _ in rows_used:
        tot[name] = tot.get(name, 0) + v
        cnt[name] += 1
    T = sum(tot.values())
    print(f'total {T / steps:.1f} us/step over {len(rows_used)} launches ({steps} whole iterations)')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f'{v / steps:9.1f} us/step {100 * v / T:5.1f}%  n/step={cnt[k] / steps:5.1f}  {k}')
    if order and nsteps:
        print('--- the last step in launch order (us, grid) ---')
        per = len(rows) // nsteps
        for i, (name, v, grid) in enumerate(rows[-per:]):
            print(f'{i:4d} {v:8.1f}  {grid:>14s}  {name}')
    elif order and len(starts) >= 2:
        print('--- one iteration in launch order (us, grid) ---')
        for i, (name, v, grid) in enumerate(rows[starts[0]:starts[1]]):
            print(f'{i:4d} {v:8.1f}  {grid:>14s}  {name}')


if __name__ == '__main__':
    n = int(sys.argv[sys.argv.index('--steps') + 1]) if '--steps' in sys.argv else 0
    main(sys.argv[1], '--order' in sys.argv or True, n)
