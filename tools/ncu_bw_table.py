"""Per-launch table of one iteration from an ncu CSV with several metrics per launch
(gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum, ...):
    python tools/ncu_bw_table.py launches.csv
One line per launch of ONE whole iteration (pack_weights_kernel starts one): time, DRAM read / write MB, GB/s."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row['ID'])
        d = launches.setdefault(i, {'name': re.sub(r'^void |dsr::|\(anonymous namespace\)::|<unnamed>::', '',
                                                   re.sub(r'\(.*', '', row['Kernel Name'])),
                                    'grid': row['Grid Size'], 'block': row['Block Size']})
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        name = row['Metric Name']
        if name == 'gpu__time_duration.sum':
            v = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
        elif unit in ('Kbyte',):
            v *= 1e3
        elif unit in ('Mbyte',):
            v *= 1e6
        elif unit in ('Gbyte',):
            v *= 1e9
        d[name] = v
    rows = list(launches.values())
    starts = [i for i, r in enumerate(rows) if r['name'].startswith('pack_weights')]
    if len(starts) >= 2:
        rows = rows[starts[0]:starts[1]]
    T = sum(r.get('gpu__time_duration.sum', 0) for r in rows)
    print(f'one iteration: {len(rows)} launches, {T:.1f} us serialised')
    print('  #     us   rdMB   wrMB   GB/s  occ%   grid  name')
    for i, r in enumerate(rows):
        t = r.get('gpu__time_duration.sum', 0)
        rd = r.get('dram__bytes_read.sum', 0) / 1e6
        wr = r.get('dram__bytes_write.sum', 0) / 1e6
        occ = r.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0)
        print(f'{i:3d} {t:7.1f} {rd:6.1f} {wr:6.1f} {1e3 * (rd + wr) / max(t, 1e-9):6.0f} {occ:5.1f} '
              f'{r["grid"]:>14s}  {r["name"][:60]}')


if __name__ == '__main__':
    main(sys.argv[1])
