"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name and share."""
import collections
import csv
import re
import sys


def main(path, steps):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void |dsr::|\(anonymous namespace\)::|<unnamed>::', '', name)
        tot[name] = tot.get(name, 0) + v
        cnt[name] += 1
    T = sum(tot.values())
    print(f'total {T / steps:.1f} us/step over {sum(cnt.values())} launches ({steps} steps)')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f'{v / steps:9.1f} us/step {100 * v / T:5.1f}%  n/step={cnt[k] / steps:5.1f}  {k}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
