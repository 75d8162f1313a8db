"""Pretty-prints a DSR_TIMELINE=2 dump (tools/scales_exp.py with the kstamp build): record i carries the name of host
launch i - 1 (the ring lags by one), so names are shifted here.   python tools/ks_table.py file [file2]"""
import re, sys
def load(path):
    rows = []
    for l in open(path):
        m = re.match(r'\s*(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)/(\d+)\s+(\S+)', l)
        if m: rows.append([int(m[1]), float(m[2]), float(m[3]), float(m[4]), int(m[5]), int(m[6]), m[7]])
    return rows
def short(n):
    n = re.sub(r'^_ZN3dsr\d+', '', n); n = re.sub(r'^_GLOBAL__N__\w+?_cu_\w{8}\d\d', '', n); return n[:40]
files = [load(p) for p in sys.argv[1:]]
a = files[0]
print('rec    lead   total    body   grid/blk' + ('  |  total2' if len(files) > 1 else '') + '  name')
for i in range(1, len(a)):
    r = a[i]
    extra = ''
    if len(files) > 1 and i < len(files[1]): extra = f'  | {files[1][i][2]:7.2f}'
    print(f'{i:3d} {r[1]:7.2f} {r[2]:7.2f} {r[3]:7.2f} {r[4]:6d}/{r[5]:<3d}{extra}  {short(a[i-1][6])}')
print('sum of totals', round(sum(r[2] for r in a[1:]), 1), [round(sum(r[2] for r in f[1:]), 1) for f in files[1:]])
