"""Top stalled SASS instructions of an .ncu-rep captured with --import-source on:  python tools/ncu_stalls.py rep [n]"""
import csv, io, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16
det = subprocess.run(['ncu', '-i', rep, '--page', 'details'], capture_output=True, text=True).stdout
for l in det.splitlines():
    if any(k in l for k in ('Duration', 'SM Frequency', 'DRAM Throughput', 'Registers Per', 'Grid Size', 'Executed Ipc Active')):
        print(l.strip())
raw = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
h, u, v = raw[0], raw[1], raw[2]
for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):
    if k in h: print(k, v[h.index(k)], u[h.index(k)])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
si, src = h.index('# Samples'), h.index('Source')
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for r in rows[2:]:
    try: data.append((int(r[si]), r))
    except Exception: pass
tot = sum(d[0] for d in data)
print('samples', tot)
for c, r in sorted(data, key=lambda t: -t[0])[:n]:
    st = sorted([(int(r[i] or 0), h[i]) for i in stall], reverse=True)[:2]
    print(f'{c:6d} {100 * c / tot:5.1f}%  {r[src].strip()[:64]:64s} {st}')
