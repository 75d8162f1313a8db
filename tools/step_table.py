"""Warm, in-order CUDA-event timing of every launch of one DIP iteration (profile mode 2), grouped by call site
and by pixel-level size.   python tools/step_table.py [size]"""
import collections, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
import torch, dsr_b200
from dsr_b200._lib import lib, check
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode='bilinear').cuda()
z = (torch.rand(1, 32, size, size) * 0.1).cuda()
g = torch.randn(1, 3, size, size).cuda() * 1e-6
for _ in range(3):
    net.zero_grad(); out = net(z); out.backward(g)
plan = net._plans[(size, size)]
check(lib.dsr_plan_set_profile(plan.handle, 2))
N = 5
for _ in range(N):
    net.zero_grad(); out = net(z); out.backward(g)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 20)
n = lib.dsr_plan_profile_dump(plan.handle, buf, len(buf))
rows = [l.split('\t') for l in buf.value.decode().strip().splitlines()]
tot = collections.OrderedDict(); cnt = collections.Counter(); small = 0.0
for t, name in rows:
    t = float(t) / N
    tot[name] = tot.get(name, 0.0) + t; cnt[name] += 1
    if float(t) * N < 8.0: small += t
T = sum(tot.values())
print(f'{size}^2: {T:.0f} us per fwd+bwd over {len(rows) // N} launches; launches under 8 us sum to {small:.0f} us')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'{v:9.1f} us {100 * v / T:5.1f}%  n={cnt[k] // N:3d}  {k}')
if os.environ.get('STEP_TABLE_ORDERED', '1') != '0':
    per = len(rows) // N
    print('--- launches in order (mean us over %d iterations) ---' % N)
    for i in range(per):
        t = sum(float(rows[k * per + i][0]) for k in range(N)) / N
        print(f'{i:4d} {t:8.1f}  {rows[i][1][:110]}')
