"""Stage-by-stage comparison of the CUDA generator with the CPU oracle (debug helper).
    python tools/gan_diag.py [blocks] [B] [h] [w] [factor]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200')); sys.path.insert(0, ROOT)
import torch, dsr_b200
from dsr_b200._lib import lib, check
from oracle import gan_oracle as go
blocks, B, h, w, factor = [int(a) for a in (sys.argv[1:6] + ['1', '2', '20', '24', '8'][len(sys.argv) - 1:])]
torch.manual_seed(1)
g = dsr_b200.Generator(factor, blocks)
sd = g.state_dict(); go.perturb_trained_state(sd, 3); g.load_state_dict(sd)
g = g.cuda().eval()
x = torch.rand(B, 3, h, w, generator=torch.Generator().manual_seed(2))
y = g(x.cuda()); torch.cuda.synchronize()
rec = {}
yo = go.generator_forward({k: v.cpu() for k, v in g.state_dict().items()}, x, factor, blocks, record=rec)
plan = next(iter(g._plans.values()))
err = C.c_int()
check(lib.dsr_gen_device_error(plan.handle, C.byref(err))); print('device error word', err.value)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
def fetch(name):
    ptr, rows, ir, H, W = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib.dsr_gen_debug_tensor(plan.handle, name.encode(), C.byref(ptr), C.byref(rows), C.byref(ir), C.byref(H), C.byref(W)))
    t = torch.empty((rows.value, W.value, 64), dtype=torch.float16, device='cuda')
    check(lib.dsr_debug_copy(t.data_ptr(), ptr, t.numel() * 2, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
    t = t.cpu().float()
    return torch.stack([t[b * ir.value: b * ir.value + H.value] for b in range(B)]).permute(0, 3, 1, 2)
names = [('x0', 'x0'), ('t1', f'block{blocks - 1}_t')]
bufs = ['xa', 'xb']
names.append((bufs[(blocks - 1) % 2], f'block{blocks - 1}'))
names.append((bufs[blocks % 2], 'trunk'))
names += [(f's{i}', f's{i}') for i in range(go.SHUFFLES[factor])]
for dev, ora in names:
    a, b = fetch(dev), rec[ora]
    print(f'{dev:4s} vs {ora:10s} rel L2 {rel(a, b):.3e}  max abs {float((a - b).abs().max()):.3e}  ref rms {float(b.pow(2).mean().sqrt()):.3e}')
print(f'y    rel L2 {rel(y.cpu(), yo):.3e} max abs {float((y.cpu() - yo).abs().max()):.3e}  ref rms {float(yo.pow(2).mean().sqrt()):.3e}')
