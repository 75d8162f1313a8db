"""Times individual tensor-core launches at 512^2 with CUDA events (experiment helper; warm L2)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
import torch, dsr_b200
from dsr_b200._lib import lib, check
torch.manual_seed(0)
net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode='bilinear').cuda()
z = (torch.rand(1, 32, 512, 512) * 0.1).cuda()
out = net(z); out.backward(torch.randn_like(out) * 1e-6)
plan = net._plans[(512, 512)]
s = torch.cuda.current_stream().cuda_stream
for layer, what in (('L0.u1', 0), ('L0.d2', 0), ('L1.u1', 0), ('L0.u1', 1), ('L0.d2', 1), ('L0.u1', 2), ('L0.u2', 0), ('L0.d1', 0), ('L1.d1', 1)):
    for _ in range(3):
        check(lib.dsr_plan_debug_replay(plan.handle, layer.encode(), what, 0, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        check(lib.dsr_plan_debug_replay(plan.handle, layer.encode(), what, 0, s))
    e1.record(); torch.cuda.synchronize()
    print(f'dbg={os.environ.get("DSR_HALO_DBG","0")} {layer} what={what}: {e0.elapsed_time(e1) * 100:.1f} us')
