"""Two generator forwards on a 32-image chunk of 96x96 patches (for ncu -k captures):  python tools/gan_one.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
import torch, dsr_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
g = dsr_b200.Generator(8).cuda().eval()
x = torch.rand(B, 3, 96, 96, device='cuda')
for _ in range(2):
    y = g(x)
torch.cuda.synchronize()
print('ok', float(y.abs().mean()))
