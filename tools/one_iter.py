"""One warm fwd+bwd at the given size (for ncu -k captures):  python tools/one_iter.py [size] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
import torch, dsr_b200
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode='bilinear').cuda()
z = (torch.rand(1, 32, size, size) * 0.1).cuda()
g = torch.randn(1, 3, size, size).cuda() * 1e-6
for _ in range(iters):
    net.zero_grad(); out = net(z); out.backward(g)
torch.cuda.synchronize()
print('ok')
