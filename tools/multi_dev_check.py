"""One process, two GPUs: the library's per-device function attributes / plans work on every visible device."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200')); sys.path.insert(0, ROOT)
import torch, dsr_b200
from oracle import gan_oracle as G
n = torch.cuda.device_count()
print('devices', n)
torch.manual_seed(0)
gen = dsr_b200.Generator(8)
x = torch.rand(1, 3, 24, 24)
want = G.generator_forward({k: v.clone() for k, v in gen.state_dict().items()}, x, 8)
for d in range(n):
    dev = torch.device('cuda', d)
    with torch.cuda.device(dev):
        got = gen.to(dev).eval()(x.to(dev)).cpu()
        print(f'cuda:{d} generator rel L2 {float((got - want).norm() / want.norm()):.2e}')
        torch.manual_seed(1)
        net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode='bilinear').to(dev)
        ds = dsr_b200.Downsampler(3, 4, 'lanczos2', phase=0.5, preserve_size=True).to(dev)
        z = (torch.rand(1, 32, 64, 64) * 0.1).to(dev)
        out = net(z); loss = ds(out).pow(2).mean(); loss.backward()
        torch.cuda.synchronize(dev)
        print(f'cuda:{d} DIP step loss {float(loss):.5f} grad finite {bool(torch.isfinite(net._gflat).all())}')
