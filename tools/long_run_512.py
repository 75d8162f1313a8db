"""BASELINE configs[1]: DIP 4x SR of one synthetic 512x512 image, 3000 fused iterations on one B200.
Prints the loss / PSNR trajectory and the gradient-scale bookkeeping (evidence that the fp16 path is stable)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200')); sys.path.insert(0, ROOT)
import torch, dsr_b200
from oracle import dip_oracle as O
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
lr_img, hr = O.synthetic_pair(0, size)
torch.manual_seed(0)
net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, upsample_mode='bilinear')
cfg = {'learning_rate': 0.01, 'num_iter': iters, 'reg_noise_std': 0.05}
hr_c = hr.unsqueeze(0).cuda()
log = []
def cb(t, out):
    if t % 250 == 0 or t == iters:
        log.append((t, float(10 * torch.log10(1 / ((out - hr_c) ** 2).mean()))))
torch.cuda.synchronize(); t0 = time.time()
out, losses = dsr_b200.dip_sr_fused(net, lr_img, (size, size), 4, cfg, 'cuda:0', seed=1, callback=cb, callback_from=1)
torch.cuda.synchronize(); dt = time.time() - t0
ls = losses.cpu()
print(f'{iters} iterations at {size}^2 in {dt:.2f} s = {iters / dt:.1f} it/s (with a per-iteration host callback)')
print('loss every 250:', [f'{float(ls[i]):.2e}' for i in range(249, iters, 250)])
print('PSNR vs HR (dB):', [(t, round(p, 2)) for t, p in log])
gs = net.debug_tensor('gscale', (size, size)).flatten().cpu()
print(f'gradient scale now {float(gs[0]):.0f}, last amax {float(gs[4]):.1f}, non-finite passes {int(gs[5])}, finite losses {bool(torch.isfinite(ls).all())}')
