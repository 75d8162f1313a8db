"""End-of-round functional evidence on one B200 (beyond bench.py):
 (1) BASELINE configs[1] to completion: 3000 fused DIP iterations at 512^2 (loss / PSNR trajectory, stability);
 (2) BASELINE configs[2] flavour: 8 independent 512^2 images x 300 iterations, 1 vs 2 vs 4 in flight on the GPU;
 (3) generator on DIV2K-like LR sizes (x8: 255x170 -> 2040x1360) against the CPU oracle on a crop, with timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200')); sys.path.insert(0, ROOT)
import torch, dsr_b200
from dsr_b200 import sharder
from oracle import dip_oracle as O
from oracle import gan_oracle as G

def make_net():
    return dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                            upsample_mode='bilinear')

# ---- (1) ----
iters, size = 3000, 512
lr_img, hr = O.synthetic_pair(0, size)
torch.manual_seed(0)
net = make_net()
hr_c = hr.unsqueeze(0).cuda()
log = []
def cb(t, out):
    if t % 500 == 0 or t == iters:
        log.append((t, round(float(10 * torch.log10(1 / ((out - hr_c) ** 2).mean())), 2)))
torch.cuda.synchronize(); t0 = time.time()
out, losses = dsr_b200.dip_sr_fused(net, lr_img, (size, size), 4, {'learning_rate': 0.01, 'num_iter': iters, 'reg_noise_std': 0.05},
                                    'cuda:0', seed=1, callback=cb, callback_from=1)
torch.cuda.synchronize(); dt = time.time() - t0
ls = losses.cpu()
print(f'(1) {iters} iterations at {size}^2 in {dt:.2f} s = {iters / dt:.1f} it/s with a host callback per iteration; '
      f'all losses finite: {bool(torch.isfinite(ls).all())}')
print('    loss every 500:', [f'{float(ls[i]):.2e}' for i in range(499, iters, 500)])
print('    PSNR vs HR (dB):', log)
del net, out, losses

# ---- (2) ----
n_img, it2 = 8, 300
cfg = {'learning_rate': 0.01, 'num_iter': it2, 'reg_noise_std': 0.05}
def prepare():
    jobs = {}
    for i in range(n_img):
        lr_i, hr_i = O.synthetic_pair(i, size)
        torch.manual_seed(i)
        jobs[i] = (make_net(), lr_i, hr_i, dsr_b200.get_noise(32, 'noise', (size, size)))
    return jobs
for k in (1, 2, 4):
    jobs = prepare()
    def run_image(i):
        net_i, lr_i, hr_i, z = jobs[i]
        out_i, _ = dsr_b200.dip_sr_fused(net_i, lr_i, (size, size), 4, cfg, 'cuda:0', seed=10 + i, net_input=z)
        torch.cuda.current_stream().synchronize()
        return {'psnr': float(10 * torch.log10(1 / ((out_i.cpu()[0] - hr_i) ** 2).mean()))}
    torch.cuda.synchronize(); t0 = time.time()
    res = sharder.run_sharded(n_img, run_image, 0, 1, in_flight=k)
    torch.cuda.synchronize(); dt = time.time() - t0
    ps = [res[i]['psnr'] for i in range(n_img)]
    print(f'(2) {n_img} images x {it2} iterations, {k} in flight: {dt:.2f} s = {n_img * it2 / dt:.1f} it/s aggregate '
          f'(includes per-image set-up); mean PSNR {sum(ps) / len(ps):.2f} dB')
    del jobs

# ---- (3) ----
torch.manual_seed(3)
gen = dsr_b200.Generator(8)
sd = gen.state_dict(); G.perturb_trained_state(sd, 2); gen.load_state_dict(sd)
x = torch.rand(1, 3, 170, 255)
gen_c = gen.cuda().eval()
y = gen_c(x.cuda()); torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    y = gen_c(x.cuda())
torch.cuda.synchronize(); dt = (time.time() - t0) / 5
# oracle on a crop whose receptive field stays inside the image: generous margin, compare the crop's centre
cx = x[:, :, 40:120, 60:180]
want = G.generator_forward({k: v.cpu() for k, v in gen_c.state_dict().items()}, cx, 8)
got = y.cpu()[:, :, 40 * 8:120 * 8, 60 * 8:180 * 8]
m = 8 * 36                                   # receptive-field margin: 4 + 33 * 1 LR pixels + output convs
a, b = got[..., m:-m, m:-m], want[..., m:-m, m:-m]
rel = float((a - b).norm() / b.norm())
print(f'(3) generator x8 on a 170x255 LR image -> {tuple(y.shape)} in {dt * 1e3:.1f} ms per image; interior of an 80x120 crop vs '
      f'the CPU oracle: rel L2 {rel:.2e}')
