"""A/B of the upsample+concat+BN(132) paths on identical inputs: low-resolution-domain kernels (product) vs the
per-pixel high-resolution kernels (DSR_NO_LOWRES_UPCAT=1).  Prints per-tensor and whole-gradient agreement.
python tools/ab_upcat.py [H] [W]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
import torch, dsr_b200
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else H


def run(env):
    if env:
        os.environ['DSR_NO_LOWRES_UPCAT'] = '1'
    else:
        os.environ.pop('DSR_NO_LOWRES_UPCAT', None)
    torch.manual_seed(0)
    net = dsr_b200.get_net(32, 'skip', 'reflection', skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                           upsample_mode='bilinear').cuda()
    g = torch.Generator().manual_seed(1)
    z = (torch.rand(1, 32, H, W, generator=g) * 0.1).cuda()
    go = (torch.randn(1, 3, H, W, generator=g) * 1e-4).cuda()
    out = net(z)
    out.backward(go)
    torch.cuda.synchronize()
    return out.detach().cpu(), {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}


o1, g1 = run(False)
o2, g2 = run(True)
o3, g3 = run(False)
print('out rel diff lowres vs hires', float((o1 - o2).norm() / o2.norm()), ' lowres vs lowres', float((o1 - o3).norm() / o3.norm()))
def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))
worst = sorted(((cos(g1[k], g2[k]), k) for k in g1 if float(g2[k].norm()) > 0))[:8]
print('worst per-tensor cosines lowres vs hires:', worst)
worst = sorted(((cos(g1[k], g3[k]), k) for k in g1 if float(g3[k].norm()) > 0))[:4]
print('worst per-tensor cosines lowres vs lowres (run-to-run):', worst)
cat = lambda g: torch.cat([v.flatten() for v in g.values()])
print('whole gradient cosine: lowres vs hires', cos(cat(g1), cat(g2)), ' run-to-run', cos(cat(g1), cat(g3)))
