"""N fused SRGAN training steps and nothing else (profiling target for ncu):  python tools/gant_step.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
sys.path.insert(0, ROOT)
import dsr_b200                                   # noqa: E402
from dsr_b200 import gan_train as GT              # noqa: E402
from oracle import gan_train_oracle as O          # noqa: E402  (harness: the synthetic batch)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0')
torch.manual_seed(0)
G, D = dsr_b200.Generator(8).train(), GT.Discriminator((192, 192)).train()
V = GT.Vgg19Loss(pretrained=False).to(dev)
step = GT.GanTrainStep(G, D, V, 1e-4, 8, (24, 24), dev)
LR, HR = O.synthetic_batch(100, 8, (24, 24), 8)
LR, HR = LR.to(dev), HR.to(dev)
torch.cuda.synchronize()
print('GANT_STEPS_BEGIN', flush=True)
for _ in range(steps):
    step.do_epoch(LR, HR)
torch.cuda.synchronize()
print('launches per step', step.launches_per_step)
