"""Two fused SRGAN training steps on a small batch (2 patches of 8x8 -> 64x64): target for compute-sanitizer."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-super-resolution_b200'))
sys.path.insert(0, ROOT)
import dsr_b200                                   # noqa: E402
from dsr_b200 import gan_train as GT              # noqa: E402
from oracle import gan_train_oracle as O          # noqa: E402  (harness: the synthetic batch)

dev = torch.device('cuda:0')
torch.manual_seed(0)
G, D = dsr_b200.Generator(8).train(), GT.Discriminator((64, 64)).train()
V = GT.Vgg19Loss(pretrained=False).to(dev)
step = GT.GanTrainStep(G, D, V, 1e-4, 2, (8, 8), dev)
LR, HR = O.synthetic_batch(100, 2, (8, 8), 8)
for _ in range(2):
    lD, lG = step.do_epoch(LR.to(dev), HR.to(dev))
torch.cuda.synchronize()
print('losses', float(lD), float(lG), 'device error', step.tr.device_error())
