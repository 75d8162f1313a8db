"""Drop-in for the reference's `utils.common` (utils/common.py:1-108): checkpoint / PNG / log IO and the array
conversions the CLIs star-import (`from utils.common import *`: eval_GAN.py:14, DIP.py, train_GAN.py:15) -- the module
names those scripts pick up through the star import (np, torch, os, Image, datetime, re, OrderedDict) are exported too.
Host-side IO by nature; the one device step of this path, the float -> uint8 HWC conversion of a resolved image
(eval_GAN.py:50-53), is `dsr_b200.evalgan.to_uint8_hwc` (csrc/dsr_metrics.cu) and `save_image` accepts its result."""
import os
import re
from datetime import datetime
from typing import OrderedDict

import numpy as np
import torch
from PIL import Image


def save_model(model, name, out_dir):                       # utils/common.py:11-18
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f'{name}.pth')
    torch.save(model.state_dict(), path)
    print(f'Model saved to {path}')


def save_image(image, image_name, out_dir):                 # utils/common.py:20-33; image: uint8 H x W x C (numpy or tensor)
    if isinstance(image, torch.Tensor):
        image = image.detach().cpu().numpy()
    folder = os.path.join(out_dir, 'images/')
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, f'{image_name}.png')
    Image.fromarray(image).save(path)
    print(f'Saved to {path}')


def save_log(out_dir, **kwargs):                            # utils/common.py:35-43
    path = os.path.join(out_dir, f'{datetime.now().strftime("%Y_%m_%d_%p%I_%M")}_log.txt')
    with open(path, 'w') as f:
        for key, value in kwargs.items():
            f.write(f'{key}: {str(value)}\n')
    print(f'Log file saved to {path}')


def load_model(model, model_path):                          # utils/common.py:46-60: DataParallel's 'module.' prefix is dropped
    state = torch.load(model_path, weights_only=True)
    if any('module' in k for k in state):
        state = OrderedDict((k.replace('module.', '') if 'module' in k else k, v) for k, v in state.items())
    model.load_state_dict(state)
    return model


def pil_to_np(img_PIL):                                     # utils/common.py:62-74: W x H x C [0..255] -> C x W x H [0..1]
    ar = np.array(img_PIL)
    ar = ar.transpose(2, 0, 1) if ar.ndim == 3 else ar[None, ...]
    return ar.astype(np.float32) / 255.


def np_to_pil(img_np):                                      # utils/common.py:76-88: C x W x H [0..1] -> W x H x C [0..255]
    ar = np.clip(img_np * 255, 0, 255).astype(np.uint8)
    ar = ar[0] if img_np.shape[0] == 1 else ar.transpose(1, 2, 0)
    return Image.fromarray(ar)


def np_to_torch(img_np):                                    # utils/common.py:90-95
    return torch.from_numpy(img_np)[None, :]


def torch_to_np(img_var):                                   # utils/common.py:97-102
    return img_var.detach().cpu().numpy()[0]


def lpips(im0, im1, lpips_model):                           # utils/common.py:105-109
    with torch.no_grad():
        return lpips_model.forward(im0, im1).item()
