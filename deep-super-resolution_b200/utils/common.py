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


def _ensure_dir(path):
    os.makedirs(path, exist_ok=True)
    return path


def save_model(model, name, out_dir):                       # utils/common.py:11-18
    target = os.path.join(_ensure_dir(out_dir), name + '.pth')
    torch.save(model.state_dict(), target)
    print(f'Model saved to {target}')


def save_image(image, image_name, out_dir):                 # utils/common.py:20-33; image: uint8 H x W x C (numpy or tensor)
    pixels = image.detach().cpu().numpy() if isinstance(image, torch.Tensor) else image
    target = os.path.join(_ensure_dir(os.path.join(out_dir, 'images/')), image_name + '.png')
    Image.fromarray(pixels).save(target)
    print(f'Saved to {target}')


def save_log(out_dir, **kwargs):                            # utils/common.py:35-43: one "key: value" line per entry
    stamp = datetime.now().strftime('%Y_%m_%d_%p%I_%M')
    target = os.path.join(out_dir, stamp + '_log.txt')
    with open(target, 'w') as fh:
        fh.writelines(f'{k}: {v}\n' for k, v in kwargs.items())
    print(f'Log file saved to {target}')


def load_model(model, model_path):                          # utils/common.py:46-60: DataParallel's 'module.' prefix is dropped
    state = torch.load(model_path, weights_only=True)
    if any('module' in key for key in state):
        state = OrderedDict((key.replace('module.', '') if 'module' in key else key, val) for key, val in state.items())
    model.load_state_dict(state)
    return model


def pil_to_np(img_PIL):                                     # utils/common.py:62-74: W x H x C [0..255] -> C x W x H [0..1]
    pixels = np.asarray(img_PIL)
    planes = np.moveaxis(pixels, -1, 0) if pixels.ndim == 3 else pixels[np.newaxis]
    return planes.astype(np.float32) / 255.


def np_to_pil(img_np):                                      # utils/common.py:76-88: C x W x H [0..1] -> W x H x C [0..255]
    bytes_chw = np.clip(img_np * 255, 0, 255).astype(np.uint8)
    single = img_np.shape[0] == 1
    return Image.fromarray(bytes_chw[0] if single else np.moveaxis(bytes_chw, 0, -1))


def np_to_torch(img_np):                                    # utils/common.py:90-95: adds the batch dimension
    return torch.from_numpy(img_np).unsqueeze(0)


def torch_to_np(img_var):                                   # utils/common.py:97-102: drops the batch dimension
    return img_var.detach().cpu().numpy()[0]


def lpips(im0, im1, lpips_model):                           # utils/common.py:105-109
    with torch.no_grad():
        return lpips_model.forward(im0, im1).item()
