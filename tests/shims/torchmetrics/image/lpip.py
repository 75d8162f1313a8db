import torch


class LearnedPerceptualImagePatchSimilarity:
    """LPIPS needs pretrained AlexNet weights (not available offline): constant stand-in, value never compared."""

    def __init__(self, *a, **kw):
        pass

    def to(self, device):
        return self

    def __call__(self, a, b):
        return torch.zeros(())
