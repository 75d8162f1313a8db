"""`from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity as LPIPS` (DIP.py:8).

LPIPS is a pretrained AlexNet feature distance: third party, needs downloaded weights, out of scope.
The name resolves so that DIP.py imports; constructing it raises unless the real torchmetrics is installed."""


class LearnedPerceptualImagePatchSimilarity:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('LPIPS needs the real torchmetrics package and pretrained AlexNet weights; '
                                  'dsr_b200 provides PSNR and SSIM only')
