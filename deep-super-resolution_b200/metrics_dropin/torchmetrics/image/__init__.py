"""`from torchmetrics.image import PeakSignalNoiseRatio as PSNR, StructuralSimilarityIndexMeasure as SSIM` (DIP.py:7)."""
from dsr_b200.metrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure  # noqa: F401
