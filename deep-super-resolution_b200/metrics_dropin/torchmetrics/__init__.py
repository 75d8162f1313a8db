"""Optional drop-in for the `torchmetrics` names DIP.py / train_GAN.py import, backed by the on-device kernels of
dsr_b200.metrics.  Put deep-super-resolution_b200/metrics_dropin on sys.path ONLY when the on-device metrics are
wanted (or torchmetrics is not installed): it shadows the real package."""
