"""TEST-ONLY stand-in for the third-party `torchmetrics` package (not installed in this image), so that the UNMODIFIED
reference DIP.py (baseline/_ref) imports on the CPU arm.  Three classes, plain torch, via oracle/metrics_oracle.py."""
