from refvsr_b200.srnet import SRNet  # noqa: F401  (replaces models/SRNet.py:11-61)
