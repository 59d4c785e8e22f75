from icon_b200.net import HGPIFuNet  # noqa: F401  (reference: lib/net/HGPIFuNet.py:34)
