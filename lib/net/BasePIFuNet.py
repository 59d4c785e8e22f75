from icon_b200.net import BasePIFuNet  # noqa: F401  (reference: lib/net/BasePIFuNet.py:23)
