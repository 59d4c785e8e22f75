from icon_b200.encoders import NormalNet  # noqa: F401  (reference: lib/net/NormalNet.py:25)
