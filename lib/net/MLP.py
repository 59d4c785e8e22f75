from icon_b200.net import MLP  # noqa: F401  (reference: lib/net/MLP.py:8)
