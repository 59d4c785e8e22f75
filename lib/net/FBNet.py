from icon_b200.encoders import GlobalGenerator, ResnetBlock  # noqa: F401  (reference: lib/net/FBNet.py:202-319)
