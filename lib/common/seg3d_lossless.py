from icon_b200.engine import Seg3dLossless  # noqa: F401  (reference: lib/common/seg3d_lossless.py:36)
