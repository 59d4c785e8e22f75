from icon_b200.encoders import Residual3D, VolumeEncoder  # noqa: F401  (reference: lib/net/VE.py:56-183)
