"""Same public names as the reference's lib/net/__init__.py:1-4."""
from icon_b200.net import BasePIFuNet, HGPIFuNet
from icon_b200.encoders import NormalNet, VolumeEncoder
