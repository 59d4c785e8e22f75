from icon_b200.net import init_net  # noqa: F401
from icon_b200.encoders import ConvBlock  # noqa: F401  (reference: lib/net/net_util.py:73-126, 224-280)
