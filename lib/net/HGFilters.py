from icon_b200.encoders import ConvBlock, HourGlass, HGFilter  # noqa: F401  (reference: lib/net/HGFilters.py)
