from icon_b200.net import query_func  # noqa: F401  (reference: lib/common/train_util.py:324-348)
