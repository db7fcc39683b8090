from oracle.third_party_ops import ContinuousConv  # noqa: F401
