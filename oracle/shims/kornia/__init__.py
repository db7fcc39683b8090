from oracle.third_party_ops import create_meshgrid  # noqa: F401
