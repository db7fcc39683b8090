from oracle.third_party_ops import ball_query  # noqa: F401
