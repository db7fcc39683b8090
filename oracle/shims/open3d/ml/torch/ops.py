from oracle.third_party_ops import reduce_subarrays_sum  # noqa: F401
