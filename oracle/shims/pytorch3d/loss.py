def chamfer_distance(*a, **k):  # only referenced by an unused trainer helper
    raise NotImplementedError
