"""Tables that rosnet.extra.autoray writes into (stand-in, generator scaffolding only)."""
_FUNCS = {}
_CUSTOM_WRAPPERS = {}


def svd_not_full_matrices_wrapper(fn):
    return fn
