"""Stub: the reference's tests call matplotlib.use(); nothing on the BA path plots."""


def use(*_a, **_k):
    return None
