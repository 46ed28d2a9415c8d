"""Test-infrastructure shim: `ipdb` is imported (never used on our path) by the
reference's evaluation_tools; it is not installed in this image."""


def set_trace(*a, **k):
    raise RuntimeError("ipdb shim: set_trace called")
