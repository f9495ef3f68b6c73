"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
from . import backend, layers  # noqa: F401


class Model:  # only imported (never instantiated) by bts_decoder.py
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs, self.name = inputs, outputs, name
