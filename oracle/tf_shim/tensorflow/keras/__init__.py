"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
from . import backend, layers  # noqa: F401


class Model:  # only imported (never instantiated) by bts_decoder.py
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs, self.name = inputs, outputs, name


# bts.py imports these names at module level; the loss (bts.py:27-41) never touches them
def Input(*a, **k):  # noqa: N802
    raise NotImplementedError("tf_shim: keras.Input is only a name for `import bts` to succeed")


class applications:  # noqa: N801
    pass


class regularizers:  # noqa: N801
    pass
