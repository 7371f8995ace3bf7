"""Minimal Keras object model: build-on-first-call Layer/Model with add_weight."""
import numpy as np

from .. import _t
from . import initializers, layers  # noqa: F401


class Model(layers.Layer):
    pass
