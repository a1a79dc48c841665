"""graphlearn/python/nn/tf/config.py"""
from ....nn.utils import Config, conf  # noqa: F401
