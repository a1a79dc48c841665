"""graphlearn/examples/tf/link_trainer.py"""
from ...engine.trainers import LinkDistTrainer  # noqa: F401
