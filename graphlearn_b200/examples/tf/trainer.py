"""graphlearn/examples/tf/trainer.py: ``from graphlearn.examples.tf.trainer import LocalTrainer`` (PyTorch trainers with the same
method names: engine/trainers.py)."""
from ...engine.trainers import DistTrainer, LocalTrainer  # noqa: F401
