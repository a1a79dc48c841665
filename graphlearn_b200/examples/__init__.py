"""``graphlearn.examples`` import path of the reference: only the pieces other scripts import from it (the trainers); the runnable
example scripts live in the repository's top-level ``examples/`` directory."""
