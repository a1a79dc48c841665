"""The flat name set of graphlearn/python/nn/pytorch/__init__.py (``import graphlearn.python.nn.pytorch as thg``)."""
from ....nn import PyGDataLoader, TemporalDataLoader, TemporalDataset  # noqa: F401
from ....nn import TorchDataset as Dataset  # noqa: F401  (the torch IterableDataset over a GSL query)
from ....nn import get_cluster_spec, get_counts, launch_server, set_client_num  # noqa: F401
