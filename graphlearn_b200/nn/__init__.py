"""PyTorch model layer (``gl.nn``): the math of graphlearn/python/nn/tf/* re-implemented as
torch Modules, with the GraphSAGE layer on the fused sm_100a kernel."""
from . import loss  # noqa: F401
from .conv import (EgoGATConv, EgoGINConv, EgoLayer, EgoRGCNConv, EgoSAGEConv, EgoTGATConv,  # noqa: F401
                   TimeEncoder)
from .norm import compute_norm, compute_saint_norm  # noqa: F401
from .data import BatchGraph, Data, EgoGraph, HeteroBatchGraph, TemporalGraph  # noqa: F401
from .dataset import Batch, Dataset, PyGDataLoader, SubGraphData, TorchDataset  # noqa: F401
from .embedding import ShardedEmbedding  # noqa: F401
from .feature import FeatureEncoder  # noqa: F401
from .hetero import (BipartiteSAGEConv, HeteroConv, HeteroSubGraph, LinkPredictor, SubGraphInducer,  # noqa: F401
                     SubGraphProcessor)
from .sparse_conv import GATConv, GCNConv, SAGEConv, segment_softmax  # noqa: F401
