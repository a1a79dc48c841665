"""The flat name set of graphlearn/python/nn/tf/__init__.py (``import graphlearn.python.nn.tf as tfg``), PyTorch underneath."""
from ....models import GAT, GCN, SEAL, EgoGNN, GraphSAGE  # noqa: F401
from ....nn import (BatchGraph, Dataset, DynamicEmbeddingColumn, DynamicSparseEmbeddingColumn, EgoConv, EgoGATConv,  # noqa: F401
                    EgoGINConv, EgoGraph, EgoLayer, EgoRGCNConv, EgoSAGEConv, EmbeddingColumn, FeatureColumn, FeatureGroup,
                    FeatureHandler, FusedEmbeddingColumn, GATConv, GCNConv, HeteroBatchGraph, HeteroConv, LinearLayer,
                    LinkPredictor, Module, NumericColumn, SAGEConv, SparseEmbeddingColumn, SubConv, SubGraphInducer,
                    SubGraphProcessor, SyncBarrierHook, TemporalGraph, TimeEncoder, compute_norm, conf,
                    unsorted_segment_softmax)
from ....nn.loss import (sigmoid_cross_entropy_loss, triplet_margin_loss, triplet_softplus_loss,  # noqa: F401
                         unsupervised_softmax_cross_entropy_loss)
