"""Model zoo: every model family of the reference's examples, in PyTorch."""
from .bipartite_sage import EgoBipartiteSAGE  # noqa: F401
from .ego_gnn import EgoGNN, make_ego_gnn  # noqa: F401
from .graphsage import EgoGraphSAGE  # noqa: F401
from .node2vec import Node2Vec, gen_pair  # noqa: F401
from .sparse_gnn import GAT, GCN, SEAL, GraphSAGE, SparseGNN, drnl_node_labeling  # noqa: F401
from .tgn import TGN, TGNMemory, TemporalAttention, TemporalBatch, TemporalBatchLoader  # noqa: F401
from .ultra_gcn import UltraGCN  # noqa: F401
