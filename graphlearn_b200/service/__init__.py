"""Server mode: clients that hold NO graph drive GSL queries on graph servers (L6 / client<->server decoupling).

Reference: ``g.init(cluster=..., job_name="server" | "client")`` (graphlearn/python/graph.py:452-494): servers load the
graph and execute DAGs continuously; clients ship a DagDef once (``RunDag``) and then pull whole batches
(``GetDagValues``) - N clients : M servers, M <= N, clients are dealt to servers round-robin
(python/client.py:29-87, src/service/dist/round_robin_balancer.cc:83-150).

Here a server is a process that has built its shard(s) on its GPU(s) (one rank of the torchrun world, or a single
process) and runs :class:`GraphServer`; a client is a plain process without CUDA / torch.distributed that connects with
:class:`RemoteGraphClient`.  The wire format is a pickled tuple over ``multiprocessing.connection`` (length-prefixed TCP
with an HMAC handshake): a serialisable DagDef goes out once, numpy batches (ids + the attributes the decoders declare)
come back - the same division of labour as the reference's gRPC ``HandleDag`` / ``GetDagValues``.
"""
from .client import RemoteDataset, RemoteGraphClient  # noqa: F401
from .server import GraphServer  # noqa: F401
