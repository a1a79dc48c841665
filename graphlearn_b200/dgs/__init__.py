"""GPU-native analogue of the Dynamic Graph Service (dynamic_graph_service/).

The reference DGS is a CPU/Kafka/RocksDB online service: graph updates stream in, a pre-installed
query decides which (vertex, sampler-op) states exist, every state keeps a fixed-capacity sample
(top-k by timestamp, replaced through a min-heap), and an inference request ``/infer?qid&vid``
becomes a few KV lookups (docs/en/dgs/intro.md:19-68).  Here the same contract is kept with the
state in HBM tables:

* ``SampleStore``     per (edge type) ``[num_vertices, K]`` neighbour / timestamp / weight tables updated
                      batch-wise with a vectorised "replace the oldest" rule (TopK-by-timestamp sampler,
                      src/core/storage/topk_sampler.cc:23-41) + latest-version vertex features
                      (sampler.cc:21-46)
* ``QueryPlan``       SOURCE -> EDGE_SAMPLER/VERTEX_SAMPLER chain (fbs/plan_node.fbs:2-6)
* ``DynamicGraphService``  install_query / apply_updates / run_query (batched) / checkpoint / restore,
                      plus the adaptive ingest rate limiter (adaptive_rate_limiter.cc:52-87)

* ``Schema`` / ``Options``  the reference's JSON schema and YAML option files (schema.py)
* ``QueryPlan.from_json``   the install-query JSON produced by the Java GSL client (plan.py)
* ``FileLoader``            pattern-file driven record ingestion (file_loader.py; dataloader SDK)
* ``HttpFrontEnd``          ``/infer?qid&vid`` + admin API (install query, checkpoint, barrier, stats)
* ``CheckpointManager`` / ``BarrierMonitor``   coordinator duties (coordinator.py)
* ``PartitionedGraphService``  vid-hash partitioned stores (one per GPU) with per-hop routing (partitioned.py)

* ``StreamingCluster``  the reference's decoupled deployment: P sampling workers (ingest partitions, sample stores,
                      subscription tables) x S serving workers (k-hop row caches fed by published samples) over
                      ``LogChannel`` topics (workers.py), driven by ``Coordinator`` (worker registry, barriers,
                      consistent cluster checkpoints)

* ``client``            GSL client for Python programs (the role of the reference's Java client: ``Graph.connect`` /
                      fluent traversal -> install-query JSON / ``install`` / ``run`` / ``EgoGraph`` hop tensors), client.py
* ``python -m graphlearn_b200.dgs``   service process entry point (restore -> install -> load -> serve -> final checkpoint),
                      scheduled by the Helm chart in ``deploy/dgs``

Record batches are dicts of arrays (``LogChannel`` persists them as columnar binary segments instead of Kafka/FlatBuffers).
"""
from .coordinator import BarrierMonitor, CheckpointManager, Coordinator, WorkerRegistry  # noqa: F401
from .file_loader import FileLoader, GroupProducer, RecordBatchBuilder, decode_record_batch, encode_record_batch  # noqa: F401
from .http_server import HttpFrontEnd  # noqa: F401
from .partitioned import PartitionedGraphService, Partitioner  # noqa: F401
from .plan import PlanNode, QueryPlan  # noqa: F401
from .schema import Options, Schema  # noqa: F401
from .workers import LogChannel, SamplingWorker, ServingWorker, StreamingCluster, SubscriptionTable  # noqa: F401
from .service import AdaptiveRateLimiter, DynamicGraphService, SampleStore  # noqa: F401
