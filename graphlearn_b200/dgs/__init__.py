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

Kafka, RocksDB, the HTTP front end, the Java client and the Helm chart are deployment glue around
this core and are out of scope; ``apply_updates`` takes record batches (dict of arrays) directly.
"""
from .service import AdaptiveRateLimiter, DynamicGraphService, QueryPlan, SampleStore  # noqa: F401
