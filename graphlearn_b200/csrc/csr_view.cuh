// Device view of a hash-partitioned CSR adjacency whose shards live in the HBM
// of every rank of the box.  All pointer tables hold IPC-mapped peer pointers,
// so a kernel reads a remote adjacency row with ordinary ld.global over
// NVLink (hot path (a) of BASELINE.json).  Replaces the reference's
// MemoryTopoStorage / CompressedMemoryAdjMatrix
// (graphlearn/src/core/graph/storage/memory_adj_matrix.cc:159-225) plus the
// per-hop gRPC fan-out.
#pragma once
#include "common.cuh"

namespace glb {

struct CsrView {
  PeerTable indptr;    // const int64_t*  [nrows[r] + 1]
  PeerTable indices;   // const int64_t*  destination vids
  PeerTable eids;      // const int64_t*  global edge ids   (may be null)
  PeerTable cumw;      // const float*    inclusive in-row prefix sums of the
                       //                 sampling weight    (may be null)
  PeerTable ts;        // const int64_t*  edge timestamps, rows sorted asc
                       //                                   (may be null)
  PeerTable sorted;    // const int64_t*  destination vids of every row sorted ASCENDING (same indptr; may be null):
                       //                 membership tests by binary search (node2vec "is x a neighbour of the parent")
  int64_t nrows[kMaxWorld];
  int world;
};

struct RowRef {
  const int64_t* indices;
  const int64_t* eids;
  const float* cumw;
  const int64_t* ts;
  const int64_t* sorted;
  int64_t beg;
  int64_t deg;
};

// Resolve the adjacency row of `vid` on its owner.  deg == 0 for unknown ids,
// matching the reference where GetNeighbors() returns an empty array for an
// id that is not a source vertex.
__device__ __forceinline__ RowRef csr_row(const CsrView& g, int64_t vid) {
  RowRef r;
  r.indices = nullptr; r.eids = nullptr; r.cumw = nullptr; r.ts = nullptr; r.sorted = nullptr;
  r.beg = 0; r.deg = 0;
  if (vid < 0) return r;
  int owner = (int)(vid % g.world);
  int64_t row = vid / g.world;
  if (row >= g.nrows[owner]) return r;
  const int64_t* ip = reinterpret_cast<const int64_t*>(g.indptr.p[owner]);
  int64_t b = __ldg(ip + row);
  int64_t e = __ldg(ip + row + 1);
  r.beg = b;
  r.deg = e - b;
  r.indices = reinterpret_cast<const int64_t*>(g.indices.p[owner]);
  r.eids = reinterpret_cast<const int64_t*>(g.eids.p[owner]);
  r.cumw = reinterpret_cast<const float*>(g.cumw.p[owner]);
  r.ts = reinterpret_cast<const int64_t*>(g.ts.p[owner]);
  r.sorted = reinterpret_cast<const int64_t*>(g.sorted.p[owner]);
  return r;
}

}  // namespace glb
