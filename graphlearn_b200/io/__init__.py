"""File IO helpers (N9/N10): thin public wrappers over the native parser and the embedding writer.

    cols = gl_io.read_table("item.tsv", gl.Decoder(labeled=True, attr_types=["float"] * 100), kind="node")
    cols["a"] (ids) / cols["b"] (dst ids, edges) / "w" / "label" / "ts" / "ia" / "fa" / "strs"
    gl_io.read_table(path, decoder, part_index=r, part_count=W)     # this rank's byte-range slice

Formats are the reference's (docs/en/gl/graph/data_loader.md:117-153): TSV with an optional
``name:type`` header line, attributes joined by ``decoder.attr_delimiter``."""
from __future__ import annotations

from ..store.graph_store import Source, _expand_paths as expand_paths, _load_source
from ..utils.checkpoint import save_embeddings  # noqa: F401
from .filesystem import FileSystem, get_file_system, register_file_system  # noqa: F401


def read_table(path: str, decoder, kind: str = "node", part_index: int = 0, part_count: int = 1) -> dict:
    """Parse a node / edge table (file, directory, comma list or file:// URL) into columnar CPU tensors."""
    assert kind in ("node", "edge")
    types = "t" if kind == "node" else ("s", "d", "e")
    return _load_source(Source(kind, path, types, decoder), part_index, part_count)


__all__ = ["read_table", "expand_paths", "save_embeddings", "FileSystem", "get_file_system", "register_file_system"]
