"""File-system registry (N9): path scheme -> file system, like the reference's ``REGISTER_FILE_SYSTEM`` /
``Env::GetFileSystem`` (graphlearn/src/platform/file_system.h, env.cc:125-145, local_file_system.cc,
hadoop/hadoop_file_system.cc; scheme handling in core/io/slice_reader.h:174-186).

    local paths, ``file://``             LocalFileSystem (the native multi-threaded parser reads them in place)
    ``hdfs://``, ``viewfs://``           HadoopFileSystem over ``pyarrow.fs`` (needs libhdfs + a Hadoop client on the box)
    ``odps://`` and anything else        ``register_file_system(scheme, fs)`` - a user supplied object with
                                         ``listdir(path) -> [paths]``, ``isdir(path)``, ``open(path) -> binary file``

Remote files are SPOOLED: the native loader (csrc/host_loader.cpp) parses byte ranges of local files with one thread per
range, so a remote object is first streamed to a local spool directory (``GLB_SPOOL_DIR`` or the system temp dir) and
parsed from there - the reference's ``ByteStreamAccessFile`` reads through libhdfs record by record instead.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import tempfile
from typing import Dict, List, Tuple

from .. import errors


class FileSystem(object):
    """Interface of a registered file system."""

    def listdir(self, path: str) -> List[str]:
        raise NotImplementedError

    def isdir(self, path: str) -> bool:
        raise NotImplementedError

    def open(self, path: str):
        raise NotImplementedError


class LocalFileSystem(FileSystem):
    @staticmethod
    def _strip(path: str) -> str:
        return path[len("file://"):] if path.startswith("file://") else path

    def listdir(self, path):
        p = self._strip(path)
        return sorted(os.path.join(p, f) for f in os.listdir(p) if not f.startswith("."))

    def isdir(self, path):
        return os.path.isdir(self._strip(path))

    def open(self, path):
        return open(self._strip(path), "rb")

    def local_path(self, path):
        return self._strip(path)


class HadoopFileSystem(FileSystem):
    """``hdfs://host:port/path`` and ``viewfs://...`` through ``pyarrow.fs.HadoopFileSystem`` (libhdfs)."""

    def __init__(self):
        self._fs: Dict[str, object] = {}

    def _split(self, path: str) -> Tuple[object, str]:
        try:
            from pyarrow import fs as pafs
        except Exception as e:      # pragma: no cover
            raise errors.UnavailableError("hdfs:// sources need pyarrow: %r" % (e,))
        scheme, rest = path.split("://", 1)
        authority, _, p = rest.partition("/")
        key = scheme + "://" + authority
        if key not in self._fs:
            try:
                self._fs[key] = pafs.FileSystem.from_uri(key + "/")[0]
            except Exception as e:
                raise errors.UnavailableError("cannot open %s (is libhdfs / a Hadoop client installed on this box?): %s" % (key, e))
        return self._fs[key], "/" + p

    def listdir(self, path):
        from pyarrow import fs as pafs
        f, p = self._split(path)
        base = path.split("://", 1)[0] + "://" + path.split("://", 1)[1].partition("/")[0]
        return sorted(base + i.path for i in f.get_file_info(pafs.FileSelector(p)) if not os.path.basename(i.path).startswith("."))

    def isdir(self, path):
        from pyarrow import fs as pafs
        f, p = self._split(path)
        return f.get_file_info(p).type == pafs.FileType.Directory

    def open(self, path):
        f, p = self._split(path)
        return f.open_input_stream(p)


_REGISTRY: Dict[str, FileSystem] = {"": LocalFileSystem(), "file": LocalFileSystem(), "hdfs": HadoopFileSystem(),
                                    "viewfs": HadoopFileSystem()}


def register_file_system(scheme: str, fs: FileSystem):
    """Make ``scheme://...`` paths loadable (``odps://`` tables, object stores, ...)."""
    _REGISTRY[scheme.lower()] = fs


def get_file_system(path: str) -> FileSystem:
    scheme = path.split("://", 1)[0].lower() if "://" in path else ""
    fs = _REGISTRY.get(scheme)
    if fs is None:
        raise errors.UnimplementedError("no file system registered for %r paths (use graphlearn_b200.io.register_file_system)" % (scheme,))
    return fs


def expand(path: str) -> List[str]:
    """comma list / directory -> list of file paths (scheme preserved)"""
    out: List[str] = []
    for p in [x.strip() for x in path.split(",") if x.strip()]:
        fs = get_file_system(p)
        out.extend(fs.listdir(p) if fs.isdir(p) else [p])
    return out


def localize(path: str) -> str:
    """Local path of a (possibly remote) file: local files as they are, remote objects spooled once per process."""
    fs = get_file_system(path)
    if isinstance(fs, LocalFileSystem):
        return fs.local_path(path)
    spool = os.environ.get("GLB_SPOOL_DIR") or os.path.join(tempfile.gettempdir(), "glb_spool")
    os.makedirs(spool, exist_ok=True)
    dst = os.path.join(spool, hashlib.sha1(path.encode()).hexdigest()[:16] + "_" + os.path.basename(path))
    if not os.path.exists(dst):
        tmp = dst + ".part%d" % os.getpid()
        with fs.open(path) as src, open(tmp, "wb") as out:
            shutil.copyfileobj(src, out, 1 << 24)
        os.replace(tmp, dst)
    return dst
