"""On-disk CSR -> row chunks for the out-of-core path.

Reads the AnnData on-disk sparse layout (`<group>/data`, `<group>/indices`, `<group>/indptr` + group attribute `shape`;
the layout `anndata.io.write_zarr` produces and `sc.read_zarr` / `sc.read_h5ad(backed='r')` consume,
src/scanpy/readwrite.py:71-156,657-738; fixture layout: SURVEY.md Appendix D) from a **zarr v3** store — a directory or
a `.zip` (the reference's own in-tree fixture `src/scanpy/datasets/10x_pbmc68k_reduced.zarr.zip` is one) — without zarr /
h5py / anndata (none of them is installed here): JSON metadata, the `bytes` codec, `zstd` through libzstd (ctypes) and the
`sharding_indexed` codec are all this layout needs.

`ZarrCSR.row_chunks(chunk_size)` decodes ONLY the 1-D chunks that cover a row range and yields
`(r0, r1, indptr int64 (re-based), indices int32, data float32)`; `scanpy_b200.pp.pca(adata, chunked=True)` feeds those
chunks to `sb2_pca_stream_accumulate_f32` / `sb2_pca_stream_project_f32`, so the matrix never has to fit in host or device
memory at once (`adata.X` may be a `ZarrCSR`; `read_zarr_backed` builds such an object).  h5ad (HDF5) is NOT read: there is
no HDF5 library in this image, and a chunk B-tree walker is outside the path.
"""
from __future__ import annotations

import ctypes
import json
import struct
import zipfile
from pathlib import Path

import numpy as np

_DTYPES = {"float32": "<f4", "float64": "<f8", "int32": "<i4", "int64": "<i8", "int16": "<i2", "int8": "i1", "uint8": "u1",
           "uint16": "<u2", "uint32": "<u4", "uint64": "<u8", "bool": "?"}
_zstd = None


def _zstd_lib():
    global _zstd
    if _zstd is None:
        lib = ctypes.CDLL("libzstd.so.1")
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        _zstd = lib
    return _zstd


def _zstd_decompress(buf: bytes, expect: int) -> bytes:
    lib = _zstd_lib()
    size = lib.ZSTD_getFrameContentSize(buf, len(buf))
    if size in (2**64 - 1, 2**64 - 2):  # unknown / error: fall back to the chunk's nominal size
        size = expect
    dst = ctypes.create_string_buffer(int(size))
    got = lib.ZSTD_decompress(dst, int(size), buf, len(buf))
    if lib.ZSTD_isError(got):
        raise OSError("zstd: corrupt chunk")
    return dst.raw[:got]


class _Store:
    """Key -> bytes over a directory or a zip archive (zarr v3 stores are flat key spaces)."""

    def __init__(self, path):
        self.path = Path(path)
        self._zip = zipfile.ZipFile(self.path) if self.path.is_file() else None
        if self._zip is None and not self.path.is_dir():
            raise FileNotFoundError(str(path))

    def get(self, key: str):
        if self._zip is not None:
            try:
                return self._zip.read(key)
            except KeyError:
                return None
        f = self.path / key
        return f.read_bytes() if f.is_file() else None


class _Array1D:
    """A 1-D zarr v3 array: codecs [bytes(, zstd)] or [sharding_indexed{[bytes(, zstd)]}]."""

    def __init__(self, store: _Store, path: str):
        raw = store.get(f"{path}/zarr.json")
        if raw is None:
            raise KeyError(f"{path}/zarr.json not found in {store.path}")
        meta = json.loads(raw)
        if meta.get("node_type") != "array" or len(meta["shape"]) != 1:
            raise ValueError(f"{path}: expected a 1-D zarr v3 array")
        self.store, self.path = store, path
        self.n = int(meta["shape"][0])
        self.dtype = np.dtype(_DTYPES[meta["data_type"]])
        self.outer = int(meta["chunk_grid"]["configuration"]["chunk_shape"][0])
        sep = meta.get("chunk_key_encoding", {}).get("configuration", {}).get("separator", "/")
        self._key = (lambda i: f"{path}/c{sep}{i}") if meta.get("chunk_key_encoding", {}).get("name", "default") == "default" \
            else (lambda i: f"{path}/{i}")
        codecs = meta["codecs"]
        self.fill = meta.get("fill_value", 0)
        if codecs and codecs[0]["name"] == "sharding_indexed":
            cfg = codecs[0]["configuration"]
            self.inner = int(cfg["chunk_shape"][0])
            self.sharded = True
            self.index_crc = any(c["name"] == "crc32c" for c in cfg.get("index_codecs", []))
            self.index_at_end = cfg.get("index_location", "end") == "end"
            inner_codecs = cfg["codecs"]
        else:
            self.inner, self.sharded, inner_codecs = self.outer, False, codecs
        names = [c["name"] for c in inner_codecs]
        if not names or names[0] != "bytes" or any(n not in ("bytes", "zstd") for n in names):
            raise NotImplementedError(f"{path}: unsupported codec chain {names} (bytes [+ zstd] only)")
        endian = inner_codecs[0].get("configuration", {}).get("endian", "little")
        if endian != "little" and self.dtype.itemsize > 1:
            raise NotImplementedError(f"{path}: big-endian chunks")
        self.zstd = "zstd" in names
        self._shard_cache: tuple[int, bytes] | None = None

    def _shard(self, s: int):
        if self._shard_cache is None or self._shard_cache[0] != s:
            self._shard_cache = (s, self.store.get(self._key(s)))
        return self._shard_cache[1]

    def _inner_chunk(self, c: int) -> np.ndarray:
        """Decoded inner chunk c (global inner-chunk index), length self.inner (fill-padded)."""
        nbytes = self.inner * self.dtype.itemsize
        if self.sharded:
            per = self.outer // self.inner
            raw = self._shard(c // per)
            if raw is None:
                return np.full(self.inner, self.fill, self.dtype)
            isz = 16 * per + (4 if self.index_crc else 0)
            index = raw[-isz:] if self.index_at_end else raw[:isz]
            off, nb = struct.unpack_from("<QQ", index, 16 * (c % per))
            if off == 2**64 - 1:
                return np.full(self.inner, self.fill, self.dtype)
            buf = raw[off:off + nb]
        else:
            buf = self.store.get(self._key(c))
            if buf is None:
                return np.full(self.inner, self.fill, self.dtype)
        if self.zstd:
            buf = _zstd_decompress(buf, nbytes)
        have = len(buf) // self.dtype.itemsize
        if have >= self.inner:
            return np.frombuffer(buf, self.dtype, count=self.inner)
        # a writer may store the last chunk of an array unpadded: fill the tail
        out = np.full(self.inner, self.fill, self.dtype)
        out[:have] = np.frombuffer(buf, self.dtype, count=have)
        return out

    def read(self, lo: int, hi: int) -> np.ndarray:
        """Elements [lo, hi): only the inner chunks overlapping the range are decoded."""
        lo, hi = max(0, lo), min(self.n, hi)
        out = np.empty(max(0, hi - lo), self.dtype)
        pos = lo
        while pos < hi:
            c = pos // self.inner
            chunk = self._inner_chunk(c)
            a = pos - c * self.inner
            take = min(self.inner - a, hi - pos)
            out[pos - lo:pos - lo + take] = chunk[a:a + take]
            pos += take
        return out


class ZarrCSR:
    """Lazy, row-sliceable view of an on-disk AnnData CSR group (zarr v3)."""

    def __init__(self, store_path, group: str = "X"):
        self._store = _Store(store_path)
        raw = self._store.get(f"{group}/zarr.json")
        if raw is None:
            raise KeyError(f"group {group!r} not found in {store_path}")
        attrs = json.loads(raw).get("attributes", {})
        enc = attrs.get("encoding-type")
        if enc != "csr_matrix":
            raise NotImplementedError(f"{group}: encoding-type {enc!r}; only 'csr_matrix' groups are streamed")
        self.shape = tuple(int(v) for v in attrs["shape"])
        self._indptr = _Array1D(self._store, f"{group}/indptr")
        self._indices = _Array1D(self._store, f"{group}/indices")
        self._data = _Array1D(self._store, f"{group}/data")
        self.nnz = self._data.n
        self.dtype = np.dtype(np.float32)

    def row_chunks(self, chunk_size: int):
        n = self.shape[0]
        chunk_size = max(1, int(chunk_size))
        for r0 in range(0, n, chunk_size):
            r1 = min(n, r0 + chunk_size)
            ip = self._indptr.read(r0, r1 + 1).astype(np.int64)
            lo, hi = int(ip[0]), int(ip[-1])
            yield r0, r1, ip - lo, self._indices.read(lo, hi).astype(np.int32, copy=False), \
                self._data.read(lo, hi).astype(np.float32, copy=False)

    def tocsr(self):
        """Materialise (small matrices / tests)."""
        from scipy import sparse

        ip = self._indptr.read(0, self.shape[0] + 1).astype(np.int64)
        return sparse.csr_matrix((self._data.read(0, self.nnz).astype(np.float32), self._indices.read(0, self.nnz).astype(np.int32), ip),
                                 shape=self.shape)


def read_zarr_backed(store_path, *, group: str = "X"):
    """MiniAnnData whose `.X` stays on disk (the role of `sc.read_zarr` / `read_h5ad(backed='r')` for the chunked path)."""
    from ._compat import MiniAnnData

    return MiniAnnData(ZarrCSR(store_path, group))
