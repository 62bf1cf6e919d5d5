"""Minimal zarr v3 WRITER for the tests of scanpy_b200._io (test infrastructure: the product only reads).
Writes the AnnData on-disk CSR group layout (`<group>/{data,indices,indptr}` + attributes encoding-type/shape) as a
directory store or a zip, with codecs [bytes], [bytes, zstd] or sharding_indexed{[bytes, zstd]} + crc32c'd shard index."""
import ctypes
import json
import struct
import zipfile
from pathlib import Path

import numpy as np

_NAMES = {"<f4": "float32", "<f8": "float64", "<i4": "int32", "<i8": "int64"}
_z = ctypes.CDLL("libzstd.so.1")
_z.ZSTD_compressBound.restype = ctypes.c_size_t
_z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
_z.ZSTD_compress.restype = ctypes.c_size_t
_z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]


def _zstd(buf: bytes) -> bytes:
    cap = _z.ZSTD_compressBound(len(buf))
    dst = ctypes.create_string_buffer(cap)
    n = _z.ZSTD_compress(dst, cap, buf, len(buf), 3)
    return dst.raw[:n]


def _crc32c(data: bytes) -> int:
    crc = 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def _array_files(path: str, a: np.ndarray, mode: str, chunk: int, inner: int):
    a = np.ascontiguousarray(a)
    dt = a.dtype.newbyteorder("<").str
    files = {}
    enc = (lambda b: _zstd(b)) if mode in ("zstd", "sharded") else (lambda b: b)
    inner_codecs = [{"name": "bytes", "configuration": {"endian": "little"}}] + ([{"name": "zstd", "configuration": {"level": 3, "checksum": False}}] if mode != "raw" else [])

    def padded(lo, size):
        part = a[lo:lo + size]
        if len(part) < size:
            part = np.concatenate([part, np.zeros(size - len(part), a.dtype)])
        return part.astype(dt).tobytes()

    if mode == "sharded":
        per = chunk // inner
        for s in range(-(-len(a) // chunk)):
            body, index = b"", b""
            for k in range(per):
                lo = s * chunk + k * inner
                if lo >= len(a):
                    index += struct.pack("<QQ", 2**64 - 1, 2**64 - 1)
                    continue
                c = enc(padded(lo, inner))
                index += struct.pack("<QQ", len(body), len(c))
                body += c
            index += struct.pack("<I", _crc32c(index))
            files[f"{path}/c/{s}"] = body + index
        codecs = [{"name": "sharding_indexed", "configuration": {
            "chunk_shape": [inner], "codecs": inner_codecs,
            "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}], "index_location": "end"}}]
    else:
        for s in range(-(-len(a) // chunk)):
            files[f"{path}/c/{s}"] = enc(padded(s * chunk, chunk))
        codecs = inner_codecs
    meta = {"zarr_format": 3, "node_type": "array", "shape": [int(len(a))], "data_type": _NAMES[dt],
            "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [chunk]}},
            "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}}, "fill_value": 0, "codecs": codecs,
            "attributes": {}}
    files[f"{path}/zarr.json"] = json.dumps(meta).encode()
    return files


def write_csr_store(target, x, *, group="X", mode="sharded", chunk=4096, inner=512, as_zip=False):
    """x: scipy CSR.  mode: 'raw' | 'zstd' | 'sharded'."""
    files = {"zarr.json": json.dumps({"zarr_format": 3, "node_type": "group", "attributes": {"encoding-type": "anndata"}}).encode(),
             f"{group}/zarr.json": json.dumps({"zarr_format": 3, "node_type": "group", "attributes": {
                 "encoding-type": "csr_matrix", "encoding-version": "0.1.0", "shape": [int(x.shape[0]), int(x.shape[1])]}}).encode()}
    files.update(_array_files(f"{group}/data", x.data.astype(np.float32), mode, chunk, inner))
    files.update(_array_files(f"{group}/indices", x.indices.astype(np.int32), mode, chunk, inner))
    files.update(_array_files(f"{group}/indptr", x.indptr.astype(np.int64), mode, chunk, inner))
    target = Path(target)
    if as_zip:
        with zipfile.ZipFile(target, "w", zipfile.ZIP_STORED) as z:
            for k, v in files.items():
                z.writestr(k, v)
    else:
        for k, v in files.items():
            f = target / k
            f.parent.mkdir(parents=True, exist_ok=True)
            f.write_bytes(v)
    return target
