"""Reader/writer for the tiny "BPK1" named-array container.

The container is what the reference-side adapter (oracle/ref/ref_harness.cc)
writes and what the C++ host layer reads: ``"BPK1"``, uint32 count, then per
array: uint32 name length, name, uint32 dtype (0 = int32, 1 = float64,
2 = int64, 3 = uint8), uint32 ndim, int64 dims[ndim], raw little-endian data.
"""
import struct

import numpy as np

_DTYPES = {0: np.dtype("<i4"), 1: np.dtype("<f8"), 2: np.dtype("<i8"), 3: np.dtype("u1")}
_CODES = {np.dtype("int32"): 0, np.dtype("float64"): 1, np.dtype("int64"): 2, np.dtype("uint8"): 3}


def read_bpk(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(4) != b"BPK1":
            raise ValueError(f"{path}: not a BPK1 file")
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (ln,) = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode()
            dtype, nd = struct.unpack("<II", f.read(8))
            dims = struct.unpack(f"<{nd}q", f.read(8 * nd))
            dt = _DTYPES[dtype]
            cnt = int(np.prod(dims)) if nd else 1
            arr = np.frombuffer(f.read(cnt * dt.itemsize), dtype=dt).reshape(dims).copy()
            out[name] = arr
    return out


def write_bpk(path, arrays):
    with open(path, "wb") as f:
        f.write(b"BPK1")
        f.write(struct.pack("<I", len(arrays)))
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            code = _CODES[arr.dtype]
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", code, arr.ndim))
            f.write(struct.pack(f"<{arr.ndim}q", *arr.shape))
            f.write(arr.tobytes())
