"""Self-describing byte container for a compress() result ("a self-describing container (shapes, min/max)").

The reference keeps the coded patch as a Python dict (`custom_compress`, eval_selfcontained_entropy.py:68-74:
strings, shape_y, shape_z, min_y, max_y, min_z, max_z) and never writes it anywhere.  Format SIC-CONT-1 serialises exactly
that dict, nothing else is needed to decode (the supports already include the tail):

    header   "SICC" | u16 version=1 | u16 flags=0 | u32 B | u32 shape_y[1:4] | u32 shape_z[1:4]          (36 bytes)
    index    per patch: i32 min_z, max_z, min_y, max_y | u32 len_z, len_y                             (24 bytes each)
    payload  per patch: z stream, y stream (format SIC-RANS-1)
    trailer  u32 CRC-32 of everything before it

All integers little-endian.  `unpack` verifies magic, version, lengths and the CRC, so a truncated or damaged file is
reported instead of being decoded into garbage (the per-stream erasure check of the rANS decoder remains underneath).
Host-side only: plain bytes in, plain dict out.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict

MAGIC = b"SICC"
VERSION = 1
_HEADER = struct.Struct("<4sHHI3I3I")
_ENTRY = struct.Struct("<4i2I")


class ContainerError(ValueError):
    pass


MAX_SUPPORT = 4096      # widest support (max - min + 1) the coder's tables hold (csrc/rans_*.c*: kMaxL)


def validate_supports(compressed: Dict, B: int) -> None:
    """Structural checks a CRC cannot give (anyone can recompute a CRC): non-empty latent shapes, min <= max and a support
    the table builder and the decoder can hold.  decompress() runs this on every dict, whatever its origin."""
    for k in ("shape_y", "shape_z"):
        shp = list(compressed[k])
        if len(shp) != 4 or shp[0] != B or any(int(d) <= 0 for d in shp[1:]) or any(int(d) > (1 << 20) for d in shp[1:]):
            raise ContainerError(f"{k} = {shp} does not describe {B} non-empty patches")
    for lo_k, hi_k in (("min_z", "max_z"), ("min_y", "max_y")):
        lo, hi = list(compressed[lo_k]), list(compressed[hi_k])
        if len(lo) != B or len(hi) != B:
            raise ContainerError(f"{lo_k}/{hi_k} have {len(lo)}/{len(hi)} entries for {B} patches")
        for b in range(B):
            width = int(hi[b]) - int(lo[b]) + 1
            if width < 1 or width > MAX_SUPPORT:
                raise ContainerError(f"patch {b}: support [{int(lo[b])}, {int(hi[b])}] ({lo_k}/{hi_k}) is empty or wider than {MAX_SUPPORT}")


def pack(compressed: Dict) -> bytes:
    strings = compressed["strings"]
    B = len(strings)
    sy, sz = list(compressed["shape_y"]), list(compressed["shape_z"])
    if len(sy) != 4 or len(sz) != 4 or sy[0] != B or sz[0] != B:
        raise ContainerError(f"shape_y/shape_z {sy}/{sz} do not describe {B} patches")
    for k in ("min_y", "max_y", "min_z", "max_z"):
        if len(compressed[k]) != B:
            raise ContainerError(f"{k} has {len(compressed[k])} entries for {B} patches")
    out = [_HEADER.pack(MAGIC, VERSION, 0, B, *sy[1:], *sz[1:])]
    for b in range(B):
        z, y = strings[b]
        out.append(_ENTRY.pack(int(compressed["min_z"][b]), int(compressed["max_z"][b]), int(compressed["min_y"][b]),
                               int(compressed["max_y"][b]), len(z), len(y)))
    for z, y in strings:
        out.append(bytes(z))
        out.append(bytes(y))
    body = b"".join(out)
    return body + struct.pack("<I", zlib.crc32(body) & 0xFFFFFFFF)


def unpack(data: bytes) -> Dict:
    data = bytes(data)
    if len(data) < _HEADER.size + 4:
        raise ContainerError("container shorter than its header")
    magic, version, flags, B, *dims = _HEADER.unpack_from(data, 0)
    if magic != MAGIC:
        raise ContainerError("not a SIC-CONT container (bad magic)")
    if version != VERSION or flags != 0:
        raise ContainerError(f"unsupported container version {version} / flags {flags}")
    body, (crc,) = data[:-4], struct.unpack("<I", data[-4:])
    if zlib.crc32(body) & 0xFFFFFFFF != crc:
        raise ContainerError("container CRC mismatch (truncated or damaged)")
    off = _HEADER.size
    if off + B * _ENTRY.size > len(body):
        raise ContainerError("container index is truncated")
    entries = [_ENTRY.unpack_from(body, off + b * _ENTRY.size) for b in range(B)]
    off += B * _ENTRY.size
    validate_supports({"shape_y": [B] + list(dims[:3]), "shape_z": [B] + list(dims[3:]),
                       "min_y": [e[2] for e in entries], "max_y": [e[3] for e in entries],
                       "min_z": [e[0] for e in entries], "max_z": [e[1] for e in entries]}, B)
    strings = []
    for (_, _, _, _, lz, ly) in entries:
        if off + lz + ly > len(body):
            raise ContainerError("container payload is truncated")
        strings.append([body[off:off + lz], body[off + lz:off + lz + ly]])
        off += lz + ly
    if off != len(body):
        raise ContainerError(f"{len(body) - off} stray bytes after the last stream")
    return {"strings": strings, "shape_y": [B] + list(dims[:3]), "shape_z": [B] + list(dims[3:]),
            "min_y": [e[2] for e in entries], "max_y": [e[3] for e in entries],
            "min_z": [e[0] for e in entries], "max_z": [e[1] for e in entries]}
