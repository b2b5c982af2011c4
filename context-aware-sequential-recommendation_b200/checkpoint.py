"""TensorFlow-1.x checkpoint ("tensor bundle") reader / writer with the reference's variable names, without TensorFlow.

The reference saves with `tf.train.Saver()` (`main.py:153-159,226-228`): `<prefix>.index` is a LevelDB-style SSTable
whose values are `BundleEntryProto`s (dtype, shape, shard, offset, size, crc32c) and `<prefix>.data-00000-of-00001`
holds the raw little-endian tensors.  This module parses / emits exactly that subset (no compression, one shard,
float32 and int32 tensors) and maps the reference's variable names to this package's role names,
so reference-trained weights load into the CUDA path (`--test_model` behaviour) and weights trained here can be read
back by the reference.

    vars = read_bundle(".../model.ckpt")                       # {tf name: ndarray}
    params = to_role_names("sasrec", vars, num_blocks=2)       # {role name: ndarray}  (Adam slots are skipped)
    model.load_state_dict(params)
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DT = {1: np.float32, 3: np.int32, 9: np.int64, 2: np.float64}
_DT_INV = {np.dtype(np.float32): 1, np.dtype(np.int32): 3, np.dtype(np.int64): 9, np.dtype(np.float64): 2}


# ------------------------------------------------------------------------------------------------ varint / proto
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Flat protobuf decode: {field number: [values]} (varints as int, length-delimited as bytes, fixed32 as int)."""
    pos, out = 0, {}
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _entry_shape(shape_bytes: bytes):
    dims = []
    for d in _parse_proto(shape_bytes).get(2, []):        # TensorShapeProto.dim
        dims.append(_parse_proto(d).get(1, [0])[0])       # Dim.size
    return tuple(int(x) for x in dims)


# ------------------------------------------------------------------------------------------------ SSTable
def _block_entries(block: bytes):
    """(key, value) pairs of one table block (prefix-compressed keys, restart array at the end)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_block(buf: bytes, offset: int, size: int) -> bytes:
    if buf[offset + size] != 0:
        raise ValueError("compressed table blocks are not supported (TF writes bundles uncompressed)")
    return buf[offset:offset + size]


def read_bundle(prefix: str) -> Dict[str, np.ndarray]:
    """All tensors of `<prefix>.index` / `<prefix>.data-00000-of-00001`."""
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    if struct.unpack_from("<Q", idx, len(idx) - 8)[0] != _MAGIC:
        raise ValueError("not a tensor-bundle index (bad table magic)")
    footer = idx[-48:]
    pos = 0
    _, pos = _varint(footer, pos)     # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)  # index handle
    isize, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(_read_block(idx, ioff, isize)):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        for k, v in _block_entries(_read_block(idx, boff, bsize)):
            entries[k.decode()] = v
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        data = f.read()
    out = {}
    for name, raw in entries.items():
        if name == "":
            continue                   # BundleHeaderProto
        e = _parse_proto(raw)
        dt = _DT[e.get(1, [1])[0]]
        shape = _entry_shape(e[2][0]) if 2 in e else ()
        off, size = e.get(4, [0])[0], e.get(5, [0])[0]
        out[name] = np.frombuffer(data, dtype=dt, count=size // np.dtype(dt).itemsize, offset=off).reshape(shape).copy()
    return out


# ------------------------------------------------------------------------------------------------ writer
_CRC_TABLE = None


def _crc32c(data: bytes) -> int:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = np.array(t, dtype=np.uint32)
    crc = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _block(pairs) -> bytes:
    """One uncompressed block with a restart point at every entry (no prefix sharing)."""
    body, restarts = bytearray(), []
    for k, v in pairs:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    for r in restarts or [0]:
        body += struct.pack("<I", r)
    body += struct.pack("<I", max(1, len(restarts)))
    return bytes(body)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", _mask_crc(_crc32c(block + b"\x00")))


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], checksum_data: bool = False):
    """Writes `<prefix>.index` + `<prefix>.data-00000-of-00001` readable by `read_bundle` (and by TF's loader; the
    per-tensor crc32c field is filled only when `checksum_data` is set — pure-Python crc32c is slow on big tables)."""
    names = sorted(tensors)
    data, pairs = bytearray(), []
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards=1, version{producer=1}
    pairs.append((b"", header))
    for n in names:
        a = np.ascontiguousarray(tensors[n])
        raw = a.tobytes()
        shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in a.shape))
        e = b"\x08" + _put_varint(_DT_INV[a.dtype]) + b"\x12" + _put_varint(len(shape)) + shape
        if len(data):
            e += b"\x20" + _put_varint(len(data))
        e += b"\x28" + _put_varint(len(raw))
        if checksum_data:
            e += b"\x35" + struct.pack("<I", _mask_crc(_crc32c(raw)))
        pairs.append((n.encode(), e))
        data += raw
    dblock = _with_trailer(_block(pairs))
    handle = _put_varint(0) + _put_varint(len(dblock) - 5)
    iblock_off = len(dblock)
    iblock = _with_trailer(_block([(pairs[-1][0] + b"\x00", handle)]))
    meta_off = iblock_off + len(iblock)
    mblock = _with_trailer(_block([]))
    footer = _put_varint(meta_off) + _put_varint(len(mblock) - 5) + _put_varint(iblock_off) + _put_varint(len(iblock) - 5)
    footer = footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(dblock + iblock + mblock + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# ------------------------------------------------------------------------------------------------ name mapping
def _block_map(tf_scope: str, role: str):
    """One transformer block's variables (`feedforward` lives under scope `multihead_attention`)."""
    m = {f"{tf_scope}/ln/Variable": f"{role}.ln1.beta", f"{tf_scope}/ln/Variable_1": f"{role}.ln1.gamma",
         f"{tf_scope}/ln_1/Variable": f"{role}.ln2.beta", f"{tf_scope}/ln_1/Variable_1": f"{role}.ln2.gamma"}
    for tf_d, r in (("dense", "q"), ("dense_1", "k"), ("dense_2", "v")):
        m[f"{tf_scope}/self_attention/{tf_d}/kernel"] = f"{role}.{r}.w"
        m[f"{tf_scope}/self_attention/{tf_d}/bias"] = f"{role}.{r}.b"
    for tf_c, r in (("conv1d", "ffn1"), ("conv1d_1", "ffn2")):
        m[f"{tf_scope}/multihead_attention/{tf_c}/kernel"] = f"{role}.{r}.w"   # stored [1, H, H]
        m[f"{tf_scope}/multihead_attention/{tf_c}/bias"] = f"{role}.{r}.b"
    return m


def name_map(model: str, num_blocks: int) -> Dict[str, str]:
    """TF variable name -> role name for the SASRec variants (models/sasrec.py scopes).  CAST checkpoints shipped with
    the reference were written by earlier revisions of the model files, so only the SASRec family is
    mapped by name."""
    if model not in ("sasrec", "sasrec_static"):
        raise ValueError("name mapping is defined for sasrec / sasrec_static")
    m = {"SASRec/input_embeddings/lookup_table": "item_emb", "SASRec/ln/Variable": "main.lnf.beta",
         "SASRec/ln/Variable_1": "main.lnf.gamma"}
    if model == "sasrec":
        m["SASRec/dec_pos/lookup_table"] = "pos_emb"
    for i in range(num_blocks):
        m.update(_block_map(f"SASRec/num_blocks_{i}", f"main.{i}"))
    return m


def to_role_names(model: str, tf_vars: Dict[str, np.ndarray], num_blocks: int) -> Dict[str, np.ndarray]:
    out = {}
    for tf_name, role in name_map(model, num_blocks).items():
        a = tf_vars[tf_name]
        out[role] = a.reshape(a.shape[-2], a.shape[-1]) if a.ndim == 3 else a   # conv1d kernels [1,H,H] -> [H,H]
    return out


def to_tf_names(model: str, params: Dict[str, np.ndarray], num_blocks: int) -> Dict[str, np.ndarray]:
    out = {}
    for tf_name, role in name_map(model, num_blocks).items():
        a = np.asarray(params[role], dtype=np.float32)
        out[tf_name] = a[None] if "/conv1d" in tf_name and tf_name.endswith("kernel") else a
    return out
