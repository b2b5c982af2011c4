"""TensorFlow-1.x checkpoint ("tensor bundle") reader / writer with the reference's variable names, without TensorFlow.

The reference saves with `tf.train.Saver()` (`main.py:153-159,226-228`): `<prefix>.index` is a LevelDB-style SSTable
whose values are `BundleEntryProto`s (dtype, shape, shard, offset, size, crc32c) and `<prefix>.data-00000-of-00001`
holds the raw little-endian tensors.  This module parses / emits exactly that subset (no compression, one shard,
float32 and int32 tensors) and maps the reference's variable names to this package's role names,
so reference-trained weights load into the CUDA path (`--test_model` behaviour) and weights trained here can be read
back by the reference.

    vars = read_bundle(".../model.ckpt")                       # {tf name: ndarray}
    params = to_role_names("cast_4", vars, num_blocks=2)       # {role name: ndarray}
    model.load_state_dict(params)
    save_model(prefix, model, "cast_4", num_blocks=2)          # weights + Adam slots + beta powers + global_step
    restore_model(prefix, model, "cast_4", num_blocks=2)       # ... and back (training resumes bit-exactly)
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


def _index_entries(prefix: str) -> Dict[str, bytes]:
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
    return entries


def read_bundle_crcs(prefix: str) -> Dict[str, int]:
    """the masked crc32c TensorFlow stored for every tensor (BundleEntryProto.crc32c, field 6)"""
    out = {}
    for name, raw in _index_entries(prefix).items():
        if name:
            e = _parse_proto(raw)
            if 6 in e:
                out[name] = int(e[6][0])
    return out


def read_bundle(prefix: str) -> Dict[str, np.ndarray]:
    """All tensors of `<prefix>.index` / `<prefix>.data-00000-of-00001`."""
    entries = _index_entries(prefix)
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
    lib = _crc_lib()
    if lib is not None:
        return int(lib.cast_crc32c(data, len(data), 0))
    return _crc32c_py(data)


_CRC_LIB = False


def _crc_lib():
    """the C library's slicing-by-8 crc32c (host code; loads without a GPU).  None if the library is not built."""
    global _CRC_LIB
    if _CRC_LIB is False:
        try:
            from . import _lib
            _CRC_LIB = _lib.load_library()
        except Exception:
            _CRC_LIB = None
    return _CRC_LIB


def _crc32c_py(data: bytes) -> int:
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


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], checksum_data: bool = True):
    """Writes `<prefix>.index` + `<prefix>.data-00000-of-00001` in TensorFlow's tensor-bundle format: uncompressed
    SSTable of BundleEntryProtos with the masked crc32c of every tensor (TF's BundleReader verifies it on every read;
    tests/test_checkpoint.py checks the field byte for byte against the reference's own index)."""
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


DEAD = "<dead>"   # variables the reference creates but never trains (kept at their init, written back as such)


def _context_block_map(i: int):
    """One block of the CAST time-context tower (`CONTEXT/timeseq_num_blocks_i`, models/cast_1.py:42-60).  The block
    calls `normalize` three times: the first result (`self.timeseq_queries`, cast_1.py:45) is never used, so scope
    `ln` is a dead beta/gamma pair; `ln_1` normalises the queries and `ln_2` the feed-forward input."""
    sc, role = f"CONTEXT/timeseq_num_blocks_{i}", f"time.{i}"
    m = {f"{sc}/ln/Variable": DEAD + "beta", f"{sc}/ln/Variable_1": DEAD + "gamma",
         f"{sc}/ln_1/Variable": f"{role}.ln1.beta", f"{sc}/ln_1/Variable_1": f"{role}.ln1.gamma",
         f"{sc}/ln_2/Variable": f"{role}.ln2.beta", f"{sc}/ln_2/Variable_1": f"{role}.ln2.gamma"}
    for tf_d, r in (("dense", "q"), ("dense_1", "k"), ("dense_2", "v")):
        m[f"{sc}/self_attention/{tf_d}/kernel"] = f"{role}.{r}.w"
        m[f"{sc}/self_attention/{tf_d}/bias"] = f"{role}.{r}.b"
    for tf_c, r in (("conv1d", "ffn1"), ("conv1d_1", "ffn2")):
        m[f"{sc}/multihead_attention/{tf_c}/kernel"] = f"{role}.{r}.w"
        m[f"{sc}/multihead_attention/{tf_c}/bias"] = f"{role}.{r}.b"
    return m


CKPT_MODELS = ("sasrec", "sasrec_static", "cast_1", "cast_2", "cast_3", "cast_4", "cast_5", "cast_6")


def name_map(model: str, num_blocks: int) -> Dict[str, str]:
    """TF variable name -> role name (or DEAD...) for the models whose TensorFlow scopes are pinned by a checkpoint the
    reference ships (saved_models/ml-1m.txt/{sasrec*, cast_1..6}_*): SASRec (models/sasrec.py) and the time-context
    CAST variants (models/cast_1.py ... cast_6.py).  Variables are mapped BY ROLE: the context tower's extra `ln`
    pair is dead weight in the graph, the MLP kernels follow the concat order of engine.model_plan."""
    if model not in CKPT_MODELS:
        raise ValueError(f"TensorFlow variable names are defined for {CKPT_MODELS}; {model} checkpoints are .npz")
    m = {"SASRec/input_embeddings/lookup_table": "item_emb", "SASRec/ln/Variable": "main.lnf.beta",
         "SASRec/ln/Variable_1": "main.lnf.gamma"}
    if model == "sasrec":
        m["SASRec/dec_pos/lookup_table"] = "pos_emb"
    for i in range(num_blocks):
        m.update(_block_map(f"SASRec/num_blocks_{i}", f"main.{i}"))
    if model.startswith("cast_"):
        n = int(model.split("_")[1])
        m["CONTEXT/time_embeddings/lookup_table"] = "time_emb"
        m["CONTEXT/ln/Variable"] = "time.lnf.beta"
        m["CONTEXT/ln/Variable_1"] = "time.lnf.gamma"
        for i in range(num_blocks):
            m.update(_context_block_map(i))
        if n >= 3:
            m["INPUT-CONTEXT/hours_embeddings/lookup_table"] = "hours_emb"
            m["INPUT-CONTEXT/days_embeddings/lookup_table"] = "days_emb"
        if n >= 2:
            m.update({"SASRec/MLP/dense/kernel": "mlp.0.w", "SASRec/MLP/dense/bias": "mlp.0.b",
                      "SASRec/MLP/dense_1/kernel": "mlp.1.w", "SASRec/MLP/dense_1/bias": "mlp.1.b"})
    return m


def _from_tf(a: np.ndarray) -> np.ndarray:
    return a.reshape(a.shape[-2], a.shape[-1]) if a.ndim == 3 else a   # conv1d kernels [1,H,H] -> [H,H]


def to_role_names(model: str, tf_vars: Dict[str, np.ndarray], num_blocks: int) -> Dict[str, np.ndarray]:
    out = {}
    for tf_name, role in name_map(model, num_blocks).items():
        if not role.startswith(DEAD):
            out[role] = _from_tf(tf_vars[tf_name])
    return out


def to_tf_names(model: str, params: Dict[str, np.ndarray], num_blocks: int) -> Dict[str, np.ndarray]:
    out = {}
    for tf_name, role in name_map(model, num_blocks).items():
        if role.startswith(DEAD):     # never trained: still at the initialiser (beta = 0, gamma = 1)
            H = np.asarray(params["main.lnf.beta"]).shape[0]
            out[tf_name] = (np.ones if role.endswith("gamma") else np.zeros)(H, np.float32)
            continue
        a = np.asarray(params[role], dtype=np.float32)
        out[tf_name] = a[None] if "/conv1d" in tf_name and tf_name.endswith("kernel") else a
    return out


# ------------------------------------------------------------------------------------------------ whole training state
def save_model(prefix: str, model, model_name: str, num_blocks: int):
    """Everything `tf.train.Saver()` writes for the reference (main.py:153-159,226-228): the variables, their Adam
    slots `<var>/Adam` (m) and `<var>/Adam_1` (v) — for the variables that receive gradients —, `beta1_power`,
    `beta2_power` and `global_step`, under the reference's names."""
    eng = model.engine
    P = {k: v.detach().cpu().numpy() for k, v in eng.P.items()}
    out = to_tf_names(model_name, P, num_blocks)
    m_flat, v_flat = eng.m.detach().cpu().numpy(), eng.v.detach().cpu().numpy()
    for tf_name, role in name_map(model_name, num_blocks).items():
        if role.startswith(DEAD):
            continue
        off, n = eng.offsets[role], int(np.prod(eng.P[role].shape))
        shape = out[tf_name].shape
        out[tf_name + "/Adam"] = m_flat[off:off + n].reshape(shape).copy()
        out[tf_name + "/Adam_1"] = v_flat[off:off + n].reshape(shape).copy()
    st = eng.adam_state.detach().cpu().numpy().view(np.uint8)
    out["beta1_power"] = st[0:4].view(np.float32)[0].copy().reshape(())
    out["beta2_power"] = st[4:8].view(np.float32)[0].copy().reshape(())
    out["global_step"] = np.asarray(int(st[8:16].view(np.uint64)[0]), dtype=np.int32)   # tf.Variable(0) is int32
    write_bundle(prefix, out)
    return prefix


def restore_model(prefix: str, model, model_name: str, num_blocks: int, optimizer: bool = True):
    """Loads a bundle written by the reference or by `save_model`: weights by role; with `optimizer` also the Adam
    slots, beta powers and the step counter when the bundle has them (resume = the same bits as never stopping)."""
    import torch
    eng = model.engine
    tfv = read_bundle(prefix)
    model.load_state_dict(to_role_names(model_name, tfv, num_blocks))
    if not optimizer or "beta1_power" not in tfv:
        return tfv
    m_flat, v_flat = eng.m.detach().cpu().numpy().copy(), eng.v.detach().cpu().numpy().copy()
    for tf_name, role in name_map(model_name, num_blocks).items():
        if role.startswith(DEAD) or tf_name + "/Adam" not in tfv:
            continue
        off, n = eng.offsets[role], int(np.prod(eng.P[role].shape))
        m_flat[off:off + n] = _from_tf(tfv[tf_name + "/Adam"]).reshape(-1)
        v_flat[off:off + n] = _from_tf(tfv[tf_name + "/Adam_1"]).reshape(-1)
    eng.m.copy_(torch.from_numpy(m_flat).to(eng.device))
    eng.v.copy_(torch.from_numpy(v_flat).to(eng.device))
    st = np.zeros(16, np.uint8)
    st[0:4] = np.asarray(tfv["beta1_power"], np.float32).reshape(1).view(np.uint8)
    st[4:8] = np.asarray(tfv["beta2_power"], np.float32).reshape(1).view(np.uint8)
    st[8:16] = np.asarray([int(np.asarray(tfv["global_step"]).reshape(-1)[0])], np.uint64).view(np.uint8)
    eng.adam_state.copy_(torch.from_numpy(st.view(np.int64).copy()).to(eng.device))
    return tfv
