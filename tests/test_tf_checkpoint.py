"""TensorFlow V2 checkpoint reader, CPU only.  No TensorFlow and no TF-written file exist here, so
the reader is exercised against a WRITER of the same published formats (leveldb table, tensor_bundle.proto, snappy) kept in this
test, plus known-answer vectors for the primitives.  Unpinned against TensorFlow itself -- see merlot_b200/tf_checkpoint.py."""
import struct

import numpy as np
import pytest
import torch

from merlot_b200 import tf_checkpoint as T


# ---- an independent writer of the formats -------------------------------------------------------------------------
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def _entry_proto(dtype, shape, shard, offset, size, crc):
    dims = b"".join(_field(2, 2, _varint(len(_field(1, 0, _varint(d)))) + _field(1, 0, _varint(d))) for d in shape)
    return (_field(1, 0, _varint(dtype)) + _field(2, 2, _varint(len(dims)) + dims) + _field(3, 0, _varint(shard)) +
            _field(4, 0, _varint(offset)) + _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc)))


def _snappy_literals(data):  # a valid (if useless) snappy stream: literals of at most 60 bytes
    out = bytearray(_varint(len(data)))
    for i in range(0, len(data), 60):
        chunk = data[i:i + 60]
        out.append((len(chunk) - 1) << 2)
        out += chunk
    return bytes(out)


def _block(items, restart_interval=4):
    buf, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        buf += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def _write_table(path, items, per_block=3, compress=False):
    items = sorted(items)
    out, index = bytearray(), []

    def emit(block):
        payload, ctype = (_snappy_literals(block), 1) if compress else (block, 0)
        off = len(out)
        out.extend(payload)
        out.append(ctype)
        out.extend(struct.pack("<I", T.mask_crc(T.crc32c(payload + bytes([ctype])))))
        return off, len(payload)

    for i in range(0, len(items), per_block):
        chunk = items[i:i + per_block]
        off, size = emit(_block(chunk))
        index.append((chunk[-1][0] + b"\x00", _varint(off) + _varint(size)))  # any separator >= the block's last key
    meta_off, meta_size = emit(_block([]))
    idx_off, idx_size = emit(_block(index, restart_interval=1))
    footer = _varint(meta_off) + _varint(meta_size) + _varint(idx_off) + _varint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.TABLE_MAGIC)
    with open(path, "wb") as f:
        f.write(bytes(out) + footer)


def _write_checkpoint(prefix, tensors, compress=False):
    data, items = bytearray(), [(b"", _field(1, 0, _varint(1)) + _field(2, 0, _varint(0)))]  # header: 1 shard, little endian
    codes = {torch.float32: 1, torch.int64: 9, torch.bfloat16: 14, torch.int32: 3}
    for name, t in tensors.items():
        raw = t.contiguous().view(torch.uint8).numpy().tobytes() if t.dtype == torch.bfloat16 else t.contiguous().numpy().tobytes()
        items.append((name.encode(), _entry_proto(codes[t.dtype], list(t.shape), 0, len(data), len(raw), T.mask_crc(T.crc32c(raw)))))
        data += raw
    _write_table(prefix + ".index", items, compress=compress)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


def test_crc32c_vectorised_path_matches_the_byte_loop():
    """Buffers >= 64 KiB take the lane-parallel NumPy path (tensor payloads of a real checkpoint): same register as the byte
    loop for lengths around the lane and threshold boundaries, with and without a running crc; and RFC 3720's all-zero /
    all-ones / ascending vectors scaled past the threshold agree with the byte loop too."""
    import numpy as np
    rng = np.random.default_rng(0)
    for n in (65535, 65536, 65537, 4096 * 17 + 5, 300001):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(d) == T._crc32c_scalar(d, 0xFFFFFFFF) ^ 0xFFFFFFFF
        assert T.crc32c(d, crc=0x1234ABCD) == T._crc32c_scalar(d, 0x1234ABCD ^ 0xFFFFFFFF) ^ 0xFFFFFFFF
        assert T.crc32c(np.frombuffer(d, dtype=np.uint8)) == T.crc32c(d)  # memmap slices arrive as arrays
    for d in (bytes(70000), b"\xff" * 70000, bytes(range(256)) * 300):
        assert T.crc32c(d) == T._crc32c_scalar(d, 0xFFFFFFFF) ^ 0xFFFFFFFF


# ---- known answers for the primitives ---------------------------------------------------------------------------
def test_primitive_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283            # the standard CRC-32C check value
    assert T.crc32c(b"") == 0 and T.crc32c(bytes(32)) == 0x8A9136AA   # RFC 3720 B.4: 32 zero bytes
    assert T.read_varint(bytes([0xAC, 0x02]), 0) == (300, 2) and T.read_varint(b"\x00", 0) == (0, 1)
    assert T.read_varint(b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01", 0)[0] == 2 ** 64 - 1
    # snappy: length 16, literal "abcd", then an overlapping 12-byte copy from offset 4 (2-byte-offset element)
    assert T.snappy_uncompress(bytes([0x10, 0x0C]) + b"abcd" + bytes([0x2E, 0x04, 0x00])) == b"abcd" * 4
    # 1-byte-offset copy: length 7 from offset 3 after literal "xyz" -> "xyzxyzxyzx"
    assert T.snappy_uncompress(bytes([10, 0x08]) + b"xyz" + bytes([((7 - 4) << 2) | 1, 3])) == b"xyzxyzxyzx"
    with pytest.raises(T.CheckpointFormatError):
        T.snappy_uncompress(bytes([4, 0x2E, 0x09, 0x00]))  # copy reaching before the start of the output


@pytest.mark.parametrize("compress", [False, True])
def test_checkpoint_roundtrip(tmp_path, compress):
    g = torch.Generator().manual_seed(0)
    tensors = {
        "encoder/layer00/query_layer/kernel": torch.randn(8, 8, generator=g),
        "encoder/layer00/query_layer/kernel/adam_m": torch.randn(8, 8, generator=g).bfloat16(),
        "encoder/layer00/query_layer/bias": torch.randn(8, generator=g),
        "encoder/layer01/query_layer/kernel": torch.randn(8, 8, generator=g),
        "global_step": torch.tensor(460000, dtype=torch.int64),
        "vision_backbone/vision_transformer/conv2d/kernel": torch.randn(2, 2, 3, 8, generator=g),
        "word_embeddings/word_embeddings": torch.randn(11, 8, generator=g),
        "z/scalar_bf16": torch.tensor([1.5], dtype=torch.bfloat16),
    }
    prefix = str(tmp_path / "model.ckpt-1")
    _write_checkpoint(prefix, tensors, compress)
    assert T.list_variables(prefix) == sorted((n, list(t.shape)) for n, t in tensors.items())
    got = T.load_checkpoint(prefix)
    assert set(got) == {n for n in tensors if not n.endswith("adam_m") and n != "global_step"}   # model variables only
    for n, t in got.items():
        assert t.dtype == tensors[n].dtype and torch.equal(t, tensors[n]), n
    everything = T.load_checkpoint(prefix, names=list(tensors))
    assert int(everything["global_step"]) == 460000 and everything["encoder/layer00/query_layer/kernel/adam_m"].dtype == torch.bfloat16
    # corruption is detected: flip one byte of a tensor, then one byte of an index block
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(5)
        b = f.read(1)
        f.seek(5)
        f.write(bytes([b[0] ^ 0xFF]))
    with pytest.raises(T.CheckpointFormatError, match="checksum"):
        T.load_checkpoint(prefix)
    assert T.load_checkpoint(prefix, verify_checksums=False)  # still readable when asked not to verify
    with pytest.raises(FileNotFoundError):
        T.list_variables(str(tmp_path / "nope"))


def test_init_from_checkpoint_by_name(tmp_path, tiny_cfg):
    """The reference's rule (utils/model_utils.py:388-413): variables found in the checkpoint are restored, the rest keep their
    values.  A checkpoint without the temporal heads leaves exactly those entries untouched."""
    from merlot_b200.params import ParamStore
    from oracle import merlot_oracle as O
    params = O.init_params(tiny_cfg, seed=3, perturb=0.1)
    partial = {k: v for k, v in params.items() if "_temporal/" not in k}
    prefix = str(tmp_path / "model.ckpt-7")
    _write_checkpoint(prefix, partial)
    st = ParamStore(tiny_cfg, device="cpu")
    st.p.fill_(7.0)
    missing = st.load_checkpoint(prefix)
    assert missing and all("_temporal/" in m for m in missing)
    back = st.to_tf_dict("p")
    for k, v in partial.items():
        assert torch.equal(back[k], v), k
    assert torch.all(back["lang_viz_temporal/intermediate/kernel"] == 7.0)
