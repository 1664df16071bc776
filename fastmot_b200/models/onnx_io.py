"""Minimal ONNX file reader / writer (protobuf wire format, no `onnx` package — it is not in the image and the
importer must not depend on it).

The reference feeds ONNX files to TensorRT's parser (`fastmot/models/reid.py:47-63`, `trt.OnnxParser.parse`); this is
the replacement's front end: it decodes exactly the subset of `onnx.proto` an inference graph uses —

  ModelProto   : ir_version=1, producer_name=2, graph=7, opset_import=8 {domain=1, version=2}
  GraphProto   : node=1, name=2, initializer=5, input=11, output=12
  NodeProto    : input=1, output=2, name=3, op_type=4, attribute=5
  AttributeProto: name=1, f=2, i=3, s=4, t=5, floats=7, ints=8, type=20
  TensorProto  : dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9, double_data=10
  ValueInfoProto: name=1, type=2 {tensor_type=1 {elem_type=1, shape=2 {dim=1 {dim_value=1, dim_param=2}}}}

— into plain Python objects (`Graph`, `Node`, numpy initialisers), and encodes the same subset (used to write test
and export files).  Field numbers follow the public onnx.proto (IR version 3-9); unknown fields are skipped.
"""
import struct
from dataclasses import dataclass, field

import numpy as np

# TensorProto.DataType
FLOAT, UINT8, INT8, INT32, INT64, FLOAT16, DOUBLE = 1, 2, 3, 6, 7, 10, 11
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, INT32: np.int32, INT64: np.int64,
          FLOAT16: np.float16, DOUBLE: np.float64}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}
# AttributeProto.AttributeType
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS = 1, 2, 3, 4, 6, 7


# ------------------------------------------------------------------------------------------------ wire decoding
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _fields(buf):
    """Yields (field number, wire type, value) of one message; value is an int (varint / fixed) or a memoryview."""
    buf = memoryview(buf)
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if pos + n > end:
                raise ValueError("truncated length-delimited field")
            val = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _signed(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _packed_varints(wt, val):
    if wt == 0:
        return [_signed(val)]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(_signed(v))
    return out


def _packed_fixed(wt, val, fmt):
    size = struct.calcsize(fmt)
    if wt in (1, 5):
        return [struct.unpack("<" + fmt, val)[0]]
    return list(struct.unpack(f"<{len(val) // size}{fmt}", bytes(val)))


# ------------------------------------------------------------------------------------------------ object model
@dataclass
class Node:
    op_type: str
    inputs: list
    outputs: list
    name: str = ""
    attrs: dict = field(default_factory=dict)


@dataclass
class ValueInfo:
    name: str
    elem_type: int = FLOAT
    shape: tuple = ()          # ints, or strings for symbolic dimensions


@dataclass
class Graph:
    nodes: list
    initializers: dict          # name -> numpy array
    inputs: list                # ValueInfo, graph inputs that are not initialisers
    outputs: list
    name: str = "graph"
    opset: int = 11
    producer: str = ""


def _parse_tensor(buf):
    dims, dt, name, raw = [], FLOAT, "", None
    floats, i32, i64, f64 = [], [], [], []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            dims += _packed_varints(wt, val)
        elif fno == 2:
            dt = val
        elif fno == 4:
            floats += _packed_fixed(wt, val, "f")
        elif fno == 5:
            i32 += _packed_varints(wt, val)
        elif fno == 7:
            i64 += _packed_varints(wt, val)
        elif fno == 8:
            name = bytes(val).decode()
        elif fno == 9:
            raw = bytes(val)
        elif fno == 10:
            f64 += _packed_fixed(wt, val, "d")
    if dt not in _NP_OF:
        raise ValueError(f"tensor '{name}': unsupported ONNX data type {dt}")
    npdt = _NP_OF[dt]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(npdt).newbyteorder("<")).astype(npdt)
    elif dt == FLOAT:
        arr = np.asarray(floats, np.float32)
    elif dt == DOUBLE:
        arr = np.asarray(f64, np.float64)
    elif dt == INT64:
        arr = np.asarray(i64, np.int64)
    elif dt == FLOAT16:
        arr = np.asarray(i32, np.uint16).view(np.float16)
    else:
        arr = np.asarray(i32).astype(npdt)
    n = int(np.prod(dims)) if dims else 1
    if arr.size != n:
        raise ValueError(f"tensor '{name}': {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims)


def _parse_attr(buf):
    name, typ = "", None
    f = i = s = t = None
    floats, ints = [], []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            name = bytes(val).decode()
        elif fno == 2:
            f = struct.unpack("<f", val)[0]
        elif fno == 3:
            i = _signed(val)
        elif fno == 4:
            s = bytes(val)
        elif fno == 5:
            t = _parse_tensor(val)[1]
        elif fno == 7:
            floats += _packed_fixed(wt, val, "f")
        elif fno == 8:
            ints += _packed_varints(wt, val)
        elif fno == 20:
            typ = val
    if typ is None:     # IR < 3 files carry no type tag: first populated member wins
        typ = (A_INTS if ints else A_FLOATS if floats else A_TENSOR if t is not None else A_STRING if s is not None
               else A_FLOAT if f is not None else A_INT)
    value = {A_FLOAT: f, A_INT: i, A_STRING: s.decode() if s is not None else None, A_TENSOR: t,
             A_FLOATS: floats, A_INTS: ints}.get(typ)
    return name, value


def _parse_node(buf):
    n = Node("", [], [])
    for fno, wt, val in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(val).decode())
        elif fno == 2:
            n.outputs.append(bytes(val).decode())
        elif fno == 3:
            n.name = bytes(val).decode()
        elif fno == 4:
            n.op_type = bytes(val).decode()
        elif fno == 5:
            k, v = _parse_attr(val)
            n.attrs[k] = v
    return n


def _parse_value_info(buf):
    vi = ValueInfo("")
    for fno, wt, val in _fields(buf):
        if fno == 1:
            vi.name = bytes(val).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(val):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:
                        dims = []
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            d = "?"
                            for f5, w5, v5 in _fields(v4):
                                if f5 == 1:
                                    d = _signed(v5)
                                elif f5 == 2:
                                    d = bytes(v5).decode()
                            dims.append(d)
                        vi.shape = tuple(dims)
    return vi


def parse_model(data):
    """bytes of an .onnx file -> Graph."""
    graph_buf, opset, producer = None, 0, ""
    for fno, wt, val in _fields(data):
        if fno == 7:
            graph_buf = val
        elif fno == 2:
            producer = bytes(val).decode()
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(val):
                if f2 == 1:
                    dom = bytes(v2).decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
    if graph_buf is None:
        raise ValueError("not an ONNX ModelProto: no graph field")
    g = Graph([], {}, [], [], opset=opset, producer=producer)
    for fno, wt, val in _fields(graph_buf):
        if fno == 1:
            g.nodes.append(_parse_node(val))
        elif fno == 2:
            g.name = bytes(val).decode()
        elif fno == 5:
            name, arr = _parse_tensor(val)
            g.initializers[name] = arr
        elif fno == 11:
            g.inputs.append(_parse_value_info(val))
        elif fno == 12:
            g.outputs.append(_parse_value_info(val))
    g.inputs = [vi for vi in g.inputs if vi.name not in g.initializers]   # IR 3 lists weights as inputs too
    return g


def load(path):
    with open(path, "rb") as fh:
        return parse_model(fh.read())


# ------------------------------------------------------------------------------------------------ wire encoding
def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno, wt):
    return _enc_varint(fno << 3 | wt)


def _ld(fno, payload):
    return _key(fno, 2) + _enc_varint(len(payload)) + bytes(payload)


def _vi(fno, v):
    return _key(fno, 0) + _enc_varint(v)


def _enc_tensor(name, arr, raw=True):
    arr = np.asarray(arr)
    shape = arr.shape                       # ascontiguousarray would turn a 0-d scalar into shape (1,)
    arr = np.ascontiguousarray(arr)
    dt = _DT_OF[arr.dtype]
    out = b"".join(_vi(1, int(d)) for d in shape) + _vi(2, dt)
    if raw or dt != FLOAT:
        out += _ld(8, name.encode()) + _ld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    else:       # typed float_data (packed) — what some exporters emit for small tensors
        out += _ld(4, struct.pack(f"<{arr.size}f", *arr.ravel().tolist())) + _ld(8, name.encode())
    return out


def _enc_attr(name, v):
    out = _ld(1, name.encode())
    if isinstance(v, float):
        return out + _key(2, 5) + struct.pack("<f", v) + _vi(20, A_FLOAT)
    if isinstance(v, (bool, int, np.integer)):
        return out + _vi(3, int(v)) + _vi(20, A_INT)
    if isinstance(v, str):
        return out + _ld(4, v.encode()) + _vi(20, A_STRING)
    if isinstance(v, np.ndarray):
        return out + _ld(5, _enc_tensor("", v)) + _vi(20, A_TENSOR)
    v = list(v)
    if v and isinstance(v[0], float):
        return out + b"".join(_key(7, 5) + struct.pack("<f", x) for x in v) + _vi(20, A_FLOATS)
    return out + b"".join(_vi(8, int(x)) for x in v) + _vi(20, A_INTS)


def _enc_node(n):
    out = b"".join(_ld(1, s.encode()) for s in n.inputs) + b"".join(_ld(2, s.encode()) for s in n.outputs)
    out += _ld(3, n.name.encode()) + _ld(4, n.op_type.encode())
    return out + b"".join(_ld(5, _enc_attr(k, v)) for k, v in n.attrs.items())


def _enc_value_info(vi):
    dims = b""
    for d in vi.shape:
        dims += _ld(1, _ld(2, d.encode()) if isinstance(d, str) else _vi(1, int(d)))
    tensor = _vi(1, vi.elem_type) + _ld(2, dims)
    return _ld(1, vi.name.encode()) + _ld(2, _ld(1, tensor))


def serialize(g, ir_version=6, typed_float_data=False):
    """Graph -> bytes of an .onnx file (ModelProto)."""
    body = b"".join(_ld(1, _enc_node(n)) for n in g.nodes) + _ld(2, g.name.encode())
    body += b"".join(_ld(5, _enc_tensor(k, v, raw=not typed_float_data)) for k, v in g.initializers.items())
    body += b"".join(_ld(11, _enc_value_info(v)) for v in g.inputs)
    body += b"".join(_ld(12, _enc_value_info(v)) for v in g.outputs)
    model = _vi(1, ir_version) + _ld(2, (g.producer or "fastmot_b200").encode()) + _ld(7, body)
    model += _ld(8, _ld(1, b"") + _vi(2, g.opset))
    return model


def save(g, path, **kw):
    with open(path, "wb") as fh:
        fh.write(serialize(g, **kw))
