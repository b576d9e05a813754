"""Minimal FlatBuffers / FlexBuffers codec (the `flatbuffers` Python module is not installed
in this image, and the model files of the reference are FlatBuffers: `models/*.tflite`,
track.py:68,93).  Only what `.tflite` needs: tables, scalars, strings, vectors of scalars /
tables, and FlexBuffer maps of scalars (the custom-op options of TFLite_Detection_PostProcess).

Wire format restated from the public FlatBuffers internals documentation [3P-MEM]:
  * file = uoffset32 to the root table, optional 4-byte identifier at byte 4;
  * table = soffset32 (table_pos - vtable_pos) followed by inline fields; vtable = u16 vtable
    bytes, u16 table bytes, then one u16 per field id: offset of the field inside the table,
    0 = absent (default value);
  * references (sub-table, vector, string) are uoffset32 RELATIVE TO THE FIELD's own position
    and always point forward; vector = u32 length + elements; string = vector of bytes + NUL.
"""
from __future__ import annotations

import struct

import numpy as np

_SCALAR = {'bool': '<?', 'i8': '<b', 'u8': '<B', 'i16': '<h', 'u16': '<H', 'i32': '<i', 'u32': '<I',
           'i64': '<q', 'u64': '<Q', 'f32': '<f', 'f64': '<d'}
_NP = {'i8': np.int8, 'u8': np.uint8, 'i16': np.int16, 'i32': np.int32, 'u32': np.uint32,
       'i64': np.int64, 'f32': np.float32, 'f64': np.float64}


# ---------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------

class Table:
    """A table inside `buf` at absolute position `pos`."""

    def __init__(self, buf, pos):
        self.buf, self.pos = buf, pos
        self.vt = pos - struct.unpack_from('<i', buf, pos)[0]
        self.vt_len = struct.unpack_from('<H', buf, self.vt)[0]

    def _field(self, fid):
        o = 4 + 2 * fid
        if o + 2 > self.vt_len:
            return 0
        off = struct.unpack_from('<H', self.buf, self.vt + o)[0]
        return self.pos + off if off else 0

    def has(self, fid):
        return self._field(fid) != 0

    def scalar(self, fid, kind, default=0):
        p = self._field(fid)
        return struct.unpack_from(_SCALAR[kind], self.buf, p)[0] if p else default

    def _indirect(self, fid):
        p = self._field(fid)
        return p + struct.unpack_from('<I', self.buf, p)[0] if p else 0

    def table(self, fid):
        p = self._indirect(fid)
        return Table(self.buf, p) if p else None

    def string(self, fid, default=''):
        p = self._indirect(fid)
        if not p:
            return default
        n = struct.unpack_from('<I', self.buf, p)[0]
        return bytes(self.buf[p + 4:p + 4 + n]).decode('utf-8')

    def vector(self, fid, kind):
        """numpy array of scalars (a copy-free view of the buffer); empty when absent."""
        p = self._indirect(fid)
        if not p:
            return np.zeros(0, _NP[kind])
        n = struct.unpack_from('<I', self.buf, p)[0]
        return np.frombuffer(self.buf, dtype=np.dtype(_NP[kind]).newbyteorder('<'), count=n, offset=p + 4)

    def tables(self, fid):
        p = self._indirect(fid)
        if not p:
            return []
        n = struct.unpack_from('<I', self.buf, p)[0]
        out = []
        for i in range(n):
            e = p + 4 + 4 * i
            out.append(Table(self.buf, e + struct.unpack_from('<I', self.buf, e)[0]))
        return out


def root(buf, identifier=None):
    if len(buf) < 8:
        raise ValueError('not a FlatBuffer: file shorter than 8 bytes')
    if identifier is not None and bytes(buf[4:8]) != identifier:
        raise ValueError(f'FlatBuffer identifier {bytes(buf[4:8])!r} != {identifier!r}')
    return Table(buf, struct.unpack_from('<I', buf, 0)[0])


# ---------------------------------------------------------------------------------------------
# builder: objects are described as Python values and laid out front to back (parents before
# children, so every uoffset points forward)
# ---------------------------------------------------------------------------------------------

class T:
    """Table to build: fields = {field id: value}; value kinds:
       ('i32', 5) scalar | T(...) sub-table | 'text' string | V('i32', [..]) scalar vector |
       [T(...), ...] vector of tables | U(type_field_value, T(...)) handled by the caller as two fields."""

    def __init__(self, **fields):
        self.fields = {int(k[1:]): v for k, v in fields.items() if v is not None}


class V:
    def __init__(self, kind, values):
        self.kind = kind
        self.data = np.ascontiguousarray(np.asarray(values, dtype=_NP[kind])).tobytes()
        self.n = len(self.data) // np.dtype(_NP[kind]).itemsize
        self.align = max(4, np.dtype(_NP[kind]).itemsize)


def build(root_table, identifier=b'\0\0\0\0'):
    out = bytearray(8)
    out[4:8] = identifier
    pending = [(0, root_table)]                      # (position of the uoffset to patch, object)

    def align(n):
        while len(out) % n:
            out.append(0)

    while pending:
        patch, obj = pending.pop(0)
        if isinstance(obj, T):
            ids = sorted(obj.fields)
            nf = (ids[-1] + 1) if ids else 0
            # inline layout: 4-byte soffset, then fields by decreasing size
            sized = []
            for fid in ids:
                v = obj.fields[fid]
                if isinstance(v, tuple):
                    sized.append((struct.calcsize(_SCALAR[v[0]]), fid))
                else:
                    sized.append((4, fid))
            sized.sort(key=lambda t: (-t[0], t[1]))
            vt_bytes = 4 + 2 * nf
            align(2)
            # choose the table position so that it is 8-aligned (covers every field alignment)
            while (len(out) + vt_bytes) % 8:
                out.append(0)
            vt_pos = len(out)
            tpos = vt_pos + vt_bytes
            offs, cur = {}, 4
            for sz, fid in sized:
                cur = (cur + sz - 1) // sz * sz
                offs[fid] = cur
                cur += sz
            tbytes = cur
            out.extend(struct.pack('<HH', vt_bytes, tbytes))
            for fid in range(nf):
                out.extend(struct.pack('<H', offs.get(fid, 0)))
            assert len(out) == tpos
            out.extend(b'\0' * tbytes)
            struct.pack_into('<i', out, tpos, tpos - vt_pos)
            for fid in ids:
                v = obj.fields[fid]
                p = tpos + offs[fid]
                if isinstance(v, tuple):
                    struct.pack_into(_SCALAR[v[0]], out, p, v[1])
                else:
                    pending.append((p, v))
            target = tpos
        elif isinstance(obj, str):
            b = obj.encode('utf-8')
            align(4)
            target = len(out)
            out.extend(struct.pack('<I', len(b)) + b + b'\0')
        elif isinstance(obj, V):
            align(4)
            while (len(out) + 4) % obj.align:
                out.append(0)
            target = len(out)
            out.extend(struct.pack('<I', obj.n) + obj.data)
        elif isinstance(obj, list):
            align(4)
            target = len(out)
            out.extend(struct.pack('<I', len(obj)))
            base = len(out)
            out.extend(b'\0' * (4 * len(obj)))
            for i, t in enumerate(obj):
                pending.append((base + 4 * i, t))
        else:
            raise TypeError(type(obj))
        struct.pack_into('<I', out, patch, target - patch)
    return bytes(out)


# ---------------------------------------------------------------------------------------------
# FlexBuffers: maps of scalars
# ---------------------------------------------------------------------------------------------

FBT_INT, FBT_UINT, FBT_FLOAT, FBT_KEY, FBT_STRING = 1, 2, 3, 4, 5
FBT_INDIRECT_INT, FBT_INDIRECT_UINT, FBT_INDIRECT_FLOAT, FBT_MAP, FBT_BOOL = 6, 7, 8, 9, 26


def _flex_read(buf, pos, width, signed=False, flt=False):
    if flt:
        return struct.unpack_from({4: '<f', 8: '<d'}[width], buf, pos)[0]
    fmt = {1: 'b', 2: 'h', 4: 'i', 8: 'q'}[width]
    return struct.unpack_from('<' + (fmt if signed else fmt.upper()), buf, pos)[0]


def flex_map(buf):
    """Decode a FlexBuffer whose root is a map of scalar values -> dict."""
    buf = bytes(buf)
    if len(buf) < 3:
        return {}
    root_w = buf[-1]
    root_t = buf[-2]
    if root_t >> 2 != FBT_MAP:
        raise ValueError('FlexBuffer root is not a map')
    bw = 1 << (root_t & 3)
    rpos = len(buf) - 2 - root_w
    loc = rpos - _flex_read(buf, rpos, root_w)
    keys_pos = loc - 3 * bw
    keys_loc = keys_pos - _flex_read(buf, keys_pos, bw)
    kw = _flex_read(buf, loc - 2 * bw, bw)
    n = _flex_read(buf, loc - bw, bw)
    out = {}
    for i in range(n):
        kp = keys_loc + i * kw
        ks = kp - _flex_read(buf, kp, kw)
        key = buf[ks:buf.index(b'\0', ks)].decode()
        t = buf[loc + n * bw + i]
        ty, w = t >> 2, 1 << (t & 3)
        vp = loc + i * bw
        if ty == FBT_INT:
            val = _flex_read(buf, vp, bw, signed=True)
        elif ty == FBT_UINT:
            val = _flex_read(buf, vp, bw)
        elif ty == FBT_BOOL:
            val = bool(_flex_read(buf, vp, bw))
        elif ty == FBT_FLOAT:
            val = _flex_read(buf, vp, bw, flt=True)
        elif ty in (FBT_INDIRECT_INT, FBT_INDIRECT_UINT, FBT_INDIRECT_FLOAT):
            ip = vp - _flex_read(buf, vp, bw)
            val = _flex_read(buf, ip, w, signed=(ty == FBT_INDIRECT_INT), flt=(ty == FBT_INDIRECT_FLOAT))
        else:
            continue
        out[key] = val
    return out


def flex_build_map(d):
    """dict of str -> int | float | bool  ->  FlexBuffer bytes (4-byte slots)."""
    keys = sorted(d)
    out = bytearray()
    kpos = []
    for k in keys:
        kpos.append(len(out))
        out.extend(k.encode() + b'\0')
    while len(out) % 4:
        out.append(0)
    out.extend(struct.pack('<I', len(keys)))                   # keys vector: length, then offsets
    keys_loc = len(out)
    for p in kpos:
        out.extend(struct.pack('<I', len(out) - p))
    out.extend(struct.pack('<I', len(out) - keys_loc))         # map prefix: keys vector, its width, length
    out.extend(struct.pack('<I', 4))
    out.extend(struct.pack('<I', len(keys)))
    loc = len(out)
    types = bytearray()
    for k in keys:
        v = d[k]
        if isinstance(v, bool):
            out.extend(struct.pack('<I', int(v))); types.append((FBT_BOOL << 2) | 2)
        elif isinstance(v, int):
            out.extend(struct.pack('<i', v)); types.append((FBT_INT << 2) | 2)
        else:
            out.extend(struct.pack('<f', float(v))); types.append((FBT_FLOAT << 2) | 2)
    out.extend(types)
    while len(out) % 4:                                        # root offset is read with its own width
        out.append(0)
    out.extend(struct.pack('<I', len(out) - loc))
    out.append((FBT_MAP << 2) | 2)
    out.append(4)
    return bytes(out)
