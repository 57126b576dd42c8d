"""Minimal ROS bag (format "#ROSBAG V2.0") reader / writer for the two message types the path consumes:
sensor_msgs/PointCloud2 (`/lidar`) and sensor_msgs/Imu (`/imu/*`), so that whole-bag reprocessing runs on real bags
without a ROS installation.

The reference reads bags through the `rosbag` Python package (vil_fusion/python/downsample_pointcloud.py:43-62,
make_prettier_graphs.py:411-474, carla_tools/src/fix_rosbag_time.py); none of ROS exists in this image, so the
format is restated from its public specification: a bag is a sequence of records
`<header_len u32><header><data_len u32><data>`, a header a sequence of `<field_len u32>name=value`; op 0x05 = chunk
(compression none / bz2 / lz4) holding connection (0x07) and message-data (0x02) records.  Index records are ignored
on reading (the chunks are walked in file order) and written minimally.  Host-side IO only: nothing here is on the
device path.
"""
from __future__ import annotations

import bz2
import struct
from typing import Iterator

import numpy as np

MAGIC = b"#ROSBAG V2.0\n"
OP_MSG, OP_BAG_HEADER, OP_INDEX, OP_CHUNK, OP_CHUNK_INFO, OP_CONNECTION = 0x02, 0x03, 0x04, 0x05, 0x06, 0x07

POINTCLOUD2_MD5 = "1158d486dd51d683ce2f1be655c3c181"
IMU_MD5 = "6a62c6daae103f4ff57a132d6f95cec2"


# ------------------------------------------------------------------------------------------------ records
def _parse_header(buf: bytes) -> dict:
    out, i = {}, 0
    while i < len(buf):
        (n,) = struct.unpack_from("<I", buf, i)
        field = buf[i + 4:i + 4 + n]
        k, _, v = field.partition(b"=")
        out[k.decode()] = v
        i += 4 + n
    return out


def _records(buf: bytes, pos: int = 0) -> Iterator[tuple[dict, bytes]]:
    n = len(buf)
    while pos + 4 <= n:
        (hl,) = struct.unpack_from("<I", buf, pos)
        hdr = _parse_header(buf[pos + 4:pos + 4 + hl])
        pos += 4 + hl
        (dl,) = struct.unpack_from("<I", buf, pos)
        data = buf[pos + 4:pos + 4 + dl]
        pos += 4 + dl
        yield hdr, data


# ---- LZ4 frame format (what rosbag's "lz4" chunks hold: roslz4 writes standard LZ4 frames).  The lz4 module is not in
# this image, so the codec is written out: a block decoder (token / literals / offset / match length sequences, overlapping
# matches copied byte-wise) under the frame layer (magic, FLG / BD, optional content size, block sizes with the
# "stored uncompressed" bit, optional block / content checksums -- skipped, not verified).
LZ4_MAGIC = 0x184D2204


def _lz4_block_decode(src: bytes, out: bytearray) -> None:
    i, n = 0, len(src)
    while i < n:
        token = src[i]
        i += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]
        i += lit
        if i >= n:
            break                                   # the last sequence has literals only
        offset = src[i] | (src[i + 1] << 8)
        i += 2
        if offset == 0:
            raise RuntimeError("corrupt lz4 block (zero offset)")
        ml = token & 15
        if ml == 15:
            while True:
                b = src[i]
                i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - offset
        if start < 0:
            raise RuntimeError("corrupt lz4 block (offset before start)")
        if offset >= ml:
            out += out[start:start + ml]
        else:                                       # overlapping match: the pattern repeats
            for k in range(ml):
                out.append(out[start + k])


def lz4_frame_decompress(data: bytes) -> bytes:
    pos, out = 0, bytearray()
    while pos < len(data):                          # concatenated frames are allowed
        (magic,) = struct.unpack_from("<I", data, pos)
        if magic != LZ4_MAGIC:
            raise RuntimeError("not an lz4 frame (magic %08x)" % magic)
        flg, pos = data[pos + 4], pos + 6           # FLG, BD
        if (flg >> 6) != 1:
            raise RuntimeError("unsupported lz4 frame version")
        block_checksum, content_size, content_checksum, dict_id = (flg >> 4) & 1, (flg >> 3) & 1, (flg >> 2) & 1, flg & 1
        pos += 8 * content_size + 4 * dict_id + 1   # optional content size, dictionary id; header checksum byte
        independent = (flg >> 5) & 1
        frame_start = len(out)
        while True:
            (bs,) = struct.unpack_from("<I", data, pos)
            pos += 4
            if bs == 0:
                break                               # end mark
            stored, bs = bs >> 31, bs & 0x7FFFFFFF
            block = data[pos:pos + bs]
            pos += bs + 4 * block_checksum
            if stored:
                out += block
            elif independent:
                part = bytearray()
                _lz4_block_decode(block, part)
                out += part
            else:                                   # linked blocks: matches may reach into the previous blocks of the frame
                _lz4_block_decode(block, out)
        pos += 4 * content_checksum
        del frame_start
    return bytes(out)


def _lz4_block_encode(src: bytes) -> bytes:
    """Greedy single-pass LZ4 block encoder (4-byte hash chain of depth 1): enough to write bags and to exercise every
    branch of the decoder; format rules kept: last 5 bytes are literals, no match starts within the last 12 bytes."""
    n, out, anchor, i, table = len(src), bytearray(), 0, 0, {}

    def emit(lit_end, mlen, offset):
        lit = lit_end - anchor
        token = (min(lit, 15) << 4) | (min(mlen - 4, 15) if mlen else 0)
        out.append(token)
        if lit >= 15:
            r = lit - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(src[anchor:lit_end])
        if mlen:
            out.extend(struct.pack("<H", offset))
            if mlen - 4 >= 15:
                r = mlen - 4 - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    while i + 12 < n:
        key = src[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 0xFFFF:
            m = 4
            limit = n - 5 - i
            while m < limit and src[cand + m] == src[i + m]:
                m += 1
            emit(i, m, i - cand)
            i += m
            anchor = i
        else:
            i += 1
    emit(n, 0, 0)
    return bytes(out)


def lz4_frame_compress(data: bytes, block_size: int = 1 << 16) -> bytes:
    """One LZ4 frame: version 1, independent blocks, no checksums, no content size (FLG 0x60, BD 64 KB)."""
    out = bytearray(struct.pack("<I", LZ4_MAGIC) + bytes([0x60, 0x40, 0x82]))       # 0x82 = (xxh32(FLG BD) >> 8) & 0xff for 60 40
    for a in range(0, len(data), block_size):
        raw = data[a:a + block_size]
        enc = _lz4_block_encode(raw)
        if len(enc) < len(raw):
            out += struct.pack("<I", len(enc)) + enc
        else:
            out += struct.pack("<I", len(raw) | 0x80000000) + raw
    out += struct.pack("<I", 0)
    return bytes(out)


def _decompress(kind: bytes, data: bytes, size: int) -> bytes:
    if kind == b"none":
        return data
    if kind == b"bz2":
        return bz2.decompress(data)
    if kind == b"lz4":
        out = lz4_frame_decompress(data)
        if size and len(out) != size:
            raise RuntimeError("lz4 chunk inflates to %d bytes, header says %d" % (len(out), size))
        return out
    raise RuntimeError("unknown chunk compression %r" % kind)


def read_messages(path: str, topics=None) -> Iterator[tuple[str, str, float, bytes]]:
    """Yields (topic, message type, record time [s], serialized message) in file order."""
    with open(path, "rb") as f:
        buf = f.read()
    if not buf.startswith(MAGIC):
        raise RuntimeError("%s is not a ROS bag v2.0" % path)
    conns: dict[int, tuple[str, str]] = {}

    def handle(hdr, data):
        op = hdr["op"][0]
        if op == OP_CONNECTION:
            cid = struct.unpack("<I", hdr["conn"])[0]
            ch = _parse_header(data)
            conns[cid] = (hdr["topic"].decode(), ch.get("type", b"").decode())
        elif op == OP_MSG:
            cid = struct.unpack("<I", hdr["conn"])[0]
            secs, nsecs = struct.unpack("<II", hdr["time"])
            topic, typ = conns.get(cid, ("", ""))
            if topics is None or topic in topics:
                return topic, typ, secs + nsecs * 1e-9, data
        return None

    for hdr, data in _records(buf, len(MAGIC)):
        op = hdr["op"][0]
        if op == OP_CHUNK:
            inner = _decompress(hdr["compression"], data, struct.unpack("<I", hdr["size"])[0])
            for h2, d2 in _records(inner):
                m = handle(h2, d2)
                if m:
                    yield m
        else:
            m = handle(hdr, data)
            if m:
                yield m


# ------------------------------------------------------------------------------------------------ message codecs
def _read_string(buf: bytes, i: int) -> tuple[str, int]:
    (n,) = struct.unpack_from("<I", buf, i)
    return buf[i + 4:i + 4 + n].decode(), i + 4 + n


def parse_pointcloud2(msg: bytes) -> dict:
    """sensor_msgs/PointCloud2 -> {stamp, frame_id, height, width, fields{name: (offset, datatype, count)}, point_step,
    row_step, is_bigendian, is_dense, data (uint8 view of the payload, zero copy)}."""
    seq, secs, nsecs = struct.unpack_from("<III", msg, 0)
    frame_id, i = _read_string(msg, 12)
    height, width, nf = struct.unpack_from("<III", msg, i)
    i += 12
    fields = {}
    for _ in range(nf):
        name, i = _read_string(msg, i)
        off, dt, cnt = struct.unpack_from("<IBI", msg, i)
        i += 9
        fields[name] = (off, dt, cnt)
    big, point_step, row_step, nd = struct.unpack_from("<BIII", msg, i)
    i += 13
    data = np.frombuffer(msg, np.uint8, nd, i)
    dense = msg[i + nd]
    return dict(seq=seq, stamp=secs + nsecs * 1e-9, frame_id=frame_id, height=height, width=width, fields=fields,
                point_step=point_step, row_step=row_step, is_bigendian=bool(big), is_dense=bool(dense), data=data)


def pointcloud2_for_upload(pc: dict) -> dict:
    """The dict `api.Handle.upload_pointcloud2` takes (FLOAT32 x / y / z required, datatype 7)."""
    for k in "xyz":
        if k not in pc["fields"] or pc["fields"][k][1] != 7:
            raise RuntimeError("PointCloud2 needs FLOAT32 fields x, y, z")
    if pc["is_bigendian"]:
        raise RuntimeError("big-endian PointCloud2 payloads are not supported")
    n = pc["width"] * pc["height"]
    return dict(data=pc["data"][:n * pc["point_step"]], point_step=pc["point_step"], fields={k: pc["fields"][k][0] for k in "xyz"})


def parse_imu(msg: bytes) -> dict:
    """sensor_msgs/Imu -> {stamp, gyro[3], accel[3]} (orientation and covariances are not used by the path,
    gtsam_fusion/src/gtsam_fusion/ImuManagerRos.cpp:38-52)."""
    seq, secs, nsecs = struct.unpack_from("<III", msg, 0)
    _, i = _read_string(msg, 12)
    vals = struct.unpack_from("<37d", msg, i)      # quat 4, cov 9, ang vel 3, cov 9, lin acc 3, cov 9
    return dict(stamp=secs + nsecs * 1e-9, gyro=np.array(vals[13:16]), accel=np.array(vals[25:28]))


def _string(s: str) -> bytes:
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def _stamp(t: float) -> tuple[int, int]:
    secs = int(np.floor(t))
    nsecs = int(round((t - secs) * 1e9))
    if nsecs >= 1000000000:
        secs, nsecs = secs + 1, nsecs - 1000000000
    return secs, nsecs


def make_pointcloud2(points: np.ndarray, stamp: float, frame_id: str = "lidar", field_names=("x", "y", "z", "intensity"), seq: int = 0) -> bytes:
    pts = np.ascontiguousarray(points, np.float32)
    n, k = pts.shape
    names = list(field_names)[:k] + ["f%d" % j for j in range(len(field_names), k)]
    secs, nsecs = _stamp(stamp)
    out = [struct.pack("<III", seq, secs, nsecs), _string(frame_id), struct.pack("<III", 1, n, k)]
    for j, nm in enumerate(names):
        out += [_string(nm), struct.pack("<IBI", 4 * j, 7, 1)]
    out += [struct.pack("<BIII", 0, 4 * k, 4 * k * n, 4 * k * n), pts.tobytes(), b"\x01"]
    return b"".join(out)


def make_imu(stamp: float, gyro, accel, frame_id: str = "imu", seq: int = 0) -> bytes:
    secs, nsecs = _stamp(stamp)
    vals = [0.0, 0.0, 0.0, 1.0] + [0.0] * 9 + [float(v) for v in gyro] + [0.0] * 9 + [float(v) for v in accel] + [0.0] * 9
    return struct.pack("<III", seq, secs, nsecs) + _string(frame_id) + struct.pack("<37d", *vals)


# ------------------------------------------------------------------------------------------------ writer
def _header(fields: dict) -> bytes:
    body = b"".join(struct.pack("<I", len(k) + 1 + len(v)) + k.encode() + b"=" + v for k, v in fields.items())
    return struct.pack("<I", len(body)) + body


def _record(fields: dict, data: bytes) -> bytes:
    return _header(fields) + struct.pack("<I", len(data)) + data


def write_bag(path: str, messages, compression: str = "none", chunk_messages: int = 64) -> None:
    """messages: iterable of (topic, type, md5sum, time [s], serialized bytes).  Writes chunks of `chunk_messages`
    messages (compression "none", "bz2" or "lz4") with their connection records, then connection + chunk-info records."""
    conns: dict[str, int] = {}
    conn_recs: dict[int, bytes] = {}
    chunks, cur, cur_times = [], [], []

    def flush():
        if not cur:
            return
        raw = b"".join(cur)
        comp = raw if compression == "none" else (bz2.compress(raw) if compression == "bz2" else lz4_frame_compress(raw))
        chunks.append((_record({"op": bytes([OP_CHUNK]), "compression": compression.encode(), "size": struct.pack("<I", len(raw))}, comp),
                       min(cur_times), max(cur_times), len(cur_times)))
        cur.clear()
        cur_times.clear()

    for topic, typ, md5, t, data in messages:
        if topic not in conns:
            cid = len(conns)
            conns[topic] = cid
            ch = _header({"topic": topic.encode(), "type": typ.encode(), "md5sum": md5.encode(), "message_definition": b""})[4:]
            conn_recs[cid] = _record({"op": bytes([OP_CONNECTION]), "conn": struct.pack("<I", cid), "topic": topic.encode()}, ch)
            cur.append(conn_recs[cid])
        secs, nsecs = _stamp(t)
        cur.append(_record({"op": bytes([OP_MSG]), "conn": struct.pack("<I", conns[topic]), "time": struct.pack("<II", secs, nsecs)}, data))
        cur_times.append((secs, nsecs))
        if len(cur_times) >= chunk_messages:
            flush()
    flush()
    body = b"".join(c[0] for c in chunks)
    index_pos = len(MAGIC) + 4096 + len(body)
    bag_hdr = _header({"op": bytes([OP_BAG_HEADER]), "index_pos": struct.pack("<Q", index_pos), "conn_count": struct.pack("<I", len(conns)),
                       "chunk_count": struct.pack("<I", len(chunks))})
    pad = 4096 - len(bag_hdr) - 4                       # the bag header record is padded to 4096 bytes
    tail = b"".join(conn_recs[c] for c in sorted(conn_recs))
    pos = len(MAGIC) + 4096
    for rec, t0, t1, cnt in chunks:
        tail += _record({"op": bytes([OP_CHUNK_INFO]), "ver": struct.pack("<I", 1), "chunk_pos": struct.pack("<Q", pos),
                         "start_time": struct.pack("<II", *t0), "end_time": struct.pack("<II", *t1), "count": struct.pack("<I", 0)}, b"")
        pos += len(rec)
    with open(path, "wb") as f:
        f.write(MAGIC + bag_hdr + struct.pack("<I", pad) + b" " * pad + body + tail)


# ------------------------------------------------------------------------------------------------ the path's view of a bag
def load_bag(path: str, cloud_topic: str = "/lidar", imu_topic: str | None = None):
    """Returns (clouds, imu): clouds = list of upload dicts (+ 'stamp') for api.Handle.upload_pointcloud2 in file
    order; imu = dict(t, accel (n,3), gyro (n,3)) or None."""
    clouds, it, ia, ig = [], [], [], []
    topics = {cloud_topic} | ({imu_topic} if imu_topic else set())
    for topic, typ, _, msg in read_messages(path, topics):
        if topic == cloud_topic:
            pc = parse_pointcloud2(msg)
            up = pointcloud2_for_upload(pc)
            up["stamp"] = pc["stamp"]
            clouds.append(up)
        elif imu_topic and topic == imu_topic:
            m = parse_imu(msg)
            it.append(m["stamp"])
            ia.append(m["accel"])
            ig.append(m["gyro"])
    imu = dict(t=np.array(it), accel=np.array(ia).reshape(-1, 3), gyro=np.array(ig).reshape(-1, 3)) if imu_topic else None
    return clouds, imu
