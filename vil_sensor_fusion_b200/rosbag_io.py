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


def _decompress(kind: bytes, data: bytes, size: int) -> bytes:
    if kind == b"none":
        return data
    if kind == b"bz2":
        return bz2.decompress(data)
    if kind == b"lz4":
        try:
            import lz4.frame                                   # optional
        except ImportError as e:
            raise RuntimeError("this bag uses lz4 chunks and the lz4 module is not installed") from e
        return lz4.frame.decompress(data)
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
    messages (compression "none" or "bz2") with their connection records, then connection + chunk-info records."""
    conns: dict[str, int] = {}
    conn_recs: dict[int, bytes] = {}
    chunks, cur, cur_times = [], [], []

    def flush():
        if not cur:
            return
        raw = b"".join(cur)
        comp = raw if compression == "none" else bz2.compress(raw)
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
