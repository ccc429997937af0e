"""Packet sources and scan batching: the step before the odometry path for recorded data.

Mirrors /root/reference/src/ptudes/data.py:13-77 (`OusterLidarData.withScanIdx`), the packet sources it is
fed with - `pcap.Pcap(file, meta)` and `OusterRawBagSource` (/root/reference/src/ptudes/utils.py:171-187,
/root/reference/src/ptudes/bag.py:21-97) - and the three ouster-sdk pieces they use (`PacketFormat.from_info`,
`ScanBatcher`, `LidarPacket` / `ImuPacket`).  ouster-sdk and rosbags are not installable here, so the packet
layouts, the pcap reader and the ROS-bag (format 2.0) reader are restated from the published formats
[UPSTREAM-UNVERIFIED]; the per-pixel work (channel data -> staggered field images) runs on the GPU
(`ptk_batcher_decode`), the fields stay in HBM and `KissICPWrapper.register_frame` consumes the RANGE image
there.  No CPU fallback: decoding needs `libptk.so` and a CUDA device.
"""
import bz2
import ctypes as C
import glob
import struct
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

from . import _ffi
from ._ffi import PtkPacketFormat, PtkScanFields
from .ins.data import GRAV, IMU

PROFILE_LIDAR_LEGACY = 1
PROFILE_LIDAR_RNG19_RFL8_SIG16_NIR16_DUAL = 2
PROFILE_LIDAR_RNG19_RFL8_SIG16_NIR16 = 3
PROFILE_LIDAR_RNG15_RFL8_NIR8 = 4
PROFILES = {"LEGACY": 1, "RNG19_RFL8_SIG16_NIR16_DUAL": 2, "RNG19_RFL8_SIG16_NIR16": 3, "RNG15_RFL8_NIR8": 4}
IMU_PACKET_SIZE = 48
LIDAR_PORT, IMU_PORT = 7502, 7503

# Ouster ROS PacketMsg MD5 sum (bag.py:19)
OUSTER_PACKETMSG_MD5 = "4f7b5949e76f86d01e96b0e33ba9b5e3"


def _check(rc):
    if rc < 0:
        raise _ffi.PtkError(rc, (_ffi.load().ptk_ingest_last_error() or b"").decode())
    return rc


class PacketFormat:
    """`_client.PacketFormat.from_info(metadata)` (data.py:48): byte layout of the sensor's lidar packets."""

    def __init__(self, profile, pixels_per_column, columns_per_packet, columns_per_frame):
        if isinstance(profile, str):
            profile = PROFILES[profile.replace("PROFILE_LIDAR_", "")]
        self.c = PtkPacketFormat()
        _check(_ffi.load().ptk_packet_format_init(C.byref(self.c), int(profile), int(pixels_per_column),
                                                  int(columns_per_packet), int(columns_per_frame)))

    @classmethod
    def from_info(cls, metadata) -> "PacketFormat":
        f = metadata.format
        profile = getattr(f, "udp_profile_lidar", PROFILE_LIDAR_LEGACY)
        profile = getattr(profile, "value", profile)
        return cls(profile, f.pixels_per_column, getattr(f, "columns_per_packet", 16), f.columns_per_frame)

    def __getattr__(self, name):
        return getattr(self.c, name)

    def frame_id(self, buf) -> int:
        b = np.frombuffer(buf, dtype=np.uint8)
        return _ffi.load().ptk_packet_frame_id(C.byref(self.c), b.ctypes.data)


class LidarPacket:
    """`client.LidarPacket(buf, metadata, ts)` as far as data.py uses it."""

    def __init__(self, buf, metadata=None, capture_timestamp: float = 0.0):
        self.buf = buf
        self.capture_timestamp = capture_timestamp


class ImuPacket:
    """`client.ImuPacket`: 48 bytes - three u64 timestamps (ns), linear acceleration (g), angular velocity (deg/s)."""

    def __init__(self, buf, metadata=None, capture_timestamp: float = 0.0):
        self.buf = bytes(buf[:IMU_PACKET_SIZE])
        self.capture_timestamp = capture_timestamp
        self.sys_ts, self.accel_ts, self.gyro_ts = struct.unpack_from("<3Q", self.buf, 0)
        v = struct.unpack_from("<6f", self.buf, 24)
        self.accel = np.array(v[:3], dtype=np.float64)
        self.angular_vel = np.array(v[3:], dtype=np.float64)


def imu_from_packet(p: ImuPacket, dt: float = 0.01, _intr_rot=None) -> IMU:
    """`IMU.from_packet` (/root/reference/src/ptudes/ins/data.py:19-31)."""
    imu = IMU()
    imu.ts = p.sys_ts / 10**9
    imu.lacc = GRAV * p.accel
    imu.avel = np.pi * p.angular_vel / 180.0
    if _intr_rot is not None:
        imu.lacc = _intr_rot @ imu.lacc
        imu.avel = _intr_rot @ imu.avel
    imu.dt = dt
    return imu


class DeviceLidarScan:
    """`client.LidarScan(h, w, fields, columns_per_packet)` whose field images live in HBM.

    `field(ChanField.RANGE)` is a (H, W) torch uint32 tensor on the device (what `Odometry.register_scan` takes by
    address); the per-column headers `timestamp`, `status`, `measurement_id` are host arrays, fetched once per scan
    (they are what `client.last_valid_column_ts` reads, kiss.py:56)."""

    FIELDS = {"RANGE": "uint32", "RANGE2": "uint32", "REFLECTIVITY": "uint16", "SIGNAL": "uint16", "NEAR_IR": "uint16"}

    def __init__(self, h, w, fields=("RANGE",), device=0):
        import torch
        self.h, self.w = h, w
        self.frame_id = -1
        self.n_packets = 0
        dev = torch.device("cuda", device)
        self._f = {}
        for name in fields:
            name = getattr(name, "name", name)
            if name in self.FIELDS:
                self._f[name] = torch.empty((h, w), dtype=getattr(torch, self.FIELDS[name]), device=dev)
        if "RANGE" not in self._f:
            self._f["RANGE"] = torch.empty((h, w), dtype=torch.uint32, device=dev)
        self._d_ts = torch.empty(w, dtype=torch.uint64, device=dev)
        self._d_status = torch.empty(w, dtype=torch.uint32, device=dev)
        self._d_mid = torch.empty(w, dtype=torch.uint16, device=dev)
        self._host = None

    def _fields_struct(self) -> PtkScanFields:
        s = PtkScanFields()
        s.range = self._f["RANGE"].data_ptr()
        for cname, fname in (("range2", "RANGE2"), ("reflectivity", "REFLECTIVITY"), ("signal", "SIGNAL"), ("near_ir", "NEAR_IR")):
            if fname in self._f:
                setattr(s, cname, self._f[fname].data_ptr())
        s.timestamp = self._d_ts.data_ptr()
        s.status = self._d_status.data_ptr()
        s.measurement_id = self._d_mid.data_ptr()
        return s

    def field(self, f):
        return self._f[getattr(f, "name", f)]

    def _headers(self):
        if self._host is None:
            self._host = (self._d_ts.cpu().numpy().view(np.uint64), self._d_status.cpu().numpy().view(np.uint32),
                          self._d_mid.cpu().numpy().view(np.uint16))
        return self._host

    @property
    def timestamp(self):
        return self._headers()[0]

    @property
    def status(self):
        return self._headers()[1]

    @property
    def measurement_id(self):
        return self._headers()[2]


class ScanBatcher:
    """`_client.ScanBatcher(w, pf)` (data.py:49): `batch(packet, ls)` returns True when `ls` holds a finished
    frame.  The packets of a frame are only grouped on the host; the frame is decoded on the device, into the
    field images of `ls`, when it is complete."""

    def __init__(self, w: int, pf: PacketFormat, device: int = 0, frames: int = 4, stream: int = 0):
        assert w == pf.columns_per_frame
        self.pf = pf
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        _check(self._lib.ptk_batcher_create(C.byref(self._h), device, C.byref(pf.c), frames))
        self.device = device
        self.stream = stream

    def close(self):
        if self._h:
            self._lib.ptk_batcher_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _decode_into(self, ls: DeviceLidarScan):
        fid, npk = C.c_int(), C.c_int()
        fs = ls._fields_struct()
        _check(self._lib.ptk_batcher_decode(self._h, C.byref(fs), C.byref(fid), C.byref(npk), self.stream))
        ls.frame_id, ls.n_packets, ls._host = fid.value, npk.value, None

    def __call__(self, packet, ls: DeviceLidarScan) -> bool:
        buf = packet.buf if hasattr(packet, "buf") else packet
        b = np.frombuffer(buf, dtype=np.uint8)
        if b.size != self.pf.lidar_packet_size:
            raise ValueError(f"lidar packet of {b.size} bytes, the format says {self.pf.lidar_packet_size}")
        ready = C.c_int()
        _check(self._lib.ptk_batcher_push(self._h, b.ctypes.data, C.byref(ready)))
        if ready.value:
            self._decode_into(ls)
            return True
        return False

    def flush(self, ls: DeviceLidarScan) -> bool:
        """End of the stream: the partial frame, if any (data.py:52-56 yields `ls_write` as it is)."""
        ready = C.c_int()
        _check(self._lib.ptk_batcher_flush(self._h, C.byref(ready)))
        if ready.value:
            self._decode_into(ls)
            return True
        return False


def decode_frames(pf: PacketFormat, packets, n_frames: int, device: int = 0, fields=("RANGE",), stream: int = 0, out=None):
    """Decode `n_frames` frames of packet slots (torch uint8 tensor on the device, or host bytes/ndarray) in ONE
    launch; returns {field name: (n_frames, H, W) device tensor} + the column headers (n_frames, W).  `out`: the
    dict of an earlier call with the same shapes, to decode into again."""
    import torch
    dev = torch.device("cuda", device)
    H, W = pf.pixels_per_column, pf.columns_per_frame
    if out is None:
        out = {"RANGE": torch.empty((n_frames, H, W), dtype=torch.uint32, device=dev)}
        for name in fields:
            if name != "RANGE":
                out[name] = torch.empty((n_frames, H, W), dtype=getattr(torch, DeviceLidarScan.FIELDS[name]), device=dev)
        out["timestamp"] = torch.empty((n_frames, W), dtype=torch.uint64, device=dev)
        out["status"] = torch.empty((n_frames, W), dtype=torch.uint32, device=dev)
        out["measurement_id"] = torch.empty((n_frames, W), dtype=torch.uint16, device=dev)
    s = PtkScanFields()
    s.range = out["RANGE"].data_ptr()
    for cname, fname in (("range2", "RANGE2"), ("reflectivity", "REFLECTIVITY"), ("signal", "SIGNAL"), ("near_ir", "NEAR_IR")):
        if fname in out:
            setattr(s, cname, out[fname].data_ptr())
    s.timestamp, s.status, s.measurement_id = out["timestamp"].data_ptr(), out["status"].data_ptr(), out["measurement_id"].data_ptr()
    if hasattr(packets, "data_ptr"):
        ptr = packets.data_ptr()
    else:
        packets = np.ascontiguousarray(np.frombuffer(packets, dtype=np.uint8))
        ptr = packets.ctypes.data
    _check(_ffi.load().ptk_decode_packets(C.byref(pf.c), device, ptr, n_frames, C.byref(s), stream))
    return out


# ---- packet sources ------------------------------------------------------------------------------------
class PcapSource:
    """`pcap.Pcap(file_path, meta)` (utils.py:179): lidar and IMU packets of a pcap file in capture order.
    Native reader (`ptk_pcap_*`): classic pcap and pcapng, Ethernet / VLAN / cooked / raw-IP link types, IPv4
    reassembly."""

    def __init__(self, file_path: str, metadata, lidar_port: Optional[int] = None, imu_port: Optional[int] = None):
        self._path = str(file_path)
        self._metadata = metadata
        self._pf = PacketFormat.from_info(metadata)
        self.lidar_port = lidar_port if lidar_port is not None else getattr(metadata, "udp_port_lidar", LIDAR_PORT)
        self.imu_port = imu_port if imu_port is not None else getattr(metadata, "udp_port_imu", IMU_PORT)
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        _check(self._lib.ptk_pcap_open(C.byref(self._h), self._path.encode()))

    def __iter__(self) -> Iterator[Union[LidarPacket, ImuPacket]]:
        cap = max(self._pf.lidar_packet_size, 65536)
        buf = np.empty(cap, dtype=np.uint8)
        n, port, ts = C.c_int(), C.c_int(), C.c_double()
        while True:
            rc = _check(self._lib.ptk_pcap_next(self._h, buf.ctypes.data, cap, C.byref(n), C.byref(port), C.byref(ts)))
            if rc == 0:
                return
            # classify by size first (ports differ between recordings), then by port
            if n.value == self._pf.lidar_packet_size and port.value != self.imu_port:
                yield LidarPacket(buf[:n.value].tobytes(), self._metadata, ts.value)
            elif n.value == IMU_PACKET_SIZE and port.value != self.lidar_port:
                yield ImuPacket(buf[:n.value].tobytes(), self._metadata, ts.value)

    @property
    def metadata(self):
        return self._metadata

    def close(self) -> None:
        if self._h:
            self._lib.ptk_pcap_close(self._h)
            self._h = C.c_void_p()

    __del__ = close


def _bag_fields(hdr: bytes) -> Dict[str, bytes]:
    out, i = {}, 0
    while i + 4 <= len(hdr):
        n = struct.unpack_from("<I", hdr, i)[0]
        kv = hdr[i + 4:i + 4 + n]
        k, _, v = kv.partition(b"=")
        out[k.decode()] = v
        i += 4 + n
    return out


def _bag_records(buf, pos: int, end: int):
    """(header fields, data view, next position) of the records in buf[pos:end] (ROS bag format 2.0)."""
    while pos + 8 <= end:
        hl = struct.unpack_from("<I", buf, pos)[0]
        hdr = bytes(buf[pos + 4:pos + 4 + hl])
        dl = struct.unpack_from("<I", buf, pos + 4 + hl)[0]
        d0 = pos + 8 + hl
        yield _bag_fields(hdr), buf[d0:d0 + dl]
        pos = d0 + dl


def lz4_frame_decompress(data, size: int) -> bytes:
    """An LZ4 frame (what `rosbag record --lz4` compresses chunks with) -> `size` bytes; native (`ptk_lz4_frame_decompress`)."""
    src = np.frombuffer(data, dtype=np.uint8)
    dst = np.empty(max(int(size), 1), dtype=np.uint8)
    n = C.c_ulonglong()
    _check(_ffi.load().ptk_lz4_frame_decompress(src.ctypes.data, src.size, dst.ctypes.data, dst.size, C.byref(n)))
    if n.value != size:
        raise ValueError(f"lz4 chunk: {n.value} bytes after decompression, the chunk header says {size}")
    return dst[:n.value].tobytes()


class BagReader:
    """Messages of ROS1 bag file(s) (format 2.0) - the part of `rosbags.highlevel.AnyReader` bag.py uses.
    Chunks with `none`, `bz2` or `lz4` compression; messages come chunk by chunk, in time order within a chunk (a
    recorder writes its chunks in arrival order).  `connections`: {id: (topic, msgtype, md5sum)} seen so far."""

    def __init__(self, paths):
        self.paths = [Path(p) for p in paths] if isinstance(paths, (list, tuple)) else [Path(paths)]
        self.connections: Dict[Tuple[int, int], Tuple[str, str, str]] = {}

    def _scan(self, fi, path, buf, pos, end, want):
        msgs = []
        for h, d in _bag_records(buf, pos, end):
            op = h["op"][0]
            if op == 0x07:                                  # connection
                cid = struct.unpack("<I", h["conn"])[0]
                ch = _bag_fields(bytes(d))
                topic = h.get("topic", ch.get("topic", b"")).decode()
                self.connections[(fi, cid)] = (topic, ch.get("type", b"").decode(), ch.get("md5sum", b"").decode())
            elif op == 0x02:                                # message data
                cid = struct.unpack("<I", h["conn"])[0]
                sec, nsec = struct.unpack("<II", h["time"])
                msgs.append((sec * 10**9 + nsec, cid, d))
            elif op == 0x05:                                # chunk
                comp = h["compression"].decode()
                if comp == "none":
                    inner = d
                elif comp == "bz2":
                    inner = memoryview(bz2.decompress(bytes(d)))
                elif comp == "lz4":
                    inner = memoryview(lz4_frame_decompress(d, struct.unpack("<I", h["size"])[0]))
                else:
                    raise ValueError(f"{path}: chunk compression '{comp}' is not supported (none, bz2, lz4)")
                yield from self._scan(fi, path, inner, 0, len(inner), want)
        msgs.sort(key=lambda m: m[0])
        for t, cid, d in msgs:
            conn = self.connections.get((fi, cid), ("", "", ""))
            if want(conn):
                yield conn, t, d

    def messages(self, want):
        """(connection, timestamp ns, serialized message) of every message whose connection passes `want`."""
        for fi, path in enumerate(self.paths):
            data = memoryview(np.memmap(path, dtype=np.uint8, mode="r"))
            if bytes(data[:13]) != b"#ROSBAG V2.0\n":
                raise ValueError(f"{path}: not a ROS bag (format 2.0)")
            yield from self._scan(fi, path, data, 13, len(data), want)

    def scan_connections(self):
        """Read every connection record (a pass over the file: bag.py looks at `connections` before iterating)."""
        for _ in self.messages(lambda c: False):
            pass
        return list(self.connections.values())


def _packetmsg_buf(d) -> bytes:
    n = struct.unpack_from("<I", d, 0)[0]                   # ouster_ros/PacketMsg: uint8[] buf
    return bytes(d[4:4 + n])


class OusterRawBagSource:
    """`OusterRawBagSource(data_path, info)` (bag.py:21-97): Ouster raw packets out of ROS1 bag(s): the
    `ouster_ros/PacketMsg` messages (md5 4f7b...b5e3, `uint8[] buf`) of the topics ending in `lidar_packets` /
    `imu_packets` (or the two named topics).  Own reader of the bag format (rosbags is absent)."""

    def __init__(self, data_path, info, *, rate: float = 0.0, lidar_topic: str = "", imu_topic: str = ""):
        self._reader = BagReader(data_path)
        self._metadata = info
        self._rate = rate
        self._lidar_topic, self._imu_topic = lidar_topic, imu_topic
        self._topics: List[str] = []

    def _wanted(self, topic: str) -> Optional[str]:
        if not self._lidar_topic and not self._imu_topic:
            if topic.endswith("lidar_packets"):
                return "lidar"
            if topic.endswith("imu_packets"):
                return "imu"
            return None
        if topic == self._lidar_topic:
            return "lidar"
        if topic == self._imu_topic:
            return "imu"
        return None

    def __iter__(self) -> Iterator[Union[LidarPacket, ImuPacket]]:
        import time
        real_start, bag_start = time.monotonic(), None
        for (topic, _, md5), t, d in self._reader.messages(lambda c: self._wanted(c[0]) is not None):
            if topic not in self._topics:
                self._topics.append(topic)
            ts = t / 10**9
            if self._rate:
                bag_start = ts if bag_start is None else bag_start
                time.sleep(max(0.0, (ts - bag_start) / self._rate - (time.monotonic() - real_start)))
            if md5 != OUSTER_PACKETMSG_MD5:
                continue
            if self._wanted(topic) == "lidar":
                yield LidarPacket(_packetmsg_buf(d), self._metadata, ts)
            else:
                yield ImuPacket(_packetmsg_buf(d), self._metadata, ts)

    @property
    def topics(self) -> List[str]:
        return list(self._topics)

    @property
    def metadata(self):
        return self._metadata

    def close(self) -> None:
        pass


class IMUBagSource:
    """`IMUBagSource(data_path, imu_topic)` (bag.py:99-150): IMU samples of ROS bags, from `sensor_msgs/Imu`
    messages (header stamp, angular_velocity, linear_acceleration) or Ouster `imu_packets`."""

    def __init__(self, data_path, imu_topic: Optional[str] = None):
        self._reader = BagReader(data_path)
        conns = [c for c in self._reader.scan_connections()
                 if c[1] == "sensor_msgs/Imu" or (c[1] == "ouster_ros/PacketMsg" and c[0].endswith("imu_packets"))]
        assert len(conns), "Expect any topic with msgtype: sensor_msgs/msg/Imu or Ouster imu_packets types but found None"
        if imu_topic is not None:
            self._conns = [c for c in conns if c[0] == imu_topic]
            assert len(self._conns), f"Expect a topic with msgtype: sensor_msgs/msg/Imu and '{imu_topic}' name but found None"
        else:
            self._conns = [conns[0]]

    def __iter__(self) -> Iterator[IMU]:
        for (topic, typ, _), t, d in self._reader.messages(lambda c: c in self._conns):
            if typ == "sensor_msgs/Imu":
                # std_msgs/Header: seq u32, stamp (sec u32, nsec u32), frame_id string; then orientation (4 f64) +
                # covariance (9), angular_velocity (3) + covariance (9), linear_acceleration (3) + covariance (9)
                _, sec, nsec, n = struct.unpack_from("<IIII", d, 0)
                v = struct.unpack_from("<37d", d, 16 + n)
                yield IMU(np.array(v[25:28]), np.array(v[13:16]), sec + nsec * 1e-9)
            else:
                yield imu_from_packet(ImuPacket(_packetmsg_buf(d), None, t * 1e-9))


def read_metadata_json(meta_path: str):
    """`read_metadata_json` (utils.py:157-168): SensorInfo from a sensor metadata file, with the reference's backfill
    of `lidar_mode` for the Newer College 2020 metadata."""
    import json
    from .ouster_compat import SensorInfo
    with open(meta_path) as f:
        js = json.load(f)
    if "beam_altitude_angles" in js and "beam_azimuth_angles" in js and "lidar_mode" not in js:
        print(f"WARNING: lidar_mode is not present in legacy metadata '{meta_path}' so using lidar_mode: 1024x10")
        js["lidar_mode"] = "1024x10"
    if hasattr(SensorInfo, "from_json"):
        return SensorInfo.from_json(json.dumps(js))
    return SensorInfo(json.dumps(js))           # the real ouster-sdk class


def read_packet_source(file_path: str, meta=None):
    """`read_packet_source` (utils.py:171-187): pcap file, bag file, or a directory of bags."""
    file = Path(file_path)
    if file.is_file():
        if file.suffix == ".pcap":
            return PcapSource(file_path, meta)
        if file.suffix == ".bag":
            return OusterRawBagSource(file, meta)
    elif file.is_dir():
        return OusterRawBagSource(sorted(Path(p) for p in glob.glob(str(file / "*.bag"))), meta)
    return None


class OusterLidarData:
    """Lidar data source: LidarScan + IMUs iterator with scan index (data.py:13-77), scans batched on the GPU."""

    def __init__(self, source, *, fields=None, device: int = 0) -> None:
        self._source = source
        self._fields = tuple(fields) if fields is not None else ("RANGE",)
        self._device = device
        self._scan_idx = 0

    def withScanIdx(self, *, start_scan: int = 0, end_scan: Optional[int] = None):
        """(scan index, DeviceLidarScan | IMU) in packet order, with the reference's semantics (data.py:31-77): IMU
        samples carry the index of the scan being assembled; a scan is handed out when the first packet of the next
        frame arrives (or the stream ends while one is being assembled); scans before `start_scan` are batched and
        counted but not handed out; iteration stops once a scan past `end_scan` has been closed."""
        info = self._source.metadata
        w, h = info.format.columns_per_frame, info.format.pixels_per_column
        batcher = ScanBatcher(w, PacketFormat.from_info(info), device=self._device)
        scan_idx, scan = 0, None
        try:
            for packet in self._source:
                if isinstance(packet, ImuPacket):
                    if scan_idx >= start_scan:
                        yield scan_idx, imu_from_packet(packet)
                    continue
                if not isinstance(packet, LidarPacket):
                    continue
                if scan is None:
                    scan = DeviceLidarScan(h, w, self._fields, self._device)
                if not batcher(packet, scan):
                    continue
                # the frame is complete (the packet that closed it already sits in the next frame's slots)
                if scan_idx >= start_scan:
                    yield scan_idx, scan
                scan_idx += 1
                scan = None
                if end_scan is not None and scan_idx > end_scan:
                    return
            if scan is not None and batcher.flush(scan):
                yield scan_idx, scan
        finally:
            batcher.close()

    def __iter__(self):
        """Make an iterator just data"""
        for scan_idx, d in self.withScanIdx():
            yield scan_idx, d

    def close(self) -> None:
        """Close the underlying PacketSource."""
        self._source.close()

    @property
    def metadata(self):
        """Return metadata from the underlying PacketSource."""
        return self._source.metadata


def last_valid_packet_ts(scan) -> int:
    """data.py:95-99 for scans without a packet_timestamp array: timestamp of the last valid column."""
    valid = np.flatnonzero(np.asarray(scan.status) & 1)
    return int(scan.timestamp[valid[-1]]) if valid.size else 0


def encode_scan_packets(pf: PacketFormat, frame_id: int, range_mm, timestamp_ns, signal=None) -> np.ndarray:
    """Synthetic sensor output: the lidar packets (packets_per_frame, lidar_packet_size) uint8 of one scan in the
    LEGACY or RNG19_RFL8_SIG16_NIR16 format - the input side of the synthetic pipeline (what a sensor would have
    sent for `SynthSequence.scan(k)`), vectorised so that a bench can produce frames quickly."""
    H, W, cpp = pf.pixels_per_column, pf.columns_per_frame, pf.columns_per_packet
    legacy = pf.profile == PROFILE_LIDAR_LEGACY
    if not legacy and pf.profile != PROFILE_LIDAR_RNG19_RFL8_SIG16_NIR16:
        raise ValueError("encode_scan_packets: LEGACY and RNG19_RFL8_SIG16_NIR16 only")
    px = np.dtype([("range", "<u4"), ("refl", "<u2"), ("signal", "<u2"), ("nir", "<u2"), ("pad", "<u2")])
    if legacy:
        col = np.dtype([("ts", "<u8"), ("mid", "<u2"), ("fid", "<u2"), ("enc", "<u4"), ("px", px, (H,)), ("status", "<u4")])
        pkt = np.dtype([("col", col, (cpp,))])
    else:
        col = np.dtype([("ts", "<u8"), ("mid", "<u2"), ("status", "<u2"), ("px", px, (H,))])
        pkt = np.dtype([("type", "<u2"), ("fid", "<u2"), ("rest", "u1", (28,)), ("col", col, (cpp,)), ("footer", "u1", (32,))])
    assert pkt.itemsize == pf.lidar_packet_size
    out = np.zeros(W // cpp, dtype=pkt)
    c = out["col"]
    c["ts"] = np.asarray(timestamp_ns, dtype=np.uint64).reshape(-1, cpp)
    c["mid"] = np.arange(W, dtype=np.uint16).reshape(-1, cpp)
    r = np.asarray(range_mm, dtype=np.uint32)
    c["px"]["range"] = (r & (0xFFFFF if legacy else 0x7FFFF)).T.reshape(-1, cpp, H)
    if signal is not None:
        c["px"]["signal"] = np.asarray(signal, dtype=np.uint16).T.reshape(-1, cpp, H)
    if legacy:
        c["fid"] = frame_id & 0xFFFF
        c["status"] = 0xFFFFFFFF
    else:
        out["type"] = 1
        out["fid"] = frame_id & 0xFFFF
        c["status"] = 1
    return out.view(np.uint8).reshape(W // cpp, pf.lidar_packet_size)
