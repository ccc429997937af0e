"""TEST INFRASTRUCTURE - CPU oracle of the ingest step (packets -> LidarScan fields), plus the writers the
tests use to make packet streams, pcap files and ROS bags.  Only tests/ may import it.

Restates what /root/reference/src/ptudes/data.py:45-60 gets from ouster-sdk (`PacketFormat.from_info`,
`ScanBatcher.__call__`, `LidarScan`), which is absent from /root/reference and not installable: the packet
layouts are the published Ouster sensor UDP formats (firmware user manual, "Lidar data packet format"):

  LEGACY                        no packet header; per column 16 B header (timestamp u64, measurement id u16,
                                frame id u16, encoder u32), H x 12 B channel blocks (range u32 & 0xfffff,
                                reflectivity u16, signal u16, near-ir u16, 2 B unused), 4 B status
                                (0xffffffff valid / 0)                                -> 24896 B for H=128
  RNG19_RFL8_SIG16_NIR16        32 B packet header (type u16, frame id u16, init id u24, serial u40, ...),
                                per column 12 B header (timestamp u64, measurement id u16, status u16 bit 0),
                                H x 12 B (range u32 & 0x7ffff, reflectivity u8 @4, signal u16 @6, near-ir
                                u16 @8), 32 B packet footer                            -> 24832 B
  RNG15_RFL8_NIR8               same framing, H x 4 B (range u16 & 0x7fff in 8 mm units, reflectivity u8,
                                near-ir u8 in units of 16)                              -> 8448 B
  RNG19_RFL8_SIG16_NIR16_DUAL   same framing, H x 16 B (range u32 & 0x7ffff, reflectivity u8 @3, range2 u32
                                @4, reflectivity2 u8 @7, signal u16 @8, signal2 u16 @10, near-ir u16 @12)
                                                                                        -> 33024 B
**parity unpinned**: no reference fixture or SDK to check these against; the four packet sizes are the
known answers.  ScanBatcher rules restated: a packet of another frame closes the current one unless it is a
straggler of the previous frame (dropped); columns without the valid bit are skipped; columns nobody wrote
read zero in every field and header.
"""
import struct

import numpy as np

LEGACY, DUAL, RNG19, RNG15 = 1, 2, 3, 4


class Format:
    def __init__(self, profile, H, cpp, W):
        self.profile, self.H, self.cpp, self.W = profile, H, cpp, W
        if profile == LEGACY:
            self.pkt_hdr, self.col_hdr, self.ch, self.col_ftr, self.pkt_ftr = 0, 16, 12, 4, 0
        else:
            self.pkt_hdr, self.col_hdr, self.col_ftr, self.pkt_ftr = 32, 12, 0, 32
            self.ch = {RNG19: 12, RNG15: 4, DUAL: 16}[profile]
        self.col_size = self.col_hdr + H * self.ch + self.col_ftr
        self.size = self.pkt_hdr + cpp * self.col_size + self.pkt_ftr
        self.ppf = W // cpp


def frame_id(F, pkt):
    return struct.unpack_from("<H", pkt, 10 if F.profile == LEGACY else 2)[0]


def encode_frame(F, fid, fields, timestamp, valid=None, first_col=0, n_cols=None):
    """Packets (list of bytes) of columns [first_col, first_col + n_cols) of one frame.  fields: dict of (H, W)
    integer arrays (RANGE required; RANGE is given in mm and must be representable in the profile)."""
    n_cols = F.W - first_col if n_cols is None else n_cols
    valid = np.ones(F.W, bool) if valid is None else valid
    get = lambda n: fields.get(n, np.zeros((F.H, F.W), np.uint32))  # noqa: E731
    out = []
    for c0 in range(first_col, first_col + n_cols, F.cpp):
        b = bytearray(F.size)
        if F.profile != LEGACY:
            struct.pack_into("<HH", b, 0, 1, fid)
        for c in range(F.cpp):
            m = c0 + c
            o = F.pkt_hdr + c * F.col_size
            if F.profile == LEGACY:
                struct.pack_into("<QHHI", b, o, int(timestamp[m]), m, fid, 0)
                struct.pack_into("<I", b, o + F.col_size - 4, 0xFFFFFFFF if valid[m] else 0)
            else:
                struct.pack_into("<QHH", b, o, int(timestamp[m]), m, 1 if valid[m] else 0)
            for p in range(F.H):
                q = o + F.col_hdr + p * F.ch
                r = int(get("RANGE")[p, m])
                if F.profile == LEGACY:
                    struct.pack_into("<IHHH", b, q, r & 0xFFFFF, int(get("REFLECTIVITY")[p, m]), int(get("SIGNAL")[p, m]),
                                     int(get("NEAR_IR")[p, m]))
                elif F.profile == RNG19:
                    struct.pack_into("<IBxHH", b, q, r & 0x7FFFF, int(get("REFLECTIVITY")[p, m]) & 0xFF, int(get("SIGNAL")[p, m]),
                                     int(get("NEAR_IR")[p, m]))
                elif F.profile == RNG15:
                    struct.pack_into("<HBB", b, q, (r >> 3) & 0x7FFF, int(get("REFLECTIVITY")[p, m]) & 0xFF,
                                     (int(get("NEAR_IR")[p, m]) >> 4) & 0xFF)
                else:
                    w0 = (r & 0x7FFFF) | ((int(get("REFLECTIVITY")[p, m]) & 0xFF) << 24)
                    w1 = int(get("RANGE2")[p, m]) & 0x7FFFF
                    struct.pack_into("<IIHHH", b, q, w0, w1, int(get("SIGNAL")[p, m]), 0, int(get("NEAR_IR")[p, m]))
        out.append(bytes(b))
    return out


def decode_frame(F, packets):
    """ScanBatcher over the packets of ONE frame, in the given order (later packets overwrite earlier columns);
    returns dict of fields + headers, zero where nothing valid arrived."""
    z32 = lambda: np.zeros((F.H, F.W), np.uint32)  # noqa: E731
    z16 = lambda: np.zeros((F.H, F.W), np.uint16)  # noqa: E731
    out = {"RANGE": z32(), "RANGE2": z32(), "REFLECTIVITY": z16(), "SIGNAL": z16(), "NEAR_IR": z16(),
           "timestamp": np.zeros(F.W, np.uint64), "status": np.zeros(F.W, np.uint32), "measurement_id": np.zeros(F.W, np.uint16)}
    for pkt in packets:
        for c in range(F.cpp):
            o = F.pkt_hdr + c * F.col_size
            ts, m = struct.unpack_from("<QH", pkt, o)
            st = struct.unpack_from("<I", pkt, o + F.col_size - 4)[0] if F.profile == LEGACY else struct.unpack_from("<H", pkt, o + 10)[0]
            if not (st & 1) or m >= F.W:
                continue
            out["timestamp"][m], out["status"][m], out["measurement_id"][m] = ts, st, m
            for p in range(F.H):
                q = o + F.col_hdr + p * F.ch
                if F.profile == LEGACY:
                    r, refl, sig, nir = struct.unpack_from("<IHHH", pkt, q)
                    out["RANGE"][p, m], out["REFLECTIVITY"][p, m], out["SIGNAL"][p, m], out["NEAR_IR"][p, m] = r & 0xFFFFF, refl, sig, nir
                elif F.profile == RNG19:
                    r, refl, sig, nir = struct.unpack_from("<IBxHH", pkt, q)
                    out["RANGE"][p, m], out["REFLECTIVITY"][p, m], out["SIGNAL"][p, m], out["NEAR_IR"][p, m] = r & 0x7FFFF, refl, sig, nir
                elif F.profile == RNG15:
                    r, refl, nir = struct.unpack_from("<HBB", pkt, q)
                    out["RANGE"][p, m], out["REFLECTIVITY"][p, m], out["NEAR_IR"][p, m] = (r & 0x7FFF) << 3, refl, nir << 4
                else:
                    w0, w1, sig, _, nir = struct.unpack_from("<IIHHH", pkt, q)
                    out["RANGE"][p, m], out["REFLECTIVITY"][p, m] = w0 & 0x7FFFF, w0 >> 24
                    out["RANGE2"][p, m], out["SIGNAL"][p, m], out["NEAR_IR"][p, m] = w1 & 0x7FFFF, sig, nir
    return out


def batch_stream(F, packets):
    """ScanBatcher's frame grouping over a packet stream: list of (frame id, [packets of the frame in arrival
    order]); the last entry is the partial frame the end of the stream leaves (data.py:52-56)."""
    frames, cur, cur_id = [], [], None
    for pkt in packets:
        fid = frame_id(F, pkt)
        if cur_id is not None and fid != cur_id:
            if cur_id == ((fid + 1) & 0xFFFF):
                continue                      # straggler of the previous frame
            frames.append((cur_id, cur))
            cur, cur_id = [], None
        if cur_id is None:
            cur_id = fid
        cur.append(pkt)
    if cur_id is not None:
        frames.append((cur_id, cur))
    return frames


def imu_packet(sys_ts, accel_ts, gyro_ts, accel_g, gyro_dps):
    return struct.pack("<3Q6f", sys_ts, accel_ts, gyro_ts, *accel_g, *gyro_dps)


# ---- writers (test inputs) ---------------------------------------------------------------------------
def write_pcap(path, datagrams, mtu=1500, nanos=False, vlan=False, linktype=1, shuffle_fragments=False):
    """datagrams: list of (ts seconds, dst port, payload).  IPv4 + UDP under an Ethernet (linktype 1, optionally
    VLAN-tagged), Linux cooked (113) or raw-IP (101) link layer; payloads larger than the MTU are sent as IP fragments
    (what a sensor's 24 kB lidar packets look like on the wire), optionally with the fragments of a datagram out of
    order."""
    with open(path, "wb") as f:
        f.write(struct.pack("<IHHiIII", 0xA1B23C4D if nanos else 0xA1B2C3D4, 2, 4, 0, 0, 65535, linktype))
        ident = 1
        for ts, port, payload in datagrams:
            udp = struct.pack(">HHHH", 40000, port, 8 + len(payload), 0) + payload
            step = (mtu - 20) // 8 * 8
            offs = list(range(0, len(udp), step))
            if shuffle_fragments and len(offs) > 2:
                offs = [offs[-1]] + offs[1:-1] + [offs[0]]
            for off in offs:
                part = udp[off:off + step]
                more = off + step < len(udp)
                ip = struct.pack(">BBHHHBBH4s4s", 0x45, 0, 20 + len(part), ident, (0x2000 if more else 0) | (off // 8), 64, 17, 0,
                                 bytes([192, 168, 1, 10]), bytes([192, 168, 1, 2]))
                if linktype == 1:
                    link = b"\x02" * 6 + b"\x04" * 6 + (b"\x81\x00\x00\x05" if vlan else b"") + b"\x08\x00"
                elif linktype == 113:
                    link = struct.pack(">HHH8sH", 0, 1, 6, b"\x02" * 6 + b"\x00\x00", 0x0800)
                else:
                    link = b""
                rec = link + ip + part
                sec = int(ts)
                frac = int(round((ts - sec) * (1e9 if nanos else 1e6)))
                f.write(struct.pack("<IIII", sec, frac, len(rec), len(rec)) + rec)
            ident = (ident + 1) & 0xFFFF


def pcap_to_pcapng(src, dst, tsresol=9, big_endian=False):
    """Rewrite a classic little-endian pcap as a pcapng file (section header, one interface description with an
    if_tsresol option, enhanced packet blocks) - the format Wireshark writes by default."""
    e = ">" if big_endian else "<"
    raw = open(src, "rb").read()
    magic, _, _, _, _, _, linktype = struct.unpack_from("<IHHiIII", raw, 0)
    nanos = magic == 0xA1B23C4D

    def block(btype, body):
        pad = (-len(body)) % 4
        total = 12 + len(body) + pad
        return struct.pack(e + "II", btype, total) + body + b"\x00" * pad + struct.pack(e + "I", total)

    out = block(0x0A0D0D0A, struct.pack(e + "IHHq", 0x1A2B3C4D, 1, 0, -1))
    opt = struct.pack(e + "HHB3x", 9, 1, tsresol) + struct.pack(e + "HH", 0, 0)
    out += block(1, struct.pack(e + "HHI", linktype, 0, 65535) + opt)
    pos = 24
    while pos + 16 <= len(raw):
        sec, frac, incl, orig = struct.unpack_from("<IIII", raw, pos)
        data = raw[pos + 16:pos + 16 + incl]
        pos += 16 + incl
        ticks = (sec * 10**9 + frac * (1 if nanos else 1000)) * 10**tsresol // 10**9
        out += block(6, struct.pack(e + "IIIII", 0, ticks >> 32, ticks & 0xFFFFFFFF, incl, orig) + data)
    open(dst, "wb").write(out)


def lz4_block_compress(data: bytes) -> bytes:
    """A plain greedy LZ4 block compressor (hash of 4-byte windows; the format's end-of-block rules: the last five
    bytes are literals, no match starts in the last twelve).  Test input for the native decoder."""
    n, out, anchor, i, table = len(data), bytearray(), 0, 0, {}

    def emit(lit: bytes, mlen: int, offset: int):
        ll = len(lit)
        token = (min(ll, 15) << 4) | (min(mlen - 4, 15) if mlen else 0)
        out.append(token)
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255); r -= 255
            out.append(r)
        out.extend(lit)
        if mlen:
            out.extend(struct.pack("<H", offset))
            if mlen - 4 >= 15:
                r = mlen - 4 - 15
                while r >= 255:
                    out.append(255); r -= 255
                out.append(r)

    while i + 12 < n:
        key = data[i:i + 4]
        j = table.get(key)
        table[key] = i
        if j is not None and i - j <= 65535:
            m = 4
            while i + m < n - 5 and data[j + m] == data[i + m]:
                m += 1
            emit(data[anchor:i], m, i - j)
            i += m
            anchor = i
        else:
            i += 1
    emit(data[anchor:], 0, 0)
    return bytes(out)


def lz4_frame(data: bytes, block=65536, stored_every=0) -> bytes:
    """LZ4 frame with independent blocks (FLG 0x60, BD 0x40 = 64 KB blocks); every `stored_every`-th block is written
    uncompressed (high bit of the block size).  The header checksum byte is not computed (our decoder skips it)."""
    out = bytearray(struct.pack("<IBBB", 0x184D2204, 0x60, 0x40, 0))
    for k, o in enumerate(range(0, len(data), block)):
        raw = data[o:o + block]
        comp = lz4_block_compress(raw)
        if (stored_every and k % stored_every == stored_every - 1) or len(comp) >= len(raw):
            out += struct.pack("<I", len(raw) | 0x80000000) + raw
        else:
            out += struct.pack("<I", len(comp)) + comp
    out += struct.pack("<I", 0)
    return bytes(out)


def _bag_header(**kv):
    b = b""
    for k, v in kv.items():
        e = k.encode() + b"=" + v
        b += struct.pack("<I", len(e)) + e
    return b


def _bag_record(hdr, data):
    return struct.pack("<I", len(hdr)) + hdr + struct.pack("<I", len(data)) + data


def imu_msg(seq, ts, frame_id, avel, lacc):
    """serialized sensor_msgs/Imu (ROS1)"""
    sec = int(ts)
    fid = frame_id.encode()
    return (struct.pack("<IIII", seq, sec, int(round((ts - sec) * 1e9)), len(fid)) + fid +
            struct.pack("<37d", 0, 0, 0, 1, *([0.0] * 9), *avel, *([0.0] * 9), *lacc, *([0.0] * 9)))


def write_bag(path, messages, compression="none", per_chunk=50, imu_topics=()):
    """messages: list of (ts seconds, topic, payload bytes) -> ROS bag 2.0.  Topics are ouster_ros/PacketMsg
    (payload = the packet, wrapped as `uint8[] buf`) unless listed in `imu_topics` (payload = a serialized
    sensor_msgs/Imu, written as is)."""
    import bz2
    md5 = b"4f7b5949e76f86d01e96b0e33ba9b5e3"
    topics = sorted({m[1] for m in messages})
    cid = {t: i for i, t in enumerate(topics)}
    with open(path, "wb") as f:
        f.write(b"#ROSBAG V2.0\n")
        bh = _bag_header(op=b"\x03", index_pos=struct.pack("<Q", 0), conn_count=struct.pack("<I", len(topics)),
                         chunk_count=struct.pack("<I", 0))
        f.write(struct.pack("<I", len(bh)) + bh + struct.pack("<I", 4096 - len(bh) - 8) + b" " * (4096 - len(bh) - 8))
        seen = set()
        for i in range(0, len(messages), per_chunk):
            body = b""
            for ts, topic, payload in messages[i:i + per_chunk]:
                if topic not in seen:
                    seen.add(topic)
                    if topic in imu_topics:
                        ch = _bag_header(topic=topic.encode(), type=b"sensor_msgs/Imu", md5sum=b"6a62c6daae103f4ff57a132d6f95cec2",
                                         message_definition=b"Header header\n")
                    else:
                        ch = _bag_header(topic=topic.encode(), type=b"ouster_ros/PacketMsg", md5sum=md5, message_definition=b"uint8[] buf\n")
                    body += _bag_record(_bag_header(op=b"\x07", conn=struct.pack("<I", cid[topic]), topic=topic.encode()), ch)
                sec = int(ts)
                nsec = int(round((ts - sec) * 1e9))
                body += _bag_record(_bag_header(op=b"\x02", conn=struct.pack("<I", cid[topic]), time=struct.pack("<II", sec, nsec)),
                                    payload if topic in imu_topics else struct.pack("<I", len(payload)) + payload)
            data = bz2.compress(body) if compression == "bz2" else (lz4_frame(body, stored_every=3) if compression == "lz4" else body)
            f.write(_bag_record(_bag_header(op=b"\x05", compression=compression.encode(), size=struct.pack("<I", len(body))), data))
