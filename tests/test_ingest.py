"""Ingest: packet formats, frame grouping, pcap / ROS-bag readers on the CPU; the device decode
(`ptk_batcher_decode`, `ptk_decode_packets`) against the oracle on the GPU, and packets -> scans -> poses
end to end.  Reference: /root/reference/src/ptudes/data.py:31-77, bag.py:21-97, utils.py:171-187."""
import ctypes as C

import numpy as np
import pytest

from oracle import ingest_oracle as io
from ptudes_lab_b200 import _ffi, ingest

PROFILES = [io.LEGACY, io.RNG19, io.RNG15, io.DUAL]


def _fields(F, seed, full=True):
    rng = np.random.default_rng(seed)
    rmax = {io.LEGACY: 1 << 20, io.RNG19: 1 << 19, io.DUAL: 1 << 19, io.RNG15: 1 << 18}[F.profile]
    r = rng.integers(0, rmax, size=(F.H, F.W), dtype=np.uint32)
    r[rng.random((F.H, F.W)) < 0.2] = 0
    if F.profile == io.RNG15:
        r &= ~np.uint32(7)
    f = {"RANGE": r}
    if full:
        f["REFLECTIVITY"] = rng.integers(0, 256, size=(F.H, F.W), dtype=np.uint32)
        f["SIGNAL"] = rng.integers(0, 65536, size=(F.H, F.W), dtype=np.uint32)
        nir = rng.integers(0, 65536 if F.profile != io.RNG15 else 4096, size=(F.H, F.W), dtype=np.uint32)
        f["NEAR_IR"] = nir & ~np.uint32(15) if F.profile == io.RNG15 else nir
        f["RANGE2"] = rng.integers(0, 1 << 19, size=(F.H, F.W), dtype=np.uint32)
    ts = (1_000_000_000 + np.arange(F.W) * 97_656).astype(np.uint64)
    return f, ts


def test_packet_sizes_are_the_published_ones():
    for prof, size in ((io.LEGACY, 24896), (io.RNG19, 24832), (io.RNG15, 8448), (io.DUAL, 33024)):
        pf = ingest.PacketFormat(prof, 128, 16, 1024)
        assert pf.lidar_packet_size == size == io.Format(prof, 128, 16, 1024).size
        assert pf.packets_per_frame == 64
    assert ingest.PacketFormat("LEGACY", 64, 16, 2048).lidar_packet_size == 16 * (16 + 64 * 12 + 4)
    with pytest.raises(_ffi.PtkError):
        ingest.PacketFormat(7, 128, 16, 1024)
    with pytest.raises(_ffi.PtkError):
        ingest.PacketFormat(io.LEGACY, 128, 16, 1000)          # not a whole number of packets


@pytest.mark.parametrize("prof", PROFILES)
def test_oracle_encode_decode_round_trip(prof):
    F = io.Format(prof, 8, 16, 64)
    f, ts = _fields(F, prof)
    pk = io.encode_frame(F, 5, f, ts)
    assert len(pk) == F.ppf and all(len(p) == F.size for p in pk)
    d = io.decode_frame(F, pk)
    assert np.array_equal(d["RANGE"], f["RANGE"])
    assert np.array_equal(d["timestamp"], ts) and np.array_equal(d["measurement_id"], np.arange(F.W))
    assert np.array_equal(d["NEAR_IR"], f["NEAR_IR"]) and np.array_equal(d["REFLECTIVITY"], f["REFLECTIVITY"])
    if prof != io.RNG15:
        assert np.array_equal(d["SIGNAL"], f["SIGNAL"])
    if prof == io.DUAL:
        assert np.array_equal(d["RANGE2"], f["RANGE2"])
    assert io.frame_id(F, pk[3]) == 5


def _stream(F, n_frames, seed=0, first_id=65534):
    """frames with wrap-around ids; returns (list of packets per frame, fields per frame)"""
    frames, truth = [], []
    for k in range(n_frames):
        f, ts = _fields(F, seed + k, full=False)
        frames.append(io.encode_frame(F, (first_id + k) & 0xFFFF, f, ts + np.uint64(k * 100_000_000)))
        truth.append(f)
    return frames, truth


@pytest.mark.parametrize("prof", [io.LEGACY, io.RNG19])
def test_host_batcher_groups_frames_like_the_oracle(prof):
    """ptk_batcher_* with device = -1: frame grouping only.  Stream with id wrap-around, a lost packet, a
    repeated packet, out-of-order arrival and a straggler of the previous frame."""
    F = io.Format(prof, 4, 16, 128)
    frames, _ = _stream(F, 4)
    s = []
    s += frames[0]
    f1 = list(frames[1]); del f1[2]; f1[0], f1[1] = f1[1], f1[0]; f1.append(frames[1][5])
    s += f1
    s += [frames[2][0], frames[1][7]] + frames[2][1:]          # frames[1][7] arrives late: dropped
    s += frames[3][:3]                                          # partial last frame
    want = io.batch_stream(F, s)
    assert [w[0] for w in want] == [65534, 65535, 0, 1] and len(want[1][1]) == F.ppf and len(want[2][1]) == F.ppf

    pf = ingest.PacketFormat(prof, F.H, F.cpp, F.W)
    lib = _ffi.load()
    h = C.c_void_p()
    assert lib.ptk_batcher_create(C.byref(h), -1, C.byref(pf.c), 3) == 0
    got = []

    def drain():
        p, fid, n = C.c_void_p(), C.c_int(), C.c_int()
        assert lib.ptk_batcher_peek(h, C.byref(p), C.byref(fid), C.byref(n)) == 0
        raw = C.string_at(p.value, F.ppf * F.size)
        got.append((fid.value, n.value, [raw[i * F.size:(i + 1) * F.size] for i in range(F.ppf)]))
        assert lib.ptk_batcher_pop(h) == 0

    ready = C.c_int()
    for pkt in s:
        b = np.frombuffer(pkt, dtype=np.uint8)
        assert lib.ptk_batcher_push(h, b.ctypes.data, C.byref(ready)) == 0
        while ready.value:
            drain()
            ready.value -= 1
    assert lib.ptk_batcher_flush(h, C.byref(ready)) == 0 and ready.value == 1
    drain()
    assert lib.ptk_batcher_peek(h, None, None, None) == _ffi_code("PTK_E_STATE")
    lib.ptk_batcher_destroy(h)

    assert [g[0] for g in got] == [w[0] for w in want]
    for (fid, n, slots), (_, pkts) in zip(got, want):
        ref = io.decode_frame(F, pkts)
        dec = io.decode_frame(F, slots)                         # zero slots decode to nothing
        assert n == len({io.struct.unpack_from("<H", p, F.pkt_hdr + 8)[0] for p in pkts})
        for k in ref:
            assert np.array_equal(ref[k], dec[k]), k


def test_packet_with_a_blank_first_column_goes_to_its_own_slot():
    """A column without the valid bit may carry a blank header (measurement id 0): the packet's position in the frame
    follows from its valid columns, not from that header."""
    F = io.Format(io.RNG19, 4, 16, 64)
    f, ts = _fields(F, 5, full=False)
    valid = np.ones(F.W, bool)
    valid[32] = False
    pk = io.encode_frame(F, 3, f, ts, valid=valid)
    blank = bytearray(pk[2])
    blank[F.pkt_hdr:F.pkt_hdr + F.col_hdr] = bytes(F.col_hdr)         # first column of packet 2: header all zero
    pk[2] = bytes(blank)
    pf = ingest.PacketFormat(io.RNG19, F.H, F.cpp, F.W)
    lib = _ffi.load()
    h = C.c_void_p()
    assert lib.ptk_batcher_create(C.byref(h), -1, C.byref(pf.c), 2) == 0
    ready = C.c_int()
    for p in pk + [bytes(F.size)[:2] + (3).to_bytes(2, "little") + bytes(F.size - 4)]:      # + a packet without any valid column
        assert lib.ptk_batcher_push(h, np.frombuffer(p, dtype=np.uint8).ctypes.data, C.byref(ready)) == 0
    assert lib.ptk_batcher_flush(h, C.byref(ready)) == 0 and ready.value == 1
    ptr, n = C.c_void_p(), C.c_int()
    assert lib.ptk_batcher_peek(h, C.byref(ptr), None, C.byref(n)) == 0 and n.value == F.ppf
    raw = C.string_at(ptr.value, F.ppf * F.size)
    slots = [raw[i * F.size:(i + 1) * F.size] for i in range(F.ppf)]
    assert slots == pk
    want, got = io.decode_frame(F, pk), io.decode_frame(F, slots)
    assert np.array_equal(want["RANGE"], got["RANGE"]) and want["status"][32] == 0 and want["RANGE"][:, 32].sum() == 0
    lib.ptk_batcher_destroy(h)


def _ffi_code(name):
    return {v: k for k, v in _ffi.ERRORS.items()}[name]


def test_batcher_refuses_to_overrun_its_ring():
    F = io.Format(io.LEGACY, 4, 16, 64)
    frames, _ = _stream(F, 4, first_id=10)
    pf = ingest.PacketFormat(io.LEGACY, F.H, F.cpp, F.W)
    lib = _ffi.load()
    h = C.c_void_p()
    assert lib.ptk_batcher_create(C.byref(h), -1, C.byref(pf.c), 2) == 0
    rc = 0
    for fr in frames:
        for pkt in fr:
            rc = lib.ptk_batcher_push(h, np.frombuffer(pkt, dtype=np.uint8).ctypes.data, None)
            if rc:
                break
        if rc:
            break
    assert rc == _ffi_code("PTK_E_STATE") and b"decode or pop" in lib.ptk_ingest_last_error()
    lib.ptk_batcher_destroy(h)


@pytest.mark.parametrize("nanos,vlan", [(False, False), (True, True)])
def test_pcap_reader_reassembles_fragmented_datagrams(tmp_path, nanos, vlan):
    F = io.Format(io.LEGACY, 16, 16, 64)
    f, ts = _fields(F, 3, full=False)
    pk = io.encode_frame(F, 9, f, ts)
    imu = io.imu_packet(123456789, 123450000, 123460000, (0.01, -0.02, 1.0), (1.5, -2.5, 0.25))
    dg = [(10.0 + 0.001 * i, 7502, p) for i, p in enumerate(pk)]
    dg.insert(2, (10.0015, 7503, imu))
    dg.append((10.5, 9999, b"x" * 100))                          # unrelated traffic
    path = tmp_path / "a.pcap"
    io.write_pcap(path, dg, nanos=nanos, vlan=vlan)

    class Meta:
        class format:
            pixels_per_column, columns_per_frame, columns_per_packet, udp_profile_lidar = F.H, F.W, F.cpp, io.LEGACY

    src = ingest.read_packet_source(str(path), Meta)
    out = list(src)
    src.close()
    assert [type(p).__name__ for p in out] == ["LidarPacket"] * 2 + ["ImuPacket"] + ["LidarPacket"] * 2
    assert [p.buf for p in out if isinstance(p, ingest.LidarPacket)] == pk
    ip = out[2]
    assert (ip.sys_ts, ip.accel_ts, ip.gyro_ts) == (123456789, 123450000, 123460000)
    assert np.allclose(ip.accel, (0.01, -0.02, 1.0), atol=1e-7) and np.allclose(ip.angular_vel, (1.5, -2.5, 0.25))
    assert abs(out[0].capture_timestamp - 10.0) < 1e-6 and abs(out[3].capture_timestamp - 10.002) < 1e-6
    m = ingest.imu_from_packet(ip)
    assert m.ts == 123456789 / 10**9 and np.allclose(m.lacc, ingest.GRAV * ip.accel) and np.allclose(m.avel, np.pi * ip.angular_vel / 180)


@pytest.mark.parametrize("linktype", [113, 101])
def test_pcap_reader_other_link_layers_and_fragment_order(tmp_path, linktype):
    """Linux cooked and raw-IP captures; fragments of a datagram arriving last-first."""
    F = io.Format(io.RNG15, 64, 16, 64)                            # 4352-byte packets: three fragments each
    f, ts = _fields(F, 8, full=False)
    pk = io.encode_frame(F, 1, f, ts)
    assert F.size > 2 * 1480
    path = tmp_path / "b.pcap"
    io.write_pcap(path, [(1.0 + i, 7502, p) for i, p in enumerate(pk)], linktype=linktype, shuffle_fragments=True)

    class Meta:
        class format:
            pixels_per_column, columns_per_frame, columns_per_packet, udp_profile_lidar = F.H, F.W, F.cpp, io.RNG15

    out = list(ingest.PcapSource(str(path), Meta))
    assert [p.buf for p in out] == pk and [round(p.capture_timestamp) for p in out] == [1, 2, 3, 4]
    # the same capture as pcapng (what Wireshark writes), little and big endian, micro- and nanosecond ticks
    for be, res in ((False, 6), (True, 9)):
        ng = tmp_path / f"ng{int(be)}.pcap"
        io.pcap_to_pcapng(path, ng, tsresol=res, big_endian=be)
        out = list(ingest.PcapSource(str(ng), Meta))
        assert [p.buf for p in out] == pk and [round(p.capture_timestamp) for p in out] == [1, 2, 3, 4]
    bad = tmp_path / "c.pcap"
    bad.write_bytes(b"not a capture file" + bytes(40))
    with pytest.raises(_ffi.PtkError):
        ingest.PcapSource(str(bad), Meta)


def test_lz4_frame_decoder():
    """ptk_lz4_frame_decompress against a plain LZ4 compressor (oracle): random, repetitive and empty inputs, stored
    blocks, long literal / match runs, a hand-made frame with DEPENDENT blocks, malformed frames."""
    rng = np.random.default_rng(5)
    cases = [b"", b"a", b"abcdefgh" * 5000, bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
             bytes(rng.integers(0, 4, 200000, dtype=np.uint8)), b"x" * 100000 + bytes(range(256)) * 40]
    for data in cases:
        for stored in (0, 2):
            fr = io.lz4_frame(data, stored_every=stored)
            assert ingest.lz4_frame_decompress(fr, len(data)) == data
    assert len(io.lz4_frame(b"abcdefgh" * 5000)) < 400                      # it does compress
    # dependent blocks (FLG 0x40): the second block is one match reaching 10 bytes back into the first + 5 literals
    first = b"0123456789"
    blk1 = bytes([len(first) << 4]) + first
    blk2 = bytes([(0 << 4) | (8 - 4)]) + (10).to_bytes(2, "little") + bytes([5 << 4]) + b"ABCDE"
    fr = (0x184D2204).to_bytes(4, "little") + bytes([0x40, 0x40, 0]) + len(blk1).to_bytes(4, "little") + blk1 + \
        len(blk2).to_bytes(4, "little") + blk2 + bytes(4)
    assert ingest.lz4_frame_decompress(fr, 23) == b"0123456789" + b"01234567" + b"ABCDE"
    with pytest.raises(_ffi.PtkError):
        ingest.lz4_frame_decompress(b"\x00" * 16, 4)                          # not a frame
    with pytest.raises(_ffi.PtkError):
        ingest.lz4_frame_decompress(io.lz4_frame(b"abcdefgh" * 100)[:-9], 800)  # truncated
    with pytest.raises(ValueError):
        ingest.lz4_frame_decompress(io.lz4_frame(b"abcdefgh" * 100), 900)      # size mismatch with the chunk header


@pytest.mark.parametrize("compression", ["none", "bz2", "lz4"])
def test_bag_source_yields_the_packet_messages(tmp_path, compression):
    F = io.Format(io.RNG19, 8, 16, 64)
    f, ts = _fields(F, 4, full=False)
    pk = io.encode_frame(F, 2, f, ts)
    imu = io.imu_packet(5, 6, 7, (0, 0, 1), (0, 0, 0))
    msgs = []
    for i, p in enumerate(pk):
        msgs.append((100.0 + 0.01 * i, "/os_node/lidar_packets", p))
        msgs.append((100.005 + 0.01 * i, "/os_node/imu_packets", imu))
    msgs.append((100.2, "/tf", b"junk"))
    path = tmp_path / "a.bag"
    io.write_bag(path, msgs, compression=compression, per_chunk=3)
    src = ingest.read_packet_source(str(path), None)
    out = list(src)
    assert [p.buf for p in out if isinstance(p, ingest.LidarPacket)] == pk
    assert sum(isinstance(p, ingest.ImuPacket) for p in out) == len(pk)
    assert [type(p).__name__ for p in out[:4]] == ["LidarPacket", "ImuPacket"] * 2
    assert sorted(src.topics) == ["/os_node/imu_packets", "/os_node/lidar_packets"]
    only = ingest.OusterRawBagSource(path, None, lidar_topic="/os_node/lidar_packets")
    assert all(isinstance(p, ingest.LidarPacket) for p in only) and len(list(only)) == len(pk)
    d = ingest.read_packet_source(str(tmp_path), None)           # a directory of bags
    assert len(list(d)) == len(out)


def test_imu_bag_source(tmp_path):
    """bag.py:99-150: sensor_msgs/Imu topics, or Ouster imu_packets when that is what the bag has."""
    msgs = []
    for i in range(7):
        t = 50.0 + 0.01 * i
        msgs.append((t, "/alphasense/imu", io.imu_msg(i, t, "imu_link", (0.1 * i, -0.2, 0.3), (0.0, 9.8, 0.01 * i))))
        msgs.append((t + 0.001, "/os/imu_packets", io.imu_packet(int(t * 1e9), 0, 0, (0, 0, 1), (90.0, 0, 0))))
    path = tmp_path / "imu.bag"
    io.write_bag(path, msgs, per_chunk=4, imu_topics=("/alphasense/imu",))
    out = list(ingest.IMUBagSource(path, imu_topic="/alphasense/imu"))
    assert len(out) == 7
    assert np.allclose(out[3].avel, (0.3, -0.2, 0.3)) and np.allclose(out[3].lacc, (0.0, 9.8, 0.03)) and abs(out[3].ts - 50.03) < 1e-6
    pk = list(ingest.IMUBagSource(path, imu_topic="/os/imu_packets"))
    assert len(pk) == 7 and np.allclose(pk[0].avel, (np.pi / 2, 0, 0)) and np.allclose(pk[0].lacc, (0, 0, ingest.GRAV))
    assert len(list(ingest.IMUBagSource(path))) == 7              # the first IMU connection of the bag
    with pytest.raises(AssertionError):
        ingest.IMUBagSource(path, imu_topic="/nope")


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("prof", PROFILES)
def test_device_decode_equals_the_oracle(prof):
    """Three frames in one launch: complete; with a lost packet, invalid columns and a packet whose columns
    are shifted in measurement id; empty (all packets lost)."""
    import torch
    F = io.Format(prof, 32, 16, 256)
    pf = ingest.PacketFormat(prof, F.H, F.cpp, F.W)
    slots, want = [], []
    f0, ts0 = _fields(F, 10)
    p0 = io.encode_frame(F, 1, f0, ts0)
    slots += p0
    want.append(io.decode_frame(F, p0))
    f1, ts1 = _fields(F, 11)
    valid = np.ones(F.W, bool)
    valid[[5, 40, 41, 255]] = False
    p1 = io.encode_frame(F, 2, f1, ts1, valid=valid)
    p1[3] = bytes(F.size)                                         # lost packet
    slots += p1
    want.append(io.decode_frame(F, p1))
    slots += [bytes(F.size)] * F.ppf
    want.append(io.decode_frame(F, []))
    buf = np.frombuffer(b"".join(slots), dtype=np.uint8)
    names = ("RANGE", "RANGE2", "REFLECTIVITY", "SIGNAL", "NEAR_IR")
    for packets in (buf, torch.from_numpy(buf.copy()).cuda()):      # host bytes and device-resident packets
        # poison the outputs: every pixel must be written (decoded or zero-filled)
        out = ingest.decode_frames(pf, packets, 3, fields=names)
        torch.cuda.synchronize()
        for k in range(3):
            for n in names:
                if (n == "RANGE2" and prof != io.DUAL) or (n == "SIGNAL" and prof == io.RNG15):
                    continue
                assert np.array_equal(out[n][k].cpu().numpy().astype(np.uint32), want[k][n].astype(np.uint32)), (k, n)
            assert np.array_equal(out["timestamp"][k].cpu().numpy().view(np.uint64), want[k]["timestamp"])
            assert np.array_equal(out["status"][k].cpu().numpy().view(np.uint32), want[k]["status"])
            assert np.array_equal(out["measurement_id"][k].cpu().numpy().view(np.uint16), want[k]["measurement_id"])


@pytest.mark.gpu
def test_decode_overwrites_stale_images():
    import torch
    F = io.Format(io.LEGACY, 16, 16, 64)
    pf = ingest.PacketFormat(io.LEGACY, F.H, F.cpp, F.W)
    f, ts = _fields(F, 1, full=False)
    pk = io.encode_frame(F, 1, f, ts)
    pk[1] = bytes(F.size)
    b = ingest.ScanBatcher(F.W, pf, device=0)
    ls = ingest.DeviceLidarScan(F.H, F.W, ("RANGE", "SIGNAL"))
    ls.field("RANGE").fill_(0x7FFFFFFF)
    ls.field("SIGNAL").fill_(999)
    for p in pk:
        assert not b(ingest.LidarPacket(p), ls)
    assert b.flush(ls)
    torch.cuda.synchronize()
    want = io.decode_frame(F, pk)
    assert np.array_equal(ls.field("RANGE").cpu().numpy().view(np.uint32), want["RANGE"])
    assert np.array_equal(ls.field("SIGNAL").cpu().numpy().view(np.uint16), want["SIGNAL"])
    assert ls.frame_id == 1 and ls.n_packets == F.ppf - 1
    assert np.array_equal(ls.status, want["status"]) and np.array_equal(ls.timestamp, want["timestamp"])
    b.close()


@pytest.mark.gpu
def test_packets_to_poses_equals_range_images_to_poses(tmp_path):
    """pcap -> OusterLidarData.withScanIdx -> KissICPWrapper.register_frame on device-resident scans: same
    event order as the reference loop sees, same poses as feeding the range images."""
    from ptudes_lab_b200 import synth
    from ptudes_lab_b200.kiss import KissICPWrapper
    from ptudes_lab_b200.ouster_compat import scan_from_synth, sensor_info_from_synth
    seq = synth.make_sequence("tiny", 2)
    H, W = seq.sensor.H, seq.sensor.W
    F = io.Format(io.RNG19, H, 16, W)
    dg, n_scans = [], 5
    for k in range(n_scans):
        sc = seq.scan(k)
        pk = io.encode_frame(F, 100 + k, {"RANGE": sc.range_mm}, sc.timestamp_ns.astype(np.uint64))
        for i, p in enumerate(pk):
            t = k * 0.1 + i * 0.1 / len(pk)
            dg.append((t, 7502, p))
            if i % 4 == 0:
                dg.append((t + 1e-4, 7503, io.imu_packet(int(t * 1e9), int(t * 1e9), int(t * 1e9), (0, 0, 1), (0, 0, 0.5))))
    path = tmp_path / "seq.pcap"
    io.write_pcap(path, dg)
    meta = sensor_info_from_synth(seq.sensor, seq.dirs)
    meta.format.columns_per_packet = 16
    meta.format.udp_profile_lidar = io.RNG19

    ref = KissICPWrapper(meta, _min_range=5, _max_range=100)
    for k in range(n_scans):
        ref.register_frame(scan_from_synth(seq.scan(k)))

    data = ingest.OusterLidarData(ingest.read_packet_source(str(path), meta))
    w = KissICPWrapper(meta, _min_range=5, _max_range=100)
    events = []
    for idx, d in data.withScanIdx():
        if isinstance(d, ingest.DeviceLidarScan):
            events.append(("scan", idx))
            assert np.array_equal(d.field("RANGE").cpu().numpy().view(np.uint32), seq.scan(idx).range_mm)
            w.register_frame(d)
        else:
            events.append(("imu", idx))
    data.close()
    assert [e for e in events if e[0] == "scan"] == [("scan", k) for k in range(n_scans)]
    # a scan is yielded when the first packet of the next frame arrives: the IMU sample right after that
    # packet already carries the next index
    first_imu_after = events.index(("scan", 0)) + 1
    assert events[first_imu_after] == ("imu", 1)
    assert len(w.poses) == n_scans
    for a, b in zip(w.poses, ref.poses):
        assert np.array_equal(a, b)
    assert w.poses_ts == ref.poses_ts


@pytest.mark.parametrize("prof", [io.LEGACY, io.RNG19])
def test_synthetic_packet_encoder_equals_the_oracle_encoder(prof):
    F = io.Format(prof, 8, 16, 64)
    pf = ingest.PacketFormat(prof, 8, 16, 64)
    f, ts = _fields(F, 21)
    a = ingest.encode_scan_packets(pf, 77, f["RANGE"], ts, signal=f["SIGNAL"])
    b = io.encode_frame(F, 77, {"RANGE": f["RANGE"], "SIGNAL": f["SIGNAL"]}, ts)
    assert [bytes(x) for x in a] == b


def test_host_batcher_random_streams():
    """Property test: random packet streams (lost, repeated, locally reordered packets, stragglers of the previous
    frame, frame-id wrap-around) grouped by ptk_batcher_* decode to what the oracle's ScanBatcher rules give."""
    from hypothesis import given, settings, strategies as st_

    F = io.Format(io.RNG19, 2, 16, 64)
    pf = ingest.PacketFormat(io.RNG19, F.H, F.cpp, F.W)
    lib = _ffi.load()

    @settings(max_examples=30, deadline=None)
    @given(st_.integers(0, 2**32 - 1))
    def run(seed):
        rng = np.random.default_rng(seed)
        first = int(rng.integers(65530, 65536))
        frames, _ = _stream(F, 5, seed=seed % 1000, first_id=first)
        s = []
        for k, fr in enumerate(frames):
            fr = [p for p in fr if rng.random() > 0.15]                       # lost
            fr += [fr[i] for i in rng.integers(0, max(len(fr), 1), size=int(rng.integers(0, 2))) if fr]   # repeated
            if len(fr) > 1 and rng.random() < 0.5:                            # locally reordered
                i = int(rng.integers(0, len(fr) - 1))
                fr[i], fr[i + 1] = fr[i + 1], fr[i]
            if k and fr and rng.random() < 0.5:                               # a straggler of the previous frame
                fr.insert(1, frames[k - 1][int(rng.integers(0, F.ppf))])
            s += fr
        want = io.batch_stream(F, s)
        h = C.c_void_p()
        assert lib.ptk_batcher_create(C.byref(h), -1, C.byref(pf.c), 3) == 0
        got, ready = [], C.c_int()

        def drain():
            p, fid = C.c_void_p(), C.c_int()
            assert lib.ptk_batcher_peek(h, C.byref(p), C.byref(fid), None) == 0
            raw = C.string_at(p.value, F.ppf * F.size)
            got.append((fid.value, [raw[i * F.size:(i + 1) * F.size] for i in range(F.ppf)]))
            assert lib.ptk_batcher_pop(h) == 0

        for pkt in s:
            assert lib.ptk_batcher_push(h, np.frombuffer(pkt, dtype=np.uint8).ctypes.data, C.byref(ready)) == 0
            while ready.value:
                drain()
                ready.value -= 1
        assert lib.ptk_batcher_flush(h, C.byref(ready)) == 0
        if ready.value:
            drain()
        lib.ptk_batcher_destroy(h)
        assert [g[0] for g in got] == [w[0] for w in want]
        for (_, slots), (_, pkts) in zip(got, want):
            ref, dec = io.decode_frame(F, pkts), io.decode_frame(F, slots)
            for k in ("RANGE", "timestamp", "status", "measurement_id"):
                assert np.array_equal(ref[k], dec[k]), k

    run()


def test_sensor_info_from_json_and_xyz_lut_known_answers(tmp_path):
    """utils.py:157-168 + `client.XYZLut` without ouster-sdk: metadata JSON (legacy and firmware >= 2.5 layout) ->
    SensorInfo -> lookup tables by the documented range-to-XYZ formula.  Known answers: a level beam without azimuth
    offset looks along +x at measurement id 0 and along -y a quarter turn later (the encoder runs clockwise); the beam
    origin offset n enters as (r - n) * direction + n * (cos enc, sin enc, 0); the default lidar-to-sensor transform
    turns the frame by 180 degrees about z and lifts it 36.18 mm."""
    import json
    from ptudes_lab_b200.ouster_compat import HAVE_OUSTER, XYZLut
    if HAVE_OUSTER:
        pytest.skip("the real ouster-sdk classes are in use")
    legacy = {"beam_altitude_angles": [10.0, 0.0, -10.0, -20.0], "beam_azimuth_angles": [0.0, 0.0, 3.0, -3.0],
              "lidar_origin_to_beam_origin_mm": 15.8, "lidar_to_sensor_transform": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1],
              "data_format": {"pixels_per_column": 4, "columns_per_packet": 16, "columns_per_frame": 64,
                              "udp_profile_lidar": "RNG19_RFL8_SIG16_NIR16"}}
    p = tmp_path / "meta.json"
    p.write_text(json.dumps(legacy))
    info = ingest.read_metadata_json(str(p))                    # no lidar_mode: backfilled like the reference does
    assert info.mode == "1024x10" and (info.format.columns_per_frame, info.format.pixels_per_column) == (64, 4)
    assert ingest.PacketFormat.from_info(info).profile == io.RNG19
    lut = XYZLut(info)
    r = np.zeros((4, 64), np.uint32)
    r[1, 0] = r[1, 16] = 10000
    r[0, 32] = 5000
    xyz = lut(r)
    assert np.allclose(xyz[1, 0], (10.0, 0.0, 0.0), atol=1e-12) and np.allclose(xyz[1, 16], (0.0, -10.0, 0.0), atol=1e-12)
    n, c, s_ = 0.0158, np.cos(np.radians(10.0)), np.sin(np.radians(10.0))
    assert np.allclose(xyz[0, 32], (-(5.0 - n) * c - n, 0.0, (5.0 - n) * s_), atol=1e-12)
    assert np.all(xyz[2] == 0) and lut.range_unit == 1.0
    # firmware >= 2.5 layout, default lidar-to-sensor transform
    new = {"beam_intrinsics": {"beam_altitude_angles": legacy["beam_altitude_angles"], "beam_azimuth_angles": legacy["beam_azimuth_angles"],
                               "beam_to_lidar_transform": [1, 0, 0, 15.8, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]},
           "lidar_data_format": legacy["data_format"], "config_params": {"lidar_mode": "64x10", "udp_port_lidar": 7510},
           "sensor_info": {"prod_line": "OS-1-4"}}
    from ptudes_lab_b200.ouster_compat import SensorInfo
    info2 = SensorInfo.from_json(json.dumps(new))
    assert info2.mode == "64x10" and info2.udp_port_lidar == 7510 and info2.lidar_origin_to_beam_origin_mm == 15.8
    xyz2 = XYZLut(info2)(r)
    assert np.allclose(xyz2[1, 0], (-10.0, 0.0, 0.03618), atol=1e-12) and np.allclose(xyz2[1, 16], (0.0, 10.0, 0.03618), atol=1e-12)


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_tiny.npz"))


@pytest.mark.parametrize("prof", PROFILES)
def test_oracle_decodes_the_golden_packets(prof):
    """tests/golden/ingest_tiny.npz (make_ingest_golden.py): committed packet bytes -> committed fields."""
    g = _golden()
    F = io.Format(prof, 4, 16, 32)
    pk = [bytes(p) for p in g[f"p{prof}_packets"]]
    assert len(pk) == F.ppf and len(pk[0]) == F.size and io.frame_id(F, pk[1]) == 65535
    d = io.decode_frame(F, pk)
    for k, v in d.items():
        assert np.array_equal(v, g[f"p{prof}_{k}"]), k
    assert d["status"][19] == 0 and not d["RANGE"][:, 19].any()


@pytest.mark.gpu
@pytest.mark.parametrize("prof", PROFILES)
def test_device_decodes_the_golden_packets(prof):
    import torch
    g = _golden()
    pf = ingest.PacketFormat(prof, 4, 16, 32)
    names = ("RANGE", "RANGE2", "REFLECTIVITY", "SIGNAL", "NEAR_IR")
    out = ingest.decode_frames(pf, np.ascontiguousarray(g[f"p{prof}_packets"]).reshape(-1), 1, fields=names)
    torch.cuda.synchronize()
    for n in names:
        if (n == "RANGE2" and prof != io.DUAL) or (n == "SIGNAL" and prof == io.RNG15):
            continue
        assert np.array_equal(out[n][0].cpu().numpy().astype(np.uint32), g[f"p{prof}_{n}"].astype(np.uint32)), n
    assert np.array_equal(out["timestamp"][0].cpu().numpy().view(np.uint64), g[f"p{prof}_timestamp"])
    assert np.array_equal(out["status"][0].cpu().numpy().view(np.uint32), g[f"p{prof}_status"])
