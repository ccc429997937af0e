"""Stand-ins for the few ouster-sdk names /root/reference/src/ptudes/kiss.py uses
(`client.XYZLut`, `client.LidarScan`, `client.ChanField.RANGE`, `client.last_valid_column_ts`,
`SensorInfo.format.{columns_per_frame,pixels_per_column}`), used when ouster-sdk is absent
(it is not installable in the build/bench environment).  With the real SDK present the
real classes are used untouched.
"""
import numpy as np

try:  # pragma: no cover - exercised only where ouster-sdk exists
    import ouster.client as client  # type: ignore
    from ouster.client import ChanField, LidarScan, SensorInfo  # type: ignore
    HAVE_OUSTER = True
    XYZLut = client.XYZLut
    last_valid_column_ts = client.last_valid_column_ts
except Exception:  # ModuleNotFoundError here
    HAVE_OUSTER = False

    class ChanField:
        RANGE = "RANGE"

    class _Format:
        def __init__(self, w, h, columns_per_packet=16, udp_profile_lidar=1, pixel_shift_by_row=None):
            self.columns_per_frame = w
            self.pixels_per_column = h
            self.columns_per_packet = columns_per_packet
            self.udp_profile_lidar = udp_profile_lidar          # UDPProfileLidar value (1 = LEGACY)
            self.pixel_shift_by_row = pixel_shift_by_row if pixel_shift_by_row is not None else [0] * h

    _PROFILES = {"LEGACY": 1, "RNG19_RFL8_SIG16_NIR16_DUAL": 2, "RNG19_RFL8_SIG16_NIR16": 3, "RNG15_RFL8_NIR8": 4}
    # lidar frame -> sensor frame of every Ouster sensor unless the metadata says otherwise (mm)
    _DEFAULT_LIDAR_TO_SENSOR = [-1.0, 0, 0, 0, 0, -1.0, 0, 0, 0, 0, 1.0, 36.18, 0, 0, 0, 1.0]

    class SensorInfo:
        """Minimal `client.SensorInfo`.  Two ways to make one: from the synthetic sensor (frame geometry + a direction
        table), or `SensorInfo.from_json(text)` from a sensor's metadata file - legacy layout (flat keys) or the
        firmware >= 2.5 layout (`beam_intrinsics` / `lidar_data_format` / `config_params` ...) - which carries the beam
        angles the XYZ lookup table is computed from."""

        def __init__(self, columns_per_frame, pixels_per_column, directions=None, prod_line="OS-0-128",
                     mode="1024x10", extrinsic=None):
            self.format = _Format(columns_per_frame, pixels_per_column)
            self.directions = None if directions is None else np.asarray(directions, dtype=np.float64)
            self.prod_line = prod_line
            self.mode = mode
            self.extrinsic = np.eye(4) if extrinsic is None else np.asarray(extrinsic, dtype=np.float64)
            self.beam_altitude_angles = None
            self.beam_azimuth_angles = None
            self.lidar_origin_to_beam_origin_mm = 0.0
            self.lidar_to_sensor_transform = np.array(_DEFAULT_LIDAR_TO_SENSOR).reshape(4, 4)
            self.udp_port_lidar, self.udp_port_imu = 7502, 7503

        @classmethod
        def from_json(cls, text: str) -> "SensorInfo":
            import json
            js = json.loads(text)
            beam = js.get("beam_intrinsics", js)
            fmt = js.get("lidar_data_format", js.get("data_format", {}))
            cfg = js.get("config_params", js)
            mode = cfg.get("lidar_mode", js.get("lidar_mode", "1024x10"))
            alt = beam["beam_altitude_angles"]
            w = int(fmt.get("columns_per_frame", int(str(mode).split("x")[0])))
            h = int(fmt.get("pixels_per_column", len(alt)))
            prod = js.get("sensor_info", js).get("prod_line", "OS-1-%d" % h)
            info = cls(w, h, None, prod_line=prod, mode=mode)
            prof = fmt.get("udp_profile_lidar", cfg.get("udp_profile_lidar", "LEGACY"))
            info.format = _Format(w, h, int(fmt.get("columns_per_packet", 16)),
                                  _PROFILES[str(prof).replace("PROFILE_LIDAR_", "")] if not isinstance(prof, int) else prof,
                                  fmt.get("pixel_shift_by_row"))
            info.beam_altitude_angles = np.asarray(alt, dtype=np.float64)
            info.beam_azimuth_angles = np.asarray(beam["beam_azimuth_angles"], dtype=np.float64)
            n = beam.get("lidar_origin_to_beam_origin_mm")
            if n is None and "beam_to_lidar_transform" in beam:
                n = beam["beam_to_lidar_transform"][3]
            info.lidar_origin_to_beam_origin_mm = float(n or 0.0)
            l2s = js.get("lidar_intrinsics", js).get("lidar_to_sensor_transform", _DEFAULT_LIDAR_TO_SENSOR)
            info.lidar_to_sensor_transform = np.asarray(l2s, dtype=np.float64).reshape(4, 4)
            info.udp_port_lidar = int(cfg.get("udp_port_lidar", 7502))
            info.udp_port_imu = int(cfg.get("udp_port_imu", 7503))
            return info

    class LidarScan:
        def __init__(self, h, w, range_mm=None, timestamp=None):
            self.h, self.w = h, w
            self._range = np.zeros((h, w), dtype=np.uint32) if range_mm is None else range_mm
            self.timestamp = np.zeros(w, dtype=np.int64) if timestamp is None else timestamp
            self.status = np.ones(w, dtype=np.uint32)

        def field(self, f):
            if f != ChanField.RANGE:
                raise KeyError(f)
            return self._range

    class XYZLut:
        """range image (mm) -> (H, W, 3) float64 metres.  Like the SDK's LUT it is a per-pixel `direction` and
        `offset` table with the transforms folded in at construction: xyz = direction * (range * range_unit) + offset,
        zero where range == 0.

        From real metadata (beam angles) the tables follow the sensor documentation's range-to-XYZ formula on the
        STAGGERED image (column = measurement id): with encoder = 2 pi (1 - m / W), azimuth = -beam azimuth,
        phi = beam altitude and n = lidar_origin_to_beam_origin_mm,
            direction = (cos(encoder + azimuth) cos phi, sin(encoder + azimuth) cos phi, sin phi)
            offset    = n ((cos encoder, sin encoder, 0) - direction)
        both moved into the sensor frame by lidar_to_sensor_transform (and the extrinsic, if asked for), direction
        scaled to metres per millimetre (so range_unit = 1), offset to metres [UPSTREAM-UNVERIFIED: ouster-sdk is
        absent; restated from the published formula]."""
        range_unit = 0.001

        def __init__(self, metadata, use_extrinsics=False):
            if getattr(metadata, "directions", None) is not None:
                d = np.ascontiguousarray(metadata.directions, dtype=np.float64)
                self.offset = None
                if use_extrinsics and not np.array_equal(metadata.extrinsic, np.eye(4)):
                    E = metadata.extrinsic
                    d = np.ascontiguousarray(d @ E[:3, :3].T)
                    self.offset = np.ascontiguousarray(np.broadcast_to(E[:3, 3], d.shape))
                self.direction = d
                return
            W, H = metadata.format.columns_per_frame, metadata.format.pixels_per_column
            enc = 2.0 * np.pi * (1.0 - np.arange(W, dtype=np.float64) / W)[None, :]             # (1, W)
            azi = (-np.pi / 180.0) * metadata.beam_azimuth_angles[:, None]                      # (H, 1)
            phi = (np.pi / 180.0) * metadata.beam_altitude_angles[:, None]
            d = np.stack([np.cos(enc + azi) * np.cos(phi), np.sin(enc + azi) * np.cos(phi),
                          np.broadcast_to(np.sin(phi), (H, W))], axis=-1)
            n = metadata.lidar_origin_to_beam_origin_mm
            o = n * (np.stack([np.broadcast_to(np.cos(enc), (H, W)), np.broadcast_to(np.sin(enc), (H, W)),
                               np.zeros((H, W))], axis=-1) - d)
            T = metadata.lidar_to_sensor_transform
            if use_extrinsics:
                E = np.array(metadata.extrinsic, dtype=np.float64)
                E[:3, 3] *= 1000.0                                                              # metres -> mm, as T
                T = E @ T
            self.direction = np.ascontiguousarray((d @ T[:3, :3].T) * 0.001)
            self.offset = np.ascontiguousarray((o @ T[:3, :3].T + T[:3, 3]) * 0.001)
            self.range_unit = 1.0

        def __call__(self, scan):
            rng = scan.field(ChanField.RANGE) if hasattr(scan, "field") else scan
            r = rng.astype(np.float64) * self.range_unit
            xyz = self.direction * r[..., None]
            if self.offset is not None:
                xyz = np.where(rng[..., None] != 0, xyz + self.offset, 0.0)
            return xyz

    def last_valid_column_ts(scan):
        valid = np.flatnonzero(scan.status & 1)
        return int(scan.timestamp[valid[-1]]) if valid.size else 0


def sensor_info_from_synth(sensor, directions):
    """SensorInfo stand-in for a ptudes_lab_b200.synth.SensorModel."""
    if HAVE_OUSTER:  # pragma: no cover
        raise RuntimeError("synthetic SensorInfo is only for environments without ouster-sdk")
    return SensorInfo(sensor.W, sensor.H, directions, mode=f"{sensor.W}x10")


def scan_from_synth(synth_scan):
    h, w = synth_scan.range_mm.shape
    return LidarScan(h, w, synth_scan.range_mm, synth_scan.timestamp_ns)
