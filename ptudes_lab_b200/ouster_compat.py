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
        def __init__(self, w, h):
            self.columns_per_frame = w
            self.pixels_per_column = h

    class SensorInfo:
        """Minimal metadata: frame geometry + the direction LUT of the synthetic sensor."""

        def __init__(self, columns_per_frame, pixels_per_column, directions, prod_line="OS-0-128",
                     mode="1024x10", extrinsic=None):
            self.format = _Format(columns_per_frame, pixels_per_column)
            self.directions = np.asarray(directions, dtype=np.float64)
            self.prod_line = prod_line
            self.mode = mode
            self.extrinsic = np.eye(4) if extrinsic is None else np.asarray(extrinsic, dtype=np.float64)

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
        """range image (mm) -> (H, W, 3) float64 metres.  Like the SDK's LUT it is a per-pixel
        `direction` and `offset` table with the extrinsic folded in at construction:
        xyz = direction * (range * range_unit) + offset, zero where range == 0."""
        range_unit = 0.001

        def __init__(self, metadata, use_extrinsics=False):
            d = np.ascontiguousarray(metadata.directions, dtype=np.float64)
            self.offset = None
            if use_extrinsics and not np.array_equal(metadata.extrinsic, np.eye(4)):
                E = metadata.extrinsic
                d = np.ascontiguousarray(d @ E[:3, :3].T)
                self.offset = np.ascontiguousarray(np.broadcast_to(E[:3, 3], d.shape))
            self.direction = d

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
