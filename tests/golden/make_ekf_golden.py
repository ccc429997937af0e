"""Generate tests/golden/ekf_ref.npz by RUNNING THE REFERENCE's own ESEKF and calc_ate
(/root/reference/src/ptudes/ins/{es_ekf,data}.py) in this container, where /root/reference exists.

The reference modules do not import on Python 3.12 as they are (dataclass fields with ndarray
defaults; `ouster` absent), so they are loaded from their source files with three shims that do
not touch the filter's arithmetic:
  * `dataclass` is wrapped so that ndarray defaults become default_factory copies;
  * `ouster.client` is an empty stand-in (only used for type annotations / packet parsing);
  * `ouster.sdk.pose_util.exp_rot_vec / log_rot_mat` are SciPy's Rotation.from_rotvec().as_matrix() /
    Rotation.from_matrix().as_rotvec() - the standard SO(3) exp/log the SDK functions implement;
  * `ptudes.utils.vee` is restated (a 3x3 cross-product matrix; the real module imports the viz stack).
The vectors pin ptudes_lab_b200/ins against the reference's behaviour; tests/test_ekf.py compares
with a 1e-9 tolerance (the reference round-trips the attitude through SciPy quaternions every step,
the restatement keeps a rotation matrix, so the last bits differ).

    python tests/golden/make_ekf_golden.py
"""
import dataclasses
import os
import sys
import types

import numpy as np
from scipy.spatial.transform import Rotation

REF = "/root/reference/src/ptudes"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ekf_ref.npz")


def _patched_dataclass(cls=None, **kw):
    def wrap(c):
        for name, val in list(vars(c).items()):
            if isinstance(val, np.ndarray):
                setattr(c, name, dataclasses.field(default_factory=lambda v=val: v.copy()))
        return dataclasses.dataclass(c, **kw)
    return wrap if cls is None else wrap(cls)


def load_reference():
    ouster = types.ModuleType("ouster")
    client = types.ModuleType("ouster.client")
    client.ImuPacket = client.LidarScan = client.SensorInfo = object
    client.ChanField = types.SimpleNamespace(RANGE="RANGE")
    sdk = types.ModuleType("ouster.sdk")
    pu = types.ModuleType("ouster.sdk.pose_util")
    pu.exp_rot_vec = lambda v: Rotation.from_rotvec(np.asarray(v, dtype=float)).as_matrix()
    pu.log_rot_mat = lambda m: Rotation.from_matrix(np.asarray(m, dtype=float)).as_rotvec()
    ouster.client, ouster.sdk, sdk.pose_util = client, sdk, pu
    sys.modules.update({"ouster": ouster, "ouster.client": client, "ouster.sdk": sdk, "ouster.sdk.pose_util": pu})

    ptudes = types.ModuleType("ptudes")
    ins = types.ModuleType("ptudes.ins")
    utils = types.ModuleType("ptudes.utils")

    def vee(v):
        return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
    utils.vee = vee
    sys.modules.update({"ptudes": ptudes, "ptudes.ins": ins, "ptudes.utils": utils})

    def load(modname, path):
        src = open(path).read().replace("from dataclasses import dataclass", "from _ptk_shim import dataclass")
        mod = types.ModuleType(modname)
        mod.__file__ = path
        sys.modules[modname] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        return mod
    shim = types.ModuleType("_ptk_shim")
    shim.dataclass = _patched_dataclass
    sys.modules["_ptk_shim"] = shim
    data = load("ptudes.ins.data", os.path.join(REF, "ins", "data.py"))
    ekf = load("ptudes.ins.es_ekf", os.path.join(REF, "ins", "es_ekf.py"))
    return data, ekf


def make_inputs(seed=7, n_imu=400, pose_every=10):
    """A seeded IMU stream (sim_imu-like: piecewise-constant motion + noise + bias, cli/ekf_bench.py:44-79)
    and noisy pose measurements every `pose_every` samples."""
    rng = np.random.default_rng(seed)
    grav = 9.782940329221166 * np.array([0.0, 0.0, -1.0])
    acc_bias = np.array([0.09, -0.02, -0.04])
    gyr_bias = np.array([0.01, 0.03, -0.012])
    lacc, avel, ts = [], [], []
    acc = gyr = None
    for i in range(n_imu):
        if i % 10 == 0:
            acc = rng.normal(0.0, 0.5, 3) - grav
            gyr = rng.normal(0.0, 0.3, 3)
        lacc.append(acc + rng.normal(0, 0.05, 3) + acc_bias)
        avel.append(gyr + rng.normal(0, 0.02, 3) + gyr_bias)
        ts.append(0.01 * i)
    poses = []
    for j in range(n_imu // pose_every):
        T = np.eye(4)
        T[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.2, 3)).as_matrix()
        T[:3, 3] = rng.normal(0, 1.0, 3)
        poses.append(T)
    return np.array(lacc), np.array(avel), np.array(ts), np.array(poses)


def main():
    data, ekfm = load_reference()
    lacc, avel, ts, poses = make_inputs()
    pose_every = len(ts) // len(poses)
    out = {"lacc": lacc, "avel": avel, "ts": ts, "poses": poses, "pose_every": np.array(pose_every)}
    for tag, kw in (("a", {}), ("b", {"init_bacc": np.array([0.05, 0.0, -0.02]), "init_bgyr": np.array([0.0, 0.02, 0.0])})):
        f = ekfm.ESEKF(**kw)
        rec = {k: [] for k in ("pos", "vel", "att", "bg", "ba", "grav", "cov", "ts")}
        meas_cov = None if tag == "a" else np.diag([0.05 ** 2] * 3 + [0.02 ** 2] * 3)
        for i in range(len(ts)):
            f.processImu(data.IMU(lacc[i].copy(), avel[i].copy(), float(ts[i])))
            if i % pose_every == pose_every - 1:
                # measurement = a pose near the current estimate (so the update is in its linear regime)
                T = f.nav.pose_mat() @ poses[i // pose_every] * 1.0
                T[:3, :3] = f.nav.att_h @ Rotation.from_rotvec(0.05 * Rotation.from_matrix(poses[i // pose_every][:3, :3]).as_rotvec()).as_matrix()
                T[:3, 3] = f.nav.pos + 0.05 * poses[i // pose_every][:3, 3]
                out.setdefault(f"{tag}_meas", []).append(T.copy())
                f.processPose(T, meas_cov)
            n = f.nav
            rec["pos"].append(n.pos.copy()); rec["vel"].append(n.vel.copy()); rec["att"].append(n.att_h.copy())
            rec["bg"].append(n.bias_gyr.copy()); rec["ba"].append(n.bias_acc.copy()); rec["grav"].append(n.grav.copy())
            rec["ts"].append(f.ts)
            if i % pose_every == pose_every - 1:     # covariance after every update only (file size)
                rec["cov"].append(f._cov.copy())
        for k, v in rec.items():
            out[f"{tag}_{k}"] = np.array(v)
        out[f"{tag}_meas"] = np.array(out[f"{tag}_meas"])
        out[f"{tag}_cov0"] = f._cov_init.copy()
    # calc_ate of the reference on two trajectories
    rng = np.random.default_rng(3)
    A, G = [], []
    for k in range(30):
        T = np.eye(4); T[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.5, 3)).as_matrix(); T[:3, 3] = rng.normal(0, 5, 3)
        E = np.eye(4); E[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.02, 3)).as_matrix(); E[:3, 3] = rng.normal(0, 0.1, 3)
        G.append(T); A.append(T @ E)
    out["ate_A"], out["ate_G"] = np.array(A), np.array(G)
    out["ate_ref"] = np.array(data.calc_ate(A, G))
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
