"""What the canonical ordering rules cost against upstream's own order.

Runs the NumPy oracle twice over the same synthetic sequence: order="index" (the canonical rules every
implementation in this repo follows: downsample output in ascending input index, prune erases every far
voxel) and order="robin_map" (emulation of the tsl::robin_map iteration order kiss-icp 0.2.x really
produces: hash of SURVEY A.5, power-of-two buckets, reserve(frame.size()), robin-hood layout,
erase-while-iterating skip).  The first-grid selection SETS are identical by construction; what differs
is the order, hence which point is "first" in the second grid and in every map voxel (kiss.py:96,129).

  python profiles/r2_order_delta.py [config] [scans] [hash multiplier y]   ->  profiles/r2_order_delta[_hashyN].json

TEST INFRASTRUCTURE (imports oracle/).  The emulation is still a restatement from memory of the public
upstream sources, not a run of kiss-icp: the numbers say how sensitive the odometry is to the order, and
that the 1e-5 m bar of BASELINE.json cannot be met against upstream by ANY implementation that does not
reproduce upstream's hash-table iteration order bit for bit."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import canon, kiss_oracle as ko          # noqa: E402
from ptudes_lab_b200 import synth                    # noqa: E402
from ptudes_lab_b200.ins.data import calc_ate        # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "os0_quad"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    if len(sys.argv) > 3:       # the other candidate for upstream's second hash multiplier (see oracle/kiss_oracle.py)
        ko.UPSTREAM_HASH_Y = int(sys.argv[3])
    seq = synth.make_sequence(name, 0)
    a = ko.OracleKissICPWrapper(order="index")
    b = ko.OracleKissICPWrapper(order="robin_map")
    rows = []
    gts = []
    t0 = time.time()
    for k in range(n):
        xyz, ts, tsec, gt = seq.points(k)
        gts.append(np.array(gt, dtype=np.float64))
        pa = a.register_points(xyz, ts, tsec).copy()
        pb = b.register_points(xyz, ts, tsec).copy()
        d = canon.rigid_mul(canon.rigid_inv(pa), pb)
        row = {"abs_dt": float(np.linalg.norm(d[:3, 3])), "abs_drot": float(abs(canon.so3_log(d[:3, :3])[1]))}
        if k > 0:       # relative motion of this scan, canonical vs emulated
            ra = canon.rigid_mul(canon.rigid_inv(a.poses[-2]), a.poses[-1])
            rb = canon.rigid_mul(canon.rigid_inv(b.poses[-2]), b.poses[-1])
            dr = canon.rigid_mul(canon.rigid_inv(ra), rb)
            row["rel_dt"] = float(np.linalg.norm(dr[:3, 3]))
            row["rel_drot"] = float(abs(canon.so3_log(dr[:3, :3])[1]))
        row.update(iters=(a.last_stats["iterations"], b.last_stats["iterations"]),
                   n_ds=(a.last_counts["n_ds"], b.last_counts["n_ds"]),
                   n_src=(a.last_counts["n_src"], b.last_counts["n_src"]),
                   n_vox=(a.last_counts["n_vox"], b.last_counts["n_vox"]),
                   prune_skips=b._kiss.local_map.skipped_last_prune)
        rows.append(row)
        if k % 10 == 0:
            print(k, row, f"{time.time() - t0:.0f}s", flush=True)
    ate_rot, ate_trans = calc_ate(a.poses, b.poses)
    gt_a = calc_ate(a.poses, gts)
    gt_b = calc_ate(b.poses, gts)
    out = {
        "config": name, "scans": n,
        "same_first_grid_sizes": all(r["n_ds"][0] == r["n_ds"][1] for r in rows),
        "max_abs_dt_m": max(r["abs_dt"] for r in rows), "max_abs_drot_rad": max(r["abs_drot"] for r in rows),
        "max_rel_dt_m": max(r.get("rel_dt", 0.0) for r in rows), "max_rel_drot_rad": max(r.get("rel_drot", 0.0) for r in rows),
        "rms_rel_dt_m": float(np.sqrt(np.mean([r.get("rel_dt", 0.0) ** 2 for r in rows]))),
        "ate_between_orders": {"rot_deg2": float(ate_rot), "trans_m2": float(ate_trans),
                               "definition": "ins/data.py:124-153 calc_ate: mean SQUARED error, first pose aligned"},
        "ate_vs_ground_truth": {"index": gt_a, "robin_map": gt_b},
        "prune_skips_total": int(sum(r["prune_skips"] for r in rows)),
        "per_scan": rows,
    }
    out["upstream_hash_y"] = ko.UPSTREAM_HASH_Y
    suffix = "" if ko.UPSTREAM_HASH_Y == 19349663 else f"_hashy{ko.UPSTREAM_HASH_Y}"
    with open(os.path.join(ROOT, "profiles", f"r2_order_delta{suffix}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: v for k, v in out.items() if k != "per_scan"})


if __name__ == "__main__":
    main()
