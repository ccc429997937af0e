"""Consumer side of the odometry step: the error-state EKF that `ptudes ekf-bench` feeds with the
poses of KissICPWrapper (reference: src/ptudes/ins/{es_ekf,data}.py).  Host NumPy, as in the
reference - the filter is O(18^2) work per sample and is not part of the GPU path."""
from .data import GRAV, IMU, NavState, calc_ate, calc_ate_from_navs  # noqa: F401
from .es_ekf import ESEKF  # noqa: F401
from .es_ekf_native import ESEKFNative  # noqa: F401
