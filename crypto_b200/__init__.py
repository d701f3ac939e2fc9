"""crypto_b200: B200-native BLS12-381 MSM / fixed-base / multi-pairing backend for the
docknetwork/crypto hot path.  Host-side mirror of the reference interface lives in
crypto_b200.msm / crypto_b200.pairing_check; kernels in crypto_b200/csrc; C ABI in include/dockgpu.h."""
from . import lib  # noqa: F401
