"""ctypes loader for oracle/libcpuref.so (TEST INFRASTRUCTURE ONLY -- see cpu_ref.c).

Byte-level interface identical to the product's C ABI (include/dockgpu.h) so the
same seeded buffers can be fed to both sides of a parity test."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, 'libcpuref.so')
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ref_final_exp.restype = C.c_int
        _LIB.ref_msm_window_size.restype = C.c_int
        _LIB.ref_msm_window_size.argtypes = [C.c_size_t]
        _LIB.ref_fixed_base_window_size.restype = C.c_int
        _LIB.ref_fixed_base_window_size.argtypes = [C.c_size_t]
    return _LIB


def set_threads(n):
    """OpenMP threads of the CPU legs (torchrun exports OMP_NUM_THREADS=1; bench.py's CPU arm wants every core)."""
    lib()
    try:
        C.CDLL('libgomp.so.1').omp_set_num_threads(C.c_int(int(n)))
    except OSError:
        pass


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def _as_np(b):
    return np.frombuffer(b, dtype=np.uint8) if isinstance(b, (bytes, bytearray)) else np.asarray(b, dtype=np.uint8)


def msm_g1(bases, scalars, n=None):
    bases, scalars = _as_np(bases), _as_np(scalars)
    if n is None:
        n = min(bases.size // 96, scalars.size // 32)
    b, bp = _u8(bases); s, sp = _u8(scalars)
    out = np.zeros(144, np.uint8)
    lib().ref_msm_g1(bp, sp, C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return out


def msm_g2(bases, scalars, n=None):
    bases, scalars = _as_np(bases), _as_np(scalars)
    if n is None:
        n = min(bases.size // 192, scalars.size // 32)
    b, bp = _u8(bases); s, sp = _u8(scalars)
    out = np.zeros(288, np.uint8)
    lib().ref_msm_g2(bp, sp, C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return out


def normalize_batch_g1(jac):
    j, jp = _u8(_as_np(jac)); n = j.size // 144
    out = np.zeros(96 * n, np.uint8)
    lib().ref_normalize_batch_g1(jp, C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return out


def normalize_batch_g2(jac):
    j, jp = _u8(_as_np(jac)); n = j.size // 288
    out = np.zeros(192 * n, np.uint8)
    lib().ref_normalize_batch_g2(jp, C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return out


def batch_mul_g1(points, scalars):
    p, pp = _u8(_as_np(points)); s, sp = _u8(_as_np(scalars)); m = s.size // 32
    out = np.zeros(144 * m, np.uint8)
    lib().ref_batch_mul_g1(pp, sp, C.c_size_t(m), out.ctypes.data_as(C.c_void_p))
    return out


def g1_deserialize_compressed(data, validate=True):
    """-> (affine records, status bytes); CPU restatement of deserialize_compressed for Vec<G1Affine>."""
    d, dp = _u8(_as_np(data)); n = d.size // 48
    out = np.zeros(96 * n, np.uint8); st = np.zeros(max(n, 1), np.uint8)
    lib().ref_g1_deserialize_compressed(dp, C.c_size_t(n), C.c_int(1 if validate else 0), out.ctypes.data_as(C.c_void_p),
                                        st.ctypes.data_as(C.c_void_p))
    return out, st[:n]


def g2_deserialize_compressed(data, validate=True):
    d, dp = _u8(_as_np(data)); n = d.size // 96
    out = np.zeros(192 * n, np.uint8); st = np.zeros(max(n, 1), np.uint8)
    lib().ref_g2_deserialize_compressed(dp, C.c_size_t(n), C.c_int(1 if validate else 0), out.ctypes.data_as(C.c_void_p),
                                        st.ctypes.data_as(C.c_void_p))
    return out, st[:n]


def batch_mul_g2(points, scalars):
    p, pp = _u8(_as_np(points)); s, sp = _u8(_as_np(scalars)); m = s.size // 32
    out = np.zeros(288 * m, np.uint8)
    lib().ref_batch_mul_g2(pp, sp, C.c_size_t(m), out.ctypes.data_as(C.c_void_p))
    return out


def fixed_base_mul_many_g1(point, hint_n, scalars):
    p, pp = _u8(_as_np(point)); s, sp = _u8(_as_np(scalars)); m = s.size // 32
    out = np.zeros(144 * m, np.uint8)
    w, nw = C.c_int(0), C.c_int(0)
    lib().ref_fixed_base_mul_many_g1(pp, C.c_size_t(hint_n), sp, C.c_size_t(m), out.ctypes.data_as(C.c_void_p),
                                     C.byref(w), C.byref(nw))
    return out, w.value, nw.value


def fixed_base_mul_many_g2(point, hint_n, scalars):
    p, pp = _u8(_as_np(point)); s, sp = _u8(_as_np(scalars)); m = s.size // 32
    out = np.zeros(288 * m, np.uint8)
    w, nw = C.c_int(0), C.c_int(0)
    lib().ref_fixed_base_mul_many_g2(pp, C.c_size_t(hint_n), sp, C.c_size_t(m), out.ctypes.data_as(C.c_void_p),
                                     C.byref(w), C.byref(nw))
    return out, w.value, nw.value


def fixed_base_table_g1(point, window):
    p, pp = _u8(_as_np(point))
    nw = (255 + window - 1) // window
    out = np.zeros(96 * (nw << window), np.uint8)
    lib().ref_fixed_base_table_g1(pp, C.c_int(window), out.ctypes.data_as(C.c_void_p))
    return out


def g1_generator_muls(scalars):
    s, sp = _u8(_as_np(scalars)); m = s.size // 32
    out = np.zeros(96 * m, np.uint8)
    lib().ref_g1_generator_muls(sp, C.c_size_t(m), out.ctypes.data_as(C.c_void_p))
    return out


def g2_generator_muls(scalars):
    s, sp = _u8(_as_np(scalars)); m = s.size // 32
    out = np.zeros(192 * m, np.uint8)
    lib().ref_g2_generator_muls(sp, C.c_size_t(m), out.ctypes.data_as(C.c_void_p))
    return out


def multi_miller_loop(g1s, g2s):
    a, ap = _u8(_as_np(g1s)); b, bp = _u8(_as_np(g2s)); k = min(a.size // 96, b.size // 192)
    out = np.zeros(576, np.uint8)
    lib().ref_multi_miller_loop(ap, bp, C.c_size_t(k), out.ctypes.data_as(C.c_void_p))
    return out


def final_exp(f):
    a, ap = _u8(_as_np(f))
    out = np.zeros(576, np.uint8)
    ok = lib().ref_final_exp(ap, out.ctypes.data_as(C.c_void_p))
    return out if ok else None


def multi_pairing(g1s, g2s):
    return final_exp(multi_miller_loop(g1s, g2s))


def fp12_mul(a, b):
    x, xp = _u8(_as_np(a)); y, yp = _u8(_as_np(b))
    out = np.zeros(576, np.uint8)
    lib().ref_fp12_mul(xp, yp, out.ctypes.data_as(C.c_void_p))
    return out


def fp12_pow(a, scalar):
    x, xp = _u8(_as_np(a)); s, sp = _u8(_as_np(scalar))
    out = np.zeros(576, np.uint8)
    lib().ref_fp12_pow(xp, sp, out.ctypes.data_as(C.c_void_p))
    return out


def fp12_one():
    out = np.zeros(576, np.uint8)
    lib().ref_fp12_one(out.ctypes.data_as(C.c_void_p))
    return out


def scalar_dot_mod_r(a, b):
    """(sum_i a_i * b_i) mod r for two arrays of n x 32 B canonical scalars -> Python int."""
    x, xp = _u8(_as_np(a)); y, yp = _u8(_as_np(b))
    n = min(x.size, y.size) // 32
    out = np.zeros(72, np.uint8)
    lib().ref_scalar_dot_wide(xp, yp, C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return int.from_bytes(bytes(out), 'little') % 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def random_scalars(n, seed):
    """Seeded uniform scalars in [0, r) as n x 32 B canonical LE (numpy, vectorised
    rejection sampling; deterministic for a given (n, seed))."""
    r_mod = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros((n, 4), dtype=np.uint64)
    need = np.arange(n)
    r_l = np.array([(r_mod >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)
    while need.size:
        v = rng.integers(0, 2**64, size=(need.size, 4), dtype=np.uint64)
        v[:, 3] &= np.uint64((1 << 63) - 1)
        lt = np.zeros(need.size, bool); decided = np.zeros(need.size, bool)
        for i in (3, 2, 1, 0):
            lt |= (~decided) & (v[:, i] < r_l[i])
            decided |= v[:, i] != r_l[i]
        out[need[lt]] = v[lt]
        need = need[~lt]
    return out.view(np.uint8).reshape(-1)


def fr_ntt(data, logn, inverse=False, coset=False):
    """ark_poly Radix2EvaluationDomain fft / ifft (+ coset, offset 7) on Montgomery Fr records."""
    a = np.array(_as_np(data), dtype=np.uint8, copy=True)
    lib().ref_fr_ntt(a.ctypes.data_as(C.c_void_p), C.c_uint32(logn), C.c_int(1 if inverse else 0), C.c_int(1 if coset else 0))
    return a


def qap_h_from_abc(a, b, c, logn):
    x, xp = _u8(_as_np(a)); y, yp = _u8(_as_np(b)); z, zp = _u8(_as_np(c))
    out = np.zeros(32 << logn, np.uint8)
    lib().ref_qap_h_from_abc(xp, yp, zp, C.c_uint32(logn), out.ctypes.data_as(C.c_void_p))
    return out


def fr_spmv(row_ptr, col, coeff_mont, w_mont):
    rp = np.ascontiguousarray(row_ptr, dtype=np.uint32); cl = np.ascontiguousarray(col, dtype=np.uint32)
    co, cop = _u8(_as_np(coeff_mont)); w, wp = _u8(_as_np(w_mont))
    rows = rp.size - 1
    out = np.zeros(32 * rows, np.uint8)
    lib().ref_fr_spmv(rp.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p), cop, C.c_size_t(rows), wp, out.ctypes.data_as(C.c_void_p))
    return out
