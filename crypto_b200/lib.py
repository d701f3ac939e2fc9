"""ctypes binding of libdockgpu.so (include/dockgpu.h).

There is NO CPU fallback: if the CUDA library is missing or no B200 is visible, every entry
point raises.  The oracle under oracle/ is never imported from here.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DOCKGPU_LIB') or os.path.join(HERE, 'libdockgpu.so')      # DOCKGPU_LIB: A/B builds (tools/ab_bench.py)

G1_AFF, G1_JAC, G2_AFF, G2_JAC, FP12, SCALAR = 96, 144, 192, 288, 576, 32

# every symbol include/dockgpu.h declares (checked by tests/test_abi.py without a GPU)
EXPORTS = [
    'dg_init', 'dg_init_devices', 'dg_device_count', 'dg_shutdown', 'dg_last_error', 'dg_launch_count', 'dg_sync', 'dg_stream_status',
    'dg_bases_upload_g1', 'dg_bases_upload_g2', 'dg_bases_free', 'dg_bases_precompute',
    'dg_bases_upload_g1_sharded', 'dg_bases_upload_g2_sharded',
    'dg_msm_g1_sharded', 'dg_msm_g2_sharded', 'dg_msm_unchecked_g1_sharded',
    'dg_msm_g1_handle_device', 'dg_msm_g2_handle_device',
    'dg_msm_unchecked_g1', 'dg_msm_unchecked_g2', 'dg_fr_into_bigint',
    'dg_msm_g1', 'dg_msm_g2', 'dg_msm_g1_device', 'dg_msm_g2_device', 'dg_msm_set_window',
    'dg_msm_set_affine_rounds', 'dg_msm_plan',
    'dg_fixed_base_table_g1', 'dg_fixed_base_table_g2', 'dg_fixed_base_table_info',
    'dg_fixed_base_table_download', 'dg_fixed_base_table_free',
    'dg_fixed_base_mul_many_g1', 'dg_fixed_base_mul_many_g2',
    'dg_fixed_base_mul_many_normalized_g1', 'dg_fixed_base_mul_many_normalized_g2',
    'dg_batch_mul_g1', 'dg_batch_mul_g2', 'dg_batch_mul_add_fixed_g1', 'dg_batch_mul_add_same_g1', 'dg_compress_g1', 'dg_compress_g2',
    'dg_normalize_batch_g1', 'dg_normalize_batch_g2',
    'dg_multi_miller_loop', 'dg_final_exponentiation', 'dg_multi_pairing', 'dg_multi_pairing_is_one', 'dg_multi_pairing_batch',
    'dg_gt_pow', 'dg_fp12_mul',
    'dg_fold_g1', 'dg_fold_g1_device', 'dg_fold_g2',
    'dg_fr_ntt', 'dg_fr_ntt_device', 'dg_qap_h_from_abc', 'dg_fr_spmv',
    'dg_r1cs_upload', 'dg_r1cs_free', 'dg_groth16_prove_msms',
    'dg_g1_serialize', 'dg_g2_serialize', 'dg_g1_deserialize', 'dg_g2_deserialize',
    'dg_prof_enable', 'dg_prof_read_accumulate',
    'dg_dbg_fp_op', 'dg_dbg_fr_op', 'dg_dbg_set_tunable',
]


class DockGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libdockgpu error %d: %s' % (code, msg))
        self.code = code


_lib = None
_inited = False


def load():
    """dlopen the library without touching the GPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DockGpuError(-4, 'libdockgpu.so is not built (run `python -m crypto_b200.build`); '
                                   'there is no CPU fallback')
        _lib = C.CDLL(LIB_PATH)
        _lib.dg_launch_count.restype = C.c_uint64
        for name in EXPORTS:
            if name != 'dg_launch_count' and hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_int32
    return _lib


def _check(rc):
    if rc != 0:
        buf = C.create_string_buffer(512)
        _lib.dg_last_error(buf, C.c_size_t(512))
        raise DockGpuError(rc, buf.value.decode(errors='replace'))


def init(device=-1):
    global _inited
    lib = load()
    if not _inited:
        _check(lib.dg_init(C.c_int32(device)))
        _inited = True
    return lib


def init_devices(devices):
    """One process driving several GPUs (dg_init_devices); devices[0] is the primary device."""
    global _inited
    lib = load()
    devs = (C.c_int32 * len(devices))(*devices)
    _check(lib.dg_init_devices(devs, C.c_int32(len(devices))))
    _inited = True
    return lib


def device_count():
    n = C.c_int32(0)
    _check(init().dg_device_count(C.byref(n)))
    return n.value


def shutdown():
    global _inited
    _check(load().dg_shutdown())
    _inited = False


def stream_status(stream=0):
    """Synchronises the stream and raises DockGpuError if the last MSM queued by this thread saw a scalar >= r."""
    _check(init().dg_stream_status(C.c_void_p(stream)))


def launch_count():
    return int(load().dg_launch_count())


def sync():
    _check(init().dg_sync())


def _in(a, rec=None):
    """Contiguous uint8 view of bytes / ndarray input; returns (keepalive, void*)."""
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return a, C.c_void_p(a.ctypes.data)


def _out(nbytes):
    a = np.zeros(max(nbytes, 1), dtype=np.uint8)
    return a, C.c_void_p(a.ctypes.data)


# ---- resident bases -----------------------------------------------------------------------------
class Bases:
    """Device-resident affine bases (proving-key model)."""

    def __init__(self, affine, g2=False):
        lib = init()
        a, ap = _in(affine)
        self.g2 = g2
        self.n = a.size // (G2_AFF if g2 else G1_AFF)
        h = C.c_uint64(0)
        fn = lib.dg_bases_upload_g2 if g2 else lib.dg_bases_upload_g1
        _check(fn(ap, C.c_size_t(self.n), C.byref(h)))
        self.handle = h.value

    def precompute(self, window_bits=0):
        """Build the 2^(c*k) multiples table on the device (dg_bases_precompute)."""
        _check(load().dg_bases_precompute(C.c_uint64(self.handle), C.c_int32(window_bits)))
        return self

    def free(self):
        if self.handle:
            _check(load().dg_bases_free(C.c_uint64(self.handle)))
            self.handle = 0


class ShardedBases(Bases):
    """Bases split by contiguous ranges over every device of init_devices (dg_bases_upload_*_sharded)."""

    def __init__(self, affine, g2=False):
        lib = init()
        a, ap = _in(affine)
        self.g2 = g2
        self.n = a.size // (G2_AFF if g2 else G1_AFF)
        h = C.c_uint64(0)
        fn = lib.dg_bases_upload_g2_sharded if g2 else lib.dg_bases_upload_g1_sharded
        _check(fn(ap, C.c_size_t(self.n), C.byref(h)))
        self.handle = h.value


def msm_sharded(bases, scalars, g2=False, n=None):
    """dg_msm_*_sharded: one call, every device; `bases` is a ShardedBases handle or affine records."""
    lib = init()
    s, sp = _in(scalars)
    ns = s.size // SCALAR
    fn = lib.dg_msm_g2_sharded if g2 else lib.dg_msm_g1_sharded
    o, op = _out(G2_JAC if g2 else G1_JAC)
    if isinstance(bases, ShardedBases):
        assert bases.g2 == g2
        k = min(ns, bases.n) if n is None else n
        _check(fn(C.c_uint64(bases.handle), None, sp, C.c_size_t(k), op))
    else:
        b, bp = _in(bases)
        k = min(ns, b.size // (G2_AFF if g2 else G1_AFF)) if n is None else n
        _check(fn(C.c_uint64(0), bp if k else None, sp, C.c_size_t(k), op))
    return o[:G2_JAC if g2 else G1_JAC]


def msm_handle_device(bases, scalars_ptr, n, out_ptr, stream=0):
    """Resident bases (Bases handle), scalars/output device pointers; no copies, no sync."""
    lib = init()
    fn = lib.dg_msm_g2_handle_device if bases.g2 else lib.dg_msm_g1_handle_device
    _check(fn(C.c_uint64(bases.handle), C.c_void_p(scalars_ptr), C.c_size_t(n), C.c_void_p(out_ptr), C.c_void_p(stream)))


def msm(bases, scalars, g2=False, n=None):
    """sum s_i * P_i -> Jacobian record (144 / 288 B).  `bases` is a Bases handle or affine
    records; inputs are truncated to the shorter one like ark's msm_bigint."""
    lib = init()
    s, sp = _in(scalars)
    ns = s.size // SCALAR
    fn = lib.dg_msm_g2 if g2 else lib.dg_msm_g1
    o, op = _out(G2_JAC if g2 else G1_JAC)
    if isinstance(bases, Bases):
        assert bases.g2 == g2
        k = min(ns, bases.n) if n is None else n
        _check(fn(C.c_uint64(bases.handle), None, sp, C.c_size_t(k), op))
    else:
        b, bp = _in(bases)
        k = min(ns, b.size // (G2_AFF if g2 else G1_AFF)) if n is None else n
        _check(fn(C.c_uint64(0), bp if k else None, sp, C.c_size_t(k), op))
    return o[:G2_JAC if g2 else G1_JAC]


def msm_unchecked(bases, scalars_fr_mont, g2=False):
    """msm_unchecked semantics: scalars are Fr Montgomery records, converted on the device."""
    lib = init()
    s, sp = _in(scalars_fr_mont)
    ns = s.size // SCALAR
    fn = lib.dg_msm_unchecked_g2 if g2 else lib.dg_msm_unchecked_g1
    o, op = _out(G2_JAC if g2 else G1_JAC)
    if isinstance(bases, Bases):
        k = min(ns, bases.n)
        _check(fn(C.c_uint64(bases.handle), None, sp, C.c_size_t(k), op))
    else:
        b, bp = _in(bases)
        k = min(ns, b.size // (G2_AFF if g2 else G1_AFF))
        _check(fn(C.c_uint64(0), bp if k else None, sp, C.c_size_t(k), op))
    return o[:G2_JAC if g2 else G1_JAC]


def fr_into_bigint(fr_mont):
    lib = init()
    s, sp = _in(fr_mont)
    n = s.size // SCALAR
    o, op = _out(32 * n)
    _check(lib.dg_fr_into_bigint(sp, C.c_size_t(n), op))
    return o[:32 * n]


def msm_device(bases_ptr, scalars_ptr, n, out_ptr, stream=0, g2=False):
    """Device-pointer variant (no copies, no sync)."""
    lib = init()
    fn = lib.dg_msm_g2_device if g2 else lib.dg_msm_g1_device
    _check(fn(C.c_void_p(bases_ptr), C.c_void_p(scalars_ptr), C.c_size_t(n), C.c_void_p(out_ptr), C.c_void_p(stream)))


def prof_enable(on):
    _check(init().dg_prof_enable(C.c_int32(1 if on else 0)))


def prof_read_accumulate():
    """-> (mean ms of the k_accumulate launches since the last read, launch count)."""
    ms, cnt = C.c_double(0), C.c_int32(0)
    _check(init().dg_prof_read_accumulate(C.byref(ms), C.byref(cnt)))
    return ms.value, cnt.value


def msm_set_window(c):
    _check(init().dg_msm_set_window(C.c_int32(c)))


def dbg_set_tunable(i, v):
    _check(init().dg_dbg_set_tunable(C.c_int32(i), C.c_int32(v)))


def msm_set_affine_rounds(r):
    _check(init().dg_msm_set_affine_rounds(C.c_int32(r)))


def msm_plan(n, g2=False, precomputed_c=0):
    c, r = C.c_int32(0), C.c_int32(0)
    _check(init().dg_msm_plan(C.c_size_t(n), C.c_int32(1 if g2 else 0), C.c_int32(precomputed_c), C.byref(c), C.byref(r)))
    return c.value, r.value


# ---- fixed base ---------------------------------------------------------------------------------
class FixedBaseTable:
    def __init__(self, point_affine, hint_n, g2=False):
        lib = init()
        p, pp = _in(point_affine)
        self.g2 = g2
        h = C.c_uint64(0)
        fn = lib.dg_fixed_base_table_g2 if g2 else lib.dg_fixed_base_table_g1
        _check(fn(pp, C.c_size_t(hint_n), C.byref(h)))
        self.handle = h.value
        w, nw, is2 = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _check(lib.dg_fixed_base_table_info(C.c_uint64(self.handle), C.byref(w), C.byref(nw), C.byref(is2)))
        self.window, self.num_windows = w.value, nw.value

    def download(self):
        rec = G2_AFF if self.g2 else G1_AFF
        o, op = _out(rec * (self.num_windows << self.window))
        _check(load().dg_fixed_base_table_download(C.c_uint64(self.handle), op))
        return o

    def mul_many(self, scalars):
        lib = load()
        s, sp = _in(scalars)
        m = s.size // SCALAR
        rec = G2_JAC if self.g2 else G1_JAC
        o, op = _out(rec * m)
        fn = lib.dg_fixed_base_mul_many_g2 if self.g2 else lib.dg_fixed_base_mul_many_g1
        _check(fn(C.c_uint64(self.handle), sp, C.c_size_t(m), op))
        return o[:rec * m]

    def mul_many_normalized(self, scalars):
        """FixedBase::msm + normalize_batch fused on the device -> m affine records."""
        lib = load()
        s, sp = _in(scalars)
        m = s.size // SCALAR
        rec = G2_AFF if self.g2 else G1_AFF
        o, op = _out(rec * m)
        fn = lib.dg_fixed_base_mul_many_normalized_g2 if self.g2 else lib.dg_fixed_base_mul_many_normalized_g1
        _check(fn(C.c_uint64(self.handle), sp, C.c_size_t(m), op))
        return o[:rec * m]

    def free(self):
        if self.handle:
            _check(load().dg_fixed_base_table_free(C.c_uint64(self.handle)))
            self.handle = 0


def batch_mul(points, scalars, g2=False):
    lib = init()
    p, pp = _in(points); s, sp = _in(scalars)
    m = min(s.size // SCALAR, p.size // (G2_AFF if g2 else G1_AFF))
    rec = G2_JAC if g2 else G1_JAC
    o, op = _out(rec * m)
    fn = lib.dg_batch_mul_g2 if g2 else lib.dg_batch_mul_g1
    _check(fn(pp, sp, C.c_size_t(m), op))
    return o[:rec * m]


def batch_mul_add_fixed_g1(points, scalars_a, table, scalars_b):
    lib = init()
    p, pp = _in(points); a, ap = _in(scalars_a); b, bp = _in(scalars_b)
    m = a.size // SCALAR
    if a.size % SCALAR or b.size != a.size or p.size != G1_AFF * m:
        raise ValueError('batch_mul_add_fixed_g1: points, scalars_a and scalars_b must describe the same number of elements')
    o, op = _out(G1_AFF * m)
    _check(lib.dg_batch_mul_add_fixed_g1(pp, ap, C.c_uint64(table.handle), bp, C.c_size_t(m), op))
    return o[:G1_AFF * m]


def batch_mul_add_same_g1(points, scalars_a, v_affine, scalars_b):
    """out[i] = normalize([a_i] P_i + [b_i] V) with V given directly (no window table)."""
    lib = init()
    p, pp = _in(points); a, ap = _in(scalars_a); b, bp = _in(scalars_b); v, vp = _in(v_affine)
    m = a.size // SCALAR
    if a.size % SCALAR or b.size != a.size or p.size != G1_AFF * m or v.size != G1_AFF:
        raise ValueError('batch_mul_add_same_g1: points, scalars_a and scalars_b must describe the same number of elements')
    o, op = _out(G1_AFF * m)
    _check(lib.dg_batch_mul_add_same_g1(pp, ap, vp, bp, C.c_size_t(m), op))
    return o[:G1_AFF * m]


def compress(left, right, scalar, g2=False):
    """left[i] + [scalar] right[i], normalised (dg_compress_*)."""
    lib = init()
    l, lp = _in(left); r, rp = _in(right); s, sp = _in(scalar)
    rec = G2_AFF if g2 else G1_AFF
    m = l.size // rec
    if l.size != r.size or l.size % rec or s.size != SCALAR:
        raise ValueError('compress: left and right must hold the same number of points, scalar 32 bytes')
    o, op = _out(rec * m)
    fn = lib.dg_compress_g2 if g2 else lib.dg_compress_g1
    _check(fn(lp, rp, C.c_size_t(m), sp, op))
    return o[:rec * m]


def normalize_batch(jac, g2=False):
    lib = init()
    j, jp = _in(jac)
    m = j.size // (G2_JAC if g2 else G1_JAC)
    rec = G2_AFF if g2 else G1_AFF
    o, op = _out(rec * m)
    fn = lib.dg_normalize_batch_g2 if g2 else lib.dg_normalize_batch_g1
    _check(fn(jp, C.c_size_t(m), op))
    return o[:rec * m]


def fold(jac_points, g2=False):
    lib = init()
    j, jp = _in(jac_points)
    rec = G2_JAC if g2 else G1_JAC
    k = j.size // rec
    o, op = _out(rec)
    fn = lib.dg_fold_g2 if g2 else lib.dg_fold_g1
    _check(fn(jp, C.c_size_t(k), op))
    return o[:rec]


def fold_g1_device(jac_ptr, k, out_ptr, stream=0):
    _check(init().dg_fold_g1_device(C.c_void_p(jac_ptr), C.c_size_t(k), C.c_void_p(out_ptr), C.c_void_p(stream)))


# ---- pairing ------------------------------------------------------------------------------------
def multi_miller_loop(g1s, g2s):
    lib = init()
    a, ap = _in(g1s); b, bp = _in(g2s)
    k = min(a.size // G1_AFF, b.size // G2_AFF)
    o, op = _out(FP12)
    _check(lib.dg_multi_miller_loop(ap, bp, C.c_size_t(k), op))
    return o[:FP12]


def final_exponentiation(f):
    lib = init()
    a, ap = _in(f)
    o, op = _out(FP12)
    some = C.c_int32(0)
    _check(lib.dg_final_exponentiation(ap, op, C.byref(some)))
    return o[:FP12] if some.value else None


def multi_pairing(g1s, g2s):
    lib = init()
    a, ap = _in(g1s); b, bp = _in(g2s)
    k = min(a.size // G1_AFF, b.size // G2_AFF)
    o, op = _out(FP12)
    _check(lib.dg_multi_pairing(ap, bp, C.c_size_t(k), op))
    return o[:FP12]


def multi_pairing_batch(g1s, g2s, counts):
    """Independent pairing products in one call -> list of 576-byte GT records."""
    lib = init()
    a, ap = _in(g1s); b, bp = _in(g2s)
    nb = len(counts)
    cnt = (C.c_size_t * nb)(*counts)
    o, op = _out(FP12 * nb)
    _check(lib.dg_multi_pairing_batch(ap, bp, cnt, C.c_size_t(nb), op))
    return [o[FP12 * i:FP12 * (i + 1)] for i in range(nb)]


def multi_pairing_is_one(g1s, g2s):
    lib = init()
    a, ap = _in(g1s); b, bp = _in(g2s)
    k = min(a.size // G1_AFF, b.size // G2_AFF)
    r = C.c_int32(0)
    _check(lib.dg_multi_pairing_is_one(ap, bp, C.c_size_t(k), C.byref(r)))
    return bool(r.value)


def gt_pow(f, scalar):
    lib = init()
    a, ap = _in(f); s, sp = _in(scalar)
    o, op = _out(FP12)
    _check(lib.dg_gt_pow(ap, sp, op))
    return o[:FP12]


def fp12_mul(a, b):
    lib = init()
    x, xp = _in(a); y, yp = _in(b)
    o, op = _out(FP12)
    _check(lib.dg_fp12_mul(xp, yp, op))
    return o[:FP12]


def fr_ntt(data, logn, inverse=False, coset=False):
    """Radix2EvaluationDomain fft / ifft (+ coset) over Fr; returns a new array."""
    lib = init()
    a = np.array(_in(data)[0], dtype=np.uint8, copy=True)
    _check(lib.dg_fr_ntt(C.c_void_p(a.ctypes.data), C.c_uint32(logn), C.c_int32(1 if inverse else 0), C.c_int32(1 if coset else 0)))
    return a


def fr_ntt_device(data_ptr, tmp_ptr, logn, inverse=False, coset=False, stream=0):
    _check(init().dg_fr_ntt_device(C.c_void_p(data_ptr), C.c_void_p(tmp_ptr), C.c_uint32(logn), C.c_int32(1 if inverse else 0),
                                   C.c_int32(1 if coset else 0), C.c_void_p(stream)))


def qap_h_from_abc(a, b, c, logn):
    lib = init()
    x, xp = _in(a); y, yp = _in(b); z, zp = _in(c)
    o, op = _out(32 << logn)
    _check(lib.dg_qap_h_from_abc(xp, yp, zp, C.c_uint32(logn), op))
    return o[:32 << logn]


def fr_spmv(row_ptr, col, coeff_mont, w_mont):
    """CSR matrix (Fr Montgomery coefficients) times assignment vector (Fr Montgomery) -> rows x 32 B Montgomery."""
    lib = init()
    rp = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    cl = np.ascontiguousarray(col, dtype=np.uint32)
    co, cop = _in(coeff_mont)
    w, wp = _in(w_mont)
    if rp.size < 1 or w.size % 32 or co.size != 32 * cl.size:
        raise ValueError('fr_spmv: row_ptr must be non-empty, w a multiple of 32 bytes and coeff_mont 32 bytes per column index')
    rows, nnz, ncols = rp.size - 1, cl.size, w.size // 32
    o, op = _out(32 * rows)
    _check(lib.dg_fr_spmv(C.c_void_p(rp.ctypes.data), C.c_void_p(cl.ctypes.data), cop, C.c_size_t(rows), C.c_size_t(nnz), wp,
                          C.c_size_t(ncols), op))
    return o[:32 * rows]


class R1CS:
    """Device-resident constraint matrices (dg_r1cs_upload).  `mats` = three (row_ptr, col, coeff_mont) CSR triples."""

    def __init__(self, mats, num_constraints, num_inputs, num_vars):
        lib = init()
        keep = []
        rp, cl, co = (C.c_void_p * 3)(), (C.c_void_p * 3)(), (C.c_void_p * 3)()
        for k, (r, c, v) in enumerate(mats):
            r = np.ascontiguousarray(r, dtype=np.uint32)
            c = np.ascontiguousarray(c, dtype=np.uint32)
            v = _in(v)[0]
            if r.size != num_constraints + 1 or c.size != int(r[-1]) or v.size != 32 * c.size:
                raise ValueError('R1CS: malformed CSR matrix %d' % k)
            keep += [r, c, v]
            rp[k], cl[k], co[k] = r.ctypes.data, c.ctypes.data, v.ctypes.data
        h = C.c_uint64(0)
        _check(lib.dg_r1cs_upload(rp, cl, co, C.c_size_t(num_constraints), C.c_size_t(num_inputs), C.c_size_t(num_vars), C.byref(h)))
        self.handle, self.num_constraints, self.num_inputs, self.num_vars = h.value, num_constraints, num_inputs, num_vars

    def free(self):
        if self.handle:
            _check(load().dg_r1cs_free(C.c_uint64(self.handle)))
            self.handle = 0


def groth16_prove_msms(r1cs, full_assignment_mont, h_query, jobs, want_h=False):
    """dg_groth16_prove_msms: jobs = [(Bases, scalar offset, count)].  -> (h_acc Jacobian, [job results], h or None)."""
    lib = init()
    w, wp = _in(full_assignment_mont)
    nvars = w.size // 32
    nj = len(jobs)
    jb = (C.c_uint64 * max(nj, 1))(*[j[0].handle for j in jobs])
    jo = (C.c_uint64 * max(nj, 1))(*[j[1] for j in jobs])
    jc = (C.c_uint64 * max(nj, 1))(*[j[2] for j in jobs])
    o_h, o_hp = _out(G1_JAC)
    o_j, o_jp = _out(G2_JAC * max(nj, 1))
    logn = max((r1cs.num_constraints + r1cs.num_inputs - 1).bit_length(), 0)
    hm, hmp = _out(32 << logn) if want_h else (None, None)
    _check(lib.dg_groth16_prove_msms(C.c_uint64(r1cs.handle), wp, C.c_size_t(nvars), C.c_uint64(h_query.handle), jb, jo, jc,
                                     C.c_size_t(nj), o_hp, o_jp, hmp))
    res = [o_j[G2_JAC * i:G2_JAC * i + (G2_JAC if jobs[i][0].g2 else G1_JAC)] for i in range(nj)]
    return o_h[:G1_JAC], res, hm


# ---- ark-serialize wire formats ------------------------------------------------------------------
def serialize_points(affine, g2=False, compressed=True):
    """CanonicalSerialize::serialize_compressed / _uncompressed of a vector of affine points -> bytes array."""
    lib = init()
    a, ap = _in(affine)
    aff = G2_AFF if g2 else G1_AFF
    n = a.size // aff
    rec = (aff // 2) if compressed else aff
    o, op = _out(rec * n)
    fn = lib.dg_g2_serialize if g2 else lib.dg_g1_serialize
    _check(fn(ap, C.c_size_t(n), C.c_int32(1 if compressed else 0), op))
    return o[:rec * n]


def deserialize_points(data, g2=False, compressed=True, validate=True):
    """CanonicalDeserialize::deserialize_compressed / _uncompressed (Validate::Yes when validate) of a vector of
    points -> (affine records, per-element status bytes, number of rejected elements)."""
    lib = init()
    a, ap = _in(data)
    aff = G2_AFF if g2 else G1_AFF
    rec = (aff // 2) if compressed else aff
    n = a.size // rec
    o, op = _out(aff * n)
    st, stp = _out(n)
    bad = C.c_size_t(0)
    fn = lib.dg_g2_deserialize if g2 else lib.dg_g1_deserialize
    _check(fn(ap, C.c_size_t(n), C.c_int32(1 if compressed else 0), C.c_int32(1 if validate else 0), op, stp, C.byref(bad)))
    return o[:aff * n], st[:n], int(bad.value)


def dbg_fp_op(op_code, a, b):
    lib = init()
    x, xp = _in(a); y, yp = _in(b)
    n = x.size // 48
    o, op = _out(48 * n)
    _check(lib.dg_dbg_fp_op(C.c_int32(op_code), xp, yp, C.c_size_t(n), op))
    return o[:48 * n]


def dbg_fr_op(op_code, a, b):
    lib = init()
    x, xp = _in(a); y, yp = _in(b)
    n = x.size // 32
    o, op = _out(32 * n)
    _check(lib.dg_dbg_fr_op(C.c_int32(op_code), xp, yp, C.c_size_t(n), op))
    return o[:32 * n]
