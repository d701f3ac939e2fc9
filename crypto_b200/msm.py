"""Host-side mirror of the reference's MSM-facing interface over the C ABI.

Same names, argument meaning and error behaviour as
  * utils::msm::{WindowTable, multiply_field_elems_with_same_group_elem}   (utils/src/msm.rs:8-62)
  * ark_ec::VariableBaseMSM::{msm, msm_unchecked, msm_bigint}               
  * utils::pairs::Pairs::{msm, msm_bigint}                                  (utils/src/pairs.rs:144-156)
  * CurveGroup::normalize_batch, AffineRepr::mul_bigint
  * utils::randomized_mult_checker::RandomizedMultChecker                   (utils/src/randomized_mult_checker.rs:21-126)
so the parity tests read like the reference's own tests.  All curve arithmetic happens on the
GPU; this module only marshals bytes and does scalar-field (mod r) bookkeeping.

Types: group elements are packed Montgomery records (bytes / numpy uint8):
  affine G1 96 B, affine G2 192 B (identity = all zero), projective G1 144 B / G2 288 B.
Scalars are Python ints (canonical) or 32-byte LE records.
"""
import numpy as np

from . import lib

R_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SCALAR_SIZE = 255          # Fr::MODULUS_BIT_SIZE


class G1:
    """Group descriptor: the `G: CurveGroup` type parameter of the reference's generics."""
    g2 = False
    AFF, PROJ = lib.G1_AFF, lib.G1_JAC


class G2:
    g2 = True
    AFF, PROJ = lib.G2_AFF, lib.G2_JAC


def scalars_to_bytes(scalars):
    """ints / bytes / ndarray -> contiguous n x 32 B canonical LE (Fr::into_bigint)."""
    if isinstance(scalars, np.ndarray):
        return np.ascontiguousarray(scalars).view(np.uint8).reshape(-1)
    if isinstance(scalars, (bytes, bytearray)):
        return np.frombuffer(bytes(scalars), dtype=np.uint8)
    return np.frombuffer(b''.join(int(s % R_MODULUS).to_bytes(32, 'little') for s in scalars), dtype=np.uint8)


def points_to_bytes(points):
    if isinstance(points, np.ndarray):
        return np.ascontiguousarray(points).view(np.uint8).reshape(-1)
    if isinstance(points, (bytes, bytearray)):
        return np.frombuffer(bytes(points), dtype=np.uint8)
    return np.frombuffer(b''.join(bytes(p) for p in points), dtype=np.uint8)


def ln_without_floats(a):
    """ark_std: log2(a) * 69 / 100 with log2 = ceil."""
    return ((a - 1).bit_length() if a > 1 else 0) * 69 // 100


def is_zero(projective, group=G1):
    """Group::is_zero for a projective record: z == 0."""
    p = np.frombuffer(bytes(projective), dtype=np.uint8)
    return not p[2 * group.AFF // 2:].any()


def normalize_batch(projectives, group=G1):
    """CurveGroup::normalize_batch: m projective -> m affine records."""
    return lib.normalize_batch(points_to_bytes(projectives), g2=group.g2)


def into_affine(projective, group=G1):
    return bytes(normalize_batch(projective, group))


class LengthMismatch(ValueError):
    """ark's `Err(min_len)` from VariableBaseMSM::msm."""

    def __init__(self, min_len):
        super().__init__('bases and scalars differ in length; shorter is %d' % min_len)
        self.min_len = min_len


class VariableBaseMSM:
    """ark_ec::VariableBaseMSM for G1Projective / G2Projective."""

    def __init__(self, group=G1):
        self.group = group

    def msm_bigint(self, bases, bigints):
        """Truncates to the shorter input, never fails (ark msm_bigint)."""
        if isinstance(bases, lib.Bases):
            return bytes(lib.msm(bases, scalars_to_bytes(bigints), g2=self.group.g2))
        return bytes(lib.msm(points_to_bytes(bases), scalars_to_bytes(bigints), g2=self.group.g2))

    def msm_unchecked(self, bases, scalars):
        """into_bigint on every scalar, then msm_bigint."""
        return self.msm_bigint(bases, scalars)

    def msm(self, bases, scalars):
        """Length-checked variant: raises LengthMismatch(min_len) like `Err(min_len)`."""
        nb = bases.n if isinstance(bases, lib.Bases) else points_to_bytes(bases).size // self.group.AFF
        ns = scalars_to_bytes(scalars).size // 32
        if nb != ns:
            raise LengthMismatch(min(nb, ns))
        return self.msm_unchecked(bases, scalars)


class Pairs:
    """utils::pairs::Pairs for (affine points, scalars): equal lengths enforced at construction."""

    def __init__(self, left, right, group=G1):
        self.group = group
        self.left, self.right = points_to_bytes(left), scalars_to_bytes(right)
        nl, nr = self.left.size // group.AFF, self.right.size // 32
        if nl != nr:
            raise ValueError((nl, nr))

    def msm(self):
        return VariableBaseMSM(self.group).msm_unchecked(self.left, self.right)

    def msm_bigint(self):
        return VariableBaseMSM(self.group).msm_bigint(self.left, self.right)


class WindowTable:
    """utils::msm::WindowTable<G> (utils/src/msm.rs:8-45): fields scalar_size, window_size,
    num_windows, table (kept on the device; `table()` downloads it)."""

    def __init__(self, num_multiplications, group_elem, group=G1):
        self.group = group
        self.scalar_size = SCALAR_SIZE
        elem = points_to_bytes(group_elem)
        if elem.size == group.PROJ:                      # reference takes a projective G
            elem = normalize_batch(elem, group)
        self._t = lib.FixedBaseTable(elem, num_multiplications, g2=group.g2)
        self.window_size = self._t.window
        self.num_windows = self._t.num_windows
        assert self.window_size == WindowTable.window_size_for(num_multiplications)
        assert self.num_windows == (self.scalar_size + self.window_size - 1) // self.window_size

    @classmethod
    def new(cls, num_multiplications, group_elem, group=G1):
        return cls(num_multiplications, group_elem, group)

    @staticmethod
    def window_size_for(num_multiplications):
        """WindowTable::window_size = FixedBase::get_mul_window_size."""
        return 3 if num_multiplications < 32 else ln_without_floats(num_multiplications)

    def multiply(self, element):
        """-> one projective record."""
        return bytes(self._t.mul_many(scalars_to_bytes([element])))

    def __mul__(self, element):
        return self.multiply(element)

    def multiply_many(self, elements):
        """-> m projective records (FixedBase::msm)."""
        return self._t.mul_many(scalars_to_bytes(elements))

    def table(self):
        return self._t.download()

    def free(self):
        self._t.free()


def multiply_field_elems_with_same_group_elem(group_elem, elements, group=G1):
    """utils/src/msm.rs:56-62."""
    elements = scalars_to_bytes(elements)
    table = WindowTable(elements.size // 32, group_elem, group)
    try:
        return table.multiply_many(elements)
    finally:
        table.free()


def mul_bigint_batch(points, scalars, group=G1):
    """`cfg_iter!(points).zip(scalars).map(|(p, s)| p.mul_bigint(s))`: m projective records."""
    return lib.batch_mul(points_to_bytes(points), scalars_to_bytes(scalars), g2=group.g2)


class RandomizedMultChecker:
    """utils::randomized_mult_checker::RandomizedMultChecker<G1Affine>: folds many
    `sum p_i * s_i = t` checks into one MSM that must be zero.  Points are keyed by their
    x-coordinate so P and -P share an entry (utils/src/randomized_mult_checker.rs:107-125)."""

    def __init__(self, random, group=G1):
        self.group = group
        self.args = {}                       # x bytes -> [scalar, point bytes]
        self.random = random % R_MODULUS
        self.current_random = 1

    @classmethod
    def new(cls, random, group=G1):
        return cls(random, group)

    def _add(self, p, s):
        p = bytes(p)
        half = self.group.AFF // 2
        if not any(p):
            return                            # point at infinity does not affect the result
        x = p[:half]
        ent = self.args.get(x)
        if ent is None:
            self.args[x] = [s % R_MODULUS, p]
        elif ent[1] == p:
            ent[0] = (ent[0] + s) % R_MODULUS
        else:
            ent[0] = (ent[0] - s) % R_MODULUS

    def add_1(self, p, s, t):
        self._add(p, self.current_random * s)
        self._add(t, -self.current_random)
        self.current_random = self.current_random * self.random % R_MODULUS

    def add_2(self, p1, s1, p2, s2, t):
        self._add(p1, self.current_random * s1)
        self._add(p2, self.current_random * s2)
        self._add(t, -self.current_random)
        self.current_random = self.current_random * self.random % R_MODULUS

    def add_3(self, p1, s1, p2, s2, p3, s3, t):
        self._add(p1, self.current_random * s1)
        self._add(p2, self.current_random * s2)
        self._add(p3, self.current_random * s3)
        self._add(t, -self.current_random)
        self.current_random = self.current_random * self.random % R_MODULUS

    def add_many(self, a, b, t):
        for a_i, b_i in zip(a, b):
            self._add(a_i, self.current_random * b_i)
        self._add(t, -self.current_random)
        self.current_random = self.current_random * self.random % R_MODULUS

    def __len__(self):
        return len(self.args)

    def verify(self):
        # BTreeMap iterates in key order; the MSM result does not depend on it
        items = [self.args[k] for k in sorted(self.args)]
        points = [it[1] for it in items]
        scalars = [it[0] for it in items]
        res = VariableBaseMSM(self.group).msm_unchecked(points, scalars)
        return is_zero(res, self.group)


# ---- digit recoding of the MSM kernels (specification mirror; csrc/common.cuh msm_ndigits, csrc/msm_kernels.cuh k_digits) ----
def msm_ndigits(c):
    """Signed radix-2^c digits per scalar after the range halving s -> min(s, r - s) < 2^254."""
    return (254 + c - 1) // c + (1 if 254 % c == 0 else 0)


def msm_recode(s, c):
    """(flip, digits): the signed digits k_digits derives for the canonical scalar s, least significant first, each in
    [-(2^(c-1) - 1), 2^(c-1)]; sum(d_j 2^(c j)) == (r - s if flip else s).  Raises ValueError for s >= r (DG_ERR_BAD_ARG)."""
    if not 0 <= s < R_MODULUS:
        raise ValueError('scalar is not canonical (>= r)')
    flip = R_MODULUS - s < s
    v = R_MODULUS - s if flip else s
    half, digits, carry = 1 << (c - 1), [], 0
    for j in range(msm_ndigits(c)):
        d = ((v >> (c * j)) & ((1 << c) - 1)) + carry
        carry = 1 if d > half else 0
        digits.append(d - (1 << c) if carry else d)
    if carry:
        raise AssertionError('top digit carried out: msm_ndigits(%d) is too small' % c)
    return flip, digits


# ---- GLV split of the plain-bases path (specification mirror; csrc/msm_kernels.cuh glv_split) -----------------------
BLS_X_ABS = 0xD201000000010000
GLV_X2 = BLS_X_ABS * BLS_X_ABS                      # lambda = -x^2 mod r is the eigenvalue of phi(x, y) = (beta x, y)
GLV_LAMBDA = (-GLV_X2) % R_MODULUS


def glv_ndigits(c):
    """Signed radix-2^c digits per 127-bit half-scalar."""
    return (128 + c - 1) // c


def glv_split(s):
    """(sign1, k1, sign2, k2) with  s = sign1 * k1 + sign2 * k2 * lambda  (mod r)  and  k1, k2 < 2^127: the scalar is first
    reduced to s' = min(s, r - s), then s' = q x^2 +- k1 with k1 <= x^2 / 2 (Barrett quotient by the constant x^2)."""
    if not 0 <= s < R_MODULUS:
        raise ValueError('scalar is not canonical (>= r)')
    flip = R_MODULUS - s < s
    sp = R_MODULUS - s if flip else s
    q = (sp * ((1 << 256) // GLV_X2)) >> 256
    rem = sp - q * GLV_X2
    while rem >= GLV_X2:
        rem -= GLV_X2
        q += 1
    neg1 = rem > GLV_X2 // 2
    if neg1:
        rem, q = GLV_X2 - rem, q + 1
    sigma = -1 if flip else 1
    return (-sigma if neg1 else sigma), rem, -sigma, q


def msm_digits_per_scalar(c, plain_bases):
    """Bucket entries a non-zero scalar contributes at most: 2 x glv_ndigits for plain bases, msm_ndigits through a table."""
    return 2 * glv_ndigits(c) if plain_bases else msm_ndigits(c)
