"""Big-integer oracle for the BLS12-381 hot path (TEST INFRASTRUCTURE ONLY).

This file is a checker.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product path
(crypto_b200/) never does and fails loudly when the CUDA library is missing.

PARITY STATUS: "parity unpinned" by the reference.  docknetwork/crypto holds no
golden vector / known-answer test for MSM, fixed-base mul, Miller loop or final
exponentiation; the arithmetic lives in the un-vendored
dependencies ark-ec ^0.4.1 / ark-ff ^0.4.1 / ark-bls12-381 ^0.4.0
(/root/reference/Cargo.toml:32-46, no Cargo.lock).  The oracle is therefore
pinned on (1) the public BLS12-381 constants (generators on curve, r*G = O,
p and r derived from x), (2) mathematical uniqueness: MSM / fixed-base /
batch-mul / normalize outputs are unique group elements, so any correct
implementation is bit-identical in canonical affine form, (3) an independent
second derivation of the pairing (polynomial-ring Fp12, affine textbook Miller
loop, plain pow final exponentiation) that must agree with the arkworks-style
tower algorithm restated here, and (4) the relational tests the reference
itself uses (utils/src/msm.rs:116-308, utils/src/randomized_pairing_check.rs:
234-421).

Everything here is plain Python ints: use for small cases only.
"""

X_ABS = 0xD201000000010000          # |x|, x is negative
X_IS_NEG = True
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
assert R == X_ABS**4 - X_ABS**2 + 1
assert P == ((-X_ABS - 1) ** 2 * R) // 3 + (-X_ABS)

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)

FP_R = (1 << 384) % P          # Montgomery radix for Fp (ark MontBackend, 6x64)
FR_R = (1 << 256) % R


# ---------------------------------------------------------------- Fp ------
class FpOps:
    """Field ops object so curve code is generic over Fp / Fp2."""
    zero = 0
    one = 1

    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def mul(a, b): return (a * b) % P
    @staticmethod
    def sqr(a): return (a * a) % P
    @staticmethod
    def inv(a): return pow(a, P - 2, P)
    @staticmethod
    def is_zero(a): return a % P == 0
    @staticmethod
    def muli(a, k): return (a * k) % P


# ---------------------------------------------------------------- Fp2 -----
def fp2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def fp2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def fp2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def fp2_sqr(a): return fp2_mul(a, a)
def fp2_conj(a): return (a[0], (-a[1]) % P)
def fp2_mul_fp(a, k): return ((a[0] * k) % P, (a[1] * k) % P)
def fp2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
    return ((a[0] * n) % P, (-a[1] * n) % P)
def fp2_is_zero(a): return a[0] % P == 0 and a[1] % P == 0
def fp2_mul_xi(a):  # times (1 + u)
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)
def fp2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = fp2_mul(r, a)
        a = fp2_sqr(a)
        e >>= 1
    return r


class Fp2Ops:
    zero = (0, 0)
    one = (1, 0)
    add = staticmethod(fp2_add)
    sub = staticmethod(fp2_sub)
    neg = staticmethod(fp2_neg)
    mul = staticmethod(fp2_mul)
    sqr = staticmethod(fp2_sqr)
    inv = staticmethod(fp2_inv)
    is_zero = staticmethod(fp2_is_zero)
    muli = staticmethod(fp2_mul_fp)


FP2_ZERO, FP2_ONE = (0, 0), (1, 0)
XI = (1, 1)

# ---------------------------------------------------------------- Fp6 -----
# Fp6 = Fp2[v]/(v^3 - xi), element (c0, c1, c2)
FP6_ZERO = (FP2_ZERO, FP2_ZERO, FP2_ZERO)
FP6_ONE = (FP2_ONE, FP2_ZERO, FP2_ZERO)

def fp6_add(a, b): return tuple(fp2_add(x, y) for x, y in zip(a, b))
def fp6_sub(a, b): return tuple(fp2_sub(x, y) for x, y in zip(a, b))
def fp6_neg(a): return tuple(fp2_neg(x) for x in a)
def fp6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    t0, t1, t2 = fp2_mul(a0, b0), fp2_mul(a1, b1), fp2_mul(a2, b2)
    c0 = fp2_add(t0, fp2_mul_xi(fp2_add(fp2_mul(a1, b2), fp2_mul(a2, b1))))
    c1 = fp2_add(fp2_add(fp2_mul(a0, b1), fp2_mul(a1, b0)), fp2_mul_xi(t2))
    c2 = fp2_add(fp2_add(fp2_mul(a0, b2), fp2_mul(a2, b0)), t1)
    return (c0, c1, c2)
def fp6_mul_v(a):  # times v
    return (fp2_mul_xi(a[2]), a[0], a[1])
def fp6_inv(a):
    a0, a1, a2 = a
    t0 = fp2_sub(fp2_sqr(a0), fp2_mul_xi(fp2_mul(a1, a2)))
    t1 = fp2_sub(fp2_mul_xi(fp2_sqr(a2)), fp2_mul(a0, a1))
    t2 = fp2_sub(fp2_sqr(a1), fp2_mul(a0, a2))
    d = fp2_add(fp2_mul(a0, t0), fp2_mul_xi(fp2_add(fp2_mul(a2, t1), fp2_mul(a1, t2))))
    di = fp2_inv(d)
    return (fp2_mul(t0, di), fp2_mul(t1, di), fp2_mul(t2, di))

# ---------------------------------------------------------------- Fp12 ----
# Fp12 = Fp6[w]/(w^2 - v), element (c0, c1)
FP12_ONE = (FP6_ONE, FP6_ZERO)

def fp12_mul(a, b):
    t0 = fp6_mul(a[0], b[0])
    t1 = fp6_mul(a[1], b[1])
    c0 = fp6_add(t0, fp6_mul_v(t1))
    c1 = fp6_sub(fp6_sub(fp6_mul(fp6_add(a[0], a[1]), fp6_add(b[0], b[1])), t0), t1)
    return (c0, c1)
def fp12_sqr(a): return fp12_mul(a, a)
def fp12_conj(a): return (a[0], fp6_neg(a[1]))
def fp12_inv(a):
    d = fp6_sub(fp6_mul(a[0], a[0]), fp6_mul_v(fp6_mul(a[1], a[1])))
    di = fp6_inv(d)
    return (fp6_mul(a[0], di), fp6_neg(fp6_mul(a[1], di)))
def fp12_pow(a, e):
    r = FP12_ONE
    while e:
        if e & 1:
            r = fp12_mul(r, a)
        a = fp12_sqr(a)
        e >>= 1
    return r

# Frobenius coefficients: gamma_k[i] = xi^(i*(p^k-1)/6), i = 0..5
def _frob_coeffs(k):
    e = (P**k - 1) // 6
    g = fp2_pow(XI, e)
    out = [FP2_ONE]
    for _ in range(5):
        out.append(fp2_mul(out[-1], g))
    return out
FROB = {k: _frob_coeffs(k) for k in (1, 2, 3)}

def fp12_frobenius(a, k):
    """a^(p^k) for k in 1..3.  Basis element v^j w^i = w^(2j+i) picks up gamma_k[2j+i]."""
    g = FROB[k]
    cj = (lambda z: fp2_conj(z)) if (k & 1) else (lambda z: z)
    c0 = tuple(fp2_mul(cj(a[0][j]), g[2 * j]) for j in range(3))
    c1 = tuple(fp2_mul(cj(a[1][j]), g[2 * j + 1]) for j in range(3))
    return (c0, c1)

def fp12_mul_by_014(f, c0, c1, c4):
    """Sparse mult by c0 + c1*v + c4*v*w (ark Fp12::mul_by_014)."""
    s = ((c0, c1, FP2_ZERO), (FP2_ZERO, c4, FP2_ZERO))
    return fp12_mul(f, s)

def fp12_cyclotomic_exp_x(f):
    """f^|x| then conjugate because x < 0 (ark Bls12::exp_by_x)."""
    r = fp12_pow(f, X_ABS)
    return fp12_conj(r) if X_IS_NEG else r


# ---------------------------------------------------------------- curves --
class Curve:
    """Short Weierstrass y^2 = x^3 + b, a = 0, affine points as (x, y) or None."""

    def __init__(self, F, b):
        self.F, self.b = F, b

    def on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)) == F.zero

    def neg(self, pt):
        return None if pt is None else (pt[0], self.F.neg(pt[1]))

    def add(self, p1, p2):
        F = self.F
        if p1 is None: return p2
        if p2 is None: return p1
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            if F.is_zero(F.add(y1, y2)):
                return None
            lam = F.mul(F.muli(F.sqr(x1), 3), F.inv(F.muli(y1, 2)))
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    # Jacobian for speed in naive scalar mul
    def _jdbl(self, p):
        F = self.F
        X, Y, Z = p
        if F.is_zero(Z): return p
        A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
        D = F.muli(F.sub(F.sub(F.sqr(F.add(X, B)), A), C), 2)
        E = F.muli(A, 3); Fq = F.sqr(E)
        X3 = F.sub(Fq, F.muli(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
        Z3 = F.muli(F.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def _jadd_affine(self, p, q):
        F = self.F
        X1, Y1, Z1 = p
        if F.is_zero(Z1): return (q[0], q[1], F.one)
        Z1Z1 = F.sqr(Z1)
        U2 = F.mul(q[0], Z1Z1)
        S2 = F.mul(F.mul(q[1], Z1), Z1Z1)
        if U2 == X1:
            if S2 == Y1:
                return self._jdbl(p)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, X1); HH = F.sqr(H); HHH = F.mul(H, HH)
        rr = F.sub(S2, Y1); V = F.mul(X1, HH)
        X3 = F.sub(F.sub(F.sqr(rr), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(rr, F.sub(V, X3)), F.mul(Y1, HHH))
        Z3 = F.mul(Z1, H)
        return (X3, Y3, Z3)

    def to_affine(self, p):
        F = self.F
        X, Y, Z = p
        if F.is_zero(Z): return None
        zi = F.inv(Z); zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(F.mul(Y, zi2), zi))

    def mul(self, pt, k):
        """[k]pt for any integer k (same as ark mul_bigint on the canonical integer)."""
        if pt is None or k == 0:
            return None
        if k < 0:
            return self.mul(self.neg(pt), -k)
        F = self.F
        acc = (F.one, F.one, F.zero)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == '1':
                acc = self._jadd_affine(acc, pt)
        return self.to_affine(acc)

    def msm_naive(self, bases, scalars):
        """sum s_i * P_i, definition; truncates to the shorter input like
        ark VariableBaseMSM::msm_bigint."""
        acc = None
        for b, s in zip(bases, scalars):
            acc = self.add(acc, self.mul(b, s))
        return acc


E1 = Curve(FpOps, 4)
E2 = Curve(Fp2Ops, (4, 4))
assert E1.on_curve(G1_GEN) and E2.on_curve(G2_GEN)


# ------------------------------------------------- ark-style algorithms ---
def ln_without_floats(a):
    """ark_std::log2(a) * 69 / 100 with log2 = ceil(log2(a))."""
    return ((a - 1).bit_length() if a > 1 else 0) * 69 // 100

def msm_window_size(n):
    """ark-ec 0.4 VariableBaseMSM window rule."""
    return 3 if n < 32 else ln_without_floats(n) + 2

def fixed_base_window_size(n):
    """ark-ec 0.4 FixedBase::get_mul_window_size (used by utils/src/msm.rs:20)."""
    return 3 if n < 32 else ln_without_floats(n)

def make_digits(s, c, num_bits=255):
    """Signed radix-2^c digits, digit in [-2^(c-1), 2^(c-1)), last one unsigned (ark make_digits)."""
    n = (num_bits + c - 1) // c
    radix, half = 1 << c, 1 << (c - 1)
    out, carry = [], 0
    for i in range(n):
        d = ((s >> (i * c)) & (radix - 1)) + carry
        carry = (d + half) >> c
        d -= carry << c
        if i == n - 1:
            d += carry << c
        out.append(d)
    return out

def msm_pippenger(curve, bases, scalars):
    """Restatement of ark-ec 0.4 msm_bigint_wnaf on big ints."""
    n = min(len(bases), len(scalars))
    bases, scalars = bases[:n], scalars[:n]
    c = msm_window_size(n)
    nd = (255 + c - 1) // c
    digits = [make_digits(s, c) for s in scalars]
    window_sums = []
    for w in range(nd):
        buckets = [None] * (1 << c)    # ark allocates 1 << c: the top digit is left unsigned
        for i in range(n):
            d = digits[i][w]
            if d > 0:
                buckets[d - 1] = curve.add(buckets[d - 1], bases[i])
            elif d < 0:
                buckets[-d - 1] = curve.add(buckets[-d - 1], curve.neg(bases[i]))
        running = res = None
        for b in reversed(buckets):
            running = curve.add(running, b)
            res = curve.add(res, running)
        window_sums.append(res)
    total = None
    for ws in reversed(window_sums[1:]):
        total = curve.add(total, ws)
        for _ in range(c):
            total = curve.add(total, total)
    return curve.add(window_sums[0], total)

def fixed_base_table(curve, g, window, scalar_size=255):
    """ark FixedBase::get_window_table: table[k][j] = j * 2^(k*window) * g."""
    outerc = (scalar_size + window - 1) // window
    last = 1 << (scalar_size - (outerc - 1) * window)
    table = []
    g_outer = g
    for k in range(outerc):
        size = last if k == outerc - 1 else (1 << window)
        row, acc = [], None
        for _ in range(size):
            row.append(acc)
            acc = curve.add(acc, g_outer)
        table.append(row)
        for _ in range(window):
            g_outer = curve.add(g_outer, g_outer)
    return table

def windowed_mul(curve, table, window, s):
    """ark FixedBase::windowed_mul."""
    res = None
    for k, row in enumerate(table):
        idx = (s >> (k * window)) & ((1 << window) - 1)
        res = curve.add(res, row[idx])
    return res


# ------------------------------------------------- pairing, ark-style -----
TWO_INV = pow(2, P - 2, P)
_X_BITS = bin(X_ABS)[3:]            # BE bits of |x| without the top one

def g2_prepare(q):
    """ark bls12::G2Prepared::from: 68 line-coefficient triples (homogeneous projective)."""
    if q is None:
        return None
    rx, ry, rz = q[0], q[1], FP2_ONE
    B = (4, 4)
    coeffs = []
    def dbl():
        nonlocal rx, ry, rz
        a = fp2_mul_fp(fp2_mul(rx, ry), TWO_INV)
        b = fp2_sqr(ry)
        c = fp2_sqr(rz)
        e = fp2_mul(B, fp2_add(fp2_add(c, c), c))
        f = fp2_add(fp2_add(e, e), e)
        g = fp2_mul_fp(fp2_add(b, f), TWO_INV)
        h = fp2_sub(fp2_sqr(fp2_add(ry, rz)), fp2_add(b, c))
        i = fp2_sub(e, b)
        j = fp2_sqr(rx)
        e2 = fp2_sqr(e)
        rx = fp2_mul(a, fp2_sub(b, f))
        ry = fp2_sub(fp2_sqr(g), fp2_add(fp2_add(e2, e2), e2))
        rz = fp2_mul(b, h)
        return (i, fp2_add(fp2_add(j, j), j), fp2_neg(h))
    def add():
        nonlocal rx, ry, rz
        theta = fp2_sub(ry, fp2_mul(q[1], rz))
        lam = fp2_sub(rx, fp2_mul(q[0], rz))
        c = fp2_sqr(theta)
        d = fp2_sqr(lam)
        e = fp2_mul(lam, d)
        f = fp2_mul(rz, c)
        g = fp2_mul(rx, d)
        h = fp2_sub(fp2_add(e, f), fp2_add(g, g))
        rx = fp2_mul(lam, h)
        ry = fp2_sub(fp2_mul(theta, fp2_sub(g, h)), fp2_mul(e, ry))
        rz = fp2_mul(rz, e)
        j = fp2_sub(fp2_mul(theta, q[0]), fp2_mul(lam, q[1]))
        return (j, fp2_neg(theta), lam)
    for bit in _X_BITS:
        coeffs.append(dbl())
        if bit == '1':
            coeffs.append(add())
    return coeffs

def _ell(f, coeffs, p):
    c0, c1, c2 = coeffs
    return fp12_mul_by_014(f, c0, fp2_mul_fp(c1, p[0]), fp2_mul_fp(c2, p[1]))

def multi_miller_loop(ps, qs):
    """ark Bls12::multi_miller_loop: pairs with an identity side are dropped."""
    pairs = [(p, g2_prepare(q)) for p, q in zip(ps, qs) if p is not None and q is not None]
    f = FP12_ONE
    idx = 0
    for bit in _X_BITS:
        f = fp12_sqr(f)
        for p, co in pairs:
            f = _ell(f, co[idx], p)
        idx += 1
        if bit == '1':
            for p, co in pairs:
                f = _ell(f, co[idx], p)
            idx += 1
    return fp12_conj(f) if X_IS_NEG else f

def final_exponentiation(f):
    """ark Bls12::final_exponentiation (eprint 2020/875 chain): returns
    f^(3*(p^12-1)/r).  None iff f == 0."""
    if f == (FP6_ZERO, FP6_ZERO):
        return None
    f1 = fp12_conj(f)
    f2 = fp12_inv(f)
    r = fp12_mul(f1, f2)
    f2 = r
    r = fp12_mul(fp12_frobenius(r, 2), f2)
    y0 = fp12_sqr(r)
    y1 = fp12_cyclotomic_exp_x(r)
    y2 = fp12_conj(r)
    y1 = fp12_mul(y1, y2)
    y2 = fp12_cyclotomic_exp_x(y1)
    y1 = fp12_conj(y1)
    y1 = fp12_mul(y1, y2)
    y2 = fp12_cyclotomic_exp_x(y1)
    y1 = fp12_frobenius(y1, 1)
    y1 = fp12_mul(y1, y2)
    r = fp12_mul(r, y0)
    y0 = fp12_cyclotomic_exp_x(y1)
    y2 = fp12_cyclotomic_exp_x(y0)
    y0 = fp12_frobenius(y1, 2)
    y1 = fp12_conj(y1)
    y1 = fp12_mul(y1, y2)
    y1 = fp12_mul(y1, y0)
    r = fp12_mul(r, y1)
    return r

def multi_pairing(ps, qs):
    return final_exponentiation(multi_miller_loop(ps, qs))

def pairing(p, q):
    return multi_pairing([p], [q])


# ------------------------------------ pairing, independent derivation -----
# Fp12 as Fp[w]/(w^12 - 2 w^6 + 2): w^6 = xi = 1 + u, u^2 = -1.
def _poly_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):          # w^12 = 2 w^6 - 2
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return [v % P for v in t[:12]]

def _poly_pow(a, e):
    r = [1] + [0] * 11
    while e:
        if e & 1:
            r = _poly_mul(r, a)
        a = _poly_mul(a, a)
        e >>= 1
    return r

def fp12_to_poly(a):
    """Tower element -> polynomial: coefficient of v^j w^i sits at w^(2j+i); a+bu = (a-b) + b w^6."""
    out = [0] * 12
    for i in range(2):
        for j in range(3):
            c = a[i][j]
            k = 2 * j + i
            out[k] = (out[k] + c[0] - c[1]) % P
            out[k + 6] = (out[k + 6] + c[1]) % P
    return out

def _poly_from_fp2(c, k):
    out = [0] * 12
    out[k] = (c[0] - c[1]) % P
    out[k + 6] = c[1] % P
    return out

def _poly_inv(a):
    return _poly_pow(a, P**12 - 2)

def pairing_textbook(p, q, cube=True):
    """Ate pairing by the book: untwist Q to E(Fp12) via (x/w^2, y/w^3), affine Miller
    loop for f_{|x|,Q}(P), inverse for x<0, plain pow by (p^12-1)/r (times 3 to match
    the arkworks final-exponentiation convention)."""
    if p is None or q is None:
        return [1] + [0] * 11
    # 1/w^2 and 1/w^3 in the polynomial ring
    w = [0, 1] + [0] * 10
    w2i = _poly_inv(_poly_mul(w, w))
    w3i = _poly_mul(w2i, _poly_inv(w))
    xq = _poly_mul(_poly_from_fp2(q[0], 0), w2i)
    yq = _poly_mul(_poly_from_fp2(q[1], 0), w3i)
    xp = [p[0]] + [0] * 11
    yp = [p[1]] + [0] * 11
    add = lambda a, b: [(x + y) % P for x, y in zip(a, b)]
    sub = lambda a, b: [(x - y) % P for x, y in zip(a, b)]
    def line(t, s):
        # line through t and s (tangent if equal) evaluated at P, and t + s
        (x1, y1), (x2, y2) = t, s
        if x1 == x2 and y1 == y2:
            num = _poly_mul([3] + [0] * 11, _poly_mul(x1, x1))
            den = add(y1, y1)
        else:
            num, den = sub(y2, y1), sub(x2, x1)
        lam = _poly_mul(num, _poly_inv(den))
        x3 = sub(sub(_poly_mul(lam, lam), x1), x2)
        y3 = sub(_poly_mul(lam, sub(x1, x3)), y1)
        val = sub(sub(yp, y1), _poly_mul(lam, sub(xp, x1)))
        return val, (x3, y3)
    f = [1] + [0] * 11
    t = (xq, yq)
    for bit in _X_BITS:
        l, t2 = line(t, t)
        f = _poly_mul(_poly_mul(f, f), l)
        t = t2
        if bit == '1':
            l, t2 = line(t, (xq, yq))
            f = _poly_mul(f, l)
            t = t2
    if X_IS_NEG:
        f = _poly_inv(f)
    e = (P**12 - 1) // R
    return _poly_pow(f, 3 * e if cube else e)


# ------------------------------------------------- wire formats -----------
def fp_to_mont_bytes(a):
    """Fp canonical int -> 48 B little-endian Montgomery (ark Fp<MontBackend,6>.0.0)."""
    return ((a * FP_R) % P).to_bytes(48, 'little')

def fp_from_mont_bytes(b):
    return (int.from_bytes(b, 'little') * pow(FP_R, P - 2, P)) % P

def g1_to_bytes(pt):
    """Packed record of the C ABI: x || y Montgomery LE (96 B); identity = all zero."""
    if pt is None:
        return bytes(96)
    return fp_to_mont_bytes(pt[0]) + fp_to_mont_bytes(pt[1])

def g1_from_bytes(b):
    if b == bytes(96):
        return None
    return (fp_from_mont_bytes(b[:48]), fp_from_mont_bytes(b[48:96]))

def g2_to_bytes(pt):
    """x.c0 || x.c1 || y.c0 || y.c1 Montgomery LE (192 B); identity = all zero."""
    if pt is None:
        return bytes(192)
    return b''.join(fp_to_mont_bytes(c) for c in (pt[0][0], pt[0][1], pt[1][0], pt[1][1]))

def g2_from_bytes(b):
    if b == bytes(192):
        return None
    c = [fp_from_mont_bytes(b[i * 48:(i + 1) * 48]) for i in range(4)]
    return ((c[0], c[1]), (c[2], c[3]))

def fp12_to_bytes(a):
    """c0.c0.c0, c0.c0.c1, c0.c1.c0, ... (ark field order), 12 x 48 B Montgomery LE."""
    return b''.join(fp_to_mont_bytes(a[i][j][k]) for i in range(2) for j in range(3) for k in range(2))

def fp12_from_bytes(b):
    c = [fp_from_mont_bytes(b[i * 48:(i + 1) * 48]) for i in range(12)]
    return tuple(tuple((c[i * 6 + j * 2], c[i * 6 + j * 2 + 1]) for j in range(3)) for i in range(2))

def scalar_to_bytes(s):
    """Canonical (non-Montgomery) 32 B LE = ark BigInt<4> as msm_bigint receives it."""
    return (s % R).to_bytes(32, 'little')

def g1_compressed(pt):
    """ark-bls12-381 0.4 / Zcash compressed G1 (48 B big-endian x with flag bits)."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(pt[0].to_bytes(48, 'big'))
    b[0] |= 0x80
    if pt[1] > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


# ------------------------------------------------- ark-serialize wire formats ("next" row f4) ---
# ark-bls12-381 0.4 (not vendored in /root/reference; Cargo.toml:34) serialises points in the
# Zcash / IETF format, which the reference reaches through CanonicalSerialize::serialize_compressed /
# CanonicalDeserialize::deserialize_compressed (utils/src/serde_utils.rs:13-33, the derive on
# legogroth16/src/data_structures.rs:7-189):  big-endian coordinates; the three top bits of byte 0
# are flags  0x80 compressed, 0x40 infinity, 0x20 "y is the lexicographically larger root";
# G2 writes c1 before c0.  Deserialisation with Validate::Yes also checks subgroup membership.
def fp_sqrt(a):
    """A square root of a in Fp (p = 3 mod 4) or None."""
    s = pow(a, (P + 1) // 4, P)
    return s if (s * s) % P == a % P else None

def fp2_sqrt(a):
    """A square root of a in Fp2 = Fp[u]/(u^2+1) or None (norm method)."""
    c0, c1 = a[0] % P, a[1] % P
    if c1 == 0:
        s = fp_sqrt(c0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt((-c0) % P)            # (t u)^2 = -t^2
        return None if s is None else (0, s)
    n = fp_sqrt((c0 * c0 + c1 * c1) % P)
    if n is None:
        return None
    for nn in (n, (-n) % P):
        d = ((c0 + nn) * TWO_INV) % P
        s = fp_sqrt(d)
        if s is not None and s != 0:
            r = (s, (c1 * pow(2 * s, P - 2, P)) % P)
            if fp2_sqr(r) == (c0, c1):
                return r
    return None

def fp_lex_largest(y):
    return y % P > (P - 1) // 2

def fp2_lex_largest(y):
    return fp_lex_largest(y[1]) if y[1] % P else fp_lex_largest(y[0])

# endomorphisms behind ark-bls12-381's fast subgroup checks (eprint 2021/1130, section 6):
#   G1:  (beta x, y) == -[x^2] P        G2:  psi(Q) == [x] Q,  psi = twist o Frobenius o untwist
def _find_beta():
    g = 2
    while True:
        b = pow(g, (P - 1) // 3, P)
        if b != 1:
            break
        g += 1
    for cand in (b, (b * b) % P):
        if ((cand * G1_GEN[0]) % P, G1_GEN[1]) == E1.neg(E1.mul(G1_GEN, X_ABS * X_ABS)):
            return cand
    raise AssertionError('no cube root of unity matches the G1 endomorphism')

BETA = _find_beta()
PSI_X = fp2_inv(fp2_pow(XI, (P - 1) // 3))
PSI_Y = fp2_inv(fp2_pow(XI, (P - 1) // 2))

def g2_psi(q):
    if q is None:
        return None
    return (fp2_mul(fp2_conj(q[0]), PSI_X), fp2_mul(fp2_conj(q[1]), PSI_Y))

assert g2_psi(G2_GEN) == E2.neg(E2.mul(G2_GEN, X_ABS))          # psi acts as [x] on G2, x < 0

def g1_in_subgroup(pt):
    """Definition: [r] P == O."""
    return E1.mul(pt, R) is None

def g2_in_subgroup(pt):
    return E2.mul(pt, R) is None

def g1_in_subgroup_fast(pt):
    if pt is None:
        return True
    return ((BETA * pt[0]) % P, pt[1]) == E1.neg(E1.mul(pt, X_ABS * X_ABS))

def g2_in_subgroup_fast(pt):
    if pt is None:
        return True
    return g2_psi(pt) == E2.neg(E2.mul(pt, X_ABS))

SER_OK, SER_MALFORMED, SER_NOT_ON_CURVE, SER_NOT_IN_SUBGROUP = 0, 1, 2, 3

def g1_serialize(pt, compressed=True):
    if pt is None:
        return bytes([0xC0 if compressed else 0x40]) + bytes(47 if compressed else 95)
    b = bytearray(pt[0].to_bytes(48, 'big'))
    if compressed:
        b[0] |= 0x80
        if fp_lex_largest(pt[1]):
            b[0] |= 0x20
        return bytes(b)
    return bytes(b) + pt[1].to_bytes(48, 'big')

def g2_serialize(pt, compressed=True):
    if pt is None:
        return bytes([0xC0 if compressed else 0x40]) + bytes(95 if compressed else 191)
    b = bytearray(pt[0][1].to_bytes(48, 'big') + pt[0][0].to_bytes(48, 'big'))
    if compressed:
        b[0] |= 0x80
        if fp2_lex_largest(pt[1]):
            b[0] |= 0x20
        return bytes(b)
    return bytes(b) + pt[1][1].to_bytes(48, 'big') + pt[1][0].to_bytes(48, 'big')

def _deser_flags(b0, compressed):
    return bool(b0 & 0x80) == compressed, bool(b0 & 0x40), bool(b0 & 0x20)

def g1_deserialize(b, compressed=True, validate=True):
    """-> (status, point).  Malformed = wrong compression flag, coordinate >= p, stray bits with the infinity flag,
    sort flag on an uncompressed encoding."""
    ok, inf, largest = _deser_flags(b[0], compressed)
    body = bytes([b[0] & 0x1F]) + bytes(b[1:])
    if not ok or (not compressed and largest):
        return SER_MALFORMED, None
    if inf:
        return (SER_OK, None) if not any(body) and not largest else (SER_MALFORMED, None)
    x = int.from_bytes(body[:48], 'big')
    if x >= P:
        return SER_MALFORMED, None
    if compressed:
        y = fp_sqrt((x * x * x + 4) % P)
        if y is None:
            return SER_NOT_ON_CURVE, None
        if fp_lex_largest(y) != largest:
            y = (-y) % P
    else:
        y = int.from_bytes(body[48:96], 'big')
        if y >= P:
            return SER_MALFORMED, None
        if not E1.on_curve((x, y)):
            return SER_NOT_ON_CURVE, None
    if validate and not g1_in_subgroup((x, y)):
        return SER_NOT_IN_SUBGROUP, None
    return SER_OK, (x, y)

def g2_deserialize(b, compressed=True, validate=True):
    ok, inf, largest = _deser_flags(b[0], compressed)
    body = bytes([b[0] & 0x1F]) + bytes(b[1:])
    if not ok or (not compressed and largest):
        return SER_MALFORMED, None
    if inf:
        return (SER_OK, None) if not any(body) and not largest else (SER_MALFORMED, None)
    x1, x0 = int.from_bytes(body[:48], 'big'), int.from_bytes(body[48:96], 'big')
    if x0 >= P or x1 >= P:
        return SER_MALFORMED, None
    x = (x0, x1)
    if compressed:
        y = fp2_sqrt(fp2_add(fp2_mul(fp2_sqr(x), x), (4, 4)))
        if y is None:
            return SER_NOT_ON_CURVE, None
        if fp2_lex_largest(y) != largest:
            y = fp2_neg(y)
    else:
        y1, y0 = int.from_bytes(body[96:144], 'big'), int.from_bytes(body[144:192], 'big')
        if y0 >= P or y1 >= P:
            return SER_MALFORMED, None
        y = (y0, y1)
        if not E2.on_curve((x, y)):
            return SER_NOT_ON_CURVE, None
    if validate and not g2_in_subgroup((x, y)):
        return SER_NOT_IN_SUBGROUP, None
    return SER_OK, (x, y)

def curve_point_from_x(curve_is_g2, seed):
    """A point ON THE CURVE (in general NOT in the prime-order subgroup): first x >= seed-derived start with a root."""
    rng = SplitMix64(seed)
    while True:
        if curve_is_g2:
            x = (rng.next() * rng.next() * rng.next() % P, rng.next() * rng.next() * rng.next() % P)
            y = fp2_sqrt(fp2_add(fp2_mul(fp2_sqr(x), x), (4, 4)))
        else:
            x = rng.next() * rng.next() * rng.next() * rng.next() % P
            y = fp_sqrt((x * x * x + 4) % P)
        if y is not None:
            return (x, y)


# ------------------------------------------------- deterministic inputs ---
class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF
    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    def scalar(self):
        """Uniform in [0, r): 4 draws, clear top bit, rejection-sample."""
        while True:
            v = 0
            for i in range(4):
                v |= self.next() << (64 * i)
            v &= (1 << 255) - 1
            if v < R:
                return v


# ------------------------------------------------- Fr / NTT ------------
FR_GENERATOR = 7
FR_TWO_ADICITY = 32
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, (R - 1) >> FR_TWO_ADICITY, R)
# the constant ark-bls12-381 0.4 ships as FrConfig::TWO_ADIC_ROOT_OF_UNITY
assert FR_TWO_ADIC_ROOT == 10238227357739495823651030575849232062558860180284477541189508159991286009131

def fr_to_mont_bytes(a):
    return ((a * FR_R) % R).to_bytes(32, 'little')

def fr_from_mont_bytes(b):
    return (int.from_bytes(b, 'little') * pow(FR_R, R - 2, R)) % R

def fr_domain_generator(logn):
    return pow(FR_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - logn), R)

def fr_fft_definition(coeffs, logn, coset=False):
    """evals[i] = sum_j a_j (off * g^i)^j : ark_poly Radix2EvaluationDomain::fft (coset: offset 7)."""
    n = 1 << logn
    g = fr_domain_generator(logn)
    off = FR_GENERATOR if coset else 1
    out = []
    for i in range(n):
        x = off * pow(g, i, R) % R
        acc, xp = 0, 1
        for a in coeffs:
            acc = (acc + a * xp) % R
            xp = xp * x % R
        out.append(acc)
    return out
