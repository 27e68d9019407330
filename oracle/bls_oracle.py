"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the milagro_bls verification path.

This file is a plain big-int restatement of the reference's algorithm for the hot path
(SURVEY.md section 8).  It is imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg -- never by the product package `milagro_bls_b200`.

Abbreviations for citations:  A/ = /root/reference/incubator-milagro-crypto-rust/src/
                              M/ = /root/reference/src/

Parity status: PINNED for hash_to_curve_g2 (RFC 9380 vectors shipped by the reference,
A/test_utils/hash_to_curve_vectors/BLS12381G2_XMDSHA-256_SSWU_RO_.json, consumed by
A/bls381/core.rs:858-937) and for the compressed encodings (A/bls381/core.rs:1185-1225 =
M/src/amcl_utils.rs:83-144).  GT (Fp12) values are NOT pinned by any constant in the
reference; they are anchored on the accept/reject behaviour of the
reference's tests and on the independently known BLS12-381 GT generator coefficient
(tests/test_oracle.py::test_gt_generator_anchor).

Every value that crosses the reference's API is a canonical residue / affine point /
GT element, so plain `% p` arithmetic is used: the reference's Montgomery radix, lazy
reduction and window tables are not observable.
"""
import hashlib

from . import rom_constants as ROM

p = ROM.MODULUS                    # A/roms/rom_bls381_64.rs:28-36
r = ROM.CURVE_ORDER                # :70-78
BNX = ROM.CURVE_BNX                # :98   |x|, true x is negative (:168)
CRU = ROM.CURVE_CRU                # :100-108
G1_GEN = (ROM.CURVE_GX, ROM.CURVE_GY)
G2_GEN = ((ROM.CURVE_PXA, ROM.CURVE_PXB), (ROM.CURVE_PYA, ROM.CURVE_PYB))
DST_G2 = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"   # A/bls381/proof_of_possession.rs:38

G1_BYTES, G2_BYTES, MODBYTES = 48, 96, 48                  # A/bls381/core.rs:41-50
COMPRESSION_FLAG, INFINITY_FLAG, Y_FLAG = 0x80, 0x40, 0x20


class AmclError(Exception):
    """Mirror of A/errors.rs:1-11 (kind carried as a string)."""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


# ------------------------------------------------------------------------------------------
# Fp2 = Fp[i]/(i^2+1)                                                         A/fp2.rs
# ------------------------------------------------------------------------------------------
F2_ZERO, F2_ONE = (0, 0), (1, 0)


def f2_add(a, b):
    return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)


def f2_sub(a, b):
    return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)


def f2_neg(a):
    return ((-a[0]) % p, (-a[1]) % p)


def f2_conj(a):
    return (a[0], (-a[1]) % p)


def f2_mul(a, b):                      # A/fp2.rs:258-300
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def f2_sqr(a):                         # A/fp2.rs:237-255
    return ((a[0] + a[1]) * (a[0] - a[1]) % p, 2 * a[0] * a[1] % p)


def f2_muls(a, s):                     # pmul / imul
    return (a[0] * s % p, a[1] * s % p)


def f2_inv(a):                         # A/fp2.rs:370-383
    n = pow((a[0] * a[0] + a[1] * a[1]) % p, -1, p)
    return (a[0] * n % p, (-a[1]) * n % p)


def f2_mul_ip(a):                      # *(1+i)            A/fp2.rs:401-408
    return ((a[0] - a[1]) % p, (a[0] + a[1]) % p)


def f2_times_i(a):                     # *i
    return ((-a[1]) % p, a[0])


def f2_is_zero(a):
    return a[0] % p == 0 and a[1] % p == 0


def fp_sqrt(a):                        # A/fp.rs:717-729  (p = 3 mod 4)
    return pow(a, (p + 1) // 4, p)


def fp_is_qr(a):                       # A/fp.rs:733-737 jacobi()==1 ; (0 has jacobi 0)
    return pow(a, (p - 1) // 2, p) == 1


def f2_sqrt(a):
    """A/fp2.rs:304-339.  Returns (is_qr, root).  Which of the two roots is returned is fixed
    the same way as the reference (it is not observable on the hot path: every caller
    normalises the sign afterwards)."""
    if f2_is_zero(a):
        return True, F2_ZERO
    w1 = (a[0] * a[0] + a[1] * a[1]) % p
    if not fp_is_qr(w1):
        return False, F2_ZERO
    w1 = fp_sqrt(w1)
    inv2 = (p + 1) // 2
    w2 = (a[0] + w1) * inv2 % p
    if not fp_is_qr(w2):
        w2 = (a[0] - w1) * inv2 % p
        if not fp_is_qr(w2):
            return False, F2_ZERO
    ra = fp_sqrt(w2)
    rb = a[1] * pow(2 * ra % p, -1, p) % p
    return True, (ra, rb)


def fp_sgn0(a):                        # A/fp.rs:746-753
    return (a % p) & 1


def f2_sgn0(a):                        # A/fp2.rs:449-455
    return fp_sgn0(a[1]) if a[0] % p == 0 else fp_sgn0(a[0])


# ------------------------------------------------------------------------------------------
# Fp4 = Fp2[j]/(j^2-(1+i))                                                    A/fp4.rs
# ------------------------------------------------------------------------------------------
F4_ZERO, F4_ONE = (F2_ZERO, F2_ZERO), (F2_ONE, F2_ZERO)


def f4_add(a, b):
    return (f2_add(a[0], b[0]), f2_add(a[1], b[1]))


def f4_sub(a, b):
    return (f2_sub(a[0], b[0]), f2_sub(a[1], b[1]))


def f4_neg(a):
    return (f2_neg(a[0]), f2_neg(a[1]))


def f4_mul(a, b):                      # A/fp4.rs:275-308
    t0 = f2_mul(a[0], b[0])
    t1 = f2_mul(a[1], b[1])
    t2 = f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1]))
    return (f2_add(t0, f2_mul_ip(t1)), f2_sub(f2_sub(t2, t0), t1))


def f4_sqr(a):                         # A/fp4.rs:243-272
    return f4_mul(a, a)


def f4_times_i(a):                     # *j               A/fp4.rs:359-367
    return (f2_mul_ip(a[1]), a[0])


def f4_conj(a):                        # A/fp4.rs:184-188
    return (a[0], f2_neg(a[1]))


def f4_inv(a):                         # A/fp4.rs:340-356
    t = f2_inv(f2_sub(f2_sqr(a[0]), f2_mul_ip(f2_sqr(a[1]))))
    return (f2_mul(a[0], t), f2_neg(f2_mul(a[1], t)))


def f4_frob(a, f3):                    # A/fp4.rs:370-374
    return (f2_conj(a[0]), f2_mul(f2_conj(a[1]), f3))


def f4_pmul(a, s):                     # multiply both halves by an Fp2
    return (f2_mul(a[0], s), f2_mul(a[1], s))


# ------------------------------------------------------------------------------------------
# Fp12 = Fp4[w]/(w^3-j)                                                       A/fp12.rs
# ------------------------------------------------------------------------------------------
F12_ONE = (F4_ONE, F4_ZERO, F4_ZERO)


def f12_mul(x, y):                     # A/fp12.rs:300-366 (dense product; sparse variants 371-707
    a, b, c = x                        #  compute the same value)
    d, e, f = y
    ad, be, cf = f4_mul(a, d), f4_mul(b, e), f4_mul(c, f)
    r0 = f4_add(ad, f4_times_i(f4_add(f4_mul(b, f), f4_mul(c, e))))
    r1 = f4_add(f4_add(f4_mul(a, e), f4_mul(b, d)), f4_times_i(cf))
    r2 = f4_add(f4_add(f4_mul(a, f), f4_mul(c, d)), be)
    return (r0, r1, r2)


def f12_sqr(x):                        # A/fp12.rs:252-297
    return f12_mul(x, x)


def f12_conj(x):                       # A/fp12.rs:204-208
    return (f4_conj(x[0]), f4_neg(f4_conj(x[1])), f4_conj(x[2]))


def f12_inv(x):                        # A/fp12.rs:710-754
    a, b, c = x
    f0 = f4_sub(f4_sqr(a), f4_times_i(f4_mul(b, c)))
    f1 = f4_sub(f4_times_i(f4_sqr(c)), f4_mul(a, b))
    f2 = f4_sub(f4_sqr(b), f4_mul(a, c))
    f3 = f4_add(f4_mul(a, f0), f4_times_i(f4_add(f4_mul(c, f1), f4_mul(b, f2))))
    f3 = f4_inv(f3)
    return (f4_mul(f0, f3), f4_mul(f1, f3), f4_mul(f2, f3))


FROB = (ROM.FRA, ROM.FRB)              # A/roms/rom_bls381_64.rs:47-64 : (1+i)^((p-1)/6)


def f12_frob(x, f=FROB):               # A/fp12.rs:757-771
    f2 = f2_sqr(f)
    f3 = f2_mul(f2, f)
    return (f4_frob(x[0], f3), f4_pmul(f4_frob(x[1], f3), f), f4_pmul(f4_frob(x[2], f3), f2))


def f12_pow(x, e):                     # A/fp12.rs:954-980 (any chain gives the same value)
    res = F12_ONE
    for bit in bin(e)[2:]:
        res = f12_sqr(res)
        if bit == "1":
            res = f12_mul(res, x)
    return res


def f12_is_unity(x):                   # A/fp12.rs:162-165
    return x == F12_ONE


def f12_to_bytes(x):                   # A/fp12.rs:859-913 : a.a.a a.a.b a.b.a a.b.b b... c...
    out = b""
    for f4 in x:
        for f2 in f4:
            for c in f2:
                out += (c % p).to_bytes(MODBYTES, "big")
    return out


# ------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 4 over Fp.  Points are None (infinity) or affine (x, y).     A/ecp.rs
# ------------------------------------------------------------------------------------------
def g1_is_on_curve(P):
    return P is None or (P[1] * P[1] - P[0] ** 3 - 4) % p == 0


def g1_neg(P):                         # A/ecp.rs:294-304
    return None if P is None else (P[0], (-P[1]) % p)


def g1_add(P, Q):                      # group law computed by A/ecp.rs:743-819 (complete formulas)
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0:
            return None
        lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
    x3 = (lam * lam - P[0] - Q[0]) % p
    return (x3, (lam * (P[0] - x3) - P[1]) % p)


def g1_mul(P, e):                      # plain [e]P (A/ecp.rs:1074-1160 computes the same element)
    R = None
    for bit in bin(e)[2:] if e > 0 else "":
        R = g1_add(R, R)
        if bit == "1":
            R = g1_add(R, P)
    return R


def g1_phi(P):                         # (x,y) -> (CRU*x, y)   A/pair.rs:631-633, A/ecp.rs:309-311
    return None if P is None else (P[0] * CRU % p, P[1])


def _signed_split(u):
    """A/pair.rs:635-649 / 678-687: replace u by r-u and negate the point iff r-u has fewer bits.
    Returns (scalar, negate)."""
    t = r - (u % r)                    # Big::modneg, A/big.rs:1152-1156: for u == 0 this is r itself
    if t.bit_length() < u.bit_length():
        return t, True
    return u, False


def pair_g1mul(P, e):
    """A/pair.rs:625-656 with glv() 567-576 (BLS branch): u0 = e mod x^2, u1 = r - e div x^2,
    result = [+-u0]P + [+-u1]phi(P).  For P in G1 this equals [e]P."""
    x2 = BNX * BNX
    u0, u1 = e % x2, r - (e // x2)
    R, Q = P, g1_phi(P)
    u0, n0 = _signed_split(u0)
    if n0:
        R = g1_neg(R)
    u1, n1 = _signed_split(u1)
    if n1:
        Q = g1_neg(Q)
    return g1_add(g1_mul(R, u0), g1_mul(Q, u1))


def subgroup_check_g1(P):              # A/bls381/core.rs:116-120
    return pair_g1mul(P, r) is None


# ------------------------------------------------------------------------------------------
# G2: y^2 = x^3 + 4(1+i) over Fp2 (M-type twist).                              A/ecp2.rs
# ------------------------------------------------------------------------------------------
B2 = (4, 4)


def g2_rhs(x):                         # A/ecp2.rs:347-365
    return f2_add(f2_mul(f2_sqr(x), x), B2)


def g2_is_on_curve(P):
    return P is None or f2_sqr(P[1]) == g2_rhs(P[0])


def g2_neg(P):
    return None if P is None else (P[0], f2_neg(P[1]))


def g2_add(P, Q):                      # group law computed by A/ecp2.rs:368-527
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if f2_is_zero(f2_add(P[1], Q[1])):
            return None
        lam = f2_mul(f2_muls(f2_sqr(P[0]), 3), f2_inv(f2_muls(P[1], 2)))
    else:
        lam = f2_mul(f2_sub(Q[1], P[1]), f2_inv(f2_sub(Q[0], P[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), P[0]), Q[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(P[0], x3)), P[1]))


def g2_mul(P, e):                      # A/ecp2.rs:554-620 computes the same element
    R = None
    for bit in bin(e)[2:] if e > 0 else "":
        R = g2_add(R, R)
        if bit == "1":
            R = g2_add(R, P)
    return R


PSI_X = f2_inv(FROB)                   # A/ecp2.rs:785-789 / A/pair.rs:664-671 (M-type: 1/f)


def g2_frob(P, X=PSI_X):               # psi             A/ecp2.rs:538-548
    if P is None:
        return None
    X2 = f2_sqr(X)
    return (f2_mul(f2_conj(P[0]), X2), f2_mul(f2_mul(f2_conj(P[1]), X2), X))


def pair_g2mul(P, e):
    """A/pair.rs:661-693 with gs() 604-618 (BLS branch): base-|x| digits u0..u3, u1 and u3 replaced
    by modneg, Q[i] = psi^i(P), per-digit sign flip, then the joint multiplication ECP2::mul4
    (A/ecp2.rs:631-729) = sum of [u_i]Q_i."""
    u, w = [], e
    for _ in range(3):
        u.append(w % BNX)
        w //= BNX
    u.append(w)
    u[1] = r - (u[1] % r)              # Big::modneg, A/big.rs:1152-1156 (r - 0 = r)
    u[3] = r - (u[3] % r)
    Q = [P]
    for _ in range(3):
        Q.append(g2_frob(Q[-1]))
    acc = None
    for i in range(4):
        ui, neg = _signed_split(u[i])
        Qi = g2_neg(Q[i]) if neg else Q[i]
        acc = g2_add(acc, g2_mul(Qi, ui))
    return acc


def subgroup_check_g2(P):              # A/bls381/core.rs:123-127
    return pair_g2mul(P, r) is None


def g2_clear_cofactor(P):              # A/ecp2.rs:784-805 (BLS branch), Budroni-Pintore
    xQ = g2_mul(P, BNX)
    x2Q = g2_mul(xQ, BNX)
    xQ = g2_neg(xQ)                    # NegativeX
    x2Q = g2_add(x2Q, g2_neg(xQ))
    x2Q = g2_add(x2Q, g2_neg(P))
    xQ = g2_add(xQ, g2_neg(P))
    xQ = g2_frob(xQ)
    P2 = g2_frob(g2_frob(g2_add(P, P)))
    return g2_add(g2_add(P2, x2Q), xQ)


# ------------------------------------------------------------------------------------------
# hash_to_curve_g2                       A/hash_to_curve.rs, A/bls381/iso.rs, A/bls381/core.rs
# ------------------------------------------------------------------------------------------
H2C_L = 64                             # A/roms/rom_bls381_64.rs:175


def expand_message_xmd(msg, len_in_bytes, dst):      # A/hash_to_curve.rs:137-201
    ell = (len_in_bytes + 31) // 32
    if ell > 255:
        raise AmclError("HashToFieldError")
    if len(dst) > 255:
        dst_prime = hashlib.sha256(b"H2C-OVERSIZE-DST-" + dst).digest() + bytes([32])
    else:
        dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + len_in_bytes.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    b = [b0, hashlib.sha256(b0 + b"\x01" + dst_prime).digest()]
    out = b[1]
    for i in range(2, ell + 1):
        t = bytes(x ^ y for x, y in zip(b0, b[i - 1]))
        b.append(hashlib.sha256(t + bytes([i]) + dst_prime).digest())
        out += b[i]
    return out[:len_in_bytes]


def hash_to_field_fp2(msg, count, dst):              # A/hash_to_curve.rs:111-131
    prb = expand_message_xmd(msg, count * 2 * H2C_L, dst)
    u = []
    for i in range(count):
        e = []
        for j in range(2):
            off = H2C_L * (j + i * 2)
            e.append(int.from_bytes(prb[off:off + H2C_L], "big") % p)   # A/dbig.rs:174-207,294-310
        u.append((e[0], e[1]))
    return u


SSWU_A = (ROM.SSWU_A2_A, ROM.SSWU_A2_B)              # A/roms/rom_bls381_64.rs:202-223
SSWU_B = (ROM.SSWU_B2_A, ROM.SSWU_B2_B)
SSWU_Z = (ROM.SSWU_Z2_A, ROM.SSWU_Z2_B)


def simplified_swu_fp2(u):                           # A/hash_to_curve.rs:283-346
    tmp1 = f2_mul(f2_sqr(u), SSWU_Z)
    tv1 = f2_add(f2_sqr(tmp1), tmp1)
    a_inv = f2_inv(SSWU_A)
    if f2_is_zero(tv1):
        x = f2_mul(f2_mul(f2_inv(SSWU_Z), SSWU_B), a_inv)
    else:
        tv1 = f2_inv(tv1)
        x = f2_mul(f2_neg(f2_mul(f2_add(tv1, F2_ONE), SSWU_B)), a_inv)

    def g(xx):
        return f2_add(f2_mul(f2_add(f2_sqr(xx), SSWU_A), xx), SSWU_B)

    ok, y = f2_sqrt(g(x))
    if not ok:
        x = f2_mul(x, tmp1)
        ok, y = f2_sqrt(g(x))
        assert ok, "Hash to Curve SSWU failure - no square roots"
    if f2_sgn0(u) != f2_sgn0(y):
        y = f2_neg(y)
    return x, y


def iso3_to_ecp2(x, y):                              # A/bls381/iso.rs:177-206
    vals = []
    for coeffs in (ROM.ISO3_XNUM, ROM.ISO3_XDEN, ROM.ISO3_YNUM, ROM.ISO3_YDEN):
        v = coeffs[-1]
        for k in reversed(coeffs[:-1]):
            v = f2_add(f2_mul(v, x), k)
        vals.append(v)
    xn, xd, yn, yd = vals
    yn = f2_mul(yn, y)
    z = f2_mul(xd, yd)
    if f2_is_zero(z):
        return None
    zi = f2_inv(z)
    return (f2_mul(f2_mul(xn, yd), zi), f2_mul(f2_mul(yn, xd), zi))


def map_to_curve_g2(u):                              # A/bls381/core.rs:846-849
    return iso3_to_ecp2(*simplified_swu_fp2(u))


def hash_to_curve_g2(msg, dst=DST_G2):               # A/bls381/core.rs:831-839, M/src/amcl_utils.rs:33-35
    u = hash_to_field_fp2(msg, 2, dst)
    q0 = map_to_curve_g2(u[0])
    q1 = map_to_curve_g2(u[1])
    return g2_clear_cofactor(g2_add(q0, q1))


# ------------------------------------------------------------------------------------------
# Pairing                                                                     A/pair.rs
# ------------------------------------------------------------------------------------------
def _line(a_w0, a_w3, c_w5):
    """FP12::new_fp4s(FP4(a_w0, a_w3), 0, FP4(c_w5).times_i())   (A/pair.rs:71-83, 119-131)."""
    return ((a_w0, a_w3), F4_ZERO, f4_times_i((c_w5, F2_ZERO)))


def _linedbl(A, qx, qy):                             # A/pair.rs:35-84 with Z = 1 (A kept affine:
    X, Y = A                                         #  the Fp2 scale of a line dies in fexp, B.3)
    yz = f2_mul_ip(f2_muls(f2_neg(f2_muls(Y, 4)), qy))
    xx = f2_muls(f2_muls(f2_sqr(X), 6), qx)
    zz = f2_muls(f2_mul_ip(F2_ONE), 3 * 4 * 2)       # 3*b*Z^2 *(1+i) *2
    zz = f2_sub(zz, f2_muls(f2_sqr(Y), 2))
    return _line(yz, zz, xx), g2_add(A, A)


def _lineadd(A, B, qx, qy):                          # A/pair.rs:88-133 with Z1 = 1
    x1 = f2_sub(A[0], B[0])
    y1 = f2_sub(A[1], B[1])
    t1 = f2_mul(x1, B[1])
    t2 = f2_sub(f2_mul(y1, B[0]), t1)
    return _line(f2_mul_ip(f2_muls(x1, qy)), t2, f2_neg(f2_muls(y1, qx))), g2_add(A, B)


ATE_BITS = 65                                        # A/roms/rom_bls381_64.rs:166


def initmp():                                        # A/pair.rs:156-162
    return [F12_ONE] * ATE_BITS


def another(rr, P, Q):                               # A/pair.rs:182-238 (P in G2 affine, Q in G1 affine)
    """A pair with infinity on either side contributes only subfield factors in the reference
    (SURVEY.md B.5); here it is skipped, which yields the same GT after fexp."""
    if P is None or Q is None:
        return
    qx, qy = Q
    n, n3 = BNX, 3 * BNX
    nb = n3.bit_length()
    A, NP = P, g2_neg(P)
    for i in range(nb - 2, 0, -1):
        lv, A = _linedbl(A, qx, qy)
        bt = ((n3 >> i) & 1) - ((n >> i) & 1)
        if bt == 1:
            lv2, A = _lineadd(A, P, qx, qy)
            lv = f12_mul(lv, lv2)
        if bt == -1:
            lv2, A = _lineadd(A, NP, qx, qy)
            lv = f12_mul(lv, lv2)
        rr[i] = f12_mul(rr[i], lv)


def miller(rr):                                      # A/pair.rs:166-178
    res = F12_ONE
    for i in range(ATE_BITS - 1, 0, -1):
        res = f12_mul(f12_sqr(res), rr[i])
    res = f12_conj(res)                              # NegativeX
    return f12_mul(res, rr[0])


def fexp(m):                                         # A/pair.rs:409-541 (BLS branch 485-539)
    x = BNX
    rr = f12_mul(f12_conj(m), f12_inv(m))
    rr = f12_mul(f12_frob(f12_frob(rr)), rr)

    def powx(v, e):                                  # pow(|x|) then conj (NegativeX)
        return f12_conj(f12_pow(v, e))

    y0 = f12_sqr(rr)
    y1 = powx(y0, x)
    y2 = powx(y1, x >> 1)
    y3 = f12_conj(rr)
    y1 = f12_mul(y1, y3)
    y1 = f12_mul(f12_conj(y1), y2)
    y2 = powx(y1, x)
    y3 = powx(y2, x)
    y1 = f12_conj(y1)
    y3 = f12_mul(y3, y1)
    y1 = f12_conj(y1)
    y1 = f12_frob(f12_frob(f12_frob(y1)))
    y2 = f12_frob(f12_frob(y2))
    y1 = f12_mul(y1, y2)
    y2 = powx(y3, x)
    y2 = f12_mul(f12_mul(y2, y0), rr)
    y1 = f12_mul(y1, y2)
    y1 = f12_mul(y1, f12_frob(y3))
    return y1


def ate2(P1, Q1, R1, S1):                            # A/pair.rs:313-405 == product of two Miller loops
    rr = initmp()
    another(rr, P1, Q1)
    another(rr, R1, S1)
    return miller(rr)


def ate2_evaluation(a, b, c, d):                     # M/src/amcl_utils.rs:38-42
    return fexp(ate2(a, b, c, d)) == F12_ONE


# ------------------------------------------------------------------------------------------
# ZCash (de)serialisation                                   A/bls381/core.rs:131-486
# ------------------------------------------------------------------------------------------
def _chk_inf(b, kind="InvalidPoint"):
    if b[0] & 0x3F or any(b[1:]):
        raise AmclError(kind)


def serialize_g1(P):                                 # core.rs:145-172
    if P is None:
        return bytes([COMPRESSION_FLAG | INFINITY_FLAG]) + bytes(47)
    out = bytearray(P[0].to_bytes(48, "big"))
    if P[1] > (-P[1]) % p:
        out[0] |= Y_FLAG
    out[0] |= COMPRESSION_FLAG
    return bytes(out)


def serialize_uncompressed_g1(P):                    # core.rs:177-190
    if P is None:
        return bytes([INFINITY_FLAG]) + bytes(95)
    return P[0].to_bytes(48, "big") + P[1].to_bytes(48, "big")


def deserialize_g1(b):                               # core.rs:195-307
    if len(b) == 0:
        raise AmclError("InvalidG1Size")
    if b[0] & COMPRESSION_FLAG == 0:
        if len(b) != 96:
            raise AmclError("InvalidG1Size")
        if b[0] & INFINITY_FLAG:
            _chk_inf(b)
            return None
        if b[0] & Y_FLAG:
            raise AmclError("InvalidYFlag")
        x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
        y = int.from_bytes(b[48:], "big")
        if x >= p or y >= p or not g1_is_on_curve((x, y)):
            raise AmclError("InvalidPoint")
        return (x, y)
    if len(b) != 48:
        raise AmclError("InvalidG1Size")
    if b[0] & INFINITY_FLAG:
        _chk_inf(b)
        return None
    yflag = bool(b[0] & Y_FLAG)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= p:
        raise AmclError("InvalidPoint")
    rhs = (x * x * x + 4) % p
    if not fp_is_qr(rhs):                            # ECP::new_big, A/ecp.rs:141-155 (rhs != 0 on this curve)
        raise AmclError("InvalidPoint")
    y = fp_sqrt(rhs)
    if (y > (-y) % p) != yflag:
        y = (-y) % p
    return (x, y)


def _cmp_fp2(a, b):                                  # core.rs:131-140 (im first, then re)
    ka, kb = (a[1], a[0]), (b[1], b[0])
    return (ka > kb) - (ka < kb)


def serialize_g2(P):                                 # core.rs:312-339
    if P is None:
        return bytes([COMPRESSION_FLAG | INFINITY_FLAG]) + bytes(95)
    out = bytearray(P[0][1].to_bytes(48, "big") + P[0][0].to_bytes(48, "big"))
    if _cmp_fp2(P[1], f2_neg(P[1])) > 0:
        out[0] |= Y_FLAG
    out[0] |= COMPRESSION_FLAG
    return bytes(out)


def serialize_uncompressed_g2(P):                    # core.rs:344-364
    if P is None:
        return bytes([INFINITY_FLAG]) + bytes(191)
    return b"".join(v.to_bytes(48, "big") for v in (P[0][1], P[0][0], P[1][1], P[1][0]))


def deserialize_g2(b):                               # core.rs:369-486
    if len(b) == 0:
        raise AmclError("InvalidG2Size")
    if b[0] & COMPRESSION_FLAG == 0:
        if len(b) != 192:
            raise AmclError("InvalidG2Size")
        if b[0] & INFINITY_FLAG:
            _chk_inf(b)
            return None
        if b[0] & Y_FLAG:
            raise AmclError("InvalidYFlag")
        v = [int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")] + \
            [int.from_bytes(b[48 * k:48 * k + 48], "big") for k in (1, 2, 3)]
        if any(c >= p for c in v):
            raise AmclError("InvalidPoint")
        P = ((v[1], v[0]), (v[3], v[2]))
        if not g2_is_on_curve(P):
            raise AmclError("InvalidPoint")
        return P
    if len(b) != 96:
        raise AmclError("InvalidG2Size")
    if b[0] & INFINITY_FLAG:
        _chk_inf(b)
        return None
    yflag = bool(b[0] & Y_FLAG)
    xim = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    xre = int.from_bytes(b[48:], "big")
    if xim >= p or xre >= p:
        raise AmclError("InvalidPoint")
    x = (xre, xim)
    ok, y = f2_sqrt(g2_rhs(x))                       # ECP2::new_fp2, A/ecp2.rs:103-116
    if not ok:
        raise AmclError("InvalidPoint")
    if (_cmp_fp2(y, f2_neg(y)) > 0) != yflag:
        y = f2_neg(y)
    return (x, y)


# ------------------------------------------------------------------------------------------
# milagro_bls API semantics                    M/src/{amcl_utils,keys,signature,aggregates}.rs
# ------------------------------------------------------------------------------------------
def decompress_g1(b):                                # M/src/amcl_utils.rs:52-58
    if len(b) != G1_BYTES:
        raise AmclError("InvalidG1Size")
    return deserialize_g1(b)


def decompress_g2(b):                                # M/src/amcl_utils.rs:68-74
    if len(b) != G2_BYTES:
        raise AmclError("InvalidG2Size")
    return deserialize_g2(b)


def key_validate(P):                                 # M/src/keys.rs:181-186
    return P is not None and subgroup_check_g1(P)


def public_key_from_bytes(b):                        # M/src/keys.rs:140-147
    P = decompress_g1(b)
    if not key_validate(P):
        raise AmclError("InvalidPoint")
    return P


def sk_to_pk(sk):                                    # M/src/keys.rs:126-131 (test-input synthesis only)
    return g1_mul(G1_GEN, sk % r)


def sign(sk, msg):                                   # M/src/signature.rs:17-21 (test-input synthesis only)
    return g2_mul(hash_to_curve_g2(msg), sk % r)


def aggregate_public_keys(pks):                      # M/src/aggregates.rs:29-56
    if len(pks) == 0:
        raise AmclError("AggregateEmptyPoints")
    acc = None
    for P in pks:
        acc = g1_add(acc, P)
    return acc


def aggregate_signatures(sigs):                      # M/src/aggregates.rs:100-106
    acc = None
    for S in sigs:
        acc = g2_add(acc, S)
    return acc


NEG_G1 = g1_neg(G1_GEN)


def signature_verify(sig, msg, pk, want_gt=False):   # M/src/signature.rs:27-40
    if not subgroup_check_g2(sig):
        return (False, None) if want_gt else False
    gt = fexp(ate2(sig, NEG_G1, hash_to_curve_g2(msg), pk))
    ok = gt == F12_ONE
    return (ok, gt) if want_gt else ok


def fast_aggregate_verify_pre_aggregated(sig, msg, apk, want_gt=False):   # M/src/aggregates.rs:223-253
    if not subgroup_check_g2(sig) or apk is None:
        return (False, None) if want_gt else False
    gt = fexp(ate2(sig, NEG_G1, hash_to_curve_g2(msg), apk))
    ok = gt == F12_ONE
    return (ok, gt) if want_gt else ok


def fast_aggregate_verify(sig, msg, pks, want_gt=False):                  # M/src/aggregates.rs:177-215
    if len(pks) == 0:
        return (False, None) if want_gt else False
    if not subgroup_check_g2(sig):
        return (False, None) if want_gt else False
    return fast_aggregate_verify_pre_aggregated(sig, msg, aggregate_public_keys(pks), want_gt)


def aggregate_verify(sig, msgs, pks, want_gt=False):                      # M/src/aggregates.rs:130-170
    if len(msgs) != len(pks) or len(pks) == 0 or not subgroup_check_g2(sig):
        return (False, None) if want_gt else False
    rr = initmp()
    for m, pk in zip(msgs, pks):
        another(rr, hash_to_curve_g2(m), pk)
    another(rr, sig, NEG_G1)
    gt = fexp(miller(rr))
    ok = f12_is_unity(gt)
    return (ok, gt) if want_gt else ok


def draw_scalar(rng_fill):                           # M/src/aggregates.rs:278-287
    """rng_fill(n) -> n bytes.  8 bytes big-endian -> i64 -> abs(); zero is redrawn.
    (i64::MIN, probability 2^-64, is out of contract -- SURVEY.md C.3 -- and is redrawn.)"""
    while True:
        v = int.from_bytes(rng_fill(8), "big", signed=True)
        if v == -(1 << 63):
            continue
        v = abs(v)
        if v != 0:
            return v


def verify_multiple_aggregate_signatures(rng_fill, sets, want_gt=False):  # M/src/aggregates.rs:261-316
    """sets: iterable of (sig_point, apk_point, msg_bytes)."""
    final_agg_sig = None
    rr = initmp()
    for sig, apk, msg in sets:
        if not subgroup_check_g2(sig):
            return (False, None) if want_gt else False
        c = draw_scalar(rng_fill)
        another(rr, hash_to_curve_g2(msg), pair_g1mul(apk, c))
        final_agg_sig = g2_add(final_agg_sig, pair_g2mul(sig, c))
    another(rr, final_agg_sig, NEG_G1)
    gt = fexp(miller(rr))
    ok = f12_is_unity(gt)
    return (ok, gt) if want_gt else ok


class SeededRng:
    """Deterministic byte stream (SHA-256 in counter mode) standing in for the caller-injected
    `rand::Rng` of M/src/aggregates.rs:261.  The same stream is implemented on the host side of
    the product (milagro_bls_b200.rng) so both draw identical scalars."""

    def __init__(self, seed: bytes):
        self.seed, self.ctr, self.buf = seed, 0, b""

    def fill(self, n):
        while len(self.buf) < n:
            self.buf += hashlib.sha256(self.seed + self.ctr.to_bytes(8, "big")).digest()
            self.ctr += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        return out
