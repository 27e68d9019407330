// G1 (y^2 = x^3 + 4 over Fp) and G2 (y^2 = x^3 + 4(1+i) over Fp2, M-type twist) in Jacobian coordinates.
//
// Replaces the reference's homogeneous-projective complete formulas
//   /root/reference/incubator-milagro-crypto-rust/src/ecp.rs:552-592,743-819   (G1 dbl/add)
//   /root/reference/incubator-milagro-crypto-rust/src/ecp2.rs:368-527,538-548 (G2 dbl/add/frob)
// The coordinate system is not observable; exceptional inputs (infinity, P+P, P-P)
// are handled explicitly so the group element always equals the reference's.
#pragma once
#include "tower.cuh"

// ---- uniform field interface over fp / fp2 -----------------------------------------------------
B3_FN void f_add(fp& r, const fp& a, const fp& b) { fp_add(r, a, b); }
B3_FN void f_sub(fp& r, const fp& a, const fp& b) { fp_sub(r, a, b); }
B3_FN void f_mul(fp& r, const fp& a, const fp& b) { fp_mul(r, a, b); }
B3_FN void f_sqr(fp& r, const fp& a) { fp_sqr(r, a); }
B3_FN void f_dbl(fp& r, const fp& a) { fp_dbl(r, a); }
B3_FN void f_neg(fp& r, const fp& a) { fp_neg(r, a); }
B3_FN void f_inv(fp& r, const fp& a) { fp_inv(r, a); }
B3_FN bool f_is_zero(const fp& a) { return fp_is_zero(a); }
B3_FN bool f_eq(const fp& a, const fp& b) { return fp_eq(a, b); }
B3_FN void f_select(fp& r, bool c, const fp& a, const fp& b) { fp_select(r, c, a, b); }
B3_FN void f_one(fp& r) { r = FP_ONE; }
B3_FN void f_zero(fp& r) { r = FP_NIL; }

B3_FN void f_add(fp2& r, const fp2& a, const fp2& b) { fp2_add(r, a, b); }
B3_FN void f_sub(fp2& r, const fp2& a, const fp2& b) { fp2_sub(r, a, b); }
B3_FN void f_mul(fp2& r, const fp2& a, const fp2& b) { fp2_mul(r, a, b); }
B3_FN void f_sqr(fp2& r, const fp2& a) { fp2_sqr(r, a); }
B3_FN void f_dbl(fp2& r, const fp2& a) { fp2_dbl(r, a); }
B3_FN void f_neg(fp2& r, const fp2& a) { fp2_neg(r, a); }
B3_FN void f_inv(fp2& r, const fp2& a) { fp2_inv(r, a); }
B3_FN bool f_is_zero(const fp2& a) { return fp2_is_zero(a); }
B3_FN bool f_eq(const fp2& a, const fp2& b) { return fp2_eq(a, b); }
B3_FN void f_select(fp2& r, bool c, const fp2& a, const fp2& b) { fp2_select(r, c, a, b); }
B3_FN void f_one(fp2& r) { fp2_one(r); }
B3_FN void f_zero(fp2& r) { fp2_zero(r); }

// ---- PAIRED products ------------------------------------------------------------------------------------------------
// Two INDEPENDENT products of a formula, named as a pair.  For the one-thread (fp, fp2) and lane-pair (fp2h) representations
// they simply run one after the other; the replicated representations of quad.cuh (fpd: an Fp value held by both lanes of a
// lane pair, fp2q: a lane-pair Fp2 value held by both pairs of a lane quad) run them AT THE SAME TIME, one on each half of
// the group, and exchange the results -- twice the threads per item for the latency-bound chains of a single call.
// The point formulas below are ordered into such pairs.  No output may alias an input of the other product.
template <class F>
B3_FN void f_mul_par(F& r1, const F& a1, const F& b1, F& r2, const F& a2, const F& b2) { f_mul(r1, a1, b1); f_mul(r2, a2, b2); }
template <class F>
B3_FN void f_sqr_par(F& r1, const F& a1, F& r2, const F& a2) { f_sqr(r1, a1); f_sqr(r2, a2); }
template <class F>
B3_FN void f_mulsqr_par(F& r1, const F& a1, const F& b1, F& r2, const F& a2) { f_mul(r1, a1, b1); f_sqr(r2, a2); }

// curve constant b: 4 (G1) / 4(1+i) (G2)
B3_FN void f_mul_b(fp& r, const fp& a) { fp t; fp_dbl(t, a); fp_dbl(r, t); }
B3_FN void f_mul_b(fp2& r, const fp2& a) { fp2 t; fp2_dbl(t, a); fp2_dbl(t, t); fp2_mul_xi(r, t); }

template <class F>
struct jac {          // (X : Y : Z), x = X/Z^2, y = Y/Z^3; infinity <=> Z == 0
    F x, y, z;
};
template <class F>
struct aff {          // affine point; inf != 0 marks the point at infinity (x, y ignored)
    F x, y;
    uint32_t inf;
};
typedef jac<fp> g1_jac;
typedef aff<fp> g1_aff;
typedef jac<fp2> g2_jac;
typedef aff<fp2> g2_aff;

template <class F>
B3_FN void pt_set_inf(jac<F>& r) { f_one(r.x); f_one(r.y); f_zero(r.z); }
template <class F>
B3_FN bool pt_is_inf(const jac<F>& a) { return f_is_zero(a.z); }
template <class F>
B3_FN void pt_from_aff(jac<F>& r, const aff<F>& a) {
    if (a.inf) { pt_set_inf(r); return; }
    r.x = a.x; r.y = a.y; f_one(r.z);
}
template <class F>
B3_FN void pt_neg(jac<F>& r, const jac<F>& a) { r.x = a.x; f_neg(r.y, a.y); r.z = a.z; }
template <class F>
B3_FN void pt_select(jac<F>& r, bool c, const jac<F>& a, const jac<F>& b) {
    f_select(r.x, c, a.x, b.x); f_select(r.y, c, a.y, b.y); f_select(r.z, c, a.z, b.z);
}

// dbl-2009-l (a = 0): 2M + 5S.  Z = 0 stays Z = 0.  (Neither curve has 2-torsion, so Y != 0 on finite points.)
// Paired: (X^2, Y^2), (B^2, (X+B)^2), (Y Z, E^2), E (D - X3): four rounds instead of seven products.
template <class F>
B3_FN_NOINLINE void pt_dbl(jac<F>& r, const jac<F>& p) {
    F A, B, C, D, E, Fq, t, yz;
    f_sqr_par(A, p.x, B, p.y);
    f_add(t, p.x, B);
    f_sqr_par(C, B, D, t);
    f_sub(D, D, A);
    f_sub(D, D, C);
    f_dbl(D, D);
    f_dbl(E, A);
    f_add(E, E, A);
    f_mulsqr_par(yz, p.y, p.z, Fq, E);
    f_dbl(r.z, yz);
    f_dbl(t, D);
    f_sub(r.x, Fq, t);
    f_sub(t, D, r.x);
    f_mul(t, E, t);
    f_dbl(C, C); f_dbl(C, C); f_dbl(C, C);
    f_sub(r.y, t, C);
}

// add-2007-bl with explicit handling of infinity / doubling / inverse operands: 11M + 5S in eight paired rounds
template <class F>
B3_FN_NOINLINE void pt_add(jac<F>& r, const jac<F>& p, const jac<F>& q) {
    bool pinf = pt_is_inf(p), qinf = pt_is_inf(q);
    F Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, t2;
    f_sqr_par(Z1Z1, p.z, Z2Z2, q.z);
    f_mul_par(U1, p.x, Z2Z2, U2, q.x, Z1Z1);
    f_mul_par(t, q.z, Z2Z2, t2, p.z, Z1Z1);
    f_mul_par(S1, p.y, t, S2, q.y, t2);
    f_sub(H, U2, U1);
    f_sub(rr, S2, S1);
    if (!pinf && !qinf && f_is_zero(H) && f_is_zero(rr)) {      // P == Q
        pt_dbl(r, p);
        return;
    }
    jac<F> o;
    f_dbl(rr, rr);
    f_dbl(I, H);
    f_add(t, p.z, q.z);
    f_sqr_par(I, I, t2, t);                                      // I = (2H)^2, t2 = (Z1 + Z2)^2
    f_sub(t2, t2, Z1Z1);
    f_sub(t2, t2, Z2Z2);
    f_mul_par(J, H, I, V, U1, I);
    f_mulsqr_par(S2, S1, J, o.x, rr);                            // S2 = S1 J, o.x = rr^2
    f_sub(o.x, o.x, J);
    f_dbl(t, V);
    f_sub(o.x, o.x, t);
    f_sub(t, V, o.x);
    f_mul_par(U2, rr, t, o.z, t2, H);                            // H == 0, r != 0  =>  Z3 = 0 (infinity)
    f_dbl(S2, S2);
    f_sub(o.y, U2, S2);
    pt_select(o, qinf, p, o);
    pt_select(r, pinf, q, o);
}

// madd-2007-bl (q affine, finite or flagged infinity): 7M + 4S in six paired rounds
template <class F>
B3_FN_NOINLINE void pt_add_aff(jac<F>& r, const jac<F>& p, const aff<F>& q) {
    bool pinf = pt_is_inf(p), qinf = q.inf != 0;
    F Z1Z1, U2, S2, H, HH, I, J, rr, V, t, t2;
    f_sqr(Z1Z1, p.z);
    f_mul_par(U2, q.x, Z1Z1, t, p.z, Z1Z1);
    f_sub(H, U2, p.x);
    f_mulsqr_par(S2, q.y, t, HH, H);
    f_sub(rr, S2, p.y);
    if (!pinf && !qinf && f_is_zero(H) && f_is_zero(rr)) {
        pt_dbl(r, p);
        return;
    }
    jac<F> o, qj;
    f_dbl(rr, rr);
    f_dbl(I, HH); f_dbl(I, I);
    f_mul_par(J, H, I, V, p.x, I);
    f_add(t, p.z, H);
    f_sqr_par(o.x, rr, t2, t);                                   // o.x = rr^2, t2 = (Z1 + H)^2
    f_sub(o.x, o.x, J);
    f_dbl(t, V);
    f_sub(o.x, o.x, t);
    f_sub(t, V, o.x);
    f_mul_par(U2, rr, t, S2, p.y, J);
    f_dbl(S2, S2);
    f_sub(o.y, U2, S2);
    f_sub(t2, t2, Z1Z1);
    f_sub(o.z, t2, HH);
    qj.x = q.x; qj.y = q.y; f_one(qj.z);
    pt_select(o, qinf, p, o);
    pt_select(r, pinf, qj, o);
    if (pinf && qinf) pt_set_inf(r);
}

template <class F>
B3_FN_NOINLINE void pt_to_aff(aff<F>& r, const jac<F>& p) {
    if (pt_is_inf(p)) { f_zero(r.x); f_zero(r.y); r.inf = 1; return; }
    F zi, zi2;
    f_inv(zi, p.z);
    f_sqr(zi2, zi);
    f_mul(r.x, p.x, zi2);
    f_mul(zi2, zi2, zi);
    f_mul(r.y, p.y, zi2);
    r.inf = 0;
}

// projective equality (reference: A/ecp.rs:339-357, A/ecp2.rs:182-200)
template <class F>
B3_FN_NOINLINE bool pt_eq(const jac<F>& a, const jac<F>& b) {
    bool ai = pt_is_inf(a), bi = pt_is_inf(b);
    if (ai || bi) return ai && bi;
    F za, zb, t0, t1;
    f_sqr(za, a.z);
    f_sqr(zb, b.z);
    f_mul(t0, a.x, zb);
    f_mul(t1, b.x, za);
    if (!f_eq(t0, t1)) return false;
    f_mul(za, za, a.z);
    f_mul(zb, zb, b.z);
    f_mul(t0, a.y, zb);
    f_mul(t1, b.y, za);
    return f_eq(t0, t1);
}

// y^2 == x^3 + b
template <class F>
B3_FN_NOINLINE bool pt_on_curve_aff(const aff<F>& a) {
    if (a.inf) return true;
    F l, rr, one;
    f_sqr(l, a.y);
    f_sqr(rr, a.x);
    f_mul(rr, rr, a.x);
    f_one(one);
    f_mul_b(one, one);
    f_add(rr, rr, one);
    return f_eq(l, rr);
}

// [k]P, k a 64-bit scalar, left-to-right double-and-add (k is public on this path)
template <class F>
B3_FN_NOINLINE void pt_mul_u64(jac<F>& r, const jac<F>& p, uint64_t k) {
    jac<F> acc;
    pt_set_inf(acc);
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        if (started) pt_dbl(acc, acc);
        if ((k >> i) & 1) {
            if (started) pt_add(acc, acc, p);
            else { acc = p; started = true; }
        }
    }
    r = acc;
}
// [k]P, k a 64-bit scalar, SIGNED 4-BIT FIXED WINDOWS: k = sum_i d_i 16^i + c 16^16 with d_i in [-8, 7] (the nibbles of
// k + 0x88..8 minus 8, c = the carry out of bit 64).  Every thread of a warp adds at the same steps -- with one thread per
// set and a different scalar in every lane, double-and-add runs its additions at half the lanes (each bit is set in
// about half of them) -- and the work drops from 63 doublings + ~32 additions to a table of [1..8]P (4 doublings +
// 3 additions), 64 doublings and <= 17 additions.  pt_add is complete (infinity, doubling, inverse operands), so the
// result is [k]P for ANY point of the curve, also outside the prime-order subgroup.
template <class F>
B3_FN_NOINLINE void pt_mul_u64_w4(jac<F>& r, const jac<F>& p, uint64_t k) {
    jac<F> tbl[8];                                  // tbl[j] = [j + 1] P
    tbl[0] = p;
    pt_dbl(tbl[1], p);
    pt_add(tbl[2], tbl[1], p);
    pt_dbl(tbl[3], tbl[1]);
    pt_add(tbl[4], tbl[3], p);
    pt_dbl(tbl[5], tbl[2]);
    pt_add(tbl[6], tbl[5], p);
    pt_dbl(tbl[7], tbl[3]);
    const uint64_t off = 0x8888888888888888ull;
    const uint64_t ks = k + off;
    const int carry = ks < k ? 1 : 0;
    jac<F> acc;
    pt_set_inf(acc);
    if (carry) acc = p;
    for (int i = 15; i >= 0; i--) {
        pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc);
        const int d = (int)((ks >> (4 * i)) & 15u) - 8;
        if (d != 0) {
            jac<F> q = tbl[(d < 0 ? -d : d) - 1];
            if (d < 0) pt_neg(q, q);
            pt_add(acc, acc, q);
        }
    }
    r = acc;
}
// [k]P for affine P (mixed additions)
template <class F>
B3_FN_NOINLINE void pt_mul_u64_aff(jac<F>& r, const aff<F>& p, uint64_t k) {
    jac<F> acc;
    pt_set_inf(acc);
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        if (started) pt_dbl(acc, acc);
        if ((k >> i) & 1) {
            if (started) pt_add_aff(acc, acc, p);
            else { pt_from_aff(acc, p); started = true; }
        }
    }
    r = acc;
}
// [k]P for a 256-bit scalar given as 8 little-endian u32 words (used for input synthesis and signing-side helpers)
template <class F>
B3_FN_NOINLINE void pt_mul_u256_aff(jac<F>& r, const aff<F>& p, const uint32_t* k) {
    jac<F> acc;
    pt_set_inf(acc);
    for (int i = 255; i >= 0; i--) {
        pt_dbl(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) pt_add_aff(acc, acc, p);
    }
    r = acc;
}

// ---- endomorphisms ------------------------------------------------------------------------------
// (templates over the Fp2 representation F2: fp2 = one thread per value, fp2h = lane pairs, see fp2h.cuh)
// psi on G2 (untwist-Frobenius-twist), Jacobian: (conj X * cx, conj Y * cy, conj Z)
//   reference: A/ecp2.rs:538-548 with X = 1/FROB (A/ecp2.rs:785-789)
template <class F2>
B3_FN_NOINLINE void g2_psi(jac<F2>& r, const jac<F2>& p) {
    F2 t, c;
    f2_const(c, PSI_CX);
    fp2_conj(t, p.x); fp2_mul(r.x, t, c);
    f2_const(c, PSI_CY);
    fp2_conj(t, p.y); fp2_mul(r.y, t, c);
    fp2_conj(r.z, p.z);
}
template <class F2>
B3_FN void g2_psi2(jac<F2>& r, const jac<F2>& p) {
    fp2_mul_fp(r.x, p.x, PSI2_CX);
    fp2_mul_fp(r.y, p.y, PSI2_CY);
    r.z = p.z;
}
// phi on G1: (beta x, y); acts as [-x^2] on G1
B3_FN void g1_phi(g1_jac& r, const g1_jac& p) { fp_mul(r.x, p.x, FP_BETA); r.y = p.y; r.z = p.z; }

// Subgroup membership.  The reference tests [r]P == O through its GLV/GS ladders
// (A/bls381/core.rs:116-127 -> A/pair.rs:625-693); any exact membership test gives the same answer on
// every on-curve input.  G2: psi(P) == [x]P = -[|x|]P.  G1: phi(P) == [-x^2]P.
// (signatures are parsed to affine form: the five additions of the [|x|] ladder are mixed, 7M + 4S instead of 11M + 5S)
template <class F2>
B3_FN_NOINLINE bool g2_in_subgroup_aff(const aff<F2>& a) {
    if (a.inf) return true;
    jac<F2> xp, p, ps;
    pt_mul_u64_aff(xp, a, B3_X_ABS);
    pt_neg(xp, xp);
    pt_from_aff(p, a);
    g2_psi(ps, p);
    return pt_eq(xp, ps);
}
B3_FN_NOINLINE bool g1_in_subgroup(const g1_jac& p) {
    if (pt_is_inf(p)) return true;
    g1_jac t, ph;
    pt_mul_u64(t, p, B3_X_ABS);
    pt_mul_u64(t, t, B3_X_ABS);
    pt_neg(t, t);
    g1_phi(ph, p);
    return pt_eq(t, ph);
}

// Budroni-Pintore cofactor clearing: [x^2 - x - 1]P + [x - 1]psi(P) + psi^2(2P)   (A/ecp2.rs:784-805)
template <class F2>
B3_FN_NOINLINE void g2_clear_cofactor(jac<F2>& r, const jac<F2>& p) {
    jac<F2> xp, x2p, t, np;
    pt_mul_u64(xp, p, B3_X_ABS);            // |x| P
    pt_mul_u64(x2p, xp, B3_X_ABS);          // x^2 P
    pt_neg(xp, xp);                         // x P   (x negative)
    pt_neg(np, p);
    pt_neg(t, xp);
    pt_add(x2p, x2p, t);                    // x^2 P - x P
    pt_add(x2p, x2p, np);                   // x^2 P - x P - P
    pt_add(xp, xp, np);                     // x P - P
    g2_psi(xp, xp);
    pt_dbl(t, p);
    g2_psi2(t, t);
    pt_add(t, t, x2p);
    pt_add(r, t, xp);
}

// ---- wire formats (ZCash uncompressed; A/bls381/core.rs:177-190, 344-364) -----------------------
// G1: x || y (48-byte big-endian each); G2: x.im || x.re || y.im || y.re; infinity = 0x40 then zeros.
// Records travel as 16-byte words (fp.cuh: wire_load / wire_store -- LDG.128 / STG.128 on aligned arrays).
B3_FN_NOINLINE void g1_aff_to_wire(uint8_t* out, const g1_aff& a) {
    b3_q16 q[6];
    if (a.inf) {
#pragma unroll
        for (int i = 0; i < 6; i++) { q[i].x = 0; q[i].y = 0; q[i].z = 0; q[i].w = 0; }
        q[0].x = 0x40;
    } else {
        fp t;
        fp_from_mont(t, a.x); fp_raw_to_q(q, t);
        fp_from_mont(t, a.y); fp_raw_to_q(q + 3, t);
    }
    wire_store<6>(out, q);
}
B3_FN_NOINLINE void g2_aff_to_wire(uint8_t* out, const g2_aff& a) {
    b3_q16 q[12];
    if (a.inf) {
#pragma unroll
        for (int i = 0; i < 12; i++) { q[i].x = 0; q[i].y = 0; q[i].z = 0; q[i].w = 0; }
        q[0].x = 0x40;
    } else {
        fp t;
        fp_from_mont(t, a.x.c1); fp_raw_to_q(q, t);
        fp_from_mont(t, a.x.c0); fp_raw_to_q(q + 3, t);
        fp_from_mont(t, a.y.c1); fp_raw_to_q(q + 6, t);
        fp_from_mont(t, a.y.c0); fp_raw_to_q(q + 9, t);
    }
    wire_store<12>(out, q);
}
// status codes shared with the C ABI (mirror A/errors.rs:1-11)
#define B3_OK 0
#define B3_ERR_INVALID_POINT (-5)
#define B3_ERR_INVALID_YFLAG (-8)
// Parse a record held as 16-byte words, without the on-curve check (caller decides); returns B3_OK or an error code.
// Byte 0 of the record (the flag byte) is the low byte of q[0].x.
B3_FN_NOINLINE int g1_aff_from_q(g1_aff& r, const b3_q16* q) {
    const uint32_t b0 = q[0].x & 0xffu;
    if (b0 & 0x80) return B3_ERR_INVALID_POINT;             // compressed flag on a 96-byte buffer
    if (b0 & 0x40) {
        const uint32_t acc = (q[0].x & 0xffffff3fu) | q[0].y | q[0].z | q[0].w | b3_q16_or(q + 1, 5);
        if (acc) return B3_ERR_INVALID_POINT;
        r.x = FP_NIL; r.y = FP_NIL; r.inf = 1;
        return B3_OK;
    }
    if (b0 & 0x20) return B3_ERR_INVALID_YFLAG;
    fp x, y;
    fp_raw_from_q(x, q);
    fp_raw_from_q(y, q + 3);
    if (!fp_raw_lt_p(x) || !fp_raw_lt_p(y)) return B3_ERR_INVALID_POINT;
    fp_to_mont(r.x, x); fp_to_mont(r.y, y); r.inf = 0;
    return B3_OK;
}
B3_FN_NOINLINE int g1_aff_from_wire(g1_aff& r, const uint8_t* in) {
    b3_q16 q[6];
    wire_load<6>(q, in);
    return g1_aff_from_q(r, q);
}
B3_FN_NOINLINE int g2_aff_from_q(g2_aff& r, const b3_q16* q) {
    const uint32_t b0 = q[0].x & 0xffu;
    if (b0 & 0x80) return B3_ERR_INVALID_POINT;
    if (b0 & 0x40) {
        const uint32_t acc = (q[0].x & 0xffffff3fu) | q[0].y | q[0].z | q[0].w | b3_q16_or(q + 1, 11);
        if (acc) return B3_ERR_INVALID_POINT;
        fp2_zero(r.x); fp2_zero(r.y); r.inf = 1;
        return B3_OK;
    }
    if (b0 & 0x20) return B3_ERR_INVALID_YFLAG;
    fp v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        fp_raw_from_q(v[k], q + 3 * k);
        if (!fp_raw_lt_p(v[k])) return B3_ERR_INVALID_POINT;
    }
    fp_to_mont(r.x.c1, v[0]); fp_to_mont(r.x.c0, v[1]);
    fp_to_mont(r.y.c1, v[2]); fp_to_mont(r.y.c0, v[3]);
    r.inf = 0;
    return B3_OK;
}
B3_FN_NOINLINE int g2_aff_from_wire(g2_aff& r, const uint8_t* in) {
    b3_q16 q[12];
    wire_load<12>(q, in);
    return g2_aff_from_q(r, q);
}
