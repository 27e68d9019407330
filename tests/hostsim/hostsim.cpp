// TEST INFRASTRUCTURE ONLY.  Compiles the device headers of milagro_bls_b200/csrc with a host C++ compiler
// (B3_HOSTSIM: the PTX carry-flag primitives are emulated) so the device algorithms can be checked against
// the oracle on a machine without a GPU.  This library is loaded only by tests/test_hostsim.py; the product
// library never links it and has no CPU path.
#define B3_HOSTSIM 1
#include <string.h>
#include "../../milagro_bls_b200/csrc/h2c.cuh"
#include "../../milagro_bls_b200/csrc/pairing.cuh"
#include "test_only.cuh"

static void fp_in(fp& r, const uint8_t* b) { fp t; fp_raw_from_be(t, b); fp_to_mont(r, t); }
static void fp_out(uint8_t* b, const fp& a) { fp t; fp_from_mont(t, a); fp_raw_to_be(b, t); }
static void fp2_in(fp2& r, const uint8_t* b) { fp_in(r.c0, b); fp_in(r.c1, b + 48); }
static void fp2_out(uint8_t* b, const fp2& a) { fp_out(b, a.c0); fp_out(b + 48, a.c1); }
static void fp12_in(fp12& r, const uint8_t* b) {      // wire order w^0,w^3,w^1,w^4,w^2,w^5
    const int order[6] = {0, 3, 1, 4, 2, 5};
    for (int k = 0; k < 6; k++) fp2_in(fp12_coef(r, order[k]), b + 96 * k);
}

extern "C" {
// op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 half(a), 5 inv(a), 6 sqr(a)
void hs_fp_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    fp x, y, r;
    fp_in(x, a); fp_in(y, b);
    switch (op) {
        case 0: fp_mul(r, x, y); break;
        case 1: fp_add(r, x, y); break;
        case 2: fp_sub(r, x, y); break;
        case 3: fp_neg(r, x); break;
        case 4: fp_half(r, x); break;
        case 5: fp_inv(r, x); break;
        default: fp_sqr(r, x); break;
    }
    fp_out(out, r);
}
// raw Montgomery product of two arbitrary 384-bit operands is only defined for a < p; exposed for edge tests
void hs_fp_mont_mul_raw(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    fp x, y, r; fp_raw_from_be(x, a); fp_raw_from_be(y, b); fp_mul(r, x, y); fp_raw_to_be(out, r);
}
// raw Montgomery square a^2 / 2^384 mod p of a plain integer a < p (limb patterns chosen by the test)
void hs_fp_sqr_raw(const uint8_t* a, uint8_t* out) {
    fp x, r; fp_raw_from_be(x, a); fp_sqr(r, x); fp_raw_to_be(out, r);
}
// a^e for a plain 384-bit exponent e (sliding-window fp_pow_const)
void hs_fp_pow(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    fp x, ex, r; fp_in(x, a); fp_raw_from_be(ex, b); fp_pow_const(r, x, ex); fp_out(out, r);
}
void hs_fp2_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    fp2 x, y, r;
    fp2_in(x, a); fp2_in(y, b);
    switch (op) {
        case 0: fp2_mul(r, x, y); break;
        case 1: fp2_sqr(r, x); break;
        case 2: fp2_inv(r, x); break;
        default: fp2_mul_xi(r, x); break;
    }
    fp2_out(out, r);
}
int hs_fp2_sqrt_or_z(const uint8_t* a, uint8_t* out) {
    fp2 x, r; fp2_in(x, a);
    bool sq = fp2_sqrt_or_z(r, x);
    fp2_out(out, r);
    return sq ? 1 : 0;
}
int hs_fp2_sgn0(const uint8_t* a) {
    fp2 x; fp2_in(x, a);
    return (int)fp2_sgn0(x);
}
// op: 0 mul, 1 sqr, 2 inv, 3 frob, 4 frob2, 5 frob3, 6 conj, 7 cyclo_sqr, 8 final_exp, 9 pow_x
void hs_fp12_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    fp12 x, y, r;
    fp12_in(x, a); fp12_in(y, b);
    switch (op) {
        case 0: fp12_mul(r, x, y); break;
        case 1: fp12_sqr(r, x); break;
        case 2: fp12_inv(r, x); break;
        case 3: fp12_frob(r, x); break;
        case 4: fp12_frob2(r, x); break;
        case 5: fp12_frob3(r, x); break;
        case 6: fp12_conj(r, x); break;
        case 7: fp12_cyclo_sqr(r, x); break;
        case 8: final_exp(r, x); break;
        default: fp12_pow_x(r, x); break;
    }
    fp12_to_wire(out, r);
}
void hs_hash_to_field(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dlen, uint8_t* out192) {
    fp2 u0, u1;
    hash_to_field_fp2_x2(u0, u1, msg, len, dst, dlen);
    fp2_out(out192, u0); fp2_out(out192 + 96, u1);
}
// map_to_curve_g2(u) -> uncompressed wire point
void hs_map_to_curve_g2(const uint8_t* u96, uint8_t* out192) {
    fp2 u; fp2_in(u, u96);
    g2_jac q; map_to_curve_g2(q, u);
    g2_aff a; pt_to_aff(a, q);
    g2_aff_to_wire(out192, a);
}
void hs_sswu(const uint8_t* u96, uint8_t* out288) {
    fp2 u, xn, xd, y; fp2_in(u, u96);
    sswu_g2(xn, xd, y, u);
    fp2_out(out288, xn); fp2_out(out288 + 96, xd); fp2_out(out288 + 192, y);
}
// debug: intermediates of sswu_g2 (same statements, exported one by one)
void hs_sswu_dbg(const uint8_t* u96, uint8_t* outdbg) {
    uint8_t* out = outdbg;
    fp2 u; fp2_in(u, u96);
    fp2 tv1, tv2, gxn, gxd, t, t2, xn, xd;
    fp2_sqr(tv1, u);
    fp2_mul(tv1, tv1, SSWU_Z);
    fp2_sqr(tv2, tv1);
    fp2_add(tv2, tv2, tv1);
    bool exc = fp2_is_zero(tv2);
    fp2 one; fp2_one(one);
    fp2_add(t, tv2, one);
    fp2_mul(t, t, SSWU_B);
    fp2_neg(t, t);
    fp2_select(xn, exc, SSWU_B, t);
    fp2_mul(t, tv2, SSWU_A);
    fp2_select(xd, exc, SSWU_ZA, t);
    fp2_sqr(t, xd);
    fp2_mul(gxd, t, xd);
    fp2_mul(t, t, SSWU_A);
    fp2_sqr(t2, xn);
    fp2_add(t, t, t2);
    fp2_mul(t, t, xn);
    fp2_mul(t2, gxd, SSWU_B);
    fp2_add(gxn, t, t2);
    fp n, s2;
    fp_sqr(n, gxd.c0);
    fp_sqr(s2, gxd.c1);
    fp_add(n, n, s2);
    fp2 wv, root;
    fp2_conj(t, gxd);
    fp2_mul(wv, gxn, t);
    fp2_mul_fp(wv, wv, n);
    bool sq = fp2_sqrt_or_z(root, wv);
    fp ninv;
    fp_inv(ninv, n);
    fp2 root2;
    fp2_mul_fp(root2, root, ninv);
    fp2_out(out, tv1); fp2_out(out + 96, gxn); fp2_out(out + 192, gxd);
    fp_out(out + 288, n); fp2_out(out + 336, wv); fp2_out(out + 432, root);
    fp_out(out + 528, ninv); fp2_out(out + 576, root2);
    out[672] = sq ? 1 : 0; out[673] = (uint8_t)fp2_sgn0(u); out[674] = (uint8_t)fp2_sgn0(root2);
    // tail
    fp2 tt, root3 = root2, root4;
    if (!sq) {
        fp2_mul(tt, tv1, u);
        fp2_mul(root3, root3, tt);
    }
    fp2_out(out + 700, root3);
    out[675] = (uint8_t)fp2_sgn0(root3);
    root4 = root3;
    if (fp2_sgn0(u) != fp2_sgn0(root4)) fp2_neg(root4, root4);
    fp2_out(out + 800, root4);
    fp2 rxn, rxd, ry;
    sswu_g2(rxn, rxd, ry, u);
    fp2_out(out + 900, ry);
    out[686] = fp2_eq(root4, ry);
    // variant A: in-place scaling like the real function
    fp2 ra = root;
    fp2_mul_fp(ra, ra, ninv);
    out[676] = fp2_eq(ra, root2) ? 1 : 0;
    // variant B: in-place xn update + t reuse
    fp2 xb = xn;
    if (!sq) {
        fp2_mul(xb, xb, tv1);
        fp2_mul(t, tv1, u);
        fp2_mul(ra, ra, t);
    }
    out[677] = fp2_eq(ra, root3) ? 1 : 0;
    if (fp2_sgn0(u) != fp2_sgn0(ra)) fp2_neg(ra, ra);
    out[678] = fp2_eq(ra, root4) ? 1 : 0;
    out[679] = fp2_eq(ry, root4) ? 1 : 0;
    out[680] = fp2_eq(rxn, xb) ? 1 : 0;
    out[681] = fp2_eq(rxd, xd) ? 1 : 0;
}
void hs_iso3(const uint8_t* in288, uint8_t* out192) {
    fp2 xn, xd, y; fp2_in(xn, in288); fp2_in(xd, in288 + 96); fp2_in(y, in288 + 192);
    g2_jac q; iso3_g2(q, xn, xd, y);
    g2_aff a; pt_to_aff(a, q);
    g2_aff_to_wire(out192, a);
}
void hs_hash_to_g2(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dlen, uint8_t* out192) {
    g2_jac q; hash_to_g2_jac(q, msg, len, dst, dlen);
    g2_aff a; pt_to_aff(a, q);
    g2_aff_to_wire(out192, a);
}
// G1 ops on wire points. op: 0 add, 1 mul by k (u64), 2 dbl, 3 add via mixed, 4 phi
int hs_g1_op(int op, const uint8_t* a96, const uint8_t* b96, uint64_t k, uint8_t* out96) {
    g1_aff A, B; int e;
    if ((e = g1_aff_from_wire(A, a96))) return e;
    if ((e = g1_aff_from_wire(B, b96))) return e;
    g1_jac P, Q, R;
    pt_from_aff(P, A); pt_from_aff(Q, B);
    // de-normalise P so the generic (non-mixed) path is exercised: scale by z = 3
    fp three; fp_add(three, FP_ONE, FP_ONE); fp_add(three, three, FP_ONE);
    if (!A.inf) { fp z2, z3; fp_sqr(z2, three); fp_mul(z3, z2, three); fp_mul(P.x, P.x, z2); fp_mul(P.y, P.y, z3); P.z = three; }
    switch (op) {
        case 0: pt_add(R, P, Q); break;
        case 1: pt_mul_u64(R, P, k); break;
        case 2: pt_dbl(R, P); break;
        case 3: pt_add_aff(R, P, B); break;
        case 5: pt_mul_u64_w4(R, P, k); break;
        default: g1_phi(R, P); break;
    }
    g1_aff o; pt_to_aff(o, R);
    g1_aff_to_wire(out96, o);
    return 0;
}
int hs_g2_op(int op, const uint8_t* a192, const uint8_t* b192, uint64_t k, uint8_t* out192) {
    g2_aff A, B; int e;
    if ((e = g2_aff_from_wire(A, a192))) return e;
    if ((e = g2_aff_from_wire(B, b192))) return e;
    g2_jac P, Q, R;
    pt_from_aff(P, A); pt_from_aff(Q, B);
    switch (op) {
        case 0: pt_add(R, P, Q); break;
        case 1: pt_mul_u64(R, P, k); break;
        case 2: pt_dbl(R, P); break;
        case 3: pt_add_aff(R, P, B); break;
        case 4: g2_psi(R, P); break;
        case 5: g2_psi2(R, P); break;
        default: g2_clear_cofactor(R, P); break;
    }
    g2_aff o; pt_to_aff(o, R);
    g2_aff_to_wire(out192, o);
    return 0;
}
int hs_g1_check(const uint8_t* a96, int* on_curve, int* in_subgroup) {
    g1_aff A; int e = g1_aff_from_wire(A, a96); if (e) return e;
    *on_curve = pt_on_curve_aff(A);
    g1_jac P; pt_from_aff(P, A);
    *in_subgroup = g1_in_subgroup(P);
    return 0;
}
int hs_g2_check(const uint8_t* a192, int* on_curve, int* in_subgroup) {
    g2_aff A; int e = g2_aff_from_wire(A, a192); if (e) return e;
    *on_curve = pt_on_curve_aff(A);
    g2_jac P; pt_from_aff(P, A);
    *in_subgroup = g2_in_subgroup(P);
    if ((g2_in_subgroup_aff(A) ? 1 : 0) != *in_subgroup) return -99;      // the mixed-addition variant must agree
    return 0;
}
// GT = fexp( prod_i miller(Q_i, P_i) )
int hs_multi_pairing(const uint8_t* q192s, const uint8_t* p96s, int n, uint8_t* gt576, int* is_one) {
    fp12 acc; fp12_one(acc);
    for (int i = 0; i < n; i++) {
        g2_aff Q; g1_aff P; int e;
        if ((e = g2_aff_from_wire(Q, q192s + 192 * i))) return e;
        if ((e = g1_aff_from_wire(P, p96s + 96 * i))) return e;
        fp12 f; miller_loop_pair(f, Q, P);
        fp12_mul(acc, acc, f);
    }
    fp12 gt; final_exp(gt, acc);
    fp12_to_wire(gt576, gt);
    *is_one = fp12_is_one(gt);
    return 0;
}
}
