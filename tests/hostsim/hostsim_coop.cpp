// TEST INFRASTRUCTURE ONLY.  Host replay of the CTA-cooperative Fp12 routines (milagro_bls_b200/csrc/coop12.cuh):
// the threads of each phase are executed one after another (COOP_PHASE under B3_HOSTSIM), so the phase logic and
// the index arithmetic can be checked against the oracle without a GPU.  Loaded only by tests/test_hostsim.py.
#define B3_HOSTSIM 1
#include <string.h>
#include "../../milagro_bls_b200/csrc/coop12.cuh"

static void fp_in(fp& r, const uint8_t* b) { fp t; fp_raw_from_be(t, b); fp_to_mont(r, t); }
static void fp2_in(fp2& r, const uint8_t* b) { fp_in(r.c0, b); fp_in(r.c1, b + 48); }
static void fp12_in(fp12& r, const uint8_t* b) {      // wire order w^0,w^3,w^1,w^4,w^2,w^5
    const int order[6] = {0, 3, 1, 4, 2, 5};
    for (int k = 0; k < 6; k++) fp2_in(fp12_coef(r, order[k]), b + 96 * k);
}

extern "C" {
// out = f * (l0 + l3 w^3 + l5 w^5) through the dot-product sparse multiplication; line = 3 x 96 bytes (l0, l3, l5)
void hc_mul_by_line_dot(const uint8_t* f576, const uint8_t* line288, uint8_t* out) {
    fp12 f, r;
    fp2 l0, l3, l5;
    fp12_in(f, f576);
    fp2_in(l0, line288); fp2_in(l3, line288 + 96); fp2_in(l5, line288 + 192);
    line_ops o;
    line_ops_make(o, l0, l3, l5);
    fp12_mul_by_line_dot(r, f, o);
    fp12_to_wire(out, r);
}
// out = a1*b1 + a2*b2 mod p through the dual-product Montgomery routine (fp.cuh: fp_mul2)
void hc_fp_mul2(const uint8_t* a1, const uint8_t* b1, const uint8_t* a2, const uint8_t* b2, uint8_t* out) {
    fp x1, y1, x2, y2, r, t;
    fp_in(x1, a1); fp_in(y1, b1); fp_in(x2, a2); fp_in(y2, b2);
    fp_mul2(r, x1, y1, x2, y2);
    fp_from_mont(t, r);
    fp_raw_to_be(out, t);
}
// op: 0 mul, 1 conj, 2 frob, 3 frob2, 4 frob3, 5 pow_x, 6 pow_x_half, 7 final_exp
void hc_fp12_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    static coop_fexp_ws s;
    fp12 x, y;
    fp12_in(x, a); fp12_in(y, b);
    s.m = x; s.t = y;
    switch (op) {
        case 0: coop_fp12_mul(s.rr, s.m, s.t, s.ws); break;
        case 1: coop_fp12_conj(s.rr, s.m); break;
        case 2: coop_fp12_frob(s.rr, s.m, 1); break;
        case 3: coop_fp12_frob(s.rr, s.m, 2); break;
        case 4: coop_fp12_frob(s.rr, s.m, 3); break;
        case 5: coop_fp12_pow_x(s.rr, s.m, 0, s.ws); break;
        case 6: coop_fp12_pow_x(s.rr, s.m, 1, s.ws); break;
        default: coop_final_exp(s); break;
    }
    fp12_to_wire(out, s.rr);
}

// Split multi-Miller loop replayed on the host: lines per pair (miller_*_step_u), per-slot accumulation with the
// (xP, -yP) scaling, cooperative closing chain, cooperative final exponentiation -> GT bytes.
int hc_multi_pairing_split(const uint8_t* q192s, const uint8_t* p96s, int n, uint8_t* gt576) {
    static fp12 slots[B3_MILLER_SLOTS];
    static coop_fexp_ws s;
    for (int k = 0; k < B3_MILLER_SLOTS; k++) fp12_one(slots[k]);
    for (int i = 0; i < n; i++) {
        g2_aff Q; g1_aff P; int e;
        if ((e = g2_aff_from_wire(Q, q192s + 192 * i))) return e;
        if ((e = g1_aff_from_wire(P, p96s + 96 * i))) return e;
        if (Q.inf || P.inf) continue;
        // de-normalise both members (Jacobian with Z = 3 resp. 5) so the projective paths are exercised
        fp three, five; fp_add(three, FP_ONE, FP_ONE); fp_add(three, three, FP_ONE); fp_add(five, three, FP_ONE); fp_add(five, five, FP_ONE);
        fp2 qz; qz.c0 = three; qz.c1 = five;
        fp2 qz2, qz3, qX, qY; fp2_sqr(qz2, qz); fp2_mul(qz3, qz2, qz); fp2_mul(qX, Q.x, qz2); fp2_mul(qY, Q.y, qz3);
        fp pz2, pz3, pX, pY; fp_sqr(pz2, five); fp_mul(pz3, pz2, five); fp_mul(pX, P.x, pz2); fp_mul(pY, P.y, pz3);
        fp pxz, pny; fp_mul(pxz, pX, five); fp_neg(pny, pY);                  // g1_pp: (X Z, -Y, Z^3)
        miller_pt_t<fp2> t, Qh;
        miller_start(t, qX, qY, qz);
        Qh = t;
        int a = B3_MILLER_DBL_SLOTS;
        for (int it = 0; it < B3_MILLER_DBL_SLOTS; it++) {
            fp2 u0, l3, u5, l0, l5;
            miller_dbl_step_u(t, u0, l3, u5);
            fp2_mul_fp(l0, u0, pny); fp2_mul_fp(l3, l3, pz3); fp2_mul_fp(l5, u5, pxz);
            fp12_mul_by_line(slots[it], l0, l3, l5);
            if ((B3_X_ABS >> (62 - it)) & 1) {
                miller_add_step_u(t, u0, l3, u5, Qh.x, Qh.y, Qh.z);
                fp2_mul_fp(l0, u0, pny); fp2_mul_fp(l3, l3, pz3); fp2_mul_fp(l5, u5, pxz);
                fp12 d; fp12_from_line(d, l0, l3, l5);
                fp12_mul(slots[a], slots[a], d);
                a++;
            }
        }
    }
    coop_miller_chain(s.m, slots, s.ws);
    coop_final_exp(s);
    fp12_to_wire(gt576, s.rr);
    return 0;
}
// Per-item path (k_items_finish): line table in the layout of k_miller_lines, then coop_item_miller (sparse cooperative
// line products) and the cooperative final exponentiation -> GT bytes of prod_i e(Q_i, P_i).
int hc_item_pairing(const uint8_t* q192s, const uint8_t* p96s, int n, uint8_t* gt576) {
    static fp2 lines[B3_MILLER_SLOTS * 4 * 3];
    static coop_item_pair pr[4];
    static coop_fexp_ws s;
    static fp12 line;
    if (n > 4) return -1;
    for (int i = 0; i < n; i++) {
        g2_aff Q; g1_aff P; int e;
        if ((e = g2_aff_from_wire(Q, q192s + 192 * i))) return e;
        if ((e = g1_aff_from_wire(P, p96s + 96 * i))) return e;
        pr[i].idx = (size_t)i;
        pr[i].valid = !(Q.inf || P.inf);
        g1_jac pj; pt_from_aff(pj, P);
        g1_pp pp; g1_pp_from_jac(pp, pj);
        pr[i].ny = pp.ny; pr[i].z3 = pp.z3; pr[i].xz = pp.xz;
        if (!pr[i].valid) continue;
        miller_pt_t<fp2> t, Qh;
        fp2 one; fp2_one(one);
        miller_start(t, Q.x, Q.y, one);
        Qh = t;
        int a = B3_MILLER_DBL_SLOTS;
        for (int it = 0; it < B3_MILLER_DBL_SLOTS; it++) {
            fp2* o = lines + ((size_t)it * n + i) * 3;
            miller_dbl_step_u(t, o[0], o[1], o[2]);
            if ((B3_X_ABS >> (62 - it)) & 1) {
                o = lines + ((size_t)a * n + i) * 3;
                miller_add_step_u(t, o[0], o[1], o[2], Qh.x, Qh.y, Qh.z);
                a++;
            }
        }
    }
    coop_item_miller(s.m, line, s.ws, lines, (size_t)n, pr, n);
    coop_final_exp(s);
    fp12_to_wire(gt576, s.rr);
    return 0;
}
}
