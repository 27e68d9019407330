// TEST INFRASTRUCTURE ONLY -- single-thread reference forms of the device algorithms, used by tests/hostsim (CPU build of the
// device code) and its GPU twin to cross-check the product kernels' building blocks against the oracle.  Nothing in
// milagro_bls_b200/ includes this file.
#pragma once
#include "../../milagro_bls_b200/csrc/h2c.cuh"
#include "../../milagro_bls_b200/csrc/pairing.cuh"

// full hash_to_curve_g2, Jacobian output (not normalised)
B3_FN_NOINLINE void hash_to_g2_jac(g2_jac& r, const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
    fp2 u0, u1;
    hash_to_field_fp2_x2(u0, u1, msg, msg_len, dst, dst_len);
    g2_jac q0, q1;
    map_to_curve_g2(q0, u0);
    map_to_curve_g2(q1, u1);
    pt_add(q0, q0, q1);
    g2_clear_cofactor(r, q0);
}

// G2 membership of a Jacobian point (the product kernel uses the affine form, curve.cuh: g2_in_subgroup_aff)
template <class F2>
B3_FN_NOINLINE bool g2_in_subgroup(const jac<F2>& p) {
    if (pt_is_inf(p)) return true;
    jac<F2> xp, ps;
    pt_mul_u64(xp, p, B3_X_ABS);
    pt_neg(xp, xp);
    g2_psi(ps, p);
    return pt_eq(xp, ps);
}

// Miller-loop state of one pair: T in homogeneous projective coordinates on the twist (x = X/Z, y = Y/Z).
struct miller_pt {
    fp2 x, y, z;
};

// Doubling step (Costello-Lange-Naehrig homogeneous formulas, b' = 4 xi): T <- 2T and the tangent line
//   l = l0 + l3 w^3 + l5 w^5,  l0 = -xi (2YZ) yP,  l3 = 3 b' Z^2 - Y^2,  l5 = 3 X^2 xP
// (same line as the reference, A/pair.rs:35-84, up to the factor 2).
B3_FN_NOINLINE void miller_dbl_step(miller_pt& t, fp2& l0, fp2& l3, fp2& l5, const fp& xp, const fp& yp_neg) {
    fp2 a, b, c, e, f, g, h, j, e2, u;
    fp2_mul(a, t.x, t.y);
    fp2_half(a, a);                 // A = XY/2
    fp2_sqr(b, t.y);                // B = Y^2
    fp2_sqr(c, t.z);                // C = Z^2
    fp2_mul3(u, c);
    f_mul_b(e, u);                  // E = 3 b' C
    fp2_mul3(f, e);                 // F = 3E
    fp2_add(g, b, f);
    fp2_half(g, g);                 // G = (B+F)/2
    fp2_add(h, t.y, t.z);
    fp2_sqr(h, h);
    fp2_add(u, b, c);
    fp2_sub(h, h, u);               // H = 2YZ
    fp2_sqr(j, t.x);                // J = X^2
    fp2_sqr(e2, e);
    // line
    fp2_sub(l3, e, b);
    fp2_mul3(u, j);
    fp2_mul_fp(l5, u, xp);
    fp2_mul_xi(u, h);
    fp2_mul_fp(l0, u, yp_neg);
    // point
    fp2_sub(u, b, f);
    fp2_mul(t.x, a, u);             // X3 = A (B - F)
    fp2_sqr(g, g);
    fp2_mul3(u, e2);
    fp2_sub(t.y, g, u);             // Y3 = G^2 - 3 E^2
    fp2_mul(t.z, b, h);             // Z3 = B H
}

// Addition step T <- T + Q (Q affine) and the chord line
//   l0 = xi lambda yP,  l3 = theta xQ - lambda yQ,  l5 = -theta xP,   theta = Y - yQ Z, lambda = X - xQ Z
// (reference: A/pair.rs:88-133).
B3_FN_NOINLINE void miller_add_step(miller_pt& t, fp2& l0, fp2& l3, fp2& l5, const fp2& xq, const fp2& yq,
                                    const fp& xp_neg, const fp& yp) {
    fp2 theta, lambda, c, d, e, f, g, h, u;
    fp2_mul(u, yq, t.z);
    fp2_sub(theta, t.y, u);
    fp2_mul(u, xq, t.z);
    fp2_sub(lambda, t.x, u);
    fp2_sqr(c, theta);
    fp2_sqr(d, lambda);
    fp2_mul(e, lambda, d);
    fp2_mul(f, t.z, c);
    fp2_mul(g, t.x, d);
    fp2_add(h, e, f);
    fp2_sub(h, h, g);
    fp2_sub(h, h, g);               // H = E + F - 2G
    // line
    fp2_mul(l3, theta, xq);
    fp2_mul(u, lambda, yq);
    fp2_sub(l3, l3, u);
    fp2_mul_fp(l5, theta, xp_neg);
    fp2_mul_xi(u, lambda);
    fp2_mul_fp(l0, u, yp);
    // point
    fp2_mul(t.x, lambda, h);
    fp2_sub(u, g, h);
    fp2_mul(u, theta, u);
    fp2_mul(g, e, t.y);
    fp2_sub(t.y, u, g);
    fp2_mul(t.z, t.z, e);
}


// Miller loop of one pair (Q in G2 affine, P in G1 affine), multiplied INTO f (f <- f^(2^63..) is NOT shared
// here: this routine runs the whole loop on its own accumulator and returns conj(f_{|x|,Q}(P))).
// Pairs with an infinite member contribute 1.
B3_FN_NOINLINE void miller_loop_pair(fp12& f, const g2_aff& q, const g1_aff& p) {
    fp12_one(f);
    if (q.inf || p.inf) return;
    miller_pt t;
    t.x = q.x; t.y = q.y; fp2_one(t.z);
    fp xp_neg, yp_neg;
    fp_neg(xp_neg, p.x);
    fp_neg(yp_neg, p.y);
    fp2 l0, l3, l5;
    const uint64_t x = B3_X_ABS;
    for (int i = 62; i >= 0; i--) {
        if (i != 62) fp12_sqr(f, f);
        miller_dbl_step(t, l0, l3, l5, p.x, yp_neg);
        fp12_mul_by_line(f, l0, l3, l5);
        if ((x >> i) & 1) {
            miller_add_step(t, l0, l3, l5, q.x, q.y, xp_neg, p.y);
            fp12_mul_by_line(f, l0, l3, l5);
        }
    }
    fp12_conj(f, f);                // the curve parameter is negative
}

