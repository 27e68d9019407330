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
}
