// hash_to_curve_g2: RFC 9380 suite BLS12381G2_XMD:SHA-256_SSWU_RO_ (the reference implements draft-09 with
// identical outputs).  Replaces
//   /root/reference/incubator-milagro-crypto-rust/src/hash_to_curve.rs:111-201,283-346
//   /root/reference/incubator-milagro-crypto-rust/src/hash256.rs:86-209
//   /root/reference/incubator-milagro-crypto-rust/src/bls381/iso.rs:177-206, core.rs:831-849
//   /root/reference/incubator-milagro-crypto-rust/src/ecp2.rs:784-805
// The SSWU map is evaluated projectively without field inversions and with a branch-uniform square-root
// (both candidates of the reference's try / retry flow, hash_to_curve.rs:323-338, come out of one pair of Fp
// exponentiations); the result is identical because the sign of y is fixed by sgn0 afterwards.
#pragma once
#include "curve.cuh"

// ------------------------------------------------------------------------------------------------ SHA-256
B3_CONST uint32_t SHA256_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

struct sha256_ctx {
    uint32_t h[8];
    uint32_t w[16];     // current block, big-endian words
    uint32_t fill;      // bytes in the current block
    uint64_t total;     // total bytes absorbed
};

B3_FN uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

B3_FN_NOINLINE void sha256_compress(uint32_t* h, const uint32_t* blk) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = blk[i];
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) wi = w[i];
        else {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA256_K[i] + wi;
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
B3_FN void sha256_init(sha256_ctx& c) {
    c.h[0] = 0x6a09e667u; c.h[1] = 0xbb67ae85u; c.h[2] = 0x3c6ef372u; c.h[3] = 0xa54ff53au;
    c.h[4] = 0x510e527fu; c.h[5] = 0x9b05688cu; c.h[6] = 0x1f83d9abu; c.h[7] = 0x5be0cd19u;
    for (int i = 0; i < 16; i++) c.w[i] = 0;
    c.fill = 0; c.total = 0;
}
B3_FN void sha256_put(sha256_ctx& c, uint8_t byte) {
    uint32_t idx = c.fill >> 2, sh = 24 - 8 * (c.fill & 3);
    c.w[idx] |= (uint32_t)byte << sh;
    c.fill++; c.total++;
    if (c.fill == 64) {
        sha256_compress(c.h, c.w);
        for (int i = 0; i < 16; i++) c.w[i] = 0;
        c.fill = 0;
    }
}
// four bytes at once (the block buffer holds big-endian words): only at a word boundary of the block
B3_FN void sha256_put_word(sha256_ctx& c, uint32_t be_word) {
    c.w[c.fill >> 2] = be_word;
    c.fill += 4; c.total += 4;
    if (c.fill == 64) {
        sha256_compress(c.h, c.w);
        for (int i = 0; i < 16; i++) c.w[i] = 0;
        c.fill = 0;
    }
}
B3_FN void sha256_update(sha256_ctx& c, const uint8_t* p, uint32_t n) {
    uint32_t i = 0;
#if !defined(B3_HOSTSIM)
    if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0 && (c.fill & 3u) == 0) {      // aligned input: 32-bit loads
        if (__isGlobal(p))
            for (; i + 4 <= n; i += 4) sha256_put_word(c, b3_bswap(__ldg(reinterpret_cast<const uint32_t*>(p + i))));
        else                                                                      // message staged in shared memory (k_hash_to_g2)
            for (; i + 4 <= n; i += 4) sha256_put_word(c, b3_bswap(*reinterpret_cast<const uint32_t*>(p + i)));
    }
#endif
    for (; i < n; i++) sha256_put(c, p[i]);
}
B3_FN void sha256_final(sha256_ctx& c, uint32_t* out8) {
    uint64_t bits = c.total * 8;
    sha256_put(c, 0x80);
    while (c.fill != 56) sha256_put(c, 0);
    for (int i = 7; i >= 0; i--) sha256_put(c, (uint8_t)(bits >> (8 * i)));
    for (int i = 0; i < 8; i++) out8[i] = c.h[i];
}

// ------------------------------------------------------------------------------------------------ xmd + hash_to_field
// expand_message_xmd(msg, 256 bytes, dst) -> 64 big-endian words (b_1 || ... || b_8); dst_len <= 255
B3_FN_NOINLINE void expand_message_xmd_256(uint32_t* out64, const uint8_t* msg, uint32_t msg_len,
                                           const uint8_t* dst, uint32_t dst_len) {
    sha256_ctx c;
    uint32_t b0[8], bi[8];
    sha256_init(c);
    for (int i = 0; i < 64; i++) sha256_put(c, 0);          // Z_pad
    sha256_update(c, msg, msg_len);
    sha256_put(c, 0x01); sha256_put(c, 0x00);               // I2OSP(256, 2)
    sha256_put(c, 0x00);
    sha256_update(c, dst, dst_len);
    sha256_put(c, (uint8_t)dst_len);
    sha256_final(c, b0);
    for (int i = 0; i < 8; i++) bi[i] = 0;
    for (uint32_t k = 1; k <= 8; k++) {
        sha256_init(c);
        for (int i = 0; i < 8; i++) {
            uint32_t v = b0[i] ^ bi[i];                      // k == 1: bi = 0 -> b0
            sha256_put(c, (uint8_t)(v >> 24)); sha256_put(c, (uint8_t)(v >> 16));
            sha256_put(c, (uint8_t)(v >> 8)); sha256_put(c, (uint8_t)v);
        }
        sha256_put(c, (uint8_t)k);
        sha256_update(c, dst, dst_len);
        sha256_put(c, (uint8_t)dst_len);
        sha256_final(c, bi);
        for (int i = 0; i < 8; i++) out64[8 * (k - 1) + i] = bi[i];
    }
}
// 64 big-endian bytes (16 BE words) -> Fp element in Montgomery form: OS2IP(bytes) mod p
B3_FN_NOINLINE void fp_from_be64_words(fp& r, const uint32_t* w16) {
    fp lo, hi, t;
    for (int i = 0; i < 12; i++) lo.l[i] = w16[15 - i];
    for (int i = 0; i < 4; i++) hi.l[i] = w16[3 - i];
    for (int i = 4; i < 12; i++) hi.l[i] = 0;
    fp_mul(lo, FP_R2, lo);          // lo * R
    fp_mul(t, FP_R3, hi);           // hi * 2^384 * R
    fp_add(r, lo, t);
}
B3_FN_NOINLINE void hash_to_field_fp2_x2(fp2& u0, fp2& u1, const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
    uint32_t prb[64];
    expand_message_xmd_256(prb, msg, msg_len, dst, dst_len);
    fp_from_be64_words(u0.c0, prb);
    fp_from_be64_words(u0.c1, prb + 16);
    fp_from_be64_words(u1.c0, prb + 32);
    fp_from_be64_words(u1.c1, prb + 48);
}

// ------------------------------------------------------------------------------------------------ SSWU + 3-isogeny
// Simplified SWU onto E2': y^2 = x^3 + A' x + B', returning x = xn/xd and y (affine y), no inversions.
//   tv1 = Z u^2, tv2 = tv1^2 + tv1, x1 = (-B/A)(1 + 1/tv2) = -B (tv2 + 1) / (A tv2)   (tv2 == 0: x1 = B/(Z A))
//   gx1 = (xn^3 + A xn xd^2 + B xd^3) / xd^3;  y1 = sqrt(gx1) or, if gx1 is not a square,
//   x2 = tv1 x1 and y2 = tv1 u sqrt(Z gx1)     (RFC 9380 F.2 straight-line version)
B3_FN_NOINLINE void sswu_g2(fp2& xn, fp2& xd, fp2& y, const fp2& u) {
    fp2 tv1, tv2, gxn, gxd, t, t2;
    fp2_sqr(tv1, u);
    fp2_mul(tv1, tv1, SSWU_Z);
    fp2_sqr(tv2, tv1);
    fp2_add(tv2, tv2, tv1);
    bool exc = fp2_is_zero(tv2);
    // xn = -B (tv2 + 1)   | exceptional: B
    fp2 one;
    fp2_one(one);
    fp2_add(t, tv2, one);
    fp2_mul(t, t, SSWU_B);
    fp2_neg(t, t);
    fp2_select(xn, exc, SSWU_B, t);
    // xd = A tv2          | exceptional: Z A
    fp2_mul(t, tv2, SSWU_A);
    fp2_select(xd, exc, SSWU_ZA, t);
    // gx1 = gxn / gxd,  gxd = xd^3, gxn = xn^3 + A xn xd^2 + B xd^3
    fp2_sqr(t, xd);                  // xd^2
    fp2_mul(gxd, t, xd);             // xd^3
    fp2_mul(t, t, SSWU_A);           // A xd^2
    fp2_sqr(t2, xn);
    fp2_add(t, t, t2);               // xn^2 + A xd^2
    fp2_mul(t, t, xn);
    fp2_mul(t2, gxd, SSWU_B);
    fp2_add(gxn, t, t2);
    // y1 = sqrt(gxn / gxd) or, if that is not a square, sqrt(Z gxn / gxd): two Fp exponentiations, no inversion
    fp2 root;
    bool sq = fp2_sqrt_ratio_or_z(root, gxn, gxd);
    if (!sq) {
        fp2_mul(xn, xn, tv1);                   // x2 = tv1 x1
        fp2_mul(t, tv1, u);
        fp2_mul(root, root, t);                 // y2 = tv1 u sqrt(Z gx1)
    }
    if (fp2_sgn0(u) != fp2_sgn0(root)) fp2_neg(root, root);
    y = root;
}

// 3-isogeny E2' -> E2 evaluated at x = xn/xd, output Jacobian (A/bls381/iso.rs:177-206 computes the same point)
B3_FN_NOINLINE void iso3_g2(g2_jac& r, const fp2& xn, const fp2& xd, const fp2& y) {
    // powers of xd
    fp2 d1 = xd, d2, d3;
    fp2_sqr(d2, xd);
    fp2_mul(d3, d2, xd);
    // homogeneous Horner: poly(xn/xd) * xd^deg
    fp2 xnum, xden, ynum, yden, t;
    // x_num: degree 3
    fp2_mul(xnum, ISO3_XNUM[3], xn);
    fp2_mul(t, ISO3_XNUM[2], d1); fp2_add(xnum, xnum, t);
    fp2_mul(xnum, xnum, xn);
    fp2_mul(t, ISO3_XNUM[1], d2); fp2_add(xnum, xnum, t);
    fp2_mul(xnum, xnum, xn);
    fp2_mul(t, ISO3_XNUM[0], d3); fp2_add(xnum, xnum, t);          // * xd^3
    // x_den: degree 2 (monic)
    fp2_mul(t, ISO3_XDEN[1], d1); fp2_add(xden, xn, t);
    fp2_mul(xden, xden, xn);
    fp2_mul(t, ISO3_XDEN[0], d2); fp2_add(xden, xden, t);          // * xd^2
    // y_num: degree 3
    fp2_mul(ynum, ISO3_YNUM[3], xn);
    fp2_mul(t, ISO3_YNUM[2], d1); fp2_add(ynum, ynum, t);
    fp2_mul(ynum, ynum, xn);
    fp2_mul(t, ISO3_YNUM[1], d2); fp2_add(ynum, ynum, t);
    fp2_mul(ynum, ynum, xn);
    fp2_mul(t, ISO3_YNUM[0], d3); fp2_add(ynum, ynum, t);          // * xd^3
    // y_den: degree 3 (monic)
    fp2_mul(t, ISO3_YDEN[2], d1); fp2_add(yden, xn, t);
    fp2_mul(yden, yden, xn);
    fp2_mul(t, ISO3_YDEN[1], d2); fp2_add(yden, yden, t);
    fp2_mul(yden, yden, xn);
    fp2_mul(t, ISO3_YDEN[0], d3); fp2_add(yden, yden, t);          // * xd^3
    // x = xnum / (xden * xd),  y = y * ynum / yden.   Jacobian with Z = xden * xd * yden:
    //   X = x Z^2 = xnum * (xden xd) * yden^2,  Y = y Z^3 = y ynum * (xden xd)^3 * yden^2
    fp2 dx, z, z2;
    fp2_mul(dx, xden, xd);
    fp2_mul(z, dx, yden);
    if (fp2_is_zero(z)) { pt_set_inf(r); return; }
    fp2_sqr(z2, yden);              // yden^2
    fp2_mul(t, xnum, dx);
    fp2_mul(r.x, t, z2);
    fp2_mul(t, y, ynum);
    fp2_mul(t, t, z2);
    fp2_sqr(z2, dx);
    fp2_mul(z2, z2, dx);            // dx^3
    fp2_mul(r.y, t, z2);
    r.z = z;
}

B3_FN_NOINLINE void map_to_curve_g2(g2_jac& r, const fp2& u) {
    fp2 xn, xd, y;
    sswu_g2(xn, xd, y, u);
    iso3_g2(r, xn, xd, y);
}
