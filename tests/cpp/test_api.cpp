// C++ host-side test of the drop-in boundary: drives include/milagro_bls_b200.hpp (the compiled-language mirror of the
// reference's public API) the way the reference's own tests drive its types -- M/src/signature.rs:95-140 (sign / verify),
// M/src/keys.rs:250-350 (encodings), M/src/aggregates.rs:335-805 (aggregation, aggregate_verify, fast_aggregate_verify,
// verify_multiple_aggregate_signatures incl. the RNG contract).  Signing is not on the GPU path: signatures are synthesised
// with the library's input helpers (pk = [sk]G1, sig = [sk]H(msg)).  Needs an sm_100 device; exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/milagro_bls_b200.hpp"

using namespace milagro_bls;

static int failures = 0;
#define CHECK(cond)                                                                     \
    do {                                                                                \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

// deterministic byte stream standing in for rand::Rng (SplitMix64), counting what the callee consumes
struct CountingRng {
    uint64_t x;
    size_t consumed = 0;
    explicit CountingRng(uint64_t seed) : x(seed) {}
    uint64_t next() {
        x += 0x9E3779B97F4A7C15ull;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    void fill_bytes(uint8_t* p, size_t n) {
        consumed += n;
        for (size_t i = 0; i < n; i++) p[i] = static_cast<uint8_t>(next());
    }
};

struct Keypair {
    uint8_t sk[32];
    PublicKey pk;
};
static Keypair keypair(uint64_t k) {
    Keypair kp{};
    for (int i = 0; i < 8; i++) kp.sk[31 - i] = static_cast<uint8_t>(k >> (8 * i));
    b3_g1_mul_gen(detail::ctx(), kp.sk, 1, kp.pk.point.data());
    return kp;
}
static Signature sign(const Keypair& kp, const std::string& msg) {           // sig = [sk] hash_to_curve_g2(msg)
    const uint32_t off[2] = {0, static_cast<uint32_t>(msg.size())};
    uint8_t h[192];
    Signature s;
    b3_hash_to_g2(detail::ctx(), reinterpret_cast<const uint8_t*>(msg.data()), off, 1, nullptr, 0, h);
    b3_g2_mul(detail::ctx(), h, kp.sk, 1, s.point.data());
    return s;
}
static const uint8_t* bytes(const std::string& s) { return reinterpret_cast<const uint8_t*>(s.data()); }

int main() {
    if (!detail::ctx()) { std::printf("no sm_100 device / library: the product has no CPU path\n"); return 2; }

    // ---- encodings (M/src/keys.rs:140-186, 250-350; M/src/signature.rs:43-51)
    Keypair a = keypair(0x1234567890abcdefull), b = keypair(42), c = keypair(77777);
    auto a48 = a.pk.as_bytes();
    Result<PublicKey> back = PublicKey::from_bytes(a48.data(), a48.size());
    CHECK(back && back.value == a.pk && back.value.key_validate());
    CHECK(PublicKey::from_bytes(a48.data(), 47).err == AmclError::InvalidG1Size);
    auto unc = a.pk.as_uncompressed_bytes();
    CHECK(PublicKey::from_uncompressed_bytes(unc.data(), unc.size()).value == a.pk);
    CHECK(PublicKey::from_uncompressed_bytes(unc.data(), 95).err == AmclError::InvalidG1Size);
    uint8_t inf48[48] = {0xc0};
    CHECK(PublicKey::from_bytes_unchecked(inf48, 48).is_ok());                   // infinity decodes ...
    CHECK(PublicKey::from_bytes(inf48, 48).err == AmclError::InvalidPoint);      // ... but is not a valid key
    uint8_t bad48[48] = {0xc0, 1};
    CHECK(PublicKey::from_bytes_unchecked(bad48, 48).err == AmclError::InvalidPoint);   // infinity flag with non-zero bytes
    Signature sa = sign(a, "cats");
    auto s96 = sa.as_bytes();
    Result<Signature> sback = Signature::from_bytes(s96.data(), s96.size());
    CHECK(sback && sback.value == sa);
    CHECK(Signature::from_bytes(s96.data(), 95).err == AmclError::InvalidG2Size);

    // ---- Signature::verify (M/src/signature.rs:27-40, README example "cats")
    CHECK(sa.verify(bytes("cats"), 4, a.pk));
    CHECK(!sa.verify(bytes("dogs"), 4, a.pk));
    CHECK(!sa.verify(bytes("cats"), 4, b.pk));

    // ---- aggregation (M/src/aggregates.rs:29-124)
    CHECK(AggregatePublicKey::into_aggregate({}).err == AmclError::AggregateEmptyPoints);
    Result<AggregatePublicKey> abc = AggregatePublicKey::into_aggregate({a.pk, b.pk, c.pk});
    CHECK(abc.is_ok());
    AggregatePublicKey step = AggregatePublicKey::from_public_key(a.pk);
    step.add(b.pk);
    step.add_aggregate(AggregatePublicKey::from_public_key(c.pk));
    CHECK(step == abc.value);
    CHECK(AggregatePublicKey::aggregate({&a.pk, &b.pk, &c.pk}).value == abc.value);

    // ---- fast_aggregate_verify (M/src/aggregates.rs:177-253): one message, three signers
    Signature sb = sign(b, "cats"), sc = sign(c, "cats");
    AggregateSignature agg = AggregateSignature::aggregate({&sa, &sb, &sc});
    AggregateSignature agg2;                                                     // new() + add, one by one
    agg2.add(sa); agg2.add(sb); agg2.add_aggregate(AggregateSignature::from_signature(sc));
    CHECK(agg == agg2);
    CHECK(agg.fast_aggregate_verify(bytes("cats"), 4, {&a.pk, &b.pk, &c.pk}));
    CHECK(!agg.fast_aggregate_verify(bytes("cats"), 4, {&a.pk, &b.pk}));
    CHECK(!agg.fast_aggregate_verify(bytes("cats"), 4, {}));
    CHECK(agg.fast_aggregate_verify_pre_aggregated(bytes("cats"), 4, abc.value));
    CHECK(!agg.fast_aggregate_verify_pre_aggregated(bytes("cat"), 3, abc.value));
    auto agg96 = agg.as_bytes();
    CHECK(AggregateSignature::from_bytes(agg96.data(), agg96.size()).value == agg);

    // ---- aggregate_verify (M/src/aggregates.rs:130-170): three messages, three signers
    std::string m1 = "message one", m2 = "message two", m3 = "";
    Signature t1 = sign(a, m1), t2 = sign(b, m2), t3 = sign(c, m3);
    AggregateSignature av = AggregateSignature::aggregate({&t1, &t2, &t3});
    CHECK(av.aggregate_verify({{bytes(m1), m1.size()}, {bytes(m2), m2.size()}, {bytes(m3), m3.size()}}, {&a.pk, &b.pk, &c.pk}));
    CHECK(!av.aggregate_verify({{bytes(m2), m2.size()}, {bytes(m1), m1.size()}, {bytes(m3), m3.size()}}, {&a.pk, &b.pk, &c.pk}));
    CHECK(!av.aggregate_verify({{bytes(m1), m1.size()}}, {&a.pk, &b.pk}));                    // mismatched lengths
    CHECK(!av.aggregate_verify({}, {}));

    // ---- verify_multiple_aggregate_signatures (M/src/aggregates.rs:261-316) and its RNG contract (272-287)
    std::vector<Keypair> kps;
    std::vector<std::string> msgs;
    std::vector<AggregateSignature> sigs(5);
    std::vector<AggregatePublicKey> apks(5);
    for (int j = 0; j < 5; j++) {
        Keypair k1 = keypair(1000 + 17 * j), k2 = keypair(2000 + 31 * j);
        msgs.push_back("attestation " + std::to_string(j));
        Signature s1 = sign(k1, msgs[j]), s2 = sign(k2, msgs[j]);
        sigs[j] = AggregateSignature::aggregate({&s1, &s2});
        apks[j] = AggregatePublicKey::into_aggregate({k1.pk, k2.pk}).value;
    }
    auto sets_of = [&](const std::vector<AggregateSignature>& sg, const std::vector<std::string>& ms) {
        std::vector<AggregateSignature::Set> v;
        for (int j = 0; j < 5; j++) v.push_back({&sg[j], &apks[j], bytes(ms[j]), ms[j].size()});
        return v;
    };
    {
        CountingRng rng(1);
        auto sets = sets_of(sigs, msgs);
        CHECK(AggregateSignature::verify_multiple_aggregate_signatures(rng, sets.begin(), sets.end()));
        CHECK(rng.consumed == 5 * 8);                                            // one 8-byte draw per set
    }
    {
        CountingRng rng(2);
        std::vector<std::string> tampered = msgs;
        tampered[3][0] ^= 1;
        auto sets = sets_of(sigs, tampered);
        CHECK(!AggregateSignature::verify_multiple_aggregate_signatures(rng, sets.begin(), sets.end()));
        CHECK(rng.consumed == 5 * 8);
    }
    {   // a signature that is on the curve but outside G2 at index 2: reject, and the caller's RNG advanced by exactly 2 draws
        CountingRng pick(99);
        AggregateSignature rogue;
        bool found = false;
        for (int t = 0; t < 256 && !found; t++) {
            uint8_t x[96];
            pick.fill_bytes(x, 96);
            x[0] = 0x80 | (x[0] & 0x0f);
            x[48] &= 0x0f;
            int32_t st = 0, ok = 0;
            if (b3_g2_decompress(detail::ctx(), x, 1, rogue.point.data(), &st) == B3_OK && st == 0 &&
                b3_g2_subgroup_check(detail::ctx(), rogue.point.data(), 1, &st, &ok) == B3_OK && st == 0 && !ok)
                found = true;
        }
        CHECK(found);
        std::vector<AggregateSignature> bad = sigs;
        bad[2] = rogue;
        CountingRng rng(3);
        auto sets = sets_of(bad, msgs);
        CHECK(!AggregateSignature::verify_multiple_aggregate_signatures(rng, sets.begin(), sets.end()));
        CHECK(rng.consumed == 2 * 8);
    }
    {
        CountingRng rng(4);
        std::vector<AggregateSignature::Set> none;
        CHECK(AggregateSignature::verify_multiple_aggregate_signatures(rng, none.begin(), none.end()));      // empty batch verifies
        CHECK(rng.consumed == 0);
    }
    if (failures) std::printf("%d check(s) FAILED\n", failures);
    else std::printf("all checks passed\n");
    return failures ? 1 : 0;
}
