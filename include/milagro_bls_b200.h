/* milagro_bls_b200 -- C ABI of the B200-native BLS12-381 batch-verification engine.
 *
 * Drop-in boundary for the VERIFICATION path of sigp/milagro_bls.  The reference has no FFI: its boundary is
 * the Rust public API (M/src/lib.rs:17-22, M = /root/reference/src).  Each entry point below names the Rust
 * item it replaces; a thin Rust shim keeps the reference's signatures and calls these.
 *
 * Wire formats (exactly what the reference's own (de)serialisers produce, A = amcl src dir):
 *   G1 uncompressed  96 B : x || y, 48-byte big-endian each; infinity = 0x40 then zeros   (A/bls381/core.rs:177-190)
 *   G2 uncompressed 192 B : x.im || x.re || y.im || y.re;     infinity = 0x40 then zeros   (A/bls381/core.rs:344-364)
 *   G1 compressed    48 B, G2 compressed 96 B : ZCash flags C/I/S in the top three bits    (A/bls381/core.rs:145-172,312-339)
 *   GT              576 B : 12 x 48-byte big-endian, order of FP12::to_bytes                (A/fp12.rs:859-913)
 *   scalars               : uint64, drawn by the caller with the rule of M/src/aggregates.rs:278-287
 *
 * All pointers are HOST memory unless the name ends in _dev.  Every call is synchronous on the context's
 * stream.  A context is bound to one CUDA device and must not be used from two threads at once; create one
 * context per thread (the reference types are Send+Sync and re-entrant; so is this API across contexts).
 *
 * There is no CPU fallback: every function returns B3_ERR_CUDA if no usable sm_100 device is present.
 */
#ifndef MILAGRO_BLS_B200_H
#define MILAGRO_BLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: 0 = ok; -1..-8 mirror AmclError (A/errors.rs:1-11) in declaration order */
#define B3_OK 0
#define B3_ERR_AGGREGATE_EMPTY_POINTS (-1)
#define B3_ERR_HASH_TO_FIELD (-2)
#define B3_ERR_INVALID_SECRET_KEY_SIZE (-3)
#define B3_ERR_INVALID_SECRET_KEY_RANGE (-4)
#define B3_ERR_INVALID_POINT (-5)
#define B3_ERR_INVALID_G1_SIZE (-6)
#define B3_ERR_INVALID_G2_SIZE (-7)
#define B3_ERR_INVALID_YFLAG (-8)
#define B3_ERR_CUDA (-100)
#define B3_ERR_ARG (-101)

typedef struct b3_ctx b3_ctx;

int b3_ctx_create(int device, b3_ctx** out);
void b3_ctx_destroy(b3_ctx* ctx);
const char* b3_last_error(b3_ctx* ctx);
/* cudaStream_t of the context (as void*), so callers can order their own device work against it */
void* b3_ctx_stream(b3_ctx* ctx);
/* number of kernel launches issued by this context so far (bench.py reports it as gpu_launches) */
uint64_t b3_ctx_launch_count(b3_ctx* ctx);
/* device time in ms of the most recent call's kernels, and of its dominant (Miller-loop) kernel */
float b3_ctx_last_kernel_ms(b3_ctx* ctx, int which);
/* per-stage device time (ms, CUDA events on the context's stream) of the most recent verification / hash call;
 * stages 0 .. b3_stage_count()-1 are named by b3_stage_name() */
float b3_ctx_stage_ms(b3_ctx* ctx, int stage);
const char* b3_stage_name(int stage);
/* serial != 0: run the independent stages of a verification call one after another on the context's stream (for
 * per-stage timing); default 0: they overlap on internal streams, joined before the Miller loop */
void b3_ctx_set_serial(b3_ctx* ctx, int serial);
/* finishing kernel of b3_verify_batch: 0 = chosen by batch size (default), 1 = one CTA per item, 2 = one thread per item,
 * 3 = one lane pair per item */
void b3_ctx_set_item_kernel(b3_ctx* ctx, int which);
int b3_stage_count(void);

/* ---- (de)serialisation: PublicKey::{from_bytes, from_bytes_unchecked, as_bytes} (M/src/keys.rs:140-160),
 *      Signature::{from_bytes, as_bytes} (M/src/signature.rs:43-51), AggregateSignature::{from_bytes, as_bytes}
 *      (M/src/aggregates.rs:319-327), decompress_g1/g2 + compress_g1/g2 (M/src/amcl_utils.rs:46-74).
 *      status[i] = B3_OK or the AmclError code of item i; the call itself fails only on CUDA/argument errors. */
int b3_g1_decompress(b3_ctx*, const uint8_t* in48, size_t n, int validate /* key_validate, keys.rs:181-186 */,
                     uint8_t* out96, int32_t* status);
int b3_g2_decompress(b3_ctx*, const uint8_t* in96, size_t n, uint8_t* out192, int32_t* status);
int b3_g1_compress(b3_ctx*, const uint8_t* in96, size_t n, uint8_t* out48, int32_t* status);
int b3_g2_compress(b3_ctx*, const uint8_t* in192, size_t n, uint8_t* out96, int32_t* status);
/* PublicKey::from_uncompressed_bytes / key_validate (M/src/keys.rs:163-186): status[i] = B3_OK iff on curve;
 * valid[i] = 1 iff additionally not infinity and in G1 */
int b3_g1_validate(b3_ctx*, const uint8_t* in96, size_t n, int32_t* status, int32_t* valid);
/* subgroup_check_g2 (A/bls381/core.rs:123-127) on parsed signatures; status as above, ok[i] in {0,1} */
int b3_g2_subgroup_check(b3_ctx*, const uint8_t* in192, size_t n, int32_t* status, int32_t* ok);

/* ---- aggregation: AggregatePublicKey::{aggregate, into_aggregate} (M/src/aggregates.rs:29-56) and
 *      AggregateSignature::aggregate (M/src/aggregates.rs:100-106).  Set s owns points off[s] .. off[s+1]-1.
 *      An empty set yields status[s] = B3_ERR_AGGREGATE_EMPTY_POINTS for G1 (the reference's Err) and infinity for G2. */
int b3_g1_aggregate(b3_ctx*, const uint8_t* pks96, const uint32_t* off, size_t n_sets, uint8_t* out96, int32_t* status);
int b3_g2_aggregate(b3_ctx*, const uint8_t* sigs192, const uint32_t* off, size_t n_sets, uint8_t* out192, int32_t* status);

/* ---- hash_to_curve_g2 (M/src/amcl_utils.rs:33-35 -> A/bls381/core.rs:831-839).  Message i is
 *      msgs[off[i] .. off[i+1]).  dst == NULL selects DST_G2 (A/bls381/proof_of_possession.rs:38). */
int b3_hash_to_g2(b3_ctx*, const uint8_t* msgs, const uint32_t* off, size_t n, const uint8_t* dst, size_t dst_len,
                  uint8_t* out192);

/* ---- verification.  accept mirrors the reference's bool; gt (nullable) receives the 576-byte value the
 *      reference compares with one (FP12 after fexp). ---- */
/* Signature::verify (M/src/signature.rs:27-40) */
int b3_verify(b3_ctx*, const uint8_t sig192[192], const uint8_t pk96[96], const uint8_t* msg, size_t msg_len,
              int* accept, uint8_t* gt576);
/* AggregateSignature::fast_aggregate_verify (M/src/aggregates.rs:177-215); n_pks == 0 -> accept = 0 */
int b3_fast_aggregate_verify(b3_ctx*, const uint8_t sig192[192], const uint8_t* pks96, size_t n_pks,
                             const uint8_t* msg, size_t msg_len, int* accept, uint8_t* gt576);
/* AggregateSignature::fast_aggregate_verify_pre_aggregated (M/src/aggregates.rs:223-253) */
int b3_fast_aggregate_verify_pre_aggregated(b3_ctx*, const uint8_t sig192[192], const uint8_t apk96[96],
                                            const uint8_t* msg, size_t msg_len, int* accept, uint8_t* gt576);
/* AggregateSignature::aggregate_verify (M/src/aggregates.rs:130-170); n == 0 -> accept = 0 */
int b3_aggregate_verify(b3_ctx*, const uint8_t sig192[192], const uint8_t* pks96, const uint8_t* msgs,
                        const uint32_t* msg_off, size_t n, int* accept, uint8_t* gt576);
/* AggregateSignature::verify_multiple_aggregate_signatures (M/src/aggregates.rs:261-316).
 * Set j = (sigs192[j], apk_j, msg_j, scalars[j]).  apk_j is either apks96[j] (pk_off == NULL) or the sum of
 * pks96[pk_off[j] .. pk_off[j+1]) aggregated on the device (the C4/C5 shape: 128 keys per set).
 * first_bad (nullable) = index of the first signature failing subgroup_check_g2, or -1; if it is >= 0,
 * accept = 0 exactly as the reference returns false at that set.  n == 0 -> accept = 1. */
int b3_verify_multiple(b3_ctx*, const uint8_t* sigs192, const uint8_t* pks96, const uint32_t* pk_off,
                       const uint8_t* msgs, const uint32_t* msg_off, const uint64_t* scalars, size_t n,
                       int* accept, int64_t* first_bad, uint8_t* gt576);

/* ---- batched verification of n INDEPENDENT items with one accept bit each (locating the bad set
 *      after a batch reject, or bulk verification of unrelated signatures).  Item i is, by `mode`,
 *        B3_ITEM_VERIFY          Signature::verify(sig_i, msg_i, pk_i)                           (M/src/signature.rs:27-40)
 *        B3_ITEM_FAST_AGGREGATE  fast_aggregate_verify(sig_i, msg_i, pks96[pk_off[i] .. pk_off[i+1]))  (M/src/aggregates.rs:177-215)
 *        B3_ITEM_PRE_AGGREGATED  fast_aggregate_verify_pre_aggregated(sig_i, msg_i, apk_i)       (M/src/aggregates.rs:223-253)
 *      (pk_off is read only in mode B3_ITEM_FAST_AGGREGATE; otherwise pks96 holds one key per item).
 *      accept[i] = the reference's bool; status[i] = B3_OK or the AmclError code of a malformed input of item i
 *      (B3_ERR_AGGREGATE_EMPTY_POINTS for an empty key list; accept[i] = 0 in every such case);
 *      gt576 (nullable, n x 576 B) = each item's FP12 after fexp, all-zero for items rejected before the pairing.
 *      One final exponentiation PER ITEM: one CTA per item below 2048 items, one lane pair per item above. ---- */
#define B3_ITEM_VERIFY 0
#define B3_ITEM_FAST_AGGREGATE 1
#define B3_ITEM_PRE_AGGREGATED 2
int b3_verify_batch(b3_ctx*, int mode, const uint8_t* sigs192, const uint8_t* pks96, const uint32_t* pk_off,
                    const uint8_t* msgs, const uint32_t* msg_off, size_t n, int32_t* accept, int32_t* status,
                    uint8_t* gt576);

/* ---- device-resident variants (inputs already in HBM; used for multi-GPU sharding and resident benchmarks).
 *      partial_dev receives this rank's Miller-loop product (B3_PARTIAL_BYTES of device memory, internal
 *      Montgomery layout -- only meaningful to b3_combine_partials_dev of the same library build).
 *      The pair (sum_j [c_j] sig_j, -G1) of this rank's sets is already folded into the partial. ---- */
#define B3_PARTIAL_BYTES 592 /* 576 B Fp12 + int64 first_bad + 8 B pad */
int b3_verify_multiple_partial_dev(b3_ctx*, const uint8_t* sigs192_dev, const uint8_t* pks96_dev,
                                   const uint32_t* pk_off_dev, const uint8_t* msgs_dev, const uint32_t* msg_off_dev,
                                   const uint64_t* scalars_dev, size_t n, int64_t index_base, uint8_t* partial_dev);
/* the same with this rank's shard in HOST memory (the sharded form of b3_verify_multiple: M/src/aggregates.rs:261-316 on
 * sets [index_base, index_base + n)); the partial stays on the device for the all-gather */
int b3_verify_multiple_partial(b3_ctx*, const uint8_t* sigs192, const uint8_t* pks96, const uint32_t* pk_off,
                               const uint8_t* msgs, const uint32_t* msg_off, const uint64_t* scalars, size_t n,
                               int64_t index_base, uint8_t* partial_dev);
/* product of n_partials partials (gathered from all ranks) -> one final exponentiation -> accept, first_bad, gt */
int b3_combine_partials_dev(b3_ctx*, const uint8_t* partials_dev, size_t n_partials, int* accept, int64_t* first_bad,
                            uint8_t* gt576);
int b3_verify_batch_dev(b3_ctx*, int mode, const uint8_t* sigs192_dev, const uint8_t* pks96_dev,
                        const uint32_t* pk_off_dev, const uint8_t* msgs_dev, const uint32_t* msg_off_dev, size_t n,
                        int32_t* accept_dev, int32_t* status_dev, uint8_t* gt576_dev);
int b3_hash_to_g2_dev(b3_ctx*, const uint8_t* msgs_dev, const uint32_t* off_dev, size_t n, uint8_t* out192_dev);
int b3_g1_aggregate_dev(b3_ctx*, const uint8_t* pks96_dev, const uint32_t* off_dev, size_t n_sets, uint8_t* out96_dev,
                        int32_t* status_dev);

/* ---- signing-side helpers.  OUT of the verification path (the reference's Signature::new / SecretKey live on
 *      the CPU); exported only so tests and bench.py can synthesise valid inputs at full size quickly.
 *      scalars32: n x 32-byte big-endian integers (< 2^256). ---- */
int b3_g1_mul_gen(b3_ctx*, const uint8_t* scalars32, size_t n, uint8_t* out96);
int b3_g2_mul(b3_ctx*, const uint8_t* pts192, const uint8_t* scalars32, size_t n, uint8_t* out192);

/* ---- measurement helper: pure-IMAD roofline microbenchmark (returns 32-bit IMAD lane-ops per second) ---- */
int b3_imad_peak(b3_ctx*, int wide /* 0: IMAD, 1: IMAD.WIDE carry chain */, double* ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
