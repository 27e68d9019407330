// C ABI (include/milagro_bls_b200.h) and host-side orchestration of the verification path.
// Single translation unit: the device headers carry __constant__ tables with internal linkage.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/milagro_bls_b200.h"
#undef B3_OK
#undef B3_ERR_INVALID_POINT
#undef B3_ERR_INVALID_YFLAG
#include "kernels.cuh"

static const uint8_t kDstG2[] = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";   // A/bls381/proof_of_possession.rs:38
static const size_t kDstG2Len = 43;

#define B3_MAX_MARKS 24
// sets from which the key aggregation uses 4 lanes per set instead of 8 (less shuffle-tree overhead, twice the latency)
static const size_t kAggG4Min = getenv("B3_AGG_G4_MIN") ? (size_t)atol(getenv("B3_AGG_G4_MIN")) : 4096;
#define B3_MSM_MIN_SETS 512     // below this, S = sum [c_j] sig_j uses n separate ladders + a tree
#define B3_N_STAGES 11
// stage ids (b3_ctx_stage_ms / b3_stage_name)
enum { ST_SIG_CHECK = 0, ST_AGGREGATE, ST_G1_MUL, ST_HASH_TO_G2, ST_G2_MUL_SUM, ST_MILLER, ST_FP12_PRODUCT, ST_FINAL_EXP, ST_COPY, ST_MILLER_LINES, ST_END };
static const char* kStageNames[B3_N_STAGES] = {"g2_parse_subgroup_check", "g1_aggregate", "g1_scalar_mul_affine", "hash_to_g2_affine",
                                               "g2_scalar_mul_sum", "miller_accumulate", "miller_chain", "final_exp", "parse_copies", "miller_lines", "end"};

struct dev_buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct b3_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // 0/1: whole call, 2/3: Miller kernel
    float last_ms[2] = {0.f, 0.f};
    // stage spans of the most recent verification call: span i = [span_a[i], span_b[i]] on the stream the stage ran on
    cudaEvent_t span_a[B3_MAX_MARKS], span_b[B3_MAX_MARKS];
    int span_id[B3_MAX_MARKS];
    int n_spans = 0;
    float stage_ms[B3_N_STAGES];
    // independent stages of verify_multiple run concurrently on aux streams (fork/join by events) unless serial != 0
    cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_fork2 = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    int serial = 0;
    int item_kernel = 0;            // b3_verify_batch finishing kernel: 0 = by batch size, 1 = CTA per item, 2 = thread per item, 3 = lane pair per item
    // scratch (grown on demand, reused across calls)
    dev_buf in_a, in_b, in_c, in_d, in_e, in_f;      // staged host inputs
    dev_buf g1j, g1j2, g1a, g2a_sig, g2j, g2j2, g2j_h, g2a, g2q, g1pp, qinf, f12a, f12b, lines, status, ok, misc, outb;
    uint8_t* d_dst = nullptr;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return B3_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)
#define CKR(expr)                      \
    do {                               \
        int _r = (expr);               \
        if (_r != B3_OK) return _r;    \
    } while (0)

static int ensure(b3_ctx* ctx, dev_buf& b, size_t bytes) {
    if (bytes <= b.cap) return B3_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t cap = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return B3_OK;
}
static inline unsigned nblk(size_t n, int tpb = B3_TPB) { return (unsigned)((n + tpb - 1) / tpb); }
#define LAUNCH(kern, grid, block, ...)                       \
    do {                                                     \
        kern<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                     \
    } while (0)

#define LAUNCH_ON(strm, kern, grid, block, ...)              \
    do {                                                     \
        kern<<<(grid), (block), 0, (strm)>>>(__VA_ARGS__);   \
        ctx->launches++;                                     \
    } while (0)

extern "C" int b3_ctx_create(int device, b3_ctx** out) {
    if (!out) return B3_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return B3_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B3_ERR_CUDA;
    if (prop.major != 10) return B3_ERR_CUDA;        // sm_100a code only: no other path exists
    if (cudaSetDevice(device) != cudaSuccess) return B3_ERR_CUDA;
    b3_ctx* ctx = new b3_ctx();
    ctx->device = device;
    // Priorities (lower number = served first): the context's own stream carries the END of every call (Miller
    // accumulation, closing chain, final exponentiation -- single-CTA kernels that must not queue behind the wide kernels
    // of other contexts' batches), aux[2] the longest dependent chain of a batch (hash_to_G2 -> Miller point chains).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio_mid = prio_hi < prio_lo ? prio_hi + 1 : prio_hi;
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete ctx; return B3_ERR_CUDA; }
    for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev[i]);
    for (int i = 0; i < B3_MAX_MARKS; i++) { cudaEventCreate(&ctx->span_a[i]); cudaEventCreate(&ctx->span_b[i]); }
    for (int i = 0; i < 4; i++) {
        if (cudaStreamCreateWithPriority(&ctx->aux[i], cudaStreamNonBlocking, i == 2 ? prio_mid : prio_lo) != cudaSuccess) { delete ctx; return B3_ERR_CUDA; }
        cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming);
    for (int i = 0; i < B3_N_STAGES; i++) ctx->stage_ms[i] = 0.f;
    if (cudaMalloc((void**)&ctx->d_dst, 256) != cudaSuccess) { delete ctx; return B3_ERR_CUDA; }
    cudaMemcpy(ctx->d_dst, kDstG2, kDstG2Len, cudaMemcpyHostToDevice);
    *out = ctx;
    return B3_OK;
}
extern "C" void b3_ctx_destroy(b3_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dev_buf* bufs[] = {&ctx->in_a, &ctx->in_b, &ctx->in_c, &ctx->in_d, &ctx->in_e, &ctx->in_f, &ctx->g1j, &ctx->g1j2, &ctx->g1a,
                       &ctx->g2a_sig, &ctx->g2j, &ctx->g2j2, &ctx->g2j_h, &ctx->g2a, &ctx->g2q, &ctx->g1pp, &ctx->qinf, &ctx->lines, &ctx->f12a, &ctx->f12b, &ctx->status, &ctx->ok,
                       &ctx->misc, &ctx->outb};
    for (dev_buf* b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->d_dst) cudaFree(ctx->d_dst);
    for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < B3_MAX_MARKS; i++) { cudaEventDestroy(ctx->span_a[i]); cudaEventDestroy(ctx->span_b[i]); }
    for (int i = 0; i < 4; i++) { cudaStreamSynchronize(ctx->aux[i]); cudaStreamDestroy(ctx->aux[i]); cudaEventDestroy(ctx->ev_join[i]); }
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_fork2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char* b3_last_error(b3_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* b3_ctx_stream(b3_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t b3_ctx_launch_count(b3_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" float b3_ctx_last_kernel_ms(b3_ctx* ctx, int which) { return (ctx && which >= 0 && which < 2) ? ctx->last_ms[which] : 0.f; }
extern "C" float b3_ctx_stage_ms(b3_ctx* ctx, int stage) { return (ctx && stage >= 0 && stage < B3_N_STAGES) ? ctx->stage_ms[stage] : 0.f; }
extern "C" const char* b3_stage_name(int stage) { return (stage >= 0 && stage < B3_N_STAGES) ? kStageNames[stage] : ""; }
extern "C" int b3_stage_count(void) { return B3_N_STAGES - 1; }
extern "C" void b3_ctx_set_serial(b3_ctx* ctx, int serial) { if (ctx) ctx->serial = serial; }
extern "C" void b3_ctx_set_item_kernel(b3_ctx* ctx, int which) { if (ctx) ctx->item_kernel = which; }
static void mark_reset(b3_ctx* ctx) { ctx->n_spans = 0; }
// open a stage span on `strm`; returns the span index (or -1 when the table is full)
static int span_begin(b3_ctx* ctx, int id, cudaStream_t strm) {
    if (ctx->n_spans >= B3_MAX_MARKS) return -1;
    int i = ctx->n_spans++;
    ctx->span_id[i] = id;
    cudaEventRecord(ctx->span_a[i], strm);
    return i;
}
static void span_end(b3_ctx* ctx, int i, cudaStream_t strm) {
    if (i >= 0) cudaEventRecord(ctx->span_b[i], strm);
}
// after a stream synchronize: fold the spans into per-stage durations
static void mark_collect(b3_ctx* ctx) {
    for (int i = 0; i < B3_N_STAGES; i++) ctx->stage_ms[i] = 0.f;
    for (int i = 0; i < ctx->n_spans; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->span_a[i], ctx->span_b[i]) == cudaSuccess) ctx->stage_ms[ctx->span_id[i]] += ms;
    }
    cudaGetLastError();
}

static int h2d(b3_ctx* ctx, dev_buf& b, const void* src, size_t bytes) {
    CKR(ensure(ctx, b, bytes ? bytes : 1));
    if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B3_OK;
}
static int d2h(b3_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return B3_OK;
}
static int sync(b3_ctx* ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return B3_OK;
}
static int begin(b3_ctx* ctx) {
    if (!ctx) return B3_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return B3_OK;
}

// ---------------------------------------------------------------------------------------------- generic pieces
// product of n Fp12 values in `a` (scratch `b`), result left in *res (device pointer into a or b)
static int fp12_product(b3_ctx* ctx, fp12* a, fp12* b, size_t n, fp12** res) {
    fp12 *src = a, *dst = b;
    while (n > 1) {
        size_t m = (n + 1) / 2;
        LAUNCH(k_fp12_mul_pairs, nblk(m), B3_TPB, src, n, dst);
        fp12* t = src; src = dst; dst = t;
        n = m;
    }
    *res = src;
    return B3_OK;
}
static int g2_sum(b3_ctx* ctx, g2_jac* a, g2_jac* b, size_t n, g2_jac** res, cudaStream_t strm = nullptr) {
    g2_jac *src = a, *dst = b;
    if (!strm) strm = ctx->stream;
    while (n > 1) {
        size_t m = (n + 1) / 2;
        LAUNCH_ON(strm, k_g2_add_pairs, nblk(2 * m), B3_TPB, src, n, dst);
        g2_jac* t = src; src = dst; dst = t;
        n = m;
    }
    *res = src;
    return B3_OK;
}
// Split multi-Miller loop (pairing.cuh), step 1: point chains of pairs [first, first + count) of an n_pairs product
// -> lines in HBM (68 x n_pairs x 288 B).  May be issued on any stream as soon as those q's exist.
static int miller_lines(b3_ctx* ctx, cudaStream_t strm, const g2_jac* q, size_t n_pairs, size_t first, size_t count) {
    if (count == 0) return B3_OK;
    int sp = span_begin(ctx, ST_MILLER_LINES, strm);
    LAUNCH_ON(strm, k_miller_lines, nblk(2 * count), B3_TPB, q, n_pairs, first, count, (fp2*)ctx->lines.p, (uint32_t*)ctx->qinf.p);
    span_end(ctx, sp, strm);
    return B3_OK;
}
static int miller_reserve(b3_ctx* ctx, size_t n_pairs) {
    CKR(ensure(ctx, ctx->lines, sizeof(fp2) * 3 * B3_MILLER_SLOTS * (n_pairs ? n_pairs : 1)));
    CKR(ensure(ctx, ctx->qinf, 4 * (n_pairs ? n_pairs : 1)));
    return B3_OK;
}
// steps 2 and 3 (context stream): per-slot accumulation over all pairs, one cooperative closing chain -> *res
static int miller_finish(b3_ctx* ctx, const g1_pp* p, size_t n_pairs, fp12** res) {
    // K pairs per accumulating six-lane group, 20 groups per CTA (k_miller_accum)
    static const unsigned kTargetK = getenv("B3_ACC_K") ? (unsigned)atoi(getenv("B3_ACC_K")) : 32u;
    size_t chunks = (n_pairs + (size_t)B3_ACC_GROUPS * kTargetK - 1) / ((size_t)B3_ACC_GROUPS * kTargetK);
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    unsigned K = (unsigned)((n_pairs + chunks * B3_ACC_GROUPS - 1) / (chunks * B3_ACC_GROUPS));
    if (K < 1) K = 1;
    CKR(ensure(ctx, ctx->f12a, sizeof(fp12) * (B3_MILLER_SLOTS * (chunks + 1) + 2)));
    fp12* partial = (fp12*)ctx->f12a.p;
    fp12* slotvals = partial + B3_MILLER_SLOTS * chunks;
    fp12* out = slotvals + B3_MILLER_SLOTS;
    *res = out;
    if (n_pairs == 0) {
        LAUNCH(k_fp12_set_one, 1, 1, out);
        return B3_OK;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    int sp = span_begin(ctx, ST_MILLER, ctx->stream);
    LAUNCH(k_miller_accum, dim3((unsigned)chunks, B3_MILLER_SLOTS), B3_TPB, (const fp2*)ctx->lines.p, (const uint32_t*)ctx->qinf.p, p, n_pairs, K, partial);
    span_end(ctx, sp, ctx->stream);
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    sp = span_begin(ctx, ST_FP12_PRODUCT, ctx->stream);
    if (chunks > 1) {
        LAUNCH(k_miller_slots, B3_MILLER_SLOTS, B3_COOP_THREADS, (const fp12*)partial, (unsigned)chunks, slotvals);
        LAUNCH(k_miller_chain, 1, B3_COOP_THREADS, (const fp12*)slotvals, 1u, out);
    } else {
        LAUNCH(k_miller_chain, 1, B3_COOP_THREADS, (const fp12*)partial, 1u, out);
    }
    span_end(ctx, sp, ctx->stream);
    return B3_OK;
}
// whole product on the context stream
static int miller_product(b3_ctx* ctx, const g2_jac* q, const g1_pp* p, size_t n_pairs, fp12** res) {
    CKR(miller_reserve(ctx, n_pairs));
    CKR(miller_lines(ctx, ctx->stream, q, n_pairs, 0, n_pairs));
    return miller_finish(ctx, p, n_pairs, res);
}
// final exponentiation of *m -> accept / gt on the host
static int finish(b3_ctx* ctx, const fp12* m, int* accept, uint8_t* gt576) {
    CKR(ensure(ctx, ctx->outb, 576 + 16));
    uint8_t* d_gt = (uint8_t*)ctx->outb.p;
    int32_t* d_one = (int32_t*)(d_gt + 576);
    int sp = span_begin(ctx, ST_FINAL_EXP, ctx->stream);
    LAUNCH(k_final_exp, 1, B3_COOP_THREADS, m, d_gt, d_one);
    span_end(ctx, sp, ctx->stream);
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    int32_t one = 0;
    uint8_t gt[576];
    CKR(d2h(ctx, &one, d_one, 4));
    CKR(d2h(ctx, gt, d_gt, 576));
    CKR(sync(ctx));
    cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
    if (cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev[2], ctx->ev[3]) != cudaSuccess) ctx->last_ms[1] = 0.f;
    cudaGetLastError();
    mark_collect(ctx);
    if (gt576) memcpy(gt576, gt, 576);
    if (accept) *accept = one ? 1 : 0;
    return B3_OK;
}
// host copy of a status array; returns the first non-zero code or B3_OK
static int first_status(b3_ctx* ctx, const int32_t* d_status, size_t n) {
    std::vector<int32_t> h(n);
    CKR(d2h(ctx, h.data(), d_status, 4 * n));
    CKR(sync(ctx));
    for (size_t i = 0; i < n; i++) if (h[i]) return h[i];
    return B3_OK;
}
static int g1_aggregate_dev_impl(b3_ctx* ctx, const uint8_t* d_pks, const uint32_t* d_off, size_t n_sets, size_t total_keys,
                                 g1_jac* d_out, int32_t* d_status) {
    if (n_sets == 0) return B3_OK;
    size_t avg = total_keys / n_sets;
    // lanes per set: enough lanes to fill the chip, few enough to keep the shuffle tree short
    if (n_sets >= 16384 || avg <= 8) LAUNCH(k_g1_aggregate<4>, nblk(n_sets * 4), B3_TPB, d_pks, d_off, n_sets, d_out, d_status);
    else if (n_sets >= 2048 || avg <= 32) LAUNCH(k_g1_aggregate<8>, nblk(n_sets * 8), B3_TPB, d_pks, d_off, n_sets, d_out, d_status);
    else LAUNCH(k_g1_aggregate<32>, nblk(n_sets * 32), B3_TPB, d_pks, d_off, n_sets, d_out, d_status);
    return B3_OK;
}

// ---------------------------------------------------------------------------------------------- serialisation API
extern "C" int b3_g1_decompress(b3_ctx* ctx, const uint8_t* in48, size_t n, int validate, uint8_t* out96, int32_t* status) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!in48 || !out96 || !status) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, in48, 48 * n));
    CKR(ensure(ctx, ctx->outb, 96 * n));
    CKR(ensure(ctx, ctx->status, 4 * n));
    LAUNCH(k_g1_decompress, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, validate, (uint8_t*)ctx->outb.p, (int32_t*)ctx->status.p);
    CKR(d2h(ctx, out96, ctx->outb.p, 96 * n));
    CKR(d2h(ctx, status, ctx->status.p, 4 * n));
    return sync(ctx);
}
extern "C" int b3_g2_decompress(b3_ctx* ctx, const uint8_t* in96, size_t n, uint8_t* out192, int32_t* status) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!in96 || !out192 || !status) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, in96, 96 * n));
    CKR(ensure(ctx, ctx->outb, 192 * n));
    CKR(ensure(ctx, ctx->status, 4 * n));
    LAUNCH(k_g2_decompress, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, (uint8_t*)ctx->outb.p, (int32_t*)ctx->status.p);
    CKR(d2h(ctx, out192, ctx->outb.p, 192 * n));
    CKR(d2h(ctx, status, ctx->status.p, 4 * n));
    return sync(ctx);
}
extern "C" int b3_g1_compress(b3_ctx* ctx, const uint8_t* in96, size_t n, uint8_t* out48, int32_t* status) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!in96 || !out48 || !status) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, in96, 96 * n));
    CKR(ensure(ctx, ctx->outb, 48 * n));
    CKR(ensure(ctx, ctx->status, 4 * n));
    LAUNCH(k_g1_compress, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, (uint8_t*)ctx->outb.p, (int32_t*)ctx->status.p);
    CKR(d2h(ctx, out48, ctx->outb.p, 48 * n));
    CKR(d2h(ctx, status, ctx->status.p, 4 * n));
    return sync(ctx);
}
extern "C" int b3_g2_compress(b3_ctx* ctx, const uint8_t* in192, size_t n, uint8_t* out96, int32_t* status) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!in192 || !out96 || !status) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, in192, 192 * n));
    CKR(ensure(ctx, ctx->outb, 96 * n));
    CKR(ensure(ctx, ctx->status, 4 * n));
    LAUNCH(k_g2_compress, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, (uint8_t*)ctx->outb.p, (int32_t*)ctx->status.p);
    CKR(d2h(ctx, out96, ctx->outb.p, 96 * n));
    CKR(d2h(ctx, status, ctx->status.p, 4 * n));
    return sync(ctx);
}
extern "C" int b3_g1_validate(b3_ctx* ctx, const uint8_t* in96, size_t n, int32_t* status, int32_t* valid) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!in96 || !status || !valid) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, in96, 96 * n));
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * n));
    CKR(ensure(ctx, ctx->status, 4 * n));
    CKR(ensure(ctx, ctx->ok, 4 * n));
    LAUNCH(k_g1_parse, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, (g1_jac*)ctx->g1j.p, (int32_t*)ctx->status.p, 1);
    LAUNCH(k_g1_key_validate, nblk(n), B3_TPB, (const g1_jac*)ctx->g1j.p, (const int32_t*)ctx->status.p, n, (int32_t*)ctx->ok.p);
    CKR(d2h(ctx, status, ctx->status.p, 4 * n));
    CKR(d2h(ctx, valid, ctx->ok.p, 4 * n));
    return sync(ctx);
}
extern "C" int b3_g2_subgroup_check(b3_ctx* ctx, const uint8_t* in192, size_t n, int32_t* status, int32_t* ok) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!in192 || !status || !ok) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, in192, 192 * n));
    CKR(ensure(ctx, ctx->g2a_sig, sizeof(g2_aff) * n));
    CKR(ensure(ctx, ctx->status, 4 * n));
    CKR(ensure(ctx, ctx->ok, 4 * n));
    LAUNCH(k_g2_parse, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, (g2_aff*)ctx->g2a_sig.p, (int32_t*)ctx->status.p, 1);
    LAUNCH(k_g2_subgroup, nblk(2 * n), B3_TPB, (const g2_aff*)ctx->g2a_sig.p, (const int32_t*)ctx->status.p, n, (int32_t*)ctx->ok.p);
    CKR(d2h(ctx, status, ctx->status.p, 4 * n));
    CKR(d2h(ctx, ok, ctx->ok.p, 4 * n));
    return sync(ctx);
}

// ---------------------------------------------------------------------------------------------- aggregation API
extern "C" int b3_g1_aggregate_dev(b3_ctx* ctx, const uint8_t* pks96_dev, const uint32_t* off_dev, size_t n_sets, uint8_t* out96_dev,
                                   int32_t* status_dev) {
    CKR(begin(ctx));
    if (n_sets == 0) return B3_OK;
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * n_sets));
    CKR(ensure(ctx, ctx->g1a, sizeof(g1_aff) * n_sets));
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, off_dev + n_sets, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKR(sync(ctx));
    CKR(g1_aggregate_dev_impl(ctx, pks96_dev, off_dev, n_sets, total, (g1_jac*)ctx->g1j.p, status_dev));
    LAUNCH(k_g1_to_affine, nblk(n_sets), B3_TPB, (const g1_jac*)ctx->g1j.p, n_sets, (g1_aff*)ctx->g1a.p);
    LAUNCH(k_g1_aff_to_wire, nblk(n_sets), B3_TPB, (const g1_aff*)ctx->g1a.p, n_sets, out96_dev);
    return sync(ctx);
}
extern "C" int b3_g1_aggregate(b3_ctx* ctx, const uint8_t* pks96, const uint32_t* off, size_t n_sets, uint8_t* out96, int32_t* status) {
    CKR(begin(ctx));
    if (n_sets == 0) return B3_OK;
    if (!off || !out96 || !status) return B3_ERR_ARG;
    size_t total = off[n_sets];
    CKR(h2d(ctx, ctx->in_a, pks96, 96 * total));
    CKR(h2d(ctx, ctx->in_b, off, 4 * (n_sets + 1)));
    CKR(ensure(ctx, ctx->outb, 96 * n_sets));
    CKR(ensure(ctx, ctx->status, 4 * n_sets));
    CKR(b3_g1_aggregate_dev(ctx, (const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, n_sets, (uint8_t*)ctx->outb.p, (int32_t*)ctx->status.p));
    CKR(d2h(ctx, out96, ctx->outb.p, 96 * n_sets));
    CKR(d2h(ctx, status, ctx->status.p, 4 * n_sets));
    return sync(ctx);
}
extern "C" int b3_g2_aggregate(b3_ctx* ctx, const uint8_t* sigs192, const uint32_t* off, size_t n_sets, uint8_t* out192, int32_t* status) {
    CKR(begin(ctx));
    if (n_sets == 0) return B3_OK;
    if (!off || !out192 || !status) return B3_ERR_ARG;
    size_t total = off[n_sets];
    CKR(h2d(ctx, ctx->in_a, sigs192, 192 * total));
    CKR(h2d(ctx, ctx->in_b, off, 4 * (n_sets + 1)));
    CKR(ensure(ctx, ctx->outb, 192 * n_sets));
    CKR(ensure(ctx, ctx->status, 4 * n_sets));
    CKR(ensure(ctx, ctx->g2j, sizeof(g2_jac) * n_sets));
    CKR(ensure(ctx, ctx->g2a, sizeof(g2_aff) * n_sets));
    size_t avg = total / n_sets;
    const uint8_t* d_in = (const uint8_t*)ctx->in_a.p;
    const uint32_t* d_off = (const uint32_t*)ctx->in_b.p;
    if (n_sets >= 16384 || avg <= 8) LAUNCH(k_g2_aggregate<4>, nblk(n_sets * 4), B3_TPB, d_in, d_off, n_sets, (g2_jac*)ctx->g2j.p, (int32_t*)ctx->status.p);
    else if (n_sets >= 2048 || avg <= 32) LAUNCH(k_g2_aggregate<8>, nblk(n_sets * 8), B3_TPB, d_in, d_off, n_sets, (g2_jac*)ctx->g2j.p, (int32_t*)ctx->status.p);
    else LAUNCH(k_g2_aggregate<32>, nblk(n_sets * 32), B3_TPB, d_in, d_off, n_sets, (g2_jac*)ctx->g2j.p, (int32_t*)ctx->status.p);
    LAUNCH(k_g2_to_affine, nblk(n_sets), B3_TPB, (const g2_jac*)ctx->g2j.p, n_sets, (g2_aff*)ctx->g2a.p);
    LAUNCH(k_g2_aff_to_wire, nblk(n_sets), B3_TPB, (const g2_aff*)ctx->g2a.p, n_sets, (uint8_t*)ctx->outb.p);
    CKR(d2h(ctx, out192, ctx->outb.p, 192 * n_sets));
    CKR(d2h(ctx, status, ctx->status.p, 4 * n_sets));
    return sync(ctx);
}

// ---------------------------------------------------------------------------------------------- hash_to_g2 API
static int stage_dst(b3_ctx* ctx, const uint8_t* dst, size_t dst_len, const uint8_t** d_dst, uint32_t* len) {
    if (!dst) { *d_dst = ctx->d_dst; *len = (uint32_t)kDstG2Len; return B3_OK; }
    if (dst_len > 255) return B3_ERR_HASH_TO_FIELD;    // oversize-DST hashing is never reached through milagro_bls
    CKR(h2d(ctx, ctx->misc, dst, dst_len));
    *d_dst = (const uint8_t*)ctx->misc.p;
    *len = (uint32_t)dst_len;
    return B3_OK;
}
// hash_to_curve_g2 left in Jacobian coordinates (pair members of the Miller loop are never normalised)
static int hash_to_g2_jac_dev(b3_ctx* ctx, cudaStream_t strm, const uint8_t* d_msgs, const uint32_t* d_off, size_t n, g2_jac* d_out) {
    LAUNCH_ON(strm, k_hash_to_g2, nblk(2 * n), B3_TPB, d_msgs, d_off, n, (const uint8_t*)ctx->d_dst, (uint32_t)kDstG2Len, d_out);
    return B3_OK;
}
static int hash_to_g2_affine_dev(b3_ctx* ctx, cudaStream_t strm, const uint8_t* d_msgs, const uint32_t* d_off, size_t n, const uint8_t* d_dst,
                                 uint32_t dst_len, g2_aff* d_out) {
    CKR(ensure(ctx, ctx->g2j_h, sizeof(g2_jac) * n));
    LAUNCH_ON(strm, k_hash_to_g2, nblk(2 * n), B3_TPB, d_msgs, d_off, n, d_dst, dst_len, (g2_jac*)ctx->g2j_h.p);
    LAUNCH_ON(strm, k_g2_to_affine, nblk(n), B3_TPB, (const g2_jac*)ctx->g2j_h.p, n, d_out);
    return B3_OK;
}
extern "C" int b3_hash_to_g2_dev(b3_ctx* ctx, const uint8_t* msgs_dev, const uint32_t* off_dev, size_t n, uint8_t* out192_dev) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    CKR(ensure(ctx, ctx->g2a, sizeof(g2_aff) * n));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    int sp = span_begin(ctx, ST_HASH_TO_G2, ctx->stream);
    CKR(hash_to_g2_affine_dev(ctx, ctx->stream, msgs_dev, off_dev, n, ctx->d_dst, (uint32_t)kDstG2Len, (g2_aff*)ctx->g2a.p));
    span_end(ctx, sp, ctx->stream);
    sp = span_begin(ctx, ST_COPY, ctx->stream);
    LAUNCH(k_g2_aff_to_wire, nblk(n), B3_TPB, (const g2_aff*)ctx->g2a.p, n, out192_dev);
    span_end(ctx, sp, ctx->stream);
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CKR(sync(ctx));
    mark_collect(ctx);
    cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
    return B3_OK;
}
extern "C" int b3_hash_to_g2(b3_ctx* ctx, const uint8_t* msgs, const uint32_t* off, size_t n, const uint8_t* dst, size_t dst_len, uint8_t* out192) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!off || !out192) return B3_ERR_ARG;
    const uint8_t* d_dst;
    uint32_t dl;
    CKR(stage_dst(ctx, dst, dst_len, &d_dst, &dl));
    CKR(h2d(ctx, ctx->in_a, msgs, off[n]));
    CKR(h2d(ctx, ctx->in_b, off, 4 * (n + 1)));
    CKR(ensure(ctx, ctx->g2a, sizeof(g2_aff) * n));
    CKR(ensure(ctx, ctx->outb, 192 * n));
    CKR(hash_to_g2_affine_dev(ctx, ctx->stream, (const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, n, d_dst, dl, (g2_aff*)ctx->g2a.p));
    LAUNCH(k_g2_aff_to_wire, nblk(n), B3_TPB, (const g2_aff*)ctx->g2a.p, n, (uint8_t*)ctx->outb.p);
    CKR(d2h(ctx, out192, ctx->outb.p, 192 * n));
    return sync(ctx);
}

// ---------------------------------------------------------------------------------------------- verification
// Shared tail of Signature::verify / fast_aggregate_verify* : pairs (sig, -G1), (H(msg), key).
//   d_sig: parsed signature (g2_aff), d_key: key as g1_jac.  reject_inf_key: the aggregate-key-at-infinity rule.
static int verify_two_pairs(b3_ctx* ctx, const g2_aff* d_sig, const int32_t* d_sig_ok, const g1_jac* d_key, int reject_inf_key,
                            const uint8_t* msg, size_t msg_len, int* accept, uint8_t* gt576) {
    uint32_t off[2] = {0, (uint32_t)msg_len};
    CKR(h2d(ctx, ctx->in_c, msg, msg_len));
    CKR(h2d(ctx, ctx->in_d, off, 8));
    CKR(ensure(ctx, ctx->g2q, sizeof(g2_jac) * 2));
    CKR(ensure(ctx, ctx->g1pp, sizeof(g1_pp) * 2));
    g2_jac* q = (g2_jac*)ctx->g2q.p;
    g1_pp* p = (g1_pp*)ctx->g1pp.p;
    // pair 0: (sig, -G1)
    LAUNCH(k_g2_aff_to_jac, 1, B3_TPB, d_sig, 1, q);
    LAUNCH(k_set_neg_g1_pp, 1, 1, p);
    // pair 1: (H(msg), key)
    CKR(hash_to_g2_jac_dev(ctx, ctx->stream, (const uint8_t*)ctx->in_c.p, (const uint32_t*)ctx->in_d.p, 1, q + 1));
    LAUNCH(k_g1_jac_to_pp, 1, B3_TPB, d_key, 1, p + 1);
    fp12* res;
    CKR(miller_product(ctx, q, p, 2, &res));
    int ok = 0;
    CKR(finish(ctx, res, &ok, gt576));
    int32_t sig_ok = 0;
    g1_pp key;
    CKR(d2h(ctx, &sig_ok, d_sig_ok, 4));
    CKR(d2h(ctx, &key, p + 1, sizeof(g1_pp)));
    CKR(sync(ctx));
    if (!sig_ok) ok = 0;
    if (reject_inf_key && key.inf) ok = 0;
    if (accept) *accept = ok;
    return B3_OK;
}
static int parse_sig(b3_ctx* ctx, const uint8_t* sig192) {
    CKR(h2d(ctx, ctx->in_a, sig192, 192));
    CKR(ensure(ctx, ctx->g2a_sig, sizeof(g2_aff)));
    CKR(ensure(ctx, ctx->status, 4 * 4));
    CKR(ensure(ctx, ctx->ok, 4 * 4));
    LAUNCH(k_g2_parse, 1, B3_TPB, (const uint8_t*)ctx->in_a.p, 1, (g2_aff*)ctx->g2a_sig.p, (int32_t*)ctx->status.p, 1);
    LAUNCH(k_g2_subgroup, 1, B3_TPB, (const g2_aff*)ctx->g2a_sig.p, (const int32_t*)ctx->status.p, 1, (int32_t*)ctx->ok.p);
    return first_status(ctx, (const int32_t*)ctx->status.p, 1);
}
extern "C" int b3_fast_aggregate_verify(b3_ctx* ctx, const uint8_t sig192[192], const uint8_t* pks96, size_t n_pks, const uint8_t* msg,
                                        size_t msg_len, int* accept, uint8_t* gt576) {
    CKR(begin(ctx));
    if (accept) *accept = 0;
    if (!sig192 || (!msg && msg_len)) return B3_ERR_ARG;
    if (n_pks == 0) return B3_OK;                       // M/src/aggregates.rs:179-181
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    CKR(parse_sig(ctx, sig192));
    uint32_t off[2] = {0, (uint32_t)n_pks};
    CKR(h2d(ctx, ctx->in_b, pks96, 96 * n_pks));
    CKR(h2d(ctx, ctx->in_e, off, 8));
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * 2));
    int32_t* d_st = (int32_t*)ctx->status.p + 1;
    CKR(g1_aggregate_dev_impl(ctx, (const uint8_t*)ctx->in_b.p, (const uint32_t*)ctx->in_e.p, 1, n_pks, (g1_jac*)ctx->g1j.p, d_st));
    CKR(first_status(ctx, d_st, 1));
    return verify_two_pairs(ctx, (const g2_aff*)ctx->g2a_sig.p, (const int32_t*)ctx->ok.p, (const g1_jac*)ctx->g1j.p, 1, msg, msg_len, accept, gt576);
}
static int verify_single_key(b3_ctx* ctx, const uint8_t* sig192, const uint8_t* pk96, int reject_inf, const uint8_t* msg, size_t msg_len,
                             int* accept, uint8_t* gt576) {
    CKR(begin(ctx));
    if (accept) *accept = 0;
    if (!sig192 || !pk96 || (!msg && msg_len)) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    CKR(parse_sig(ctx, sig192));
    CKR(h2d(ctx, ctx->in_b, pk96, 96));
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * 2));
    int32_t* d_st = (int32_t*)ctx->status.p + 1;
    LAUNCH(k_g1_parse, 1, B3_TPB, (const uint8_t*)ctx->in_b.p, 1, (g1_jac*)ctx->g1j.p, d_st, 1);
    CKR(first_status(ctx, d_st, 1));
    return verify_two_pairs(ctx, (const g2_aff*)ctx->g2a_sig.p, (const int32_t*)ctx->ok.p, (const g1_jac*)ctx->g1j.p, reject_inf, msg, msg_len, accept, gt576);
}
extern "C" int b3_verify(b3_ctx* ctx, const uint8_t sig192[192], const uint8_t pk96[96], const uint8_t* msg, size_t msg_len, int* accept,
                         uint8_t* gt576) {
    return verify_single_key(ctx, sig192, pk96, 0, msg, msg_len, accept, gt576);
}
extern "C" int b3_fast_aggregate_verify_pre_aggregated(b3_ctx* ctx, const uint8_t sig192[192], const uint8_t apk96[96], const uint8_t* msg,
                                                       size_t msg_len, int* accept, uint8_t* gt576) {
    return verify_single_key(ctx, sig192, apk96, 1, msg, msg_len, accept, gt576);
}

extern "C" int b3_aggregate_verify(b3_ctx* ctx, const uint8_t sig192[192], const uint8_t* pks96, const uint8_t* msgs, const uint32_t* msg_off,
                                   size_t n, int* accept, uint8_t* gt576) {
    CKR(begin(ctx));
    if (accept) *accept = 0;
    if (n == 0) return B3_OK;                            // M/src/aggregates.rs:132-134
    if (!sig192 || !pks96 || !msg_off) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    CKR(parse_sig(ctx, sig192));
    CKR(h2d(ctx, ctx->in_b, pks96, 96 * n));
    CKR(h2d(ctx, ctx->in_c, msgs, msg_off[n]));
    CKR(h2d(ctx, ctx->in_d, msg_off, 4 * (n + 1)));
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * n));
    CKR(ensure(ctx, ctx->g1pp, sizeof(g1_pp) * (n + 1)));
    CKR(ensure(ctx, ctx->g2q, sizeof(g2_jac) * (n + 1)));
    CKR(ensure(ctx, ctx->status, 4 * (n + 4)));
    int32_t* d_st = (int32_t*)ctx->status.p + 4;
    LAUNCH(k_g1_parse, nblk(n), B3_TPB, (const uint8_t*)ctx->in_b.p, n, (g1_jac*)ctx->g1j.p, d_st, 1);
    CKR(first_status(ctx, d_st, n));
    g2_jac* q = (g2_jac*)ctx->g2q.p;
    g1_pp* p = (g1_pp*)ctx->g1pp.p;
    LAUNCH(k_g1_jac_to_pp, nblk(n), B3_TPB, (const g1_jac*)ctx->g1j.p, n, p);
    CKR(hash_to_g2_jac_dev(ctx, ctx->stream, (const uint8_t*)ctx->in_c.p, (const uint32_t*)ctx->in_d.p, n, q));
    LAUNCH(k_g2_aff_to_jac, 1, B3_TPB, (const g2_aff*)ctx->g2a_sig.p, 1, q + n);
    LAUNCH(k_set_neg_g1_pp, 1, 1, p + n);
    fp12* res;
    CKR(miller_product(ctx, q, p, n + 1, &res));
    int ok = 0;
    CKR(finish(ctx, res, &ok, gt576));
    int32_t sig_ok = 0;
    CKR(d2h(ctx, &sig_ok, ctx->ok.p, 4));
    CKR(sync(ctx));
    if (accept) *accept = (ok && sig_ok) ? 1 : 0;
    return B3_OK;
}

// ---- batched per-item verification: n independent items, one accept bit each ------------------------
#define B3_ITEMS_PAIR_MIN 2048
static int verify_batch_core(b3_ctx* ctx, int mode, const uint8_t* d_sigs, const uint8_t* d_pks, const uint32_t* d_pk_off, size_t total_keys,
                             const uint8_t* d_msgs, const uint32_t* d_msg_off, size_t n, int32_t* d_accept, int32_t* d_status, uint8_t* d_gt) {
    CKR(ensure(ctx, ctx->g2a_sig, sizeof(g2_aff) * n));
    CKR(ensure(ctx, ctx->status, 4 * (2 * n + 8)));
    CKR(ensure(ctx, ctx->ok, 4 * (n + 8)));
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * n));
    CKR(ensure(ctx, ctx->g1pp, sizeof(g1_pp) * n));
    CKR(ensure(ctx, ctx->g2q, sizeof(g2_jac) * 2 * n));
    CKR(miller_reserve(ctx, 2 * n));
    int32_t* d_st_sig = (int32_t*)ctx->status.p;
    int32_t* d_st_key = d_st_sig + n + 4;
    g2_jac* q = (g2_jac*)ctx->g2q.p;              // q[0 .. n) = signatures, q[n .. 2n) = H(msg)
    g1_pp* keys = (g1_pp*)ctx->g1pp.p;
    cudaStream_t sm = ctx->stream;
    cudaStream_t s0 = ctx->serial ? sm : ctx->aux[0], s1 = ctx->serial ? sm : ctx->aux[1], s2 = ctx->serial ? sm : ctx->aux[2];
    int sp;
    if (!ctx->serial) {
        CK(cudaEventRecord(ctx->ev_fork, sm));
        CK(cudaStreamWaitEvent(s1, ctx->ev_fork, 0));
        CK(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
    }
    // H_i = hash_to_curve_g2(msg_i) and its point chain: the longest dependent chain, issued first
    sp = span_begin(ctx, ST_HASH_TO_G2, s2);
    CKR(hash_to_g2_jac_dev(ctx, s2, d_msgs, d_msg_off, n, q + n));
    span_end(ctx, sp, s2);
    CKR(miller_lines(ctx, s2, q, 2 * n, n, n));
    // signatures: parse + on-curve, then the subgroup checks (aux0) beside their point chains (main)
    sp = span_begin(ctx, ST_COPY, sm);
    LAUNCH_ON(sm, k_g2_parse, nblk(n), B3_TPB, d_sigs, n, (g2_aff*)ctx->g2a_sig.p, d_st_sig, 1);
    LAUNCH_ON(sm, k_g2_aff_to_jac, nblk(n), B3_TPB, (const g2_aff*)ctx->g2a_sig.p, n, q);
    span_end(ctx, sp, sm);
    if (!ctx->serial) {
        CK(cudaEventRecord(ctx->ev_fork2, sm));
        CK(cudaStreamWaitEvent(s0, ctx->ev_fork2, 0));
    }
    sp = span_begin(ctx, ST_SIG_CHECK, s0);
    LAUNCH_ON(s0, k_g2_subgroup, nblk(2 * n), B3_TPB, (const g2_aff*)ctx->g2a_sig.p, (const int32_t*)d_st_sig, n, (int32_t*)ctx->ok.p);
    span_end(ctx, sp, s0);
    CKR(miller_lines(ctx, sm, q, 2 * n, 0, n));
    // keys: aggregate (fast_aggregate_verify) or parse (verify / pre-aggregated) -> pairing form
    sp = span_begin(ctx, ST_AGGREGATE, s1);
    if (mode == B3_ITEM_FAST_AGGREGATE) {
        size_t avg = total_keys / n;
        if (n >= kAggG4Min || avg <= 8) LAUNCH_ON(s1, k_g1_aggregate<4>, nblk(n * 4), B3_TPB, d_pks, d_pk_off, n, (g1_jac*)ctx->g1j.p, d_st_key);
        else if (n >= 2048 || avg <= 32) LAUNCH_ON(s1, k_g1_aggregate<8>, nblk(n * 8), B3_TPB, d_pks, d_pk_off, n, (g1_jac*)ctx->g1j.p, d_st_key);
        else LAUNCH_ON(s1, k_g1_aggregate<32>, nblk(n * 32), B3_TPB, d_pks, d_pk_off, n, (g1_jac*)ctx->g1j.p, d_st_key);
    } else {
        LAUNCH_ON(s1, k_g1_parse, nblk(n), B3_TPB, d_pks, n, (g1_jac*)ctx->g1j.p, d_st_key, 1);
    }
    LAUNCH_ON(s1, k_g1_jac_to_pp, nblk(n), B3_TPB, (const g1_jac*)ctx->g1j.p, n, keys);
    span_end(ctx, sp, s1);
    if (!ctx->serial) {
        cudaStream_t auxs[3] = {s0, s1, s2};
        for (int k = 0; k < 3; k++) {
            CK(cudaEventRecord(ctx->ev_join[k], auxs[k]));
            CK(cudaStreamWaitEvent(sm, ctx->ev_join[k], 0));
        }
    }
    CK(cudaEventRecord(ctx->ev[2], sm));
    sp = span_begin(ctx, ST_FINAL_EXP, sm);
    // CTA per item: ~1.5 ms per wave of 2 x 148 items; lane pair per item: one item's latency (~12 ms) for any batch that fits
    // the machine (17 ms at 16384 items, 33 ms at 32768) -- the crossover is near 2.3 k items
    const bool per_thread = ctx->item_kernel == 2;          // measured slower than the lane-pair kernel at every batch size; kept selectable
    const bool per_pair = ctx->item_kernel == 3 || (ctx->item_kernel == 0 && n >= B3_ITEMS_PAIR_MIN);
    if (per_pair)
        LAUNCH_ON(sm, k_items_finish_p, (unsigned)((2 * n + B3_ITEMS_PAIR_TPB - 1) / B3_ITEMS_PAIR_TPB), B3_ITEMS_PAIR_TPB, (const fp2*)ctx->lines.p,
                  (const uint32_t*)ctx->qinf.p, (const g1_pp*)keys, n, (const int32_t*)d_st_sig, (const int32_t*)d_st_key,
                  (const int32_t*)ctx->ok.p, mode == B3_ITEM_VERIFY ? 0 : 1, d_accept, d_status, d_gt);
    else if (per_thread)
        LAUNCH_ON(sm, k_items_finish_t, (unsigned)((n + B3_ITEMS_TPB - 1) / B3_ITEMS_TPB), B3_ITEMS_TPB, (const fp2*)ctx->lines.p,
                  (const uint32_t*)ctx->qinf.p, (const g1_pp*)keys, n, (const int32_t*)d_st_sig, (const int32_t*)d_st_key,
                  (const int32_t*)ctx->ok.p, mode == B3_ITEM_VERIFY ? 0 : 1, d_accept, d_status, d_gt);
    else
        LAUNCH_ON(sm, k_items_finish, (unsigned)n, B3_COOP_THREADS, (const fp2*)ctx->lines.p, (const uint32_t*)ctx->qinf.p, (const g1_pp*)keys, n,
                  (const int32_t*)d_st_sig, (const int32_t*)d_st_key, (const int32_t*)ctx->ok.p, mode == B3_ITEM_VERIFY ? 0 : 1, d_accept, d_status, d_gt);
    span_end(ctx, sp, sm);
    CK(cudaEventRecord(ctx->ev[3], sm));
    return B3_OK;
}
static int verify_batch_done(b3_ctx* ctx) {
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CKR(sync(ctx));
    mark_collect(ctx);
    cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
    if (cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev[2], ctx->ev[3]) != cudaSuccess) ctx->last_ms[1] = 0.f;
    cudaGetLastError();
    return B3_OK;
}
extern "C" int b3_verify_batch_dev(b3_ctx* ctx, int mode, const uint8_t* sigs192_dev, const uint8_t* pks96_dev, const uint32_t* pk_off_dev,
                                   const uint8_t* msgs_dev, const uint32_t* msg_off_dev, size_t n, int32_t* accept_dev, int32_t* status_dev,
                                   uint8_t* gt576_dev) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (mode < B3_ITEM_VERIFY || mode > B3_ITEM_PRE_AGGREGATED || n > 0x3fffffffu) return B3_ERR_ARG;
    if (!sigs192_dev || !pks96_dev || !msg_off_dev || !accept_dev || !status_dev || (mode == B3_ITEM_FAST_AGGREGATE && !pk_off_dev)) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    size_t total_keys = n;
    if (mode == B3_ITEM_FAST_AGGREGATE) {
        uint32_t t = 0;
        CK(cudaMemcpyAsync(&t, pk_off_dev + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CKR(sync(ctx));
        total_keys = t;
    }
    CKR(verify_batch_core(ctx, mode, sigs192_dev, pks96_dev, pk_off_dev, total_keys, msgs_dev, msg_off_dev, n, accept_dev, status_dev, gt576_dev));
    return verify_batch_done(ctx);
}
extern "C" int b3_verify_batch(b3_ctx* ctx, int mode, const uint8_t* sigs192, const uint8_t* pks96, const uint32_t* pk_off, const uint8_t* msgs,
                               const uint32_t* msg_off, size_t n, int32_t* accept, int32_t* status, uint8_t* gt576) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (mode < B3_ITEM_VERIFY || mode > B3_ITEM_PRE_AGGREGATED || n > 0x3fffffffu) return B3_ERR_ARG;
    if (!sigs192 || !pks96 || !msg_off || !accept || !status || (mode == B3_ITEM_FAST_AGGREGATE && !pk_off)) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    const size_t total_keys = mode == B3_ITEM_FAST_AGGREGATE ? pk_off[n] : n;
    CKR(h2d(ctx, ctx->in_a, sigs192, 192 * n));
    CKR(h2d(ctx, ctx->in_b, pks96, 96 * total_keys));
    if (mode == B3_ITEM_FAST_AGGREGATE) CKR(h2d(ctx, ctx->in_e, pk_off, 4 * (n + 1)));
    CKR(h2d(ctx, ctx->in_c, msgs, msg_off[n]));
    CKR(h2d(ctx, ctx->in_d, msg_off, 4 * (n + 1)));
    CKR(ensure(ctx, ctx->outb, (gt576 ? 576 * n : 0) + 8 * n + 16));
    int32_t* d_accept = (int32_t*)ctx->outb.p;
    int32_t* d_status = d_accept + n;
    uint8_t* d_gt = gt576 ? (uint8_t*)(d_status + n) : nullptr;
    CKR(verify_batch_core(ctx, mode, (const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, (const uint32_t*)ctx->in_e.p, total_keys,
                          (const uint8_t*)ctx->in_c.p, (const uint32_t*)ctx->in_d.p, n, d_accept, d_status, d_gt));
    CKR(d2h(ctx, accept, d_accept, 4 * n));
    CKR(d2h(ctx, status, d_status, 4 * n));
    if (gt576) CKR(d2h(ctx, gt576, d_gt, 576 * n));
    return verify_batch_done(ctx);
}

// core of verify_multiple on device-resident inputs; leaves this rank's Miller product in *res and the first bad index in d_first_bad
static int verify_multiple_core(b3_ctx* ctx, const uint8_t* d_sigs, const uint8_t* d_pks, const uint32_t* d_pk_off, size_t total_keys,
                                const uint8_t* d_msgs, const uint32_t* d_msg_off, const uint64_t* d_scalars, size_t n, long long index_base,
                                fp12** res, long long** d_first_bad_out, int* parse_err, const uint8_t* h_pks = nullptr) {
    CKR(ensure(ctx, ctx->g2a_sig, sizeof(g2_aff) * (n + 1)));
    CKR(ensure(ctx, ctx->status, 4 * (2 * n + 8)));
    CKR(ensure(ctx, ctx->ok, 4 * (n + 8)));
    CKR(ensure(ctx, ctx->g1j, sizeof(g1_jac) * (n + 1)));
    CKR(ensure(ctx, ctx->g1j2, sizeof(g1_jac) * (n + 1)));
    CKR(ensure(ctx, ctx->g1pp, sizeof(g1_pp) * (n + B3_MSM_WINDOWS)));
    CKR(ensure(ctx, ctx->g2q, sizeof(g2_jac) * (n + B3_MSM_WINDOWS)));
    CKR(ensure(ctx, ctx->g2j2, sizeof(g2_jac) * (n + 2)));
    CKR(ensure(ctx, ctx->misc, 64));
    int32_t* d_st_sig = (int32_t*)ctx->status.p;
    int32_t* d_st_key = d_st_sig + n + 4;
    long long* d_first_bad = (long long*)ctx->misc.p;
    long long init = 0x7fffffffffffffffLL;
    CK(cudaMemcpyAsync(d_first_bad, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
    g2_jac* q = (g2_jac*)ctx->g2q.p;
    g1_pp* p = (g1_pp*)ctx->g1pp.p;
    const size_t n_total = n == 0 ? 0 : n + (n >= B3_MSM_MIN_SETS ? B3_MSM_WINDOWS : 1);
    CKR(miller_reserve(ctx, n_total));
    if (n > 0) {
        CKR(ensure(ctx, ctx->g2j, sizeof(g2_jac) * (n + 1)));
        // The stages below are independent of each other; unless ctx->serial they run concurrently:
        //   main : parse signatures (... and, after the join, the end of the call: accumulation, closing chain)
        //   aux0 : subgroup checks of the parsed signatures               aux1 : aggregate keys -> P_j = [c_j] apk_j
        //   aux2 : H_j = hash_to_curve_g2(msg_j) -> their point chains    aux3 : S = sum_j [c_j] sig_j -> its point chains
        cudaStream_t sm = ctx->stream;
        cudaStream_t s0 = ctx->serial ? sm : ctx->aux[0], s1 = ctx->serial ? sm : ctx->aux[1], s2 = ctx->serial ? sm : ctx->aux[2];
        cudaStream_t s3 = ctx->serial ? sm : ctx->aux[3];
        int sp;
        if (!ctx->serial) {
            CK(cudaEventRecord(ctx->ev_fork, sm));
            CK(cudaStreamWaitEvent(s1, ctx->ev_fork, 0));
            CK(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
        }
        // 4. H_j = hash_to_curve_g2(msg_j) (M/src/aggregates.rs:290) and its Miller point chain: the longest dependent chain of
        //    the batch, so it is issued FIRST (blocks are dispatched in launch order) on the high-priority stream
        sp = span_begin(ctx, ST_HASH_TO_G2, s2);
        CKR(hash_to_g2_jac_dev(ctx, s2, d_msgs, d_msg_off, n, q));
        span_end(ctx, sp, s2);
        CKR(miller_lines(ctx, s2, q, n_total, 0, n));                 // the point chains need only H_j
        // 1. signatures: parse + on-curve (main), subgroup check (M/src/aggregates.rs:274-276) on aux0
        sp = span_begin(ctx, ST_COPY, sm);
        LAUNCH_ON(sm, k_g2_parse, nblk(n), B3_TPB, d_sigs, n, (g2_aff*)ctx->g2a_sig.p, d_st_sig, 1);
        span_end(ctx, sp, sm);
        if (!ctx->serial) {
            CK(cudaEventRecord(ctx->ev_fork2, sm));
            CK(cudaStreamWaitEvent(s0, ctx->ev_fork2, 0));
            CK(cudaStreamWaitEvent(s3, ctx->ev_fork2, 0));
        }
        sp = span_begin(ctx, ST_SIG_CHECK, s0);
        LAUNCH_ON(s0, k_g2_subgroup, nblk(2 * n), B3_TPB, (const g2_aff*)ctx->g2a_sig.p, (const int32_t*)d_st_sig, n, (int32_t*)ctx->ok.p);
        LAUNCH_ON(s0, k_first_bad, nblk(n), B3_TPB, (const int32_t*)ctx->ok.p, n, index_base, d_first_bad);
        span_end(ctx, sp, s0);
        // 2. aggregate public keys; 3. P_j = [c_j] apk_j (M/src/aggregates.rs:293), affine
        // host-pointer entry: the public keys (96 % of the input bytes) are copied on THIS stream, so the transfer
        // overlaps hash_to_G2 and the signature work instead of preceding them
        if (h_pks) CK(cudaMemcpyAsync((void*)d_pks, h_pks, 96 * total_keys, cudaMemcpyHostToDevice, s1));
        sp = span_begin(ctx, ST_AGGREGATE, s1);
        if (d_pk_off) {
            size_t avg = total_keys / n;
            if (n >= kAggG4Min || avg <= 8) LAUNCH_ON(s1, k_g1_aggregate<4>, nblk(n * 4), B3_TPB, d_pks, d_pk_off, n, (g1_jac*)ctx->g1j.p, d_st_key);
            else if (n >= 2048 || avg <= 32) LAUNCH_ON(s1, k_g1_aggregate<8>, nblk(n * 8), B3_TPB, d_pks, d_pk_off, n, (g1_jac*)ctx->g1j.p, d_st_key);
            else LAUNCH_ON(s1, k_g1_aggregate<32>, nblk(n * 32), B3_TPB, d_pks, d_pk_off, n, (g1_jac*)ctx->g1j.p, d_st_key);
        } else {
            LAUNCH_ON(s1, k_g1_parse, nblk(n), B3_TPB, d_pks, n, (g1_jac*)ctx->g1j.p, d_st_key, 1);
        }
        span_end(ctx, sp, s1);
        sp = span_begin(ctx, ST_G1_MUL, s1);
        LAUNCH_ON(s1, k_g1_mul_u64_pp, nblk(n), B3_TPB, (const g1_jac*)ctx->g1j.p, d_scalars, n, p);
        span_end(ctx, sp, s1);
        // 5. S = sum_j [c_j] sig_j (M/src/aggregates.rs:303), on aux3: the context's own (highest-priority) stream is left to the
        //    end of the call
        sp = span_begin(ctx, ST_G2_MUL_SUM, s3);
        if (n >= B3_MSM_MIN_SETS) {           // bucket method: 8 window sums, each its own pair against -[2^(8w)] G1
            const unsigned segs = msm_segs(n);
            const size_t np = msm_parts(segs);
            CKR(ensure(ctx, ctx->g2j, sizeof(g2_jac) * np));
            CKR(ensure(ctx, ctx->g2j2, sizeof(g2_jac) * B3_MSM_WINDOWS * 256));
            LAUNCH_ON(s3, k_msm_bucket, nblk(2 * np), B3_TPB, (const g2_aff*)ctx->g2a_sig.p, d_scalars, n, (g2_jac*)ctx->g2j.p, segs);
            LAUNCH_ON(s3, k_msm_scale, nblk(2 * B3_MSM_WINDOWS * 256), B3_TPB, (const g2_jac*)ctx->g2j.p, (g2_jac*)ctx->g2j2.p, segs);
            LAUNCH_ON(s3, k_msm_window_sum, B3_MSM_WINDOWS, 512, (g2_jac*)ctx->g2j2.p);
            LAUNCH_ON(s3, k_msm_pairs, 1, 32, (const g2_jac*)ctx->g2j2.p, q + n, p + n);
        } else {
            g2_jac* s;
            CKR(ensure(ctx, ctx->g2j, sizeof(g2_jac) * (n + 1)));
            LAUNCH_ON(s3, k_g2_mul_u64, nblk(2 * n), B3_TPB, (const g2_aff*)ctx->g2a_sig.p, d_scalars, n, (g2_jac*)ctx->g2j.p);
            CKR(g2_sum(ctx, (g2_jac*)ctx->g2j.p, (g2_jac*)ctx->g2j2.p, n, &s, s3));
            CK(cudaMemcpyAsync(q + n, s, sizeof(g2_jac), cudaMemcpyDeviceToDevice, s3));
            LAUNCH_ON(s3, k_set_neg_g1_pp, 1, 1, p + n);
        }
        span_end(ctx, sp, s3);
        CKR(miller_lines(ctx, s3, q, n_total, n, n_total - n));       // ... and the window sums / S
        if (!ctx->serial) {
            cudaStream_t auxs[4] = {s0, s1, s2, s3};
            for (int k = 0; k < 4; k++) {
                CK(cudaEventRecord(ctx->ev_join[k], auxs[k]));
                CK(cudaStreamWaitEvent(sm, ctx->ev_join[k], 0));
            }
        }
    }
    // 6. Miller loops over the n + 1 pairs, product
    CKR(miller_finish(ctx, p, n_total, res));
    *d_first_bad_out = d_first_bad;
    // wire-format errors of the inputs (cannot happen for values that came out of the reference's own types)
    *parse_err = B3_OK;
    if (n > 0) {
        std::vector<int32_t> h(2 * n + 8);
        CKR(d2h(ctx, h.data(), ctx->status.p, 4 * (2 * n + 8)));
        CKR(sync(ctx));
        for (size_t i = 0; i < n && *parse_err == B3_OK; i++) {
            if (h[i]) *parse_err = h[i];
            else if (h[n + 4 + i] && h[n + 4 + i] != B3_ERR_AGGREGATE_EMPTY_POINTS) *parse_err = h[n + 4 + i];
        }
    }
    return B3_OK;
}

extern "C" int b3_verify_multiple(b3_ctx* ctx, const uint8_t* sigs192, const uint8_t* pks96, const uint32_t* pk_off, const uint8_t* msgs,
                                  const uint32_t* msg_off, const uint64_t* scalars, size_t n, int* accept, int64_t* first_bad, uint8_t* gt576) {
    CKR(begin(ctx));
    if (accept) *accept = 0;
    if (first_bad) *first_bad = -1;
    if (n && (!sigs192 || !pks96 || !msg_off || !scalars)) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    size_t total_keys = pk_off ? pk_off[n] : n;
    if (n) {
        CKR(h2d(ctx, ctx->in_a, sigs192, 192 * n));
        CKR(ensure(ctx, ctx->in_b, 96 * total_keys + 1));             // copied inside the core, on the aggregation stream
        if (pk_off) CKR(h2d(ctx, ctx->in_e, pk_off, 4 * (n + 1)));
        CKR(h2d(ctx, ctx->in_c, msgs, msg_off[n]));
        CKR(h2d(ctx, ctx->in_d, msg_off, 4 * (n + 1)));
        CKR(h2d(ctx, ctx->in_f, scalars, 8 * n));
    }
    fp12* res;
    long long* d_fb;
    int perr;
    CKR(verify_multiple_core(ctx, (const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, pk_off ? (const uint32_t*)ctx->in_e.p : nullptr,
                             total_keys, (const uint8_t*)ctx->in_c.p, (const uint32_t*)ctx->in_d.p, (const uint64_t*)ctx->in_f.p, n, 0, &res,
                             &d_fb, &perr, pks96));
    if (perr) return perr;
    int ok = 0;
    CKR(finish(ctx, res, &ok, gt576));
    long long fb = 0;
    CKR(d2h(ctx, &fb, d_fb, 8));
    CKR(sync(ctx));
    if (fb == 0x7fffffffffffffffLL) fb = -1;
    if (first_bad) *first_bad = fb;
    if (accept) *accept = (ok && fb < 0) ? 1 : 0;
    return B3_OK;
}

struct partial_rec {
    fp12 f;
    long long first_bad;
    long long pad;
};
static_assert(sizeof(partial_rec) == B3_PARTIAL_BYTES, "partial record layout");

__global__ void k_pack_partial(const fp12* f, const long long* fb, partial_rec* out) {
    out->f = *f;
    out->first_bad = *fb;
    out->pad = 0;
}
__global__ void k_unpack_partials(const partial_rec* in, size_t n, fp12* f, long long* fb) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f[i] = in[i].f;
    atomicMin(fb, in[i].first_bad);
}

extern "C" int b3_verify_multiple_partial_dev(b3_ctx* ctx, const uint8_t* sigs192_dev, const uint8_t* pks96_dev, const uint32_t* pk_off_dev,
                                              const uint8_t* msgs_dev, const uint32_t* msg_off_dev, const uint64_t* scalars_dev, size_t n,
                                              int64_t index_base, uint8_t* partial_dev) {
    CKR(begin(ctx));
    if (!partial_dev) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    size_t total_keys = n;
    if (pk_off_dev && n) {
        uint32_t t = 0;
        CK(cudaMemcpyAsync(&t, pk_off_dev + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CKR(sync(ctx));
        total_keys = t;
    }
    fp12* res;
    long long* d_fb;
    int perr;
    CKR(verify_multiple_core(ctx, sigs192_dev, pks96_dev, pk_off_dev, total_keys, msgs_dev, msg_off_dev, scalars_dev, n, index_base, &res, &d_fb,
                             &perr));
    if (perr) return perr;
    LAUNCH(k_pack_partial, 1, 1, (const fp12*)res, (const long long*)d_fb, (partial_rec*)partial_dev);
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CKR(sync(ctx));
    mark_collect(ctx);
    cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
    if (cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev[2], ctx->ev[3]) != cudaSuccess) ctx->last_ms[1] = 0.f;
    cudaGetLastError();
    return B3_OK;
}
// host-pointer form of the sharded call: this rank's shard comes from HOST memory (the keys are copied on the aggregation
// stream, overlapped with the other stages, as in b3_verify_multiple); the partial stays on the device for the all-gather
extern "C" int b3_verify_multiple_partial(b3_ctx* ctx, const uint8_t* sigs192, const uint8_t* pks96, const uint32_t* pk_off, const uint8_t* msgs,
                                          const uint32_t* msg_off, const uint64_t* scalars, size_t n, int64_t index_base, uint8_t* partial_dev) {
    CKR(begin(ctx));
    if (!partial_dev) return B3_ERR_ARG;
    if (n && (!sigs192 || !pks96 || !msg_off || !scalars)) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    size_t total_keys = pk_off ? pk_off[n] : n;
    if (n) {
        CKR(h2d(ctx, ctx->in_a, sigs192, 192 * n));
        CKR(ensure(ctx, ctx->in_b, 96 * total_keys + 1));             // copied inside the core, on the aggregation stream
        if (pk_off) CKR(h2d(ctx, ctx->in_e, pk_off, 4 * (n + 1)));
        CKR(h2d(ctx, ctx->in_c, msgs, msg_off[n]));
        CKR(h2d(ctx, ctx->in_d, msg_off, 4 * (n + 1)));
        CKR(h2d(ctx, ctx->in_f, scalars, 8 * n));
    }
    fp12* res;
    long long* d_fb;
    int perr;
    CKR(verify_multiple_core(ctx, (const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, pk_off ? (const uint32_t*)ctx->in_e.p : nullptr,
                             total_keys, (const uint8_t*)ctx->in_c.p, (const uint32_t*)ctx->in_d.p, (const uint64_t*)ctx->in_f.p, n, index_base,
                             &res, &d_fb, &perr, pks96));
    if (perr) return perr;
    LAUNCH(k_pack_partial, 1, 1, (const fp12*)res, (const long long*)d_fb, (partial_rec*)partial_dev);
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CKR(sync(ctx));
    mark_collect(ctx);
    cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
    if (cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev[2], ctx->ev[3]) != cudaSuccess) ctx->last_ms[1] = 0.f;
    cudaGetLastError();
    return B3_OK;
}
extern "C" int b3_combine_partials_dev(b3_ctx* ctx, const uint8_t* partials_dev, size_t n_partials, int* accept, int64_t* first_bad,
                                       uint8_t* gt576) {
    CKR(begin(ctx));
    if (accept) *accept = 0;
    if (first_bad) *first_bad = -1;
    if (!partials_dev || n_partials == 0) return B3_ERR_ARG;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CKR(ensure(ctx, ctx->f12a, sizeof(fp12) * (n_partials + 1)));
    CKR(ensure(ctx, ctx->f12b, sizeof(fp12) * (n_partials / 2 + 2)));
    CKR(ensure(ctx, ctx->misc, 64));
    long long* d_fb = (long long*)ctx->misc.p;
    long long init = 0x7fffffffffffffffLL;
    CK(cudaMemcpyAsync(d_fb, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_unpack_partials, nblk(n_partials), B3_TPB, (const partial_rec*)partials_dev, n_partials, (fp12*)ctx->f12a.p, d_fb);
    fp12* res;
    CKR(fp12_product(ctx, (fp12*)ctx->f12a.p, (fp12*)ctx->f12b.p, n_partials, &res));
    int ok = 0;
    CKR(finish(ctx, res, &ok, gt576));
    long long fb = 0;
    CKR(d2h(ctx, &fb, d_fb, 8));
    CKR(sync(ctx));
    if (fb == 0x7fffffffffffffffLL) fb = -1;
    if (first_bad) *first_bad = fb;
    if (accept) *accept = (ok && fb < 0) ? 1 : 0;
    return B3_OK;
}

// ---------------------------------------------------------------------------------------------- signing-side helpers
extern "C" int b3_g1_mul_gen(b3_ctx* ctx, const uint8_t* scalars32, size_t n, uint8_t* out96) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!scalars32 || !out96) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, scalars32, 32 * n));
    CKR(ensure(ctx, ctx->outb, 96 * n));
    LAUNCH(k_g1_mul_gen_u256, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, n, (uint8_t*)ctx->outb.p);
    CKR(d2h(ctx, out96, ctx->outb.p, 96 * n));
    return sync(ctx);
}
extern "C" int b3_g2_mul(b3_ctx* ctx, const uint8_t* pts192, const uint8_t* scalars32, size_t n, uint8_t* out192) {
    CKR(begin(ctx));
    if (n == 0) return B3_OK;
    if (!pts192 || !scalars32 || !out192) return B3_ERR_ARG;
    CKR(h2d(ctx, ctx->in_a, pts192, 192 * n));
    CKR(h2d(ctx, ctx->in_b, scalars32, 32 * n));
    CKR(ensure(ctx, ctx->outb, 192 * n));
    LAUNCH(k_g2_mul_u256, nblk(n), B3_TPB, (const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, n, (uint8_t*)ctx->outb.p);
    CKR(d2h(ctx, out192, ctx->outb.p, 192 * n));
    return sync(ctx);
}

// ---------------------------------------------------------------------------------------------- roofline probe
extern "C" int b3_imad_peak(b3_ctx* ctx, int wide, double* ops_per_s) {
    CKR(begin(ctx));
    if (!ops_per_s) return B3_ERR_ARG;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    CKR(ensure(ctx, ctx->outb, (size_t)blocks * threads * 4));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    mark_reset(ctx);
        if (wide) LAUNCH(k_imad_wide_peak, blocks, threads, (uint32_t*)ctx->outb.p, iters, 12345u + rep);
        else LAUNCH(k_imad_peak, blocks, threads, (uint32_t*)ctx->outb.p, iters, 12345u + rep);
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CKR(sync(ctx));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        // per thread per iteration: IMAD probe 64 IMADs; wide probe 4*8 = 32 mad.lo/madc.hi PAIRS (= 32 32x32->64 MACs)
        double ops = (double)blocks * threads * (double)iters * (wide ? 32.0 : 64.0);
        double rate = ops / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    *ops_per_s = best;
    return B3_OK;
}
