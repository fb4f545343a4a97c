// decaf_gemm: validation + dispatch between the tcgen05 (bf16) and SIMT kernels; error plumbing.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"

namespace decaf {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch width of the persistent tensor-core kernels (gemm_tc.cu, ffn_tc.cu).  0 = every SM of the device.
static int g_gemm_sms = -1;

static int device_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

int num_sms() {
    if (g_gemm_sms < 0) {
        const char *e = getenv("DECAF_GEMM_SMS");
        g_gemm_sms = e ? atoi(e) : 0;
    }
    const int n = device_sms();
    return (g_gemm_sms >= 2 && g_gemm_sms < n) ? (g_gemm_sms & ~1) : n;
}

int gemm_simt_launch(const GemmArgs &a, int dtype, int n_group, cudaStream_t st);
int gemm_tc_launch(const GemmArgs &a, int n_group, cudaStream_t st);
const char *gemm_tc_why_not(const GemmArgs &a, int dtype);

}  // namespace decaf

using namespace decaf;

extern "C" const char *decaf_last_error(void) { return decaf::g_err; }
extern "C" int decaf_version(void) { return 101; }

extern "C" int decaf_set_gemm_sms(int32_t n_sms) {
    const int prev = decaf::g_gemm_sms < 0 ? 0 : decaf::g_gemm_sms;
    decaf::g_gemm_sms = n_sms < 0 ? 0 : n_sms;
    return prev;
}

// Strided host -> device upload (cudaMemcpy2DAsync): a column window [w0, w1) of a pinned (C, t) feature matrix goes straight
// into the device buffer, without a contiguous staging copy on the host (time-sharded ingest of hour-long videos).
extern "C" int decaf_upload_2d(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                               int64_t width_bytes, int64_t height, void *stream) {
    DECAF_CHECK(dst && src, "decaf_upload_2d: null pointers");
    DECAF_CHECK(width_bytes >= 0 && height >= 0 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
                "decaf_upload_2d: bad geometry");
    if (width_bytes == 0 || height == 0) return 0;
    DECAF_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes, (size_t)height,
                                 cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
}

extern "C" int decaf_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

extern "C" int decaf_gemm(const decaf_gemm_t *p, void *stream) {
    DECAF_CHECK(p && p->A && p->W, "decaf_gemm: null operand");
    DECAF_CHECK(p->dtype == DECAF_F32 || p->dtype == DECAF_BF16, "decaf_gemm: bad dtype %d", p->dtype);
    DECAF_CHECK(p->n_seq > 0 && p->rows_per_seq > 0 && p->N > 0 && p->K > 0, "decaf_gemm: empty problem");
    DECAF_CHECK(p->taps == 1 || p->taps == 3, "decaf_gemm: taps must be 1 or 3 (got %d)", p->taps);
    DECAF_CHECK(p->out_f32 || p->out_act, "decaf_gemm: no output");
    GemmArgs a;
    a.A = p->A; a.lda = p->lda;
    a.a_seq_stride = p->a_seq_stride ? p->a_seq_stride : p->rows_per_seq;
    a.n_seq = p->n_seq; a.rows_per_seq = p->rows_per_seq;
    a.W = p->W; a.N = p->N; a.K = p->K; a.taps = p->taps; a.dil = p->dil > 0 ? p->dil : 1;
    a.bias = p->bias; a.act = p->act; a.colscale = p->colscale;
    a.resid = p->resid; a.ldr = p->ldr; a.r_seq_stride = p->r_seq_stride ? p->r_seq_stride : p->rows_per_seq;
    a.rowmask = p->rowmask; a.m_seq_stride = p->m_seq_stride ? p->m_seq_stride : p->rows_per_seq;
    a.out_f32 = p->out_f32; a.ldo = p->ldo; a.o_seq_stride = p->o_seq_stride ? p->o_seq_stride : p->rows_per_seq;
    a.out_act = p->out_act; a.ldo2 = p->ldo2; a.o2_seq_stride = p->o2_seq_stride ? p->o2_seq_stride : p->rows_per_seq;
    a.g_stride_a = p->g_stride_a; a.g_stride_w = p->g_stride_w; a.g_stride_bias = p->g_stride_bias;
    a.g_stride_out_f32 = p->g_stride_out_f32; a.g_stride_out_act = p->g_stride_out_act;
    a.ln = p->ln; a.ln_w = p->ln_w; a.ln_b = p->ln_b; a.ln_eps = p->ln_eps > 0.f ? p->ln_eps : 1e-5f;
    a.pe = p->pe;
    DECAF_CHECK((a.ln_w == nullptr) == (a.ln_b == nullptr), "decaf_gemm: ln_w/ln_b must both be set or both null");
    const int n_group = p->n_group > 0 ? p->n_group : 1;
    DECAF_CHECK(a.lda >= a.K, "decaf_gemm: lda %lld < K %d", (long long)a.lda, a.K);
    DECAF_CHECK(!a.resid || a.ldr >= a.N, "decaf_gemm: ldr < N");
    DECAF_CHECK(!a.out_f32 || a.ldo >= a.N, "decaf_gemm: ldo < N");
    DECAF_CHECK(!a.out_act || a.ldo2 >= a.N, "decaf_gemm: ldo2 < N");

    cudaStream_t st = as_stream(stream);
    int impl = p->impl;
    if (impl == 0) impl = (gemm_tc_why_not(a, p->dtype) == nullptr) ? 2 : 1;
    if (impl == 2) {
        const char *why = gemm_tc_why_not(a, p->dtype);
        DECAF_CHECK(why == nullptr, "decaf_gemm: tcgen05 path not applicable: %s", why);
        return gemm_tc_launch(a, n_group, st);
    }
    DECAF_CHECK(!a.ln, "decaf_gemm: the fused LayerNorm epilogue exists on the tcgen05 path only (%s)",
                gemm_tc_why_not(a, p->dtype) ? gemm_tc_why_not(a, p->dtype) : "impl=1 requested");
    return gemm_simt_launch(a, p->dtype, n_group, st);
}
