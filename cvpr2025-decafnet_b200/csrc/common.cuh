// Shared device/host helpers for libdecaf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "decaf_b200.h"

namespace decaf {

// ----------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);

#define DECAF_CHECK(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            decaf::set_error(__VA_ARGS__);                      \
            return 1;                                           \
        }                                                       \
    } while (0)

#define DECAF_CUDA(expr)                                                               \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            decaf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,              \
                             cudaGetErrorString(_e));                                  \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

#define DECAF_LAUNCH_CHECK() DECAF_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------- dtypes
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ----------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-lane slice of a row of C = 32 * VEC channels: lane owns channels [lane*VEC, lane*VEC+VEC).
// Vector width chosen so a warp touches one contiguous span per instruction.
template <int VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ p, int lane, float (&v)[VEC]) {
    const float *q = p + lane * VEC;
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
            float4 t = *reinterpret_cast<const float4 *>(q + i);
            v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 2) {
            float2 t = *reinterpret_cast<const float2 *>(q + i);
            v[i] = t.x; v[i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC; i++) v[i] = q[i];
    }
}
template <int VEC>
__device__ __forceinline__ void load_row(const bf16 *__restrict__ p, int lane, float (&v)[VEC]) {
    const bf16 *q = p + lane * VEC;
    if constexpr (VEC % 8 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 8) {
            uint4 t = *reinterpret_cast<const uint4 *>(q + i);
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float2 f = __bfloat1622float2(h[j]);
                v[i + 2 * j] = f.x; v[i + 2 * j + 1] = f.y;
            }
        }
    } else if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
            uint2 t = *reinterpret_cast<const uint2 *>(q + i);
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
            float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            v[i] = f0.x; v[i + 1] = f0.y; v[i + 2] = f1.x; v[i + 3] = f1.y;
        }
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 2) {
            float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(q + i));
            v[i] = f.x; v[i + 1] = f.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC; i++) v[i] = __bfloat162float(q[i]);
    }
}
template <int VEC>
__device__ __forceinline__ void store_row(float *__restrict__ p, int lane, const float (&v)[VEC]) {
    float *q = p + lane * VEC;
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
            *reinterpret_cast<float4 *>(q + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 2) *reinterpret_cast<float2 *>(q + i) = make_float2(v[i], v[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; i++) q[i] = v[i];
    }
}
template <int VEC>
__device__ __forceinline__ void store_row(bf16 *__restrict__ p, int lane, const float (&v)[VEC]) {
    bf16 *q = p + lane * VEC;
    if constexpr (VEC % 8 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 8) {
            uint4 t;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
            for (int j = 0; j < 4; j++) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
            *reinterpret_cast<uint4 *>(q + i) = t;
        }
    } else if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
            uint2 t;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
            h[0] = __floats2bfloat162_rn(v[i], v[i + 1]);
            h[1] = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
            *reinterpret_cast<uint2 *>(q + i) = t;
        }
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 2)
            *reinterpret_cast<__nv_bfloat162 *>(q + i) = __floats2bfloat162_rn(v[i], v[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; i++) q[i] = __float2bfloat16_rn(v[i]);
    }
}

// Interleaved lane <-> channel mapping for kernels that also keep rows in shared memory: when VEC % 4 == 0, element i of
// the per-lane array is channel (i / 4) * 128 + lane * 4 + (i % 4), i.e. every 16-byte access of a warp covers 512
// contiguous bytes (conflict-free LDS.128 / STS.128, fully coalesced global accesses).  With the contiguous mapping of
// load_row (lane owns channels [lane * VEC, lane * VEC + VEC)) lanes i and i + 4 hit the same banks on every 16-byte
// shared-memory access (2-way conflicts: 42 % excess wavefronts in the pre-attention kernel).  Other VEC: contiguous.
template <int VEC>
__device__ __forceinline__ void iload_row(const float *__restrict__ p, int lane, float (&v)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(p + i * 32 + lane * 4);
            v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
    } else {
        load_row<VEC>(p, lane, v);
    }
}
template <int VEC>
__device__ __forceinline__ void iload_row(const bf16 *__restrict__ p, int lane, float (&v)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
            const uint2 t = *reinterpret_cast<const uint2 *>(p + i * 32 + lane * 4);
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
            const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            v[i] = f0.x; v[i + 1] = f0.y; v[i + 2] = f1.x; v[i + 3] = f1.y;
        }
    } else {
        load_row<VEC>(p, lane, v);
    }
}
template <int VEC>
__device__ __forceinline__ void istore_row(float *__restrict__ p, int lane, const float (&v)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
            *reinterpret_cast<float4 *>(p + i * 32 + lane * 4) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
        store_row<VEC>(p, lane, v);
    }
}
template <int VEC>
__device__ __forceinline__ void istore_row(bf16 *__restrict__ p, int lane, const float (&v)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
            uint2 t;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
            h[0] = __floats2bfloat162_rn(v[i], v[i + 1]);
            h[1] = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
            *reinterpret_cast<uint2 *>(p + i * 32 + lane * 4) = t;
        }
    } else {
        store_row<VEC>(p, lane, v);
    }
}

// Packed fp32 element-wise helpers (sm_100 FADD2 / FMUL2 / FFMA2: two channels per instruction) for the per-lane
// channel arrays of the row-wise kernels; odd VEC falls back to scalar code for the last element.
template <int VEC>
__device__ __forceinline__ void pk_mul(float (&d)[VEC], const float (&a)[VEC], const float (&b)[VEC]) {       // d = a * b
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) {
        const float2 r = __fmul2_rn(make_float2(a[i], a[i + 1]), make_float2(b[i], b[i + 1]));
        d[i] = r.x; d[i + 1] = r.y;
    }
    if (VEC & 1) d[VEC - 1] = a[VEC - 1] * b[VEC - 1];
}
template <int VEC>
__device__ __forceinline__ void pk_fma(float (&d)[VEC], const float (&a)[VEC], const float (&b)[VEC]) {       // d += a * b
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) {
        const float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(b[i], b[i + 1]), make_float2(d[i], d[i + 1]));
        d[i] = r.x; d[i + 1] = r.y;
    }
    if (VEC & 1) d[VEC - 1] = fmaf(a[VEC - 1], b[VEC - 1], d[VEC - 1]);
}
template <int VEC>
__device__ __forceinline__ void pk_affine(float (&d)[VEC], const float (&w)[VEC], const float (&b)[VEC]) {    // d = d * w + b
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) {
        const float2 r = __ffma2_rn(make_float2(d[i], d[i + 1]), make_float2(w[i], w[i + 1]), make_float2(b[i], b[i + 1]));
        d[i] = r.x; d[i + 1] = r.y;
    }
    if (VEC & 1) d[VEC - 1] = fmaf(d[VEC - 1], w[VEC - 1], b[VEC - 1]);
}
template <int VEC>
__device__ __forceinline__ float pk_sum(const float (&v)[VEC]) {
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) s2 = __fadd2_rn(s2, make_float2(v[i], v[i + 1]));
    float s = s2.x + s2.y;
    if (VEC & 1) s += v[VEC - 1];
    return s;
}
// v -= mean; returns sum of the centred squares
template <int VEC>
__device__ __forceinline__ float pk_center_sq(float (&v)[VEC], float mean) {
    float2 q2 = make_float2(0.f, 0.f);
    const float2 nm = make_float2(-mean, -mean);
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) {
        const float2 d = __fadd2_rn(make_float2(v[i], v[i + 1]), nm);
        q2 = __ffma2_rn(d, d, q2);
        v[i] = d.x; v[i + 1] = d.y;
    }
    float q = q2.x + q2.y;
    if (VEC & 1) { v[VEC - 1] -= mean; q = fmaf(v[VEC - 1], v[VEC - 1], q); }
    return q;
}
template <int VEC>
__device__ __forceinline__ void pk_scale(float (&v)[VEC], float r) {
    const float2 r2 = make_float2(r, r);
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) {
        const float2 d = __fmul2_rn(make_float2(v[i], v[i + 1]), r2);
        v[i] = d.x; v[i + 1] = d.y;
    }
    if (VEC & 1) v[VEC - 1] *= r;
}

// Two-pass channel LayerNorm statistics over a warp-distributed row (libs/modeling/blocks.py:
// 125-131: mean, then mean of centred squares, eps inside sqrt).  Leaves v centred and scaled.
template <int VEC>
__device__ __forceinline__ void warp_layernorm(float (&v)[VEC], int C, float eps) {
    const float mean = warp_sum(pk_sum<VEC>(v)) / (float)C;
    const float var = warp_sum(pk_center_sq<VEC>(v, mean)) / (float)C;
    const float r = rsqrtf(var + eps);                  // MUFU.RSQ, <= 2 ulp (vs ~40 instructions for sqrt + divide)
    pk_scale<VEC>(v, r);
}

// Absolute PE of word `pos` (0-based) of a query of `len` words, channel c, from the raw sinusoid table pe (pe_rows, C):
// the reference encodes every query alone (libs/worker_v2.py:945-955) and interpolates the table linearly
// (align_corners) to THAT query's length only when it exceeds max_seq_len (libs/modeling/text_net.py:163-172);
// otherwise the raw rows are used.  Same fp32 arithmetic as F.interpolate: src = pos * (rows - 1) / (len - 1).
__device__ __forceinline__ float text_pe_value(const float *__restrict__ pe, int pe_rows, int C, int len, int pos, int c) {
    if (len <= pe_rows) return pe[(int64_t)pos * C + c];
    const float scale = (float)(pe_rows - 1) / (float)(len - 1);
    const float src = scale * (float)pos;
    int i0 = (int)src;
    if (i0 > pe_rows - 1) i0 = pe_rows - 1;
    const int i1 = i0 + (i0 < pe_rows - 1 ? 1 : 0);
    const float l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
    return l0 * pe[(int64_t)i0 * C + c] + l1 * pe[(int64_t)i1 * C + c];
}

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// dispatch on channels-per-lane (C = 32 * VEC)
#define DECAF_DISPATCH_VEC(C, ...)                                                        \
    do {                                                                                  \
        switch ((C) / 32) {                                                               \
            case 1:  { constexpr int VEC = 1;  __VA_ARGS__; } break;                      \
            case 2:  { constexpr int VEC = 2;  __VA_ARGS__; } break;                      \
            case 3:  { constexpr int VEC = 3;  __VA_ARGS__; } break;                      \
            case 4:  { constexpr int VEC = 4;  __VA_ARGS__; } break;                      \
            case 5:  { constexpr int VEC = 5;  __VA_ARGS__; } break;                      \
            case 6:  { constexpr int VEC = 6;  __VA_ARGS__; } break;                      \
            case 8:  { constexpr int VEC = 8;  __VA_ARGS__; } break;                      \
            case 9:  { constexpr int VEC = 9;  __VA_ARGS__; } break;                      \
            case 12: { constexpr int VEC = 12; __VA_ARGS__; } break;                      \
            case 16: { constexpr int VEC = 16; __VA_ARGS__; } break;                      \
            case 17: { constexpr int VEC = 17; __VA_ARGS__; } break;                      \
            default:                                                                      \
                decaf::set_error("unsupported channel count %d (need 32*{1..6,8,9,12,16,17})", (int)(C)); \
                return 1;                                                                 \
        }                                                                                 \
    } while (0)

// attention kernels: C in {32,64,96,128,256,512}
#define DECAF_DISPATCH_VEC_ATTN(C, ...)                                                   \
    do {                                                                                  \
        switch ((C) / 32) {                                                               \
            case 1:  { constexpr int VEC = 1;  __VA_ARGS__; } break;                      \
            case 2:  { constexpr int VEC = 2;  __VA_ARGS__; } break;                      \
            case 3:  { constexpr int VEC = 3;  __VA_ARGS__; } break;                      \
            case 4:  { constexpr int VEC = 4;  __VA_ARGS__; } break;                      \
            case 8:  { constexpr int VEC = 8;  __VA_ARGS__; } break;                      \
            case 16: { constexpr int VEC = 16; __VA_ARGS__; } break;                      \
            default:                                                                      \
                decaf::set_error("attention: unsupported channel count %d", (int)(C));    \
                return 1;                                                                 \
        }                                                                                 \
    } while (0)

}  // namespace decaf
