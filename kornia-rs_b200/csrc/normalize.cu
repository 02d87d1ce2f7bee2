// normalize.cu — normalize_mean_std / normalize_rgb_u8 / find_min_max / normalize_min_max (a10)
// and std_mean (a11).
//
// Reference: normalize.rs:56-87, :123-146, :191-222, :235-263 (+ scalar leaf :407-421, AVX2 leaf with
// fmadd on the npixels&~7 bulk), core.rs:42-67.
//
// B200 design: streaming kernels over flat arrays with 16-byte vector accesses in both directions
// (4 pixels = 12 floats = 3 x float4 keep the channel phase fixed per thread); reductions use
// per-thread integer accumulators, warp shuffles and one atomic per CTA, so `std_mean`'s sums are
// exact integers independent of the order of accumulation (= the reference's f64 folds, which are
// exact below 2^53).
#include <algorithm>

#include "kb200_common.cuh"

namespace kb200 {

struct Ch4 { float v[4]; };

// (x - mean[c]) / std[c] — IEEE division (normalize.rs:76-84)
__global__ void __launch_bounds__(256) normalize_mean_std_c3_vec(const float4* __restrict__ src, float4* __restrict__ dst,
                                                                 size_t nquads, Ch4 mean, Ch4 stdv) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float m0 = mean.v[0], m1 = mean.v[1], m2 = mean.v[2], s0 = stdv.v[0], s1 = stdv.v[1], s2 = stdv.v[2];
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        float4 a = ldg_stream_f4(src + 3 * q), b = ldg_stream_f4(src + 3 * q + 1), c = ldg_stream_f4(src + 3 * q + 2);
        a.x = __fdiv_rn(a.x - m0, s0); a.y = __fdiv_rn(a.y - m1, s1); a.z = __fdiv_rn(a.z - m2, s2); a.w = __fdiv_rn(a.w - m0, s0);
        b.x = __fdiv_rn(b.x - m1, s1); b.y = __fdiv_rn(b.y - m2, s2); b.z = __fdiv_rn(b.z - m0, s0); b.w = __fdiv_rn(b.w - m1, s1);
        c.x = __fdiv_rn(c.x - m2, s2); c.y = __fdiv_rn(c.y - m0, s0); c.z = __fdiv_rn(c.z - m1, s1); c.w = __fdiv_rn(c.w - m2, s2);
        stg_stream_f4(dst + 3 * q, a); stg_stream_f4(dst + 3 * q + 1, b); stg_stream_f4(dst + 3 * q + 2, c);
    }
}

__global__ void normalize_mean_std_generic(const float* __restrict__ src, float* __restrict__ dst, size_t first, size_t n,
                                           uint32_t C, Ch4 mean, Ch4 stdv) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t c = (uint32_t)(i % C);
        dst[i] = __fdiv_rn(__ldg(src + i) - mean.v[c], stdv.v[c]);
    }
}

// u8*scale[c] + offset[c]; FMA for pixels < bulk (AVX2/NEON leaves), mul+add otherwise.
__device__ __forceinline__ float norm_u8(float v, float sc, float of, bool fused) { return fused ? fmaf(v, sc, of) : v * sc + of; }

// 4 px per thread: 3 x 32-bit loads (12 B), 3 x STG.128 out.
__global__ void __launch_bounds__(256) normalize_rgb_u8_vec(const uint32_t* __restrict__ src, float4* __restrict__ dst,
                                                            size_t nquads, Ch4 scale, Ch4 offset, size_t bulk_px) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float s0 = scale.v[0], s1 = scale.v[1], s2 = scale.v[2], o0 = offset.v[0], o1 = offset.v[1], o2 = offset.v[2];
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        const uint32_t w0 = __ldg(src + 3 * q), w1 = __ldg(src + 3 * q + 1), w2 = __ldg(src + 3 * q + 2);
        const bool f = 4 * q < bulk_px;  // bulk is a multiple of 8 px
        float4 a, b, c;
        a.x = norm_u8(byte_to_float(w0, 0), s0, o0, f); a.y = norm_u8(byte_to_float(w0, 1), s1, o1, f);
        a.z = norm_u8(byte_to_float(w0, 2), s2, o2, f); a.w = norm_u8(byte_to_float(w0, 3), s0, o0, f);
        b.x = norm_u8(byte_to_float(w1, 0), s1, o1, f); b.y = norm_u8(byte_to_float(w1, 1), s2, o2, f);
        b.z = norm_u8(byte_to_float(w1, 2), s0, o0, f); b.w = norm_u8(byte_to_float(w1, 3), s1, o1, f);
        c.x = norm_u8(byte_to_float(w2, 0), s2, o2, f); c.y = norm_u8(byte_to_float(w2, 1), s0, o0, f);
        c.z = norm_u8(byte_to_float(w2, 2), s1, o1, f); c.w = norm_u8(byte_to_float(w2, 3), s2, o2, f);
        stg_stream_f4(dst + 3 * q, a); stg_stream_f4(dst + 3 * q + 1, b); stg_stream_f4(dst + 3 * q + 2, c);
    }
}

__global__ void normalize_rgb_u8_generic(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t first_px,
                                         size_t npixels, Ch4 scale, Ch4 offset, size_t bulk_px) {
    const size_t i = first_px + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npixels) return;
    const bool f = i < bulk_px;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[3 * i + c] = norm_u8((float)src[3 * i + c], scale.v[c], offset.v[c], f);
}

// ── min / max ───────────────────────────────────────────────────────────────────────────────
// Order-preserving float <-> uint32 map so atomicMin/atomicMax on integers implement float min/max.
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(u);
}

__global__ void minmax_init_kernel(uint32_t* mm) { mm[0] = 0xFFFFFFFFu; mm[1] = 0u; }

// find_min_max (normalize.rs:123-146): strict `<` / `>` scans from the first element — the result is
// the minimum / maximum under IEEE ordering; NaNs never win a comparison (same here: fminf/fmaxf-free,
// comparisons only).
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ src, size_t n, uint32_t* mm) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float lo = __int_as_float(0x7F800000), hi = __int_as_float(0xFF800000);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(src + i);
        if (v < lo) lo = v;
        if (v > hi) hi = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, o), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, o);
        if (l2 < lo) lo = l2;
        if (h2 > hi) hi = h2;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm, f2ord(lo));
        atomicMax(mm + 1, f2ord(hi));
    }
}

__global__ void minmax_finish_kernel(uint32_t* mm) {
    float* f = reinterpret_cast<float*>(mm);
    const float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
    f[0] = lo; f[1] = hi;
}

// (x - min_val) * (max - min) / (max_val - min_val) + min   (normalize.rs:213-218)
__global__ void __launch_bounds__(256) normalize_min_max_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                size_t n, float mn, float mx,
                                                                const float* __restrict__ minmax) {
    const float min_val = minmax[0], max_val = minmax[1];
    const float range = mx - mn, denom = max_val - min_val;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __fdiv_rn((__ldg(src + i) - min_val) * range, denom) + mn;
}

// ── std_mean ────────────────────────────────────────────────────────────────────────────────
// Flat u8 stream, 12 bytes (4 px) per step so byte j of a step is channel j % 3.  Per-thread u32
// accumulators are flushed into u64 before they can overflow (255² * 4 px * 16384 steps < 2^32).
__global__ void __launch_bounds__(256) std_mean_kernel(const uint8_t* __restrict__ src, size_t npixels,
                                                       unsigned long long* __restrict__ sums) {
    unsigned long long s64[3] = {0, 0, 0}, q64[3] = {0, 0, 0};
    uint32_t s32[3] = {0, 0, 0}, q32[3] = {0, 0, 0};
    const size_t nquads = npixels / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool al4 = (reinterpret_cast<uintptr_t>(src) & 3u) == 0;
    uint32_t pending = 0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        uint32_t w[3];
        if (al4) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src) + 3 * q;
            w[0] = __ldg(p); w[1] = __ldg(p + 1); w[2] = __ldg(p + 2);
        } else {
            const uint8_t* p = src + 12 * q;
#pragma unroll
            for (int k = 0; k < 3; ++k) w[k] = p[4 * k] | (p[4 * k + 1] << 8) | (p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const uint32_t v = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            s32[j % 3] += v;
            q32[j % 3] += v * v;
        }
        if (++pending == 16384u) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { s64[c] += s32[c]; q64[c] += q32[c]; s32[c] = 0; q32[c] = 0; }
            pending = 0;
        }
    }
    // tail pixels (npixels % 4) — first threads of CTA 0
    if (blockIdx.x == 0 && threadIdx.x < (npixels & 3)) {
        const uint8_t* p = src + 3 * (nquads * 4 + threadIdx.x);
#pragma unroll
        for (int c = 0; c < 3; ++c) { const uint32_t v = p[c]; s32[c] += v; q32[c] += v * v; }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { s64[c] += s32[c]; q64[c] += q32[c]; }
    __shared__ unsigned long long red[6][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s64[c] += __shfl_xor_sync(0xFFFFFFFFu, s64[c], o);
            q64[c] += __shfl_xor_sync(0xFFFFFFFFu, q64[c], o);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { red[c][warp] = s64[c]; red[3 + c][warp] = q64[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, t);
    }
}

static inline unsigned stream_grid(size_t items, unsigned block, unsigned ctas_per_sm) {
    const size_t want = (items + block - 1) / block;
    const size_t cap = (size_t)device_info().sm_count * ctas_per_sm;
    return (unsigned)std::max<size_t>(1, std::min(want, cap));
}

}  // namespace kb200

using namespace kb200;

extern "C" {

KB200_API int kb200_normalize_mean_std_f32(kb200_stream_t stream, const float* src, float* dst, size_t npixels,
                                           uint32_t channels, const float* mean, const float* stdv) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("dst", dst));
    KB200_TRY(check_ptr("mean", mean)); KB200_TRY(check_ptr("std", stdv));
    if (channels == 0 || channels > 4) return fail(KB200_ERR_UNSUPPORTED, "normalize_mean_std supports 1..4 channels, got %u", channels);
    if (npixels == 0) return KB200_OK;
    Ch4 m{}, sd{};
    for (uint32_t c = 0; c < channels; ++c) { m.v[c] = mean[c]; sd.v[c] = stdv[c]; }
    cudaStream_t s = as_stream(stream);
    const size_t n = npixels * channels;
    size_t done = 0;
    if (channels == 3 && aligned16(src) && aligned16(dst)) {
        const size_t nquads = npixels / 4;
        if (nquads) {
            normalize_mean_std_c3_vec<<<stream_grid(nquads, 256, 8), 256, 0, s>>>(
                reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), nquads, m, sd);
            KB200_TRY(check_launch("normalize_mean_std_c3_vec"));
        }
        done = nquads * 12;
    }
    if (done < n) {
        normalize_mean_std_generic<<<stream_grid(n - done, 256, 8), 256, 0, s>>>(src, dst, done, n, channels, m, sd);
        KB200_TRY(check_launch("normalize_mean_std_generic"));
    }
    return KB200_OK;
}

KB200_API int kb200_normalize_rgb_u8_f32(kb200_stream_t stream, const uint8_t* src, float* dst, size_t npixels,
                                         const float scale[3], const float offset[3], int leaf) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("dst", dst));
    KB200_TRY(check_ptr("scale", scale)); KB200_TRY(check_ptr("offset", offset));
    if (leaf < 0 || leaf > 2) return fail(KB200_ERR_INVALID_ARGUMENT, "unknown cpu leaf %d", leaf);
    if (npixels == 0) return KB200_OK;
    Ch4 sc{}, of{};
    for (int c = 0; c < 3; ++c) { sc.v[c] = scale[c]; of.v[c] = offset[c]; }
    const size_t bulk = (leaf == KB200_LEAF_SCALAR) ? 0 : (npixels & ~(size_t)7);
    cudaStream_t s = as_stream(stream);
    size_t done = 0;
    if (aligned4(src) && aligned16(dst)) {
        const size_t nquads = npixels / 4;
        if (nquads) {
            normalize_rgb_u8_vec<<<stream_grid(nquads, 256, 8), 256, 0, s>>>(
                reinterpret_cast<const uint32_t*>(src), reinterpret_cast<float4*>(dst), nquads, sc, of, bulk);
            KB200_TRY(check_launch("normalize_rgb_u8_vec"));
        }
        done = nquads * 4;
    }
    if (done < npixels) {
        normalize_rgb_u8_generic<<<div_up(npixels - done, 256), 256, 0, s>>>(src, dst, done, npixels, sc, of, bulk);
        KB200_TRY(check_launch("normalize_rgb_u8_generic"));
    }
    return KB200_OK;
}

KB200_API int kb200_find_min_max_f32(kb200_stream_t stream, const float* src, size_t n, float* minmax_dev) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("minmax_dev", minmax_dev));
    if (n == 0) return fail(KB200_ERR_INVALID_ARGUMENT, "image data is not initialized (empty image)");  // ImageDataNotInitialized
    cudaStream_t s = as_stream(stream);
    uint32_t* mm = reinterpret_cast<uint32_t*>(minmax_dev);
    minmax_init_kernel<<<1, 1, 0, s>>>(mm);
    minmax_kernel<<<stream_grid(n, 256, 8), 256, 0, s>>>(src, n, mm);
    minmax_finish_kernel<<<1, 1, 0, s>>>(mm);
    return check_launch("minmax_kernel");
}

KB200_API int kb200_normalize_min_max_f32(kb200_stream_t stream, const float* src, float* dst, size_t n, float mn,
                                          float mx, const float* minmax_dev) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("dst", dst)); KB200_TRY(check_ptr("minmax_dev", minmax_dev));
    if (n == 0) return fail(KB200_ERR_INVALID_ARGUMENT, "image data is not initialized (empty image)");
    normalize_min_max_kernel<<<stream_grid(n, 256, 8), 256, 0, as_stream(stream)>>>(src, dst, n, mn, mx, minmax_dev);
    return check_launch("normalize_min_max_kernel");
}

KB200_API int kb200_std_mean_u8_c3(kb200_stream_t stream, const uint8_t* src, size_t npixels, uint64_t* sums_dev) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("sums_dev", sums_dev));
    cudaStream_t s = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(sums_dev, 0, 6 * sizeof(uint64_t), s);
    if (e != cudaSuccess) return fail(KB200_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    if (npixels == 0) return KB200_OK;
    const size_t nquads = std::max<size_t>(npixels / 4, 1);
    std_mean_kernel<<<stream_grid(nquads, 256, 4), 256, 0, s>>>(src, npixels, reinterpret_cast<unsigned long long*>(sums_dev));
    return check_launch("std_mean_kernel");
}

}  // extern "C"
