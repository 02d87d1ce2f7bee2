// filter.cu — separable f32 filters: separable_filter, gaussian_blur, sobel (a6, a7; config 4).
//
// Reference: filter/separable_filter.rs:87-155 (CPU engine: H pass into a full f32 `temp` image,
// then V pass; correlation; taps ascending; `acc += v * k` unfused; out-of-bounds taps SKIPPED =
// constant-zero border, no renormalisation), filter/ops.rs:116-203 (gaussian_blur, sobel),
// cuda/filter.rs:55-110, :515-565 and filter/cuda.rs:106-232 (GPU twin: 2 launches through a DRAM
// scratch image for a blur, 5 launches / 11 image traversals for sobel).
//
// B200 design: ONE kernel per op.  A CTA stages a (TH + ky-1) x (TW + kx-1) pixel tile (zero-filled
// outside the image) in shared memory, runs the H pass into a second shared tile and the V pass
// straight to global memory — the f32 intermediate never touches HBM (traffic: 1 read + 1 write of
// the image instead of 2+2; sobel: 1+1 instead of 11).  Sobel runs both gradient filters off the
// same staged tile and fuses the magnitude.
//
// Bit-exactness with the two-pass reference: each intermediate value is the same ascending,
// unfused (-fmad=false) accumulation starting from +0.0.  A zero-filled halo element contributes
// `acc += 0*k` = acc (acc can never be -0.0: it starts at +0.0 and x + (-x) rounds to +0.0), which
// equals "tap skipped" for every finite tap.  (Non-finite taps × zero halo would differ; taps come
// from finite Gaussian/Sobel tables.)
#include <algorithm>

#include "kb200_common.cuh"

namespace kb200 {

static constexpr int KB200_MAX_TAPS = 31;

struct SepTaps {
    float kx[32];
    float ky[32];
    int kxn, kyn;
};

struct SepGeom {
    uint32_t cols, rows, C;
    uint32_t tw, th;          // tile size in pixels
    uint32_t tiles_x, tiles_y;
};

// KX/KY > 0: compile-time tap counts (fully unrolled); 0: runtime loops.
// SOBEL: kx = derivative taps, ky = smoothing taps (same length); gx = H(kx)·V(ky), gy = H(ky)·V(kx),
// out = sqrt(gx*gx + gy*gy)   (filter/ops.rs:187-200)
template <int KX, int KY, bool SOBEL>
__global__ void __launch_bounds__(256) sep_filter_fused_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                               const __grid_constant__ SepTaps taps,
                                                               const __grid_constant__ SepGeom g) {
    extern __shared__ __align__(16) float smem[];
    const int kxn = KX > 0 ? KX : taps.kxn, kyn = KY > 0 ? KY : taps.kyn;
    const int hx = kxn / 2, hy = kyn / 2;  // offsets_x = i - half  (separable_filter.rs:60-69)
    const int C = (int)g.C;
    const int in_w = ((int)g.tw + kxn - 1) * C;  // floats per staged row
    const int mid_w = (int)g.tw * C;
    const int in_h = (int)g.th + kyn - 1;
    float* s_in = smem;
    float* s_mid = smem + (size_t)in_w * in_h;             // [in_h][mid_w]
    float* s_mid2 = SOBEL ? s_mid + (size_t)mid_w * in_h : nullptr;

    const uint32_t tile = blockIdx.x;
    const uint32_t tx = tile % g.tiles_x, ty = (tile / g.tiles_x) % g.tiles_y, img = tile / (g.tiles_x * g.tiles_y);
    const int x0 = (int)(tx * g.tw), y0 = (int)(ty * g.th);
    const size_t img_off = (size_t)img * g.cols * g.rows * C;
    const float* s = src + img_off;
    const int row_floats = (int)g.cols * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    // phase 1: stage the input tile, zero outside the image
    const int gx_base = (x0 - hx) * C;
    for (int r = warp; r < in_h; r += nwarps) {
        const int gy = y0 - hy + r;
        const bool row_ok = gy >= 0 && gy < (int)g.rows;
        const float* grow = s + (size_t)(row_ok ? gy : 0) * row_floats;
        float* srow = s_in + (size_t)r * in_w;
        for (int j = lane; j < in_w; j += 32) {
            const int gxf = gx_base + j;
            srow[j] = (row_ok && gxf >= 0 && gxf < row_floats) ? __ldg(grow + gxf) : 0.0f;
        }
    }
    __syncthreads();

    // phase 2: horizontal pass  temp[r][e] = Σ_t in[r][e + t*C] * kx[t]
    for (int r = warp; r < in_h; r += nwarps) {
        const float* srow = s_in + (size_t)r * in_w;
        for (int e = lane; e < mid_w; e += 32) {
            float acc = 0.0f, acc2 = 0.0f;
            if (KX > 0) {
#pragma unroll
                for (int t = 0; t < (KX > 0 ? KX : 1); ++t) {
                    const float v = srow[e + t * C];
                    acc += v * taps.kx[t];
                    if (SOBEL) acc2 += v * taps.ky[t];
                }
            } else {
                for (int t = 0; t < kxn; ++t) {
                    const float v = srow[e + t * C];
                    acc += v * taps.kx[t];
                    if (SOBEL) acc2 += v * taps.ky[t];
                }
            }
            s_mid[(size_t)r * mid_w + e] = acc;
            if (SOBEL) s_mid2[(size_t)r * mid_w + e] = acc2;
        }
    }
    __syncthreads();

    // phase 3: vertical pass straight to global memory
    const int out_w = min((int)g.tw, (int)g.cols - x0) * C;
    float* d = dst + img_off;
    for (int r = warp; r < (int)g.th; r += nwarps) {
        const int gy = y0 + r;
        if (gy >= (int)g.rows) break;
        float* drow = d + (size_t)gy * row_floats + (size_t)x0 * C;
        for (int e = lane; e < out_w; e += 32) {
            float acc = 0.0f, acc2 = 0.0f;
            if (KY > 0) {
#pragma unroll
                for (int t = 0; t < (KY > 0 ? KY : 1); ++t) {
                    acc += s_mid[(size_t)(r + t) * mid_w + e] * taps.ky[t];
                    if (SOBEL) acc2 += s_mid2[(size_t)(r + t) * mid_w + e] * taps.kx[t];
                }
            } else {
                for (int t = 0; t < kyn; ++t) {
                    acc += s_mid[(size_t)(r + t) * mid_w + e] * taps.ky[t];
                    if (SOBEL) acc2 += s_mid2[(size_t)(r + t) * mid_w + e] * taps.kx[t];
                }
            }
            drow[e] = SOBEL ? sqrtf(acc * acc + acc2 * acc2) : acc;
        }
    }
}

// cuda/filter.rs:515-530
__global__ void gradient_magnitude_kernel(const float* __restrict__ gx, const float* __restrict__ gy,
                                          float* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float a = __ldg(gx + i), b = __ldg(gy + i);
        dst[i] = sqrtf(a * a + b * b);
    }
}

template <int KX, int KY, bool SOBEL>
static int launch_sep_instance(cudaStream_t s, const float* src, float* dst, const SepTaps& taps, const SepGeom& g,
                               uint32_t batch, size_t smem_bytes) {
    auto kern = sep_filter_fused_kernel<KX, KY, SOBEL>;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return fail(KB200_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu) failed: %s", smem_bytes, cudaGetErrorString(e));
    }
    const size_t ntiles = (size_t)g.tiles_x * g.tiles_y * batch;
    if (ntiles > 0x7FFFFFFFull) return fail(KB200_ERR_DIMS_TOO_LARGE, "too many tiles (%zu)", ntiles);
    kern<<<(unsigned)ntiles, 256, smem_bytes, s>>>(src, dst, taps, g);
    return check_launch("sep_filter_fused_kernel");
}

static int launch_sep(cudaStream_t s, const float* src, size_t src_len, float* dst, size_t dst_len, const float* kx,
                      uint32_t kxn, const float* ky, uint32_t kyn, uint32_t cols, uint32_t rows, uint32_t C,
                      uint32_t batch, bool sobel) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("dst", dst));
    KB200_TRY(check_geometry(cols, rows, cols, rows, batch));
    if (C == 0 || kxn == 0 || kyn == 0) return fail(KB200_ERR_INVALID_KERNEL, "channels and tap counts must be at least 1");  // cuda/filter.rs:306-310
    if (C > 4) return fail(KB200_ERR_UNSUPPORTED, "separable filter supports 1..4 channels, got %u", C);
    if (kxn > (uint32_t)KB200_MAX_TAPS || kyn > (uint32_t)KB200_MAX_TAPS)
        return fail(KB200_ERR_UNSUPPORTED, "separable filter supports up to %d taps per axis, got (%u, %u)", KB200_MAX_TAPS, kxn, kyn);
    const size_t n = (size_t)cols * rows * C * batch;
    KB200_TRY(check_slice("src", src_len, n)); KB200_TRY(check_slice("dst", dst_len, n));
    if (src == dst) return fail(KB200_ERR_INVALID_ARGUMENT, "src and dst must not alias (tiles read a halo)");
    SepTaps taps{};
    for (uint32_t i = 0; i < kxn; ++i) taps.kx[i] = kx[i];
    for (uint32_t i = 0; i < kyn; ++i) taps.ky[i] = ky[i];
    taps.kxn = (int)kxn; taps.kyn = (int)kyn;
    SepGeom g;
    g.cols = cols; g.rows = rows; g.C = C;
    g.tw = 64; g.th = 32;
    if (cols <= 32) g.tw = 32;
    if (rows <= 16) g.th = 16;
    auto smem_for = [&](uint32_t tw, uint32_t th) {
        const size_t in_w = (size_t)(tw + kxn - 1) * C, in_h = th + kyn - 1, mid_w = (size_t)tw * C;
        return (in_w * in_h + mid_w * in_h * (sobel ? 2 : 1)) * sizeof(float);
    };
    size_t smem = smem_for(g.tw, g.th);
    const size_t cap = std::min<size_t>((size_t)device_info().max_smem_optin, 100 * 1024);
    while (smem > cap && (g.tw > 16 || g.th > 8)) {
        if (g.th > 8 && g.th >= g.tw / 2) g.th /= 2; else g.tw /= 2;
        smem = smem_for(g.tw, g.th);
    }
    if (smem > cap) return fail(KB200_ERR_UNSUPPORTED, "filter tile does not fit shared memory (%zu B)", smem);
    g.tiles_x = (cols + g.tw - 1) / g.tw;
    g.tiles_y = (rows + g.th - 1) / g.th;
    if (sobel) {
        if (kxn == 3) return launch_sep_instance<3, 3, true>(s, src, dst, taps, g, batch, smem);
        return launch_sep_instance<5, 5, true>(s, src, dst, taps, g, batch, smem);
    }
    if (kxn == 3 && kyn == 3) return launch_sep_instance<3, 3, false>(s, src, dst, taps, g, batch, smem);
    if (kxn == 5 && kyn == 5) return launch_sep_instance<5, 5, false>(s, src, dst, taps, g, batch, smem);
    if (kxn == 7 && kyn == 7) return launch_sep_instance<7, 7, false>(s, src, dst, taps, g, batch, smem);
    return launch_sep_instance<0, 0, false>(s, src, dst, taps, g, batch, smem);
}

}  // namespace kb200

using namespace kb200;

extern "C" {

KB200_API int kb200_separable_filter_f32(kb200_stream_t stream, const float* src, size_t src_len, float* dst,
                                         size_t dst_len, float* /*scratch*/, const float* kx, uint32_t kx_len,
                                         const float* ky, uint32_t ky_len, uint32_t cols, uint32_t rows,
                                         uint32_t channels, uint32_t batch) {
    KB200_TRY(check_ptr("kx", kx)); KB200_TRY(check_ptr("ky", ky));
    return launch_sep(as_stream(stream), src, src_len, dst, dst_len, kx, kx_len, ky, ky_len, cols, rows, channels, batch, false);
}

KB200_API int kb200_gaussian_blur_f32(kb200_stream_t stream, const float* src, size_t src_len, float* dst,
                                      size_t dst_len, uint32_t cols, uint32_t rows, uint32_t channels,
                                      uint32_t batch, uint32_t ksize_x, uint32_t ksize_y, float sigma_x,
                                      float sigma_y) {
    uint32_t kxn, kyn;
    float sx, sy;
    KB200_TRY(kb200_gaussian_resolve(ksize_x, ksize_y, sigma_x, sigma_y, &kxn, &kyn, &sx, &sy));
    if (kxn > (uint32_t)KB200_MAX_TAPS || kyn > (uint32_t)KB200_MAX_TAPS)
        return fail(KB200_ERR_UNSUPPORTED, "gaussian_blur supports up to %d taps per axis, got (%u, %u)", KB200_MAX_TAPS, kxn, kyn);
    float kx[32], ky[32];
    kb200_gaussian_kernel_1d(kxn, sx, kx);
    kb200_gaussian_kernel_1d(kyn, sy, ky);
    return launch_sep(as_stream(stream), src, src_len, dst, dst_len, kx, kxn, ky, kyn, cols, rows, channels, batch, false);
}

KB200_API int kb200_sobel_f32(kb200_stream_t stream, const float* src, size_t src_len, float* dst, size_t dst_len,
                              uint32_t cols, uint32_t rows, uint32_t channels, uint32_t batch, uint32_t ksize) {
    // filter/kernels.rs:55-72
    static const float d3[3] = {-1.0f, 0.0f, 1.0f}, s3[3] = {1.0f, 2.0f, 1.0f};
    static const float d5[5] = {-1.0f, -2.0f, 0.0f, 2.0f, 1.0f}, s5[5] = {1.0f, 4.0f, 6.0f, 4.0f, 1.0f};
    if (ksize != 3 && ksize != 5) return fail(KB200_ERR_INVALID_KERNEL, "invalid sobel kernel length %u (expected 3 or 5)", ksize);
    return launch_sep(as_stream(stream), src, src_len, dst, dst_len, ksize == 3 ? d3 : d5, ksize, ksize == 3 ? s3 : s5,
                      ksize, cols, rows, channels, batch, true);
}

KB200_API int kb200_gradient_magnitude_f32(kb200_stream_t stream, const float* gx, const float* gy, float* dst,
                                           size_t n) {
    KB200_TRY(check_ptr("gx", gx)); KB200_TRY(check_ptr("gy", gy)); KB200_TRY(check_ptr("dst", dst));
    if (n == 0) return KB200_OK;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)device_info().sm_count * 16);
    gradient_magnitude_kernel<<<grid, 256, 0, as_stream(stream)>>>(gx, gy, dst, n);
    return check_launch("gradient_magnitude_kernel");
}

}  // extern "C"
