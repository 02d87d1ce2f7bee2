// resize_fused.cu — fused u8 HWC → f32 CHW bilinear resize + normalize (a2; BASELINE config 2).
//
// Reference: resize/fused.rs:147-228 (general bilinear, half-pixel, non-antialiased), scalar leaf
// :273-318, AVX2+FMA leaf :414-497 (FMA form on the dst_w&~7 bulk, scalar form on the tail),
// exact-2x box path :57-127 with leaves :528-559 (scalar) / fused_row_avx2 (FMA on dst_w&~15).
//
// `fma_bulk` = number of leading destination columns whose arithmetic is the reference's FMA leaf;
// columns ≥ fma_bulk use the scalar (mul, add) leaf — so the output is bit-identical to what the
// reference produces on the chosen CPU (x86 AVX2+FMA: bulk = dst_w & ~7, or & ~15 on the 2x path).
#include "kb200_common.cuh"

namespace kb200 {

struct FusedParams {
    uint32_t sw, sh, dw, dh;
    float scale_x, scale_y;
    float scale[3], bias[3];
    uint32_t fma_bulk;
};

__device__ __forceinline__ float fused_lerp(float a, float b, float c, float d, float wx, float wy, float sc, float bi,
                                            bool fused) {
    if (fused) {  // resize/fused.rs:475-478
        const float top = fmaf(b - a, wx, a);
        const float bot = fmaf(d - c, wx, c);
        const float val = fmaf(bot - top, wy, top);
        return fmaf(val, sc, bi);
    }
    const float top = a + wx * (b - a);  // resize/fused.rs:286-317
    const float bot = c + wx * (d - c);
    const float val = top + wy * (bot - top);
    return val * sc + bi;
}

// Generic path: any size / alignment.  One thread per destination pixel, batch = grid.z.
template <bool BOX2X>
__global__ void __launch_bounds__(256) fused_resize_gather_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                                  const __grid_constant__ FusedParams p) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= p.dw || y >= p.dh) return;
    const uint8_t* s = src + (size_t)blockIdx.z * p.sw * p.sh * 3;
    const size_t plane = (size_t)p.dw * p.dh;
    float* d = dst + (size_t)blockIdx.z * plane * 3 + (size_t)y * p.dw + x;
    const bool fused = x < p.fma_bulk;
    if (BOX2X) {  // resize/fused.rs:528-559
        const uint8_t* r0 = s + ((size_t)(2 * y) * p.sw + 2 * x) * 3;
        const uint8_t* r1 = r0 + (size_t)p.sw * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t sum = (uint32_t)r0[c] + r0[3 + c] + r1[c] + r1[3 + c];
            const float s4 = p.scale[c] * 0.25f;
            d[c * plane] = fused ? fmaf((float)sum, s4, p.bias[c]) : (float)sum * s4 + p.bias[c];
        }
        return;
    }
    const float fx = fmaxf(((float)x + 0.5f) * p.scale_x - 0.5f, 0.0f);
    const float fy = fmaxf(((float)y + 0.5f) * p.scale_y - 0.5f, 0.0f);
    const uint32_t x0 = min((uint32_t)fx, p.sw - 1u), y0 = min((uint32_t)fy, p.sh - 1u);
    const uint32_t x1 = min(x0 + 1u, p.sw - 1u), y1 = min(y0 + 1u, p.sh - 1u);
    const float wx = fx - (float)x0, wy = fy - (float)y0;
    const uint8_t* row0 = s + (size_t)y0 * p.sw * 3;
    const uint8_t* row1 = s + (size_t)y1 * p.sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = (float)row0[x0 * 3 + c], b = (float)row0[x1 * 3 + c];
        const float cc = (float)row1[x0 * 3 + c], dd = (float)row1[x1 * 3 + c];
        d[c * plane] = fused_lerp(a, b, cc, dd, wx, wy, p.scale[c], p.bias[c], fused);
    }
}

int launch_fused_resize_staged(cudaStream_t s, const uint8_t* src, float* dst, const FusedParams& p, uint32_t batch,
                               bool* handled);

}  // namespace kb200

using namespace kb200;

extern "C" {

KB200_API int kb200_resize_normalize_chw_u8_f32(kb200_stream_t stream, const uint8_t* src, size_t src_len,
                                                float* dst, size_t dst_len, uint32_t sw, uint32_t sh, uint32_t dw,
                                                uint32_t dh, uint32_t batch, const float scale[3],
                                                const float bias[3], int leaf) {
    KB200_TRY(check_ptr("src", src)); KB200_TRY(check_ptr("dst", dst));
    KB200_TRY(check_ptr("scale", scale)); KB200_TRY(check_ptr("bias", bias));
    if (leaf < 0 || leaf > 2) return fail(KB200_ERR_INVALID_ARGUMENT, "unknown cpu leaf %d", leaf);
    if (batch == 0) return fail(KB200_ERR_INVALID_ARGUMENT, "batch must be non-zero");
    if (batch > 65535u) return fail(KB200_ERR_INVALID_ARGUMENT, "batch %u exceeds 65535 per call", batch);
    // InvalidChannelShape checks of resize/fused.rs:156-167 (lengths must cover the images)
    KB200_TRY(check_slice("src", src_len, (size_t)sw * sh * 3 * batch));
    KB200_TRY(check_slice("dst", dst_len, (size_t)dw * dh * 3 * batch));
    if (dw == 0 || dh == 0 || sw == 0 || sh == 0) return KB200_OK;  // resize/fused.rs:184-186: empty is a no-op
    FusedParams p;
    p.sw = sw; p.sh = sh; p.dw = dw; p.dh = dh;
    p.scale_x = (float)sw / (float)dw;
    p.scale_y = (float)sh / (float)dh;
    for (int c = 0; c < 3; ++c) { p.scale[c] = scale[c]; p.bias[c] = bias[c]; }
    const bool box2x = (sw == 2 * dw && sh == 2 * dh);
    if (leaf == KB200_LEAF_SCALAR) p.fma_bulk = 0;
    else if (box2x) p.fma_bulk = dw & ~15u;                          // fused_row_avx2 / fused_row_neon: 16 px per iter
    else p.fma_bulk = (leaf == KB200_LEAF_X86_AVX2_FMA) ? (dw & ~7u) : (dw & ~3u);  // :448 / :355
    cudaStream_t s = as_stream(stream);
    if (!box2x) {
        bool handled = false;
        KB200_TRY(launch_fused_resize_staged(s, src, dst, p, batch, &handled));
        if (handled) return KB200_OK;
    }
    dim3 block(32, 8), grid(div_up(dw, 32), div_up(dh, 8), batch);
    if (box2x) fused_resize_gather_kernel<true><<<grid, block, 0, s>>>(src, dst, p);
    else fused_resize_gather_kernel<false><<<grid, block, 0, s>>>(src, dst, p);
    return check_launch("fused_resize_gather_kernel");
}

}  // extern "C"

namespace kb200 {
// Placeholder until the row-span staged kernel lands (round-1 step 2): never handles.
int launch_fused_resize_staged(cudaStream_t, const uint8_t*, float*, const FusedParams&, uint32_t, bool* handled) {
    *handled = false;
    return KB200_OK;
}
}  // namespace kb200
