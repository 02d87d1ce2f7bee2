// u8_sampler.cuh — the Q10 bilinear blend of the u8 warps / remap (warp/common.rs:14-181) with WORD-granular taps.
//
// The byte version issues 12 `LDG.U8` per RGB pixel (4 taps x 3 channels); ncu on the round-1 kernels showed them bound
// by LSU instructions, far from DRAM (0.12-0.17 of the roofline).  The two taps of a row are six consecutive bytes, so
// three aligned 32-bit words per row cover them: 6 `LDG.32` + two funnel shifts per row instead of 12 byte loads.  The
// arithmetic is unchanged — `(top*fy1 + bot*fy + 2^19) >> 20` with `top = p0*fx1 + p1*fx` — and bit-exact.
#pragma once

#include <stdint.h>

namespace kb200 {

// `img` is 4-byte aligned, `img_bytes` = sw*sh*3.  Taps (xi, yi), (xi1, yi), (xi, yi1), (xi1, yi1) with xi1 in {xi, xi+1}.
// Returns false (nothing written) when the 12-byte window of a row would run past the image — the caller's byte path
// handles those few pixels at the very end of the image.
__device__ __forceinline__ bool q10_blend_c3_words(const uint8_t* __restrict__ img, uint32_t img_bytes, int sw, int xi, int yi, int xi1, int yi1,
                                                   uint32_t fx, uint32_t fy, uint8_t* __restrict__ d) {
    const uint32_t o0 = ((uint32_t)yi * (uint32_t)sw + (uint32_t)xi) * 3u, o1 = ((uint32_t)yi1 * (uint32_t)sw + (uint32_t)xi) * 3u;
    const uint32_t b0 = o0 & ~3u, b1 = o1 & ~3u;
    if (max(b0, b1) + 12u > img_bytes) return false;
    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(img + b0);
    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(img + b1);
    const uint32_t a0 = __ldg(w0), a1 = __ldg(w0 + 1), a2 = __ldg(w0 + 2);
    const uint32_t c0 = __ldg(w1), c1 = __ldg(w1 + 1), c2 = __ldg(w1 + 2);
    const uint32_t s0 = (o0 & 3u) * 8u, s1 = (o1 & 3u) * 8u;
    const uint32_t lo0 = __funnelshift_r(a0, a1, s0), hi0 = __funnelshift_r(a1, a2, s0);   // bytes o..o+3 | o+4..o+7
    const uint32_t lo1 = __funnelshift_r(c0, c1, s1), hi1 = __funnelshift_r(c1, c2, s1);
    const bool dup = xi1 == xi;            // right edge: the +1 tap is the pixel itself
    const uint32_t fx1 = 1024u - fx, fy1 = 1024u - fy;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const uint32_t p00 = (lo0 >> (8 * ch)) & 0xFFu, p10 = (lo1 >> (8 * ch)) & 0xFFu;
        const uint32_t n0 = ch == 0 ? (lo0 >> 24) : ((hi0 >> (8 * (ch - 1))) & 0xFFu);
        const uint32_t n1 = ch == 0 ? (lo1 >> 24) : ((hi1 >> (8 * (ch - 1))) & 0xFFu);
        const uint32_t p01 = dup ? p00 : n0, p11 = dup ? p10 : n1;
        const uint32_t top = p00 * fx1 + p01 * fx;
        const uint32_t bot = p10 * fx1 + p11 * fx;
        d[ch] = (uint8_t)((top * fy1 + bot * fy + (1u << 19)) >> 20);
    }
    return true;
}

}  // namespace kb200
