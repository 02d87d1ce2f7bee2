// oracle/kornia_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (C++17) of the kornia-rs `kornia-imgproc` pixel-kernel hot
// path named in BASELINE.json:north_star.  Every function cites the reference
// file:line it follows (paths relative to /root/reference/crates/kornia-imgproc/src).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load this library — and only as the checker / the CPU number that is
// reported beside the GPU number.  The product path (kornia-rs_b200/, the
// C-ABI libkornia_b200.so) never links or calls anything in here.
//
// PARITY STATUS: pinned.  The Rust reference cannot be built in this image (no
// cargo/rustc), so the oracle is pinned against the reference's own in-tree
// known-answer tests (transcribed into tests/test_oracle_golden.py with the
// file:line of each) and against cv2 fixtures for the NV12 Q20 decode (the
// reference states byte-parity with cv2 for that path).
//
// Build rules that matter for bit-level parity with the Rust code:
//   * -ffp-contract=off : rustc never contracts a*b+c into an fma; neither do we.
//     Where the reference's x86 leaf *does* call an FMA intrinsic
//     (_mm256_fmadd_ps), we call fmaf() explicitly ("leaf" selectors below).
//   * no -ffast-math; IEEE division and sqrt everywhere.
//   * expf from the platform libm, like Rust's f32::exp on Linux.
//
// Threading mirrors the reference's rayon chunking (16-row tasks; 8-row tasks
// in the fused resize; strip split above 1 Mpx for colour ops) with OpenMP so
// that the timed CPU baseline uses all host cores the way the reference does.
// Functions the reference runs single-threaded (separable_filter, sobel,
// std_mean, find_min_max) are single-threaded here too unless the caller asks
// for the "mt" variant.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define KO_API extern "C" __attribute__((visibility("default")))

// leaf selectors: which CPU leaf of the reference to mirror where scalar and
// SIMD leaves round differently (FMA vs mul+add).
enum { KO_LEAF_SCALAR = 0, KO_LEAF_X86_AVX2_FMA = 1, KO_LEAF_AARCH64_NEON = 2 };

KO_API int ko_version() { return 1; }

KO_API void ko_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

KO_API int ko_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ─────────────────────────────────────────────────────────────────────────────
// Deterministic input generators used by the reference's GPU-parity tests.
// cuda/color/mod.rs:303-321 (pattern_u8 / pattern_f32).  `seed` generalises the
// fixed 0x12345678 so batches can use one pattern per image.
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_pattern_u8(uint8_t* out, size_t len, uint32_t seed) {
    static const uint8_t prefix[15] = {0, 255, 255, 0, 0, 0, 255, 255, 255, 1, 254, 128, 128, 128, 64};
    size_t n = std::min<size_t>(15, len);
    for (size_t i = 0; i < n; ++i) out[i] = prefix[i];
    uint32_t state = seed;
    for (size_t i = n; i < len; ++i) {
        state = state * 1664525u + 1013904223u;
        out[i] = (uint8_t)(state >> 24);
    }
}

KO_API void ko_pattern_f32(float* out, size_t len, uint32_t seed) {
    std::vector<uint8_t> tmp(len);
    ko_pattern_u8(tmp.data(), len, seed);
    for (size_t i = 0; i < len; ++i) out[i] = (float)tmp[i] / 255.0f;
}

// ─────────────────────────────────────────────────────────────────────────────
// a8: gray_from_rgb  — color/gray/kernels.rs
// ─────────────────────────────────────────────────────────────────────────────
static const float RW_F32 = 0.299f, GW_F32 = 0.587f, BW_F32 = 0.114f;  // kernels.rs:2-4

// f32: scalar leaf kernels.rs:405-410 (`rw*r + gw*g + bw*b`, left to right,
// unfused); x86 AVX2+FMA leaf kernels.rs:338-402: fma(r,rw, fma(g,gw, b*bw)) for
// the 8-px bulk, scalar expression for the npixels%8 tail.  Strips
// (kernel_common.rs:51-78) are 8-px aligned so only the global tail is scalar.
KO_API void ko_gray_from_rgb_f32(const float* src, float* dst, size_t npixels, int leaf) {
    size_t bulk = 0;
    if (leaf == KO_LEAF_X86_AVX2_FMA || leaf == KO_LEAF_AARCH64_NEON) bulk = npixels & ~(size_t)7;
#pragma omp parallel for schedule(static) if (npixels >= 1024 * 1024)
    for (long long i = 0; i < (long long)npixels; ++i) {
        const float r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
        if ((size_t)i < bulk) {
            dst[i] = fmaf(r, RW_F32, fmaf(g, GW_F32, b * BW_F32));
        } else {
            dst[i] = RW_F32 * r + GW_F32 * g + BW_F32 * b;
        }
    }
}

// u8: kernels.rs:229-238  (4899R + 9617G + 1868B + 8192) >> 14
KO_API void ko_gray_from_rgb_u8(const uint8_t* src, uint8_t* dst, size_t npixels) {
#pragma omp parallel for schedule(static) if (npixels >= 1024 * 1024)
    for (long long i = 0; i < (long long)npixels; ++i) {
        const uint32_t r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
        dst[i] = (uint8_t)((4899u * r + 9617u * g + 1868u * b + 8192u) >> 14);
    }
}

// ─────────────────────────────────────────────────────────────────────────────
// a9: rgb_from_nv12 / rgb_from_yuyv — color/yuv/kernels.rs:707-737 (decode_px,
// yy_term), :945-966 (packed422 row), :1036-1069 (planar420 block), :1195-1221
// (chroma_at), color/yuv/mod.rs:209-241 (entry points).
// ─────────────────────────────────────────────────────────────────────────────
static inline int32_t yy_term(int32_t y) { return std::max(y - 16, 0) * 1220542; }  // :734-737

static inline void decode_px(int32_t yy, int32_t u, int32_t v, uint8_t* r, uint8_t* g, uint8_t* b) {  // :718-730
    u -= 128;
    v -= 128;
    const int32_t bb = (yy + 2116026 * u + (1 << 19)) >> 20;
    const int32_t gg = (yy + (-409993) * u + (-852492) * v + (1 << 19)) >> 20;
    const int32_t rr = (yy + 1673527 * v + (1 << 19)) >> 20;
    *r = (uint8_t)std::min(std::max(rr, 0), 255);
    *g = (uint8_t)std::min(std::max(gg, 0), 255);
    *b = (uint8_t)std::min(std::max(bb, 0), 255);
}

// src = Y plane (w*h) followed by interleaved UV (w*h/2); dst RGB8 HWC.
KO_API int ko_rgb_from_nv12_u8(const uint8_t* src, uint8_t* dst, size_t w, size_t h) {
    if ((w & 1) || (h & 1)) return -1;
    const uint8_t* y = src;
    const uint8_t* uv = src + w * h;
    const size_t cw = w / 2;
#pragma omp parallel for schedule(static) if (w * h >= 1024 * 1024)
    for (long long cy = 0; cy < (long long)(h / 2); ++cy) {
        const uint8_t* y_top = y + (2 * cy) * w;
        const uint8_t* y_bot = y + (2 * cy + 1) * w;
        uint8_t* d_top = dst + (2 * cy) * w * 3;
        uint8_t* d_bot = dst + (2 * cy + 1) * w * 3;
        for (size_t cx = 0; cx < cw; ++cx) {
            const size_t idx = cy * cw * 2 + cx * 2;
            const int32_t u = uv[idx], v = uv[idx + 1];
            const size_t x = cx * 2;
            for (size_t dx = 0; dx < 2; ++dx) {
                const size_t dt = (x + dx) * 3;
                decode_px(yy_term(y_top[x + dx]), u, v, &d_top[dt], &d_top[dt + 1], &d_top[dt + 2]);
                decode_px(yy_term(y_bot[x + dx]), u, v, &d_bot[dt], &d_bot[dt + 1], &d_bot[dt + 2]);
            }
        }
    }
    return 0;
}

// YUYV packed 4:2:2 (`Y0 U Y1 V`), 2 bytes/px.
KO_API int ko_rgb_from_yuyv_u8(const uint8_t* src, uint8_t* dst, size_t w, size_t h) {
    if (w & 1) return -1;
#pragma omp parallel for schedule(static) if (w * h >= 1024 * 1024)
    for (long long row = 0; row < (long long)h; ++row) {
        const uint8_t* s = src + row * w * 2;
        uint8_t* d = dst + row * w * 3;
        for (size_t g = 0; g < w / 2; ++g) {
            const size_t base = g * 4;
            const int32_t y0 = s[base], u = s[base + 1], y1 = s[base + 2], v = s[base + 3];
            decode_px(yy_term(y0), u, v, &d[g * 6], &d[g * 6 + 1], &d[g * 6 + 2]);
            decode_px(yy_term(y1), u, v, &d[g * 6 + 3], &d[g * 6 + 4], &d[g * 6 + 5]);
        }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// Shared samplers — interpolation/bilinear.rs:16-66, interpolation/nearest.rs:15-30
// ─────────────────────────────────────────────────────────────────────────────
static inline float bilinear_interpolation(const float* img, size_t rows, size_t cols, size_t C, float u, float v,
                                           size_t c) {
    const size_t iu = (size_t)truncf(u);
    const size_t iv = (size_t)truncf(v);
    const float frac_u = u - truncf(u);  // f32::fract
    const float frac_v = v - truncf(v);
    const float val00 = img[(iv * cols + iu) * C + c];
    const float val01 = (iu + 1 < cols) ? img[(iv * cols + iu + 1) * C + c] : val00;
    const float val10 = (iv + 1 < rows) ? img[((iv + 1) * cols + iu) * C + c] : val00;
    const float val11 = (iu + 1 < cols && iv + 1 < rows) ? img[((iv + 1) * cols + iu + 1) * C + c] : val00;
    const float frac_uu = 1.0f - frac_u;
    const float frac_vv = 1.0f - frac_v;
    const float w00 = frac_vv * frac_uu;
    const float w10 = frac_vv * frac_u;
    const float w01 = frac_v * frac_uu;
    const float w11 = frac_v * frac_u;
    return w00 * val00 + w10 * val01 + w01 * val10 + w11 * val11;
}

static inline float nearest_neighbor_interpolation(const float* img, size_t rows, size_t cols, size_t C, float u,
                                                   float v, size_t c) {
    // `u.round() as usize` (half away from zero; saturating cast) then clamp.
    float ru = roundf(u), rv = roundf(v);
    size_t iu = ru <= 0.0f ? 0 : (size_t)ru;
    size_t iv = rv <= 0.0f ? 0 : (size_t)rv;
    iu = std::min(iu, cols - 1);
    iv = std::min(iv, rows - 1);
    return img[(iv * cols + iu) * C + c];
}

enum { KO_NEAREST = 0, KO_BILINEAR = 1, KO_BICUBIC = 2, KO_LANCZOS = 3 };


// ─────────────────────────────────────────────────────────────────────────────
// §8(f) #3: bicubic (Keys a = -0.5) and Lanczos-3 samplers — interpolation/bicubic.rs:12-61,
// interpolation/lanczos.rs:16-266.  The reference keeps these byte-exact with its CUDA kernels: `mul_add` where
// the kernels say fmaf, plain mul/add elsewhere (this file is built -ffp-contract=off, so plain expressions stay
// unfused and fmaf() is the only fusion).
// ─────────────────────────────────────────────────────────────────────────────
static inline void keys_weights(float frac, float w[4]) {  // bicubic.rs:14-27
    float t;
    t = 1.0f + frac; w[0] = fmaf(fmaf(fmaf(-0.5f, t, 2.5f), t, -4.0f), t, 2.0f);
    t = frac;        w[1] = fmaf(fmaf(1.5f, t, -2.5f) * t, t, 1.0f);
    t = 1.0f - frac; w[2] = fmaf(fmaf(1.5f, t, -2.5f) * t, t, 1.0f);
    t = 2.0f - frac; w[3] = fmaf(fmaf(fmaf(-0.5f, t, 2.5f), t, -4.0f), t, 2.0f);
}

static inline long long clamp_ll(long long v, long long lo, long long hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline float bicubic_sample(const float* img, size_t rows, size_t cols, size_t C, float sx, float sy, size_t c) {  // bicubic.rs:33-61
    const float x0f = floorf(sx), y0f = floorf(sy);
    float wx[4], wy[4];
    keys_weights(sx - x0f, wx);
    keys_weights(sy - y0f, wy);
    const long long x0 = (long long)x0f, y0 = (long long)y0f;
    float acc = 0.0f;
    for (int dy = 0; dy < 4; ++dy) {
        const size_t yi = (size_t)clamp_ll(y0 + dy - 1, 0, (long long)rows - 1);
        const size_t row = yi * cols * C;
        for (int dx = 0; dx < 4; ++dx) {
            const size_t xi = (size_t)clamp_ll(x0 + dx - 1, 0, (long long)cols - 1);
            const float w = wx[dx] * wy[dy];
            acc = fmaf(w, img[row + xi * C + c], acc);
        }
    }
    return acc;
}

static inline float sin_pi(float x) {  // lanczos.rs:19-35
    const float k = roundf(x);
    const float r = x - k;
    const float z = 3.14159265358979323846f * r;
    const float z2 = z * z;
    float p = -2.5052108e-8f;
    p = p * z2 + 2.7557319e-6f;
    p = p * z2 + -1.984127e-4f;
    p = p * z2 + 8.333334e-3f;
    p = p * z2 + -1.6666667e-1f;
    const float s = z + z * z2 * p;
    return (((int)k) & 1) ? -s : s;
}

static inline float lanczos3(float x) {  // lanczos.rs:39-50
    const float PI = 3.14159265358979323846f;
    if (fabsf(x) < 1e-5f) return 1.0f;
    if (fabsf(x) >= 3.0f) return 0.0f;
    const float pix = PI * x;
    const float pix3 = pix * 0.33333334f;
    return sin_pi(x) * sin_pi(x * (1.0f / 3.0f)) / (pix * pix3);
}

KO_API float ko_sin_pi(float x) { return sin_pi(x); }
KO_API float ko_lanczos3(float x) { return lanczos3(x); }

// lanczos.rs:59-92 — per-axis tap base + six normalised weights on the half-pixel grid
KO_API void ko_lanczos_axis(size_t src_len, size_t dst_len, int32_t* x0s, float* weights) {
    if (src_len == 0 || dst_len == 0) return;
    const float a = (float)src_len / (float)dst_len;
    const float b = 0.5f * a - 0.5f;
    const float mx = (float)(src_len - 1);
    for (size_t i = 0; i < dst_len; ++i) {
        float s = a * (float)i + b;
        if (s < 0.0f) s = 0.0f;
        if (s > mx) s = mx;
        const float x0 = floorf(s);
        const float frac = s - x0;
        x0s[i] = (int32_t)x0;
        float w[6] = {lanczos3(frac + 2.0f), lanczos3(frac + 1.0f), lanczos3(frac), lanczos3(frac - 1.0f), lanczos3(frac - 2.0f), lanczos3(frac - 3.0f)};
        const float sum = w[0] + w[1] + w[2] + w[3] + w[4] + w[5];
        const float inv = 1.0f / sum;
        for (int t = 0; t < 6; ++t) weights[i * 6 + t] = w[t] * inv;
    }
}

// lanczos.rs:106-137 — six weights from four sin_pi evaluations (the warp kernels' form)
KO_API void ko_lanczos3_weights(float frac, float w[6]) {
    const float PI = 3.14159265358979323846f;
    const float s = sin_pi(frac);
    const float t0 = sin_pi(frac * (1.0f / 3.0f));
    const float t1 = sin_pi((frac - 1.0f) * (1.0f / 3.0f));
    const float t2 = sin_pi((frac - 2.0f) * (1.0f / 3.0f));
    const float st0 = s * t0, st1 = s * t1, st2 = s * t2;
    auto den = [&](float x) { const float pix = PI * x; const float pix3 = pix * 0.33333334f; return pix * pix3; };
    w[0] = -st1 / den(frac + 2.0f);
    w[1] = st2 / den(frac + 1.0f);
    w[2] = st0 / den(frac);
    w[3] = -st1 / den(frac - 1.0f);
    w[4] = st2 / den(frac - 2.0f);
    w[5] = st0 / den(frac - 3.0f);
    if (frac < 1e-5f) w[2] = 1.0f;
    if (fabsf(frac - 1.0f) < 1e-5f) w[3] = 1.0f;
}

static inline float lanczos_sample(const float* img, size_t rows, size_t cols, size_t C, float sx, float sy, size_t c) {  // lanczos.rs:143-181
    const float x0f = floorf(sx), y0f = floorf(sy);
    const float frac_x = sx - x0f, frac_y = sy - y0f;
    const long long x0 = (long long)x0f, y0 = (long long)y0f;
    float wx[6], wy[6];
    ko_lanczos3_weights(frac_x, wx);
    ko_lanczos3_weights(frac_y, wy);
    const float sum_wx = wx[0] + wx[1] + wx[2] + wx[3] + wx[4] + wx[5];
    const float sum_wy = wy[0] + wy[1] + wy[2] + wy[3] + wy[4] + wy[5];
    const float inv_x = 1.0f / sum_wx, inv_y = 1.0f / sum_wy;
    for (int i = 0; i < 6; ++i) { wx[i] *= inv_x; wy[i] *= inv_y; }
    float acc = 0.0f;
    for (int dy = 0; dy < 6; ++dy) {
        const size_t yi = (size_t)clamp_ll(y0 + dy - 2, 0, (long long)rows - 1);
        const size_t row = yi * cols * C;
        float rx = 0.0f;
        for (int dx = 0; dx < 6; ++dx) {
            const size_t xi = (size_t)clamp_ll(x0 + dx - 2, 0, (long long)cols - 1);
            rx = fmaf(wx[dx], img[row + xi * C + c], rx);
        }
        acc = fmaf(wy[dy], rx, acc);
    }
    return acc;
}

// lanczos.rs:187-236 — separable H-then-V with an f32 intermediate of dst_w x src_h
static void resize_lanczos_separable(const float* src, size_t sw, size_t sh, float* dst, size_t dw, size_t dh, size_t C) {
    std::vector<int32_t> x0s(dw), y0s(dh);
    std::vector<float> wx(dw * 6), wy(dh * 6);
    ko_lanczos_axis(sw, dw, x0s.data(), wx.data());
    ko_lanczos_axis(sh, dh, y0s.data(), wy.data());
    std::vector<float> inter(dw * sh * C);
#pragma omp parallel for schedule(static)
    for (long long sy = 0; sy < (long long)sh; ++sy) {
        const float* srow = src + (size_t)sy * sw * C;
        float* irow = inter.data() + (size_t)sy * dw * C;
        for (size_t dx = 0; dx < dw; ++dx) {
            const long long x0 = x0s[dx];
            const float* w = &wx[dx * 6];
            for (size_t k = 0; k < C; ++k) {
                float acc = 0.0f;
                for (int t = 0; t < 6; ++t) {
                    const size_t xi = (size_t)clamp_ll(x0 + t - 2, 0, (long long)sw - 1);
                    acc = fmaf(w[t], srow[xi * C + k], acc);
                }
                irow[dx * C + k] = acc;
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (long long dy = 0; dy < (long long)dh; ++dy) {
        const long long y0 = y0s[(size_t)dy];
        const float* w = &wy[(size_t)dy * 6];
        float* drow = dst + (size_t)dy * dw * C;
        for (size_t dx = 0; dx < dw; ++dx)
            for (size_t k = 0; k < C; ++k) {
                float acc = 0.0f;
                for (int t = 0; t < 6; ++t) {
                    const size_t yi = (size_t)clamp_ll(y0 + t - 2, 0, (long long)sh - 1);
                    acc = fmaf(w[t], inter[(yi * dw + dx) * C + k], acc);
                }
                drow[dx * C + k] = acc;
            }
    }
}

static inline float sample_mode(int mode, const float* img, size_t rows, size_t cols, size_t C, float u, float v, size_t c);

// ─────────────────────────────────────────────────────────────────────────────
// a1: resize::resize<C>(src,dst,mode) — resize/mod.rs:114-207
// ─────────────────────────────────────────────────────────────────────────────
KO_API int ko_resize_f32(const float* src, size_t sw, size_t sh, float* dst, size_t dw, size_t dh, size_t C,
                         int mode) {
    if (mode < KO_NEAREST || mode > KO_LANCZOS) return -1;
    if (sw == dw && sh == dh) {  // :134-137
        std::memcpy(dst, src, sw * sh * C * sizeof(float));
        return 0;
    }
    if (mode == KO_LANCZOS) {  // :142-145: separable on both backends
        resize_lanczos_separable(src, sw, sh, dst, dw, dh, C);
        return 0;
    }
    // axis_lut :169-176 — a*x + b, a = src/dst, b = 0.5a - 0.5, clamp to [0, src-1]
    auto axis_lut = [](size_t src_len, size_t dst_len) {
        std::vector<float> v(dst_len);
        const float a = (float)src_len / (float)dst_len;
        const float b = 0.5f * a - 0.5f;
        const float mx = (float)(src_len - 1);
        for (size_t i = 0; i < dst_len; ++i) {
            float s = a * (float)i + b;
            // f32::clamp(0, max)
            if (s < 0.0f) s = 0.0f;
            if (s > mx) s = mx;
            v[i] = s;
        }
        return v;
    };
    const std::vector<float> xs = axis_lut(sw, dw), ys = axis_lut(sh, dh);
    const long long nchunks = (long long)((dh + 15) / 16);  // parallel.rs ROWS_PER_TASK = 16
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ck = 0; ck < nchunks; ++ck) {
        const size_t y_end = std::min<size_t>(dh, (size_t)(ck + 1) * 16);
        for (size_t y = (size_t)ck * 16; y < y_end; ++y) {
            for (size_t x = 0; x < dw; ++x) {
                float* px = dst + (y * dw + x) * C;
                for (size_t k = 0; k < C; ++k) {
                    px[k] = sample_mode(mode, src, sh, sw, C, xs[x], ys[y], k);
                }
            }
        }
    }
    return 0;
}

static inline float sample_mode(int mode, const float* img, size_t rows, size_t cols, size_t C, float u, float v, size_t c) {
    switch (mode) {   // interpolation/interpolate.rs:53-66
        case KO_BILINEAR: return bilinear_interpolation(img, rows, cols, C, u, v, c);
        case KO_BICUBIC: return bicubic_sample(img, rows, cols, C, u, v, c);
        case KO_LANCZOS: return lanczos_sample(img, rows, cols, C, u, v, c);
        default: return nearest_neighbor_interpolation(img, rows, cols, C, u, v, c);
    }
}

// ─────────────────────────────────────────────────────────────────────────────
// a2: resize_normalize_to_tensor_u8_to_f32_bilinear — resize/fused.rs:147-228,
// scalar leaf :273-318, AVX2 leaf :414-497 (FMA form on the dst_w&~7 bulk),
// exact-2× box path :57-127 with scalar leaf :528-559 and AVX2 leaf (FMA on the
// dst_w&~15 bulk).  NormalizeParams::from_mean_std :28-37.
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_normalize_params_from_mean_std(const float mean[3], const float stdv[3], float scale[3],
                                              float bias[3]) {
    for (int c = 0; c < 3; ++c) {
        scale[c] = 1.0f / (stdv[c] * 255.0f);
        bias[c] = -mean[c] / stdv[c];
    }
}

static void fused_2x(const uint8_t* src, size_t sw, size_t /*sh*/, float* dst, size_t dw, size_t dh,
                     const float scale[3], const float bias[3], int leaf) {
    const size_t src_stride = sw * 3, plane = dw * dh;
    const float s4[3] = {scale[0] * 0.25f, scale[1] * 0.25f, scale[2] * 0.25f};
    size_t bulk = 0;
    if (leaf == KO_LEAF_X86_AVX2_FMA || leaf == KO_LEAF_AARCH64_NEON) bulk = dw & ~(size_t)15;
    const long long nchunks = (long long)((dh + 7) / 8);  // ROWS_PER_TASK = 8 (:104)
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ck = 0; ck < nchunks; ++ck) {
        const size_t y_end = std::min<size_t>(dh, (size_t)(ck + 1) * 8);
        for (size_t y = (size_t)ck * 8; y < y_end; ++y) {
            const uint8_t* r0 = src + (2 * y) * src_stride;
            const uint8_t* r1 = src + (2 * y + 1) * src_stride;
            for (size_t x = 0; x < dw; ++x) {
                const size_t b0 = 2 * x * 3, b1 = b0 + 3;
                for (int c = 0; c < 3; ++c) {
                    const uint32_t sum = (uint32_t)r0[b0 + c] + r0[b1 + c] + r1[b0 + c] + r1[b1 + c];
                    float o;
                    if (x < bulk) o = fmaf((float)sum, s4[c], bias[c]);
                    else o = (float)sum * s4[c] + bias[c];
                    dst[c * plane + y * dw + x] = o;
                }
            }
        }
    }
}

KO_API int ko_resize_normalize_u8_to_f32_chw_bilinear(const uint8_t* src, size_t sw, size_t sh, float* dst,
                                                      size_t dw, size_t dh, const float scale[3],
                                                      const float bias[3], int leaf) {
    if (sw == 2 * dw && sh == 2 * dh) {  // :181-183
        fused_2x(src, sw, sh, dst, dw, dh, scale, bias, leaf);
        return 0;
    }
    if (dw == 0 || dh == 0 || sw == 0 || sh == 0) return 0;  // :184-186
    const float scale_x = (float)sw / (float)dw;
    const float scale_y = (float)sh / (float)dh;
    const size_t src_stride = sw * 3;
    std::vector<size_t> x0b(dw), x1b(dw);
    std::vector<float> wx(dw);
    for (size_t dx = 0; dx < dw; ++dx) {  // :196-203
        float fx = ((float)dx + 0.5f) * scale_x - 0.5f;
        fx = std::max(fx, 0.0f);
        const size_t x0 = std::min((size_t)fx, sw - 1);
        const size_t x1 = std::min(x0 + 1, sw - 1);
        x0b[dx] = x0 * 3;
        x1b[dx] = x1 * 3;
        wx[dx] = fx - (float)x0;
    }
    size_t bulk = 0;
    if (leaf == KO_LEAF_X86_AVX2_FMA) bulk = dw & ~(size_t)7;       // :448
    if (leaf == KO_LEAF_AARCH64_NEON) bulk = dw & ~(size_t)3;       // :355
    const size_t plane = dw * dh;
    const long long nchunks = (long long)((dh + 7) / 8);  // ROWS_PER_TASK = 8 (:209)
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ck = 0; ck < nchunks; ++ck) {
        const size_t y_end = std::min<size_t>(dh, (size_t)(ck + 1) * 8);
        for (size_t y = (size_t)ck * 8; y < y_end; ++y) {
            float fy = ((float)y + 0.5f) * scale_y - 0.5f;
            fy = std::max(fy, 0.0f);
            const size_t y0 = std::min((size_t)fy, sh - 1);
            const size_t y1 = std::min(y0 + 1, sh - 1);
            const float wy = fy - (float)y0;
            const uint8_t* row0 = src + y0 * src_stride;
            const uint8_t* row1 = src + y1 * src_stride;
            float* out[3] = {dst + y * dw, dst + plane + y * dw, dst + 2 * plane + y * dw};
            // bulk: the reference's SIMD leaf shape (:414-497 / :320-407) — gather GW pixels' 4x3 corner samples into
            // stack arrays, then per channel a GW-wide FMA blend (the compiler vectorises the k loop)
            const size_t GW = (leaf == KO_LEAF_AARCH64_NEON) ? 4 : 8;
            size_t dx = 0;
            for (; dx < bulk; dx += GW) {
                float p00[24], p01[24], p10[24], p11[24];
                for (size_t k = 0; k < GW; ++k) {
                    const size_t o0 = x0b[dx + k], o1 = x1b[dx + k];
                    for (int c = 0; c < 3; ++c) {
                        p00[c * 8 + k] = (float)row0[o0 + c]; p01[c * 8 + k] = (float)row0[o1 + c];
                        p10[c * 8 + k] = (float)row1[o0 + c]; p11[c * 8 + k] = (float)row1[o1 + c];
                    }
                }
                for (int c = 0; c < 3; ++c) {
                    for (size_t k = 0; k < GW; ++k) {  // :475-478  fmadd(sub(b,a), wx, a) …
                        const float a = p00[c * 8 + k], b = p01[c * 8 + k], cc = p10[c * 8 + k], d = p11[c * 8 + k];
                        const float top = fmaf(b - a, wx[dx + k], a);
                        const float bot = fmaf(d - cc, wx[dx + k], cc);
                        const float val = fmaf(bot - top, wy, top);
                        out[c][dx + k] = fmaf(val, scale[c], bias[c]);
                    }
                }
            }
            for (; dx < dw; ++dx) {  // scalar leaf / tail :286-317
                const size_t o0 = x0b[dx], o1 = x1b[dx];
                const float w = wx[dx];
                for (int c = 0; c < 3; ++c) {
                    const float a = (float)row0[o0 + c], b = (float)row0[o1 + c];
                    const float cc = (float)row1[o0 + c], d = (float)row1[o1 + c];
                    const float top = a + w * (b - a);
                    const float bot = cc + w * (d - cc);
                    const float val = top + wy * (bot - top);
                    out[c][dx] = val * scale[c] + bias[c];
                }
            }
        }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// a3: u8 bilinear Q14 — resize/bilinear.rs:25-104 (bilinear_tap, axis LUT, row
// driver), resize/kernels.rs:1141-1166 (bilinear_row_u8_scalar).
// ─────────────────────────────────────────────────────────────────────────────
static inline void bilinear_tap(size_t i, double scale, size_t src_len, uint32_t* ofs, uint32_t* fq) {
    const double s = ((double)i + 0.5) * scale - 0.5;
    long long i0 = (long long)std::floor(s);
    double f = s - (double)i0;
    if (i0 < 0) {
        i0 = 0;
        f = 0.0;
    } else if (i0 >= (long long)src_len - 1) {
        i0 = (long long)src_len - 2;
        f = 1.0;
    }
    double q = std::round(f * 16384.0);
    uint32_t fqv = (uint32_t)q;
    *fq = std::min(fqv, 16384u);
    *ofs = (uint32_t)i0;
}

KO_API int ko_resize_bilinear_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst, size_t dw, size_t dh,
                                 size_t C) {
    if (!(C == 1 || C == 3 || C == 4)) return -1;
    if (sw < 2 || sh < 2) return -2;  // resize/mod.rs:318-320
    const double scale_x = (double)sw / (double)dw, scale_y = (double)sh / (double)dh;
    std::vector<uint32_t> xofs(dw), xfx(dw), xfx1(dw);
    for (size_t i = 0; i < dw; ++i) {
        bilinear_tap(i, scale_x, sw, &xofs[i], &xfx[i]);
        xfx1[i] = 16384u - xfx[i];
    }
    const size_t ss = sw * C, ds = dw * C;
    const long long nchunks = (long long)((dh + 15) / 16);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ck = 0; ck < nchunks; ++ck) {
        const size_t y_end = std::min<size_t>(dh, (size_t)(ck + 1) * 16);
        for (size_t y = (size_t)ck * 16; y < y_end; ++y) {
            uint32_t yi, fy;
            bilinear_tap(y, scale_y, sh, &yi, &fy);
            const uint64_t fy1 = 16384u - fy;
            const uint8_t* row0 = src + (size_t)yi * ss;
            const uint8_t* row1 = src + ((size_t)yi + 1) * ss;
            uint8_t* drow = dst + y * ds;
            for (size_t x = 0; x < dw; ++x) {
                const size_t off = (size_t)xofs[x] * C;
                const uint64_t fx = xfx[x], fx1 = xfx1[x];
                for (size_t ch = 0; ch < C; ++ch) {
                    const uint64_t p00 = row0[off + ch], p01 = row0[off + C + ch];
                    const uint64_t p10 = row1[off + ch], p11 = row1[off + C + ch];
                    const uint64_t top = p00 * fx1 + p01 * fx;
                    const uint64_t bot = p10 * fx1 + p11 * fx;
                    drow[x * C + ch] = (uint8_t)((top * fy1 + bot * (uint64_t)fy + (1ull << 27)) >> 28);
                }
            }
        }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f)#1 — the other arms of resize_fast_u8_aa (resize/mod.rs:283-410): exact-2x pyramid paths and nearest.
// ─────────────────────────────────────────────────────────────────────────────

// resize/pyramid.rs:18-45 + resize/kernels.rs:64-75 (pyrdown_row_rgb_u8_scalar): rounded mean of each 2x2 block.
KO_API void ko_pyrdown_2x_rgb_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst) {
    const size_t dw = sw / 2, dh = sh / 2, ss = sw * 3, ds = dw * 3;
    for (size_t y = 0; y < dh; ++y) {
        const uint8_t* r0 = src + (2 * y) * ss;
        const uint8_t* r1 = src + (2 * y + 1) * ss;
        uint8_t* d = dst + y * ds;
        for (size_t x = 0; x < dw; ++x)
            for (size_t ch = 0; ch < 3; ++ch) {
                const uint16_t sum = (uint16_t)(r0[(2 * x) * 3 + ch] + r0[(2 * x + 1) * 3 + ch] + r1[(2 * x) * 3 + ch] + r1[(2 * x + 1) * 3 + ch]);
                d[x * 3 + ch] = (uint8_t)((sum + 2) >> 2);
            }
    }
}

// resize/kernels.rs:168-181 (hinterp_row_rgb_u8_scalar)
static void hinterp_row_rgb_u8(const uint8_t* src, uint8_t* dst, size_t sw) {
    for (size_t ch = 0; ch < 3; ++ch) dst[ch] = src[ch];
    for (size_t j = 0; j + 1 < sw; ++j)
        for (size_t ch = 0; ch < 3; ++ch) {
            const uint16_t a = src[j * 3 + ch], b = src[(j + 1) * 3 + ch];
            const uint16_t avg = (uint16_t)((a + b + 1) >> 1);
            dst[(2 * j + 1) * 3 + ch] = (uint8_t)((a + avg + 1) >> 1);
            dst[(2 * j + 2) * 3 + ch] = (uint8_t)((b + avg + 1) >> 1);
        }
    const size_t tail = (2 * sw - 1) * 3;
    for (size_t ch = 0; ch < 3; ++ch) dst[tail + ch] = src[(sw - 1) * 3 + ch];
}
// resize/kernels.rs:274-281 (blend_75_25_row_scalar): round(0.75 a + 0.25 b) as two rounding half-adds
static void blend_75_25_row(const uint8_t* a, const uint8_t* b, uint8_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const uint16_t av = a[i], bv = b[i];
        const uint16_t avg = (uint16_t)((av + bv + 1) >> 1);
        dst[i] = (uint8_t)((av + avg + 1) >> 1);
    }
}
// resize/pyramid.rs:50-120: rows 0 and 2h-1 are horizontally interpolated edge rows; block I writes rows 2I+1, 2I+2.
KO_API void ko_pyrup_2x_rgb_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst) {
    const size_t dw = sw * 2, ss = sw * 3, ds = dw * 3;
    hinterp_row_rgb_u8(src, dst, sw);
    hinterp_row_rgb_u8(src + (sh - 1) * ss, dst + (2 * sh - 1) * ds, sw);
    std::vector<uint8_t> ha(ds), hb(ds);
    if (sh >= 2) hinterp_row_rgb_u8(src, ha.data(), sw);
    for (size_t i = 0; i + 1 < sh; ++i) {
        hinterp_row_rgb_u8(src + (i + 1) * ss, hb.data(), sw);
        blend_75_25_row(ha.data(), hb.data(), dst + (2 * i + 1) * ds, ds);
        blend_75_25_row(hb.data(), ha.data(), dst + (2 * i + 2) * ds, ds);
        ha.swap(hb);
    }
}

// resize/nearest.rs:18-21: clamp(floor((i + 0.5) * scale)) in f64 — pixel centre then floor, no -0.5
static size_t nearest_index(size_t i, double scale, size_t src_len) {
    long long v = (long long)std::floor(((double)i + 0.5) * scale);
    if (v < 0) v = 0;
    if (v > (long long)src_len - 1) v = (long long)src_len - 1;
    return (size_t)v;
}
// resize/nearest.rs:42-78
KO_API void ko_resize_nearest_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst, size_t dw, size_t dh, size_t C) {
    const double sy = (double)sh / (double)dh, sx = (double)sw / (double)dw;
    std::vector<size_t> xmap(dw);
    for (size_t x = 0; x < dw; ++x) xmap[x] = nearest_index(x, sx, sw);
    for (size_t y = 0; y < dh; ++y) {
        const uint8_t* srow = src + nearest_index(y, sy, sh) * sw * C;
        uint8_t* drow = dst + y * dw * C;
        for (size_t x = 0; x < dw; ++x)
            for (size_t ch = 0; ch < C; ++ch) drow[x * C + ch] = srow[xmap[x] * C + ch];
    }
}

// resize/mod.rs:283-410 — path selection of resize_fast_u8_aa for Nearest / Bilinear.
// interp: 0 = Nearest, 1 = Bilinear.  Returns 0, -1 (UnsupportedChannelCount), -2 (InvalidImageSize 2x2), -3 (mode not covered).
KO_API int ko_resize_fast_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst, size_t dw, size_t dh, size_t C, int interp) {
    if (interp == 1 && C == 3 && sw == dw * 2 && sh == dh * 2 && sw >= 2 && sh >= 2) { ko_pyrdown_2x_rgb_u8(src, sw, sh, dst); return 0; }
    if (interp == 1 && C == 3 && dw == sw * 2 && dh == sh * 2 && sw >= 2 && sh >= 2) { ko_pyrup_2x_rgb_u8(src, sw, sh, dst); return 0; }
    if (interp == 0) { ko_resize_nearest_u8(src, sw, sh, dst, dw, dh, C); return 0; }
    if (interp == 1) return ko_resize_bilinear_u8(src, sw, sh, dst, dw, dh, C);
    return -3;
}

// ─────────────────────────────────────────────────────────────────────────────
// a4: warp_affine — warp/affine.rs:18-38 (invert), :70-79 (rotation matrix),
// :123-366 (warp), warp/span.rs:36-81 (valid span).
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_invert_affine_transform(const float m[6], float out[6]) {
    const float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5];
    const float determinant = a * e - b * d;
    const float inv_determinant = (determinant != 0.0f) ? 1.0f / determinant : 0.0f;
    const float new_a = e * inv_determinant;
    const float new_b = -b * inv_determinant;
    const float new_d = -d * inv_determinant;
    const float new_e = a * inv_determinant;
    const float new_c = -(new_a * c + new_b * f);
    const float new_f = -(new_d * c + new_e * f);
    out[0] = new_a; out[1] = new_b; out[2] = new_c; out[3] = new_d; out[4] = new_e; out[5] = new_f;
}

KO_API void ko_get_rotation_matrix2d(float cx, float cy, float angle_deg, float scale, float out[6]) {
    const float PI_F = 3.14159265358979323846f;  // std::f32::consts::PI
    const float angle = angle_deg * PI_F / 180.0f;
    const float alpha = scale * cosf(angle);
    const float beta = scale * sinf(angle);
    const float tx = (1.0f - alpha) * cx - beta * cy;
    const float ty = beta * cx + (1.0f - alpha) * cy;
    out[0] = alpha; out[1] = beta; out[2] = tx; out[3] = -beta; out[4] = alpha; out[5] = ty;
}

static inline void constrain_span(float a, float b, bool ge, float eps, long long* lo, long long* hi) {  // span.rs:36-57
    if (std::fabs(a) < eps || a == 0.0f) {
        const bool feasible = ge ? (b >= 0.0f) : (b < 0.0f);
        if (!feasible) *hi = *lo;
        return;
    }
    const float k = -b / a;
    auto to_i64 = [](float v) -> long long {  // Rust `as i64` saturates; NaN -> 0
        if (std::isnan(v)) return 0;
        if (v >= 9.2233720368547758e18f) return INT64_MAX;
        if (v <= -9.2233720368547758e18f) return INT64_MIN;
        return (long long)v;
    };
    auto sat_add1 = [](long long v) -> long long { return v == INT64_MAX ? v : v + 1; };
    if (ge && a > 0.0f) *lo = std::max(*lo, to_i64(std::ceil(k)));
    else if (ge && !(a > 0.0f)) *hi = std::min(*hi, sat_add1(to_i64(std::floor(k))));
    else if (!ge && a > 0.0f) *hi = std::min(*hi, to_i64(std::ceil(k)));
    else *lo = std::max(*lo, sat_add1(to_i64(std::floor(k))));
}

static inline void affine_valid_span(float d0, float s00, float up0, float d1, float s01, float up1, size_t dst_w,
                                     float eps, size_t* out_lo, size_t* out_hi) {  // span.rs:62-81
    long long lo = 0, hi = (long long)dst_w;
    const float ds[2] = {d0, d1}, ss[2] = {s00, s01}, us[2] = {up0, up1};
    for (int i = 0; i < 2; ++i) {
        constrain_span(ds[i], ss[i], true, eps, &lo, &hi);
        constrain_span(ds[i], ss[i] - us[i], false, eps, &lo, &hi);
        if (lo >= hi) { *out_lo = 0; *out_hi = 0; return; }
    }
    lo = std::min(std::max(lo, 0ll), (long long)dst_w);
    hi = std::min(std::max(hi, 0ll), (long long)dst_w);
    if (lo >= hi) { *out_lo = 0; *out_hi = 0; }
    else { *out_lo = (size_t)lo; *out_hi = (size_t)hi; }
}

// Exposed for the span unit tests (warp/span.rs:95-138).
KO_API void ko_constrain_span(float a, float b, int ge, float eps, long long lo_in, long long hi_in, long long* lo,
                              long long* hi) {
    *lo = lo_in; *hi = hi_in;
    constrain_span(a, b, ge != 0, eps, lo, hi);
}
KO_API void ko_affine_valid_span(const float axes[6], size_t dst_w, float eps, size_t* lo, size_t* hi) {
    affine_valid_span(axes[0], axes[1], axes[2], axes[3], axes[4], axes[5], dst_w, eps, lo, hi);
}

KO_API int ko_warp_affine_f32(const float* src, size_t sw, size_t sh, float* dst, size_t dw, size_t dh, size_t C,
                              const float m[6], int mode) {
    if (mode < KO_NEAREST || mode > KO_LANCZOS) return -1;
    float mi[6];
    ko_invert_affine_transform(m, mi);
    const float dsx = mi[0], dsy = mi[3];
    const float src_w_f = (float)sw, src_h_f = (float)sh;
    std::vector<float> xstep_x(dw), xstep_y(dw);  // :185-186
    for (size_t x = 0; x < dw; ++x) { xstep_x[x] = dsx * (float)x; xstep_y[x] = dsy * (float)x; }
    auto in_bounds = [&](float sx0, float sy0, size_t x) -> bool {  // :201-215
        bool x_ok, y_ok;
        if (std::fabs(dsx) < 1e-6f) x_ok = sx0 >= 0.0f && sx0 < src_w_f;
        else { const float sx = dsx * (float)x + sx0; x_ok = sx >= 0.0f && sx < src_w_f; }
        if (std::fabs(dsy) < 1e-6f) y_ok = sy0 >= 0.0f && sy0 < src_h_f;
        else { const float sy = dsy * (float)x + sy0; y_ok = sy >= 0.0f && sy < src_h_f; }
        return x_ok && y_ok;
    };
    const size_t row_len = dw * C;
    const long long nchunks = (long long)((dh + 15) / 16);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ck = 0; ck < nchunks; ++ck) {
        const size_t y_end = std::min<size_t>(dh, (size_t)(ck + 1) * 16);
        for (size_t y = (size_t)ck * 16; y < y_end; ++y) {
            float* dst_row = dst + y * row_len;
            const float y_f = (float)y;
            const float sx0 = mi[1] * y_f + mi[2];
            const float sy0 = mi[4] * y_f + mi[5];
            size_t lo, hi;
            affine_valid_span(dsx, sx0, src_w_f, dsy, sy0, src_h_f, dw, 1e-6f, &lo, &hi);
            if (lo >= hi) { lo = 0; hi = 0; }
            else {  // refine_range :216-238
                while (lo < hi && !in_bounds(sx0, sy0, lo)) ++lo;
                while (lo > 0 && in_bounds(sx0, sy0, lo - 1)) --lo;
                while (hi > lo && !in_bounds(sx0, sy0, hi - 1)) --hi;
                while (hi < dw && in_bounds(sx0, sy0, hi)) ++hi;
                if (lo >= hi) { lo = 0; hi = 0; }
            }
            std::fill(dst_row, dst_row + lo * C, 0.0f);
            std::fill(dst_row + hi * C, dst_row + row_len, 0.0f);
            for (size_t x = lo; x < hi; ++x) {
                const float sx = xstep_x[x] + sx0;
                const float sy = xstep_y[x] + sy0;
                float* px = dst_row + x * C;
                if (mode == KO_BICUBIC || mode == KO_LANCZOS) {  // :325-361: per-pixel samplers on the unclamped coordinate
                    for (size_t k = 0; k < C; ++k) px[k] = sample_mode(mode, src, sh, sw, C, sx, sy, k);
                } else if (mode == KO_NEAREST) {  // :268-272
                    float rx = roundf(sx), ry = roundf(sy);
                    rx = std::min(std::max(rx, 0.0f), src_w_f - 1.0f);
                    ry = std::min(std::max(ry, 0.0f), src_h_f - 1.0f);
                    const size_t xi = (size_t)rx, yi = (size_t)ry;
                    for (size_t k = 0; k < C; ++k) px[k] = src[(yi * sw + xi) * C + k];
                } else {  // :292-317
                    const float sx_c = std::min(std::max(sx, 0.0f), src_w_f - 1.0f);
                    const float sy_c = std::min(std::max(sy, 0.0f), src_h_f - 1.0f);
                    const size_t x0 = (size_t)sx_c, y0 = (size_t)sy_c;
                    const size_t x1 = std::min(x0 + 1, sw - 1), y1 = std::min(y0 + 1, sh - 1);
                    const float fx = sx_c - (float)x0, fy = sy_c - (float)y0;
                    const float w00 = (1.0f - fy) * (1.0f - fx);
                    const float w10 = (1.0f - fy) * fx;
                    const float w01 = fy * (1.0f - fx);
                    const float w11 = fy * fx;
                    const size_t b00 = (y0 * sw + x0) * C, b10 = (y0 * sw + x1) * C;
                    const size_t b01 = (y1 * sw + x0) * C, b11 = (y1 * sw + x1) * C;
                    for (size_t k = 0; k < C; ++k)
                        px[k] = w00 * src[b00 + k] + w10 * src[b10 + k] + w01 * src[b01 + k] + w11 * src[b11 + k];
                }
            }
        }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// a5: warp_perspective — warp/perspective.rs:11-72 (det/adjugate/invert,
// transform_point), :115-165 (warp).  Out-of-bounds destination pixels are
// left untouched (CPU semantics); the CUDA twin writes 0, so parity tests
// zero-initialise dst (cuda/warp_perspective.rs:722-726).
// ─────────────────────────────────────────────────────────────────────────────
KO_API int ko_invert_homography(const float m[9], float inv[9]) {
    const float det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                      m[2] * (m[3] * m[7] - m[4] * m[6]);
    const float h8_sq = m[8] * m[8];
    const float FLT_EPS = 1.1920929e-07f;
    const float det_norm = (h8_sq > FLT_EPS) ? det / (h8_sq * std::fabs(m[8])) : det;
    if (std::fabs(det_norm) < 1e-10f) return -1;
    const float adj[9] = {
        m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
        m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
        m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]};
    const float inv_det = 1.0f / det;
    for (int i = 0; i < 9; ++i) inv[i] = adj[i] * inv_det;
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f)#1 — u8 warps: Q10 sampler (warp/common.rs:14-63, :80-181), affine Q16 span walk (warp/affine.rs:373-450,
// warp/kernels.rs:386-415), perspective row classification + direct coordinates (warp/perspective.rs:179-324,
// warp/kernels.rs:107-153).
// ─────────────────────────────────────────────────────────────────────────────
static inline int rust_f32_to_i32(float v) {  // `as i32`: saturating, NaN -> 0
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int)v;
}
static inline uint32_t rust_f32_to_u32(float v) {  // `as u32`: saturating, NaN -> 0
    if (std::isnan(v) || v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)v;
}
// warp/common.rs:80-181 (scalar form): taps at (xi, yi) with the +1 neighbour clamped to the last column / row.
// The reference reads without a bounds check (its callers guarantee xi, yi in range); an index that float rounding
// pushed outside is clamped here instead of read out of bounds.
static inline void sample_u8_q10(const uint8_t* src, int sw, int sh, size_t C, int xi, int yi, uint32_t fx, uint32_t fy, uint8_t* out) {
    xi = std::min(std::max(xi, 0), sw - 1); yi = std::min(std::max(yi, 0), sh - 1);
    const uint32_t fx1 = 1024u - fx, fy1 = 1024u - fy;
    const int xi1 = (xi + 1 < sw) ? xi + 1 : xi, yi1 = (yi + 1 < sh) ? yi + 1 : yi;
    const size_t stride = (size_t)sw * C;
    const size_t o00 = (size_t)yi * stride + (size_t)xi * C, o01 = (size_t)yi * stride + (size_t)xi1 * C;
    const size_t o10 = (size_t)yi1 * stride + (size_t)xi * C, o11 = (size_t)yi1 * stride + (size_t)xi1 * C;
    for (size_t ch = 0; ch < C; ++ch) {
        const uint32_t top = src[o00 + ch] * fx1 + src[o01 + ch] * fx;
        const uint32_t bot = src[o10 + ch] * fx1 + src[o11 + ch] * fx;
        out[ch] = (uint8_t)((top * fy1 + bot * fy + (1u << 19)) >> 20);
    }
}
// warp/common.rs:14-63: bounds-checked form — zeros for a non-finite or outside coordinate
static inline void sample_u8_checked(const uint8_t* src, int sw, int sh, size_t C, float xf, float yf, uint8_t* out) {
    if (!std::isfinite(xf) || !std::isfinite(yf)) { for (size_t c = 0; c < C; ++c) out[c] = 0; return; }
    const int xi = rust_f32_to_i32(std::floor(xf)), yi = rust_f32_to_i32(std::floor(yf));
    if (xi < 0 || xi >= sw || yi < 0 || yi >= sh) { for (size_t c = 0; c < C; ++c) out[c] = 0; return; }
    const uint32_t fx = rust_f32_to_u32((xf - (float)xi) * 1024.0f), fy = rust_f32_to_u32((yf - (float)yi) * 1024.0f);
    sample_u8_q10(src, sw, sh, C, xi, yi, fx, fy, out);
}

KO_API int ko_warp_affine_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst, size_t dw, size_t dh, size_t C, const float m[6]) {
    float mi[6];
    ko_invert_affine_transform(m, mi);
    const float dsx = mi[0], dsy = mi[3];
    const int dsx_q = rust_f32_to_i32(dsx * 65536.0f), dsy_q = rust_f32_to_i32(dsy * 65536.0f);
    for (size_t y = 0; y < dh; ++y) {
        uint8_t* drow = dst + y * dw * C;
        const float y_f = (float)y;
        const float sx0 = mi[1] * y_f + mi[2], sy0 = mi[4] * y_f + mi[5];
        size_t lo, hi;
        affine_valid_span(dsx, sx0, (float)sw, dsy, sy0, (float)sh, dw, 1e-12f, &lo, &hi);
        std::memset(drow, 0, lo * C);
        std::memset(drow + hi * C, 0, (dw - hi) * C);
        if (lo >= hi) continue;
        uint32_t sx_q = (uint32_t)rust_f32_to_i32((sx0 + dsx * (float)lo) * 65536.0f);
        uint32_t sy_q = (uint32_t)rust_f32_to_i32((sy0 + dsy * (float)lo) * 65536.0f);
        for (size_t x = lo; x < hi; ++x) {
            const int sxi = (int)sx_q, syi = (int)sy_q;
            sample_u8_q10(src, (int)sw, (int)sh, C, sxi >> 16, syi >> 16, ((uint32_t)(sxi & 0xFFFF)) >> 6, ((uint32_t)(syi & 0xFFFF)) >> 6,
                          drow + x * C);
            sx_q += (uint32_t)dsx_q;  // wrapping_add
            sy_q += (uint32_t)dsy_q;
        }
    }
    return 0;
}

KO_API int ko_warp_perspective_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst, size_t dw, size_t dh, size_t C, const float m[9]) {
    float inv[9];
    if (ko_invert_homography(m, inv) != 0) return -2;  // CannotComputeDeterminant
    const float src_w_f = (float)sw, src_h_f = (float)sh;
    for (size_t y = 0; y < dh; ++y) {
        uint8_t* drow = dst + y * dw * C;
        const float y_f = (float)y;
        float nx0 = inv[1] * y_f + inv[2], ny0 = inv[4] * y_f + inv[5], nd0 = inv[7] * y_f + inv[8];
        float dnx = inv[0], dny = inv[3], dnd = inv[6];
        auto coord = [&](size_t x, float* xf, float* yf) {   // warp/kernels.rs:107-122
            const float x_f = (float)x;
            const float nx = nx0 + dnx * x_f, ny = ny0 + dny * x_f, nd = nd0 + dnd * x_f;
            const float inv_nd = 1.0f / nd;
            *xf = nx * inv_nd; *yf = ny * inv_nd;
        };
        const float nd_end = nd0 + dnd * ((float)dw - 1.0f);
        const bool pos = nd0 > 1e-6f && nd_end > 1e-6f, neg = nd0 < -1e-6f && nd_end < -1e-6f;
        if (!(pos || neg)) {
            for (size_t x = 0; x < dw; ++x) { float xf, yf; coord(x, &xf, &yf); sample_u8_checked(src, (int)sw, (int)sh, C, xf, yf, drow + x * C); }
            continue;
        }
        if (neg) { nx0 = -nx0; ny0 = -ny0; nd0 = -nd0; dnx = -dnx; dny = -dny; dnd = -dnd; }
        long long lo = 0, hi = (long long)dw;
        constrain_span(dnx, nx0, true, 0.0f, &lo, &hi);
        constrain_span(dnx - src_w_f * dnd, nx0 - src_w_f * nd0, false, 0.0f, &lo, &hi);
        constrain_span(dny, ny0, true, 0.0f, &lo, &hi);
        constrain_span(dny - src_h_f * dnd, ny0 - src_h_f * nd0, false, 0.0f, &lo, &hi);
        size_t x_lo = (size_t)std::min(std::max(lo, 0ll), (long long)dw), x_hi = (size_t)std::min(std::max(hi, 0ll), (long long)dw);
        if (x_lo >= x_hi) { x_lo = 0; x_hi = 0; }
        std::memset(drow, 0, x_lo * C);
        std::memset(drow + x_hi * C, 0, (dw - x_hi) * C);
        // Inside the span the reference uses the unchecked sampler on [x_lo+1, x_hi-1) and the checked one on the two
        // margin columns; for an in-range coordinate both produce the same bytes, so every column is sampled checked
        // (this is also what the reference's own CUDA twin does, cuda/warp_perspective_u8.rs:140-171).
        for (size_t x = x_lo; x < x_hi; ++x) { float xf, yf; coord(x, &xf, &yf); sample_u8_checked(src, (int)sw, (int)sh, C, xf, yf, drow + x * C); }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f)#2 — remap (interpolation/remap.rs:43-128 f32, :157-296 u8).  Coordinates outside [0,w) x [0,h) (NaN included)
// produce 0: that is what the u8 path does explicitly (:255-264, :283-285) and what the f32 device twin does
// (cuda/remap.rs:80-85); the f32 CPU loop indexes unchecked there, so the defined behaviour is taken.
// ─────────────────────────────────────────────────────────────────────────────
KO_API int ko_remap_f32(const float* src, size_t sw, size_t sh, float* dst, size_t dw, size_t dh, size_t C, const float* map_x,
                        const float* map_y, int mode) {
    if (mode != KO_NEAREST && mode != KO_BILINEAR) return -1;
    for (size_t i = 0; i < dw * dh; ++i) {
        const float x = map_x[i], y = map_y[i];
        const bool in = x >= 0.0f && x < (float)sw && y >= 0.0f && y < (float)sh;
        for (size_t c = 0; c < C; ++c)
            dst[i * C + c] = !in ? 0.0f : (mode == KO_BILINEAR ? bilinear_interpolation(src, sh, sw, C, x, y, c)
                                                                : nearest_neighbor_interpolation(src, sh, sw, C, x, y, c));
    }
    return 0;
}

KO_API int ko_remap_u8(const uint8_t* src, size_t sw, size_t sh, uint8_t* dst, size_t dw, size_t dh, size_t C, const float* map_x,
                       const float* map_y, int mode) {
    if (mode != KO_NEAREST && mode != KO_BILINEAR) return -1;   // UnsupportedInterpolation
    for (size_t i = 0; i < dw * dh; ++i) {
        const float xf = map_x[i], yf = map_y[i];
        uint8_t* d = dst + i * C;
        if (mode == KO_BILINEAR) sample_u8_checked(src, (int)sw, (int)sh, C, xf, yf, d);   // :249-268
        else if (!(xf >= 0.0f && xf < (float)sw && yf >= 0.0f && yf < (float)sh)) { for (size_t c = 0; c < C; ++c) d[c] = 0; }
        else {   // :283-291
            const int xi = std::min(std::max(rust_f32_to_i32(roundf(xf)), 0), (int)sw - 1);
            const int yi = std::min(std::max(rust_f32_to_i32(roundf(yf)), 0), (int)sh - 1);
            for (size_t c = 0; c < C; ++c) d[c] = src[((size_t)yi * sw + (size_t)xi) * C + c];
        }
    }
    return 0;
}

KO_API int ko_warp_perspective_f32(const float* src, size_t sw, size_t sh, float* dst, size_t dw, size_t dh,
                                   size_t C, const float m[9], int mode) {
    if (mode < KO_NEAREST || mode > KO_LANCZOS) return -1;
    float im[9];
    if (ko_invert_homography(m, im) != 0) return -2;  // CannotComputeDeterminant
    const long long nchunks = (long long)((dh + 15) / 16);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ck = 0; ck < nchunks; ++ck) {
        const size_t y_end = std::min<size_t>(dh, (size_t)(ck + 1) * 16);
        for (size_t r = (size_t)ck * 16; r < y_end; ++r) {
            for (size_t c = 0; c < dw; ++c) {
                const float x = (float)c, y = (float)r;
                const float w = im[6] * x + im[7] * y + im[8];
                const float xo = (im[0] * x + im[1] * y + im[2]) / w;
                const float yo = (im[3] * x + im[4] * y + im[5]) / w;
                if (xo >= 0.0f && xo < (float)sw && yo >= 0.0f && yo < (float)sh) {
                    float* px = dst + (r * dw + c) * C;
                    for (size_t k = 0; k < C; ++k) px[k] = sample_mode(mode, src, sh, sw, C, xo, yo, k);
                }
            }
        }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// a6/a7: filters — filter/kernels.rs:25-41 (gaussian taps), :55-72 (sobel
// taps); filter/separable_filter.rs:87-155 (engine: H into f32 temp, then V;
// ascending taps; acc += v*k unfused; OOB taps skipped); filter/ops.rs:116-163
// (gaussian_blur parameter resolution), :174-203 (sobel).
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_gaussian_kernel_1d(size_t ksize, float sigma, float* out) {
    const float mean = (float)(ksize - 1) / 2.0f;
    const float sigma_sq = sigma * sigma;
    for (size_t i = 0; i < ksize; ++i) {
        const float x = (float)i - mean;
        out[i] = expf(-(x * x) / (2.0f * sigma_sq));
    }
    float norm = 0.0f;  // iter().sum::<f32>() — sequential, starting from 0.0
    for (size_t i = 0; i < ksize; ++i) norm += out[i];
    for (size_t i = 0; i < ksize; ++i) out[i] /= norm;
}

KO_API int ko_sobel_kernel_1d(size_t ksize, float* kx, float* ky) {
    if (ksize == 3) {
        const float a[3] = {-1.0f, 0.0f, 1.0f}, b[3] = {1.0f, 2.0f, 1.0f};
        std::memcpy(kx, a, sizeof a); std::memcpy(ky, b, sizeof b);
        return 0;
    }
    if (ksize == 5) {
        const float a[5] = {-1.0f, -2.0f, 0.0f, 2.0f, 1.0f}, b[5] = {1.0f, 4.0f, 6.0f, 4.0f, 1.0f};
        std::memcpy(kx, a, sizeof a); std::memcpy(ky, b, sizeof b);
        return 0;
    }
    return -1;
}

// `threads`: 1 = faithful (the reference engine is single-threaded); >1 = row-parallel
// variant with identical arithmetic (reported separately in the CPU baseline).
static void separable_filter_impl(const float* src, float* dst, size_t rows, size_t cols, size_t C, const float* kx,
                                  size_t kxn, const float* ky, size_t kyn, bool mt) {
    std::vector<float> temp(rows * cols * C, 0.0f);
    const long long half_x = (long long)(kxn / 2), half_y = (long long)(kyn / 2);
#pragma omp parallel for schedule(static) if (mt)
    for (long long r = 0; r < (long long)rows; ++r) {
        const size_t row_offset = (size_t)r * cols * C;
        std::vector<float> acc(C);
        for (size_t c = 0; c < cols; ++c) {
            std::fill(acc.begin(), acc.end(), 0.0f);
            for (size_t t = 0; t < kxn; ++t) {
                const long long x = (long long)c + (long long)t - half_x;
                if (x >= 0 && x < (long long)cols) {
                    const size_t idx = row_offset + (size_t)x * C;
                    for (size_t ch = 0; ch < C; ++ch) acc[ch] += src[idx + ch] * kx[t];
                }
            }
            for (size_t ch = 0; ch < C; ++ch) temp[row_offset + c * C + ch] = acc[ch];
        }
    }
#pragma omp parallel for schedule(static) if (mt)
    for (long long r = 0; r < (long long)rows; ++r) {
        const size_t row_offset = (size_t)r * cols * C;
        std::vector<float> acc(C);
        for (size_t c = 0; c < cols; ++c) {
            std::fill(acc.begin(), acc.end(), 0.0f);
            for (size_t t = 0; t < kyn; ++t) {
                const long long y = r + (long long)t - half_y;
                if (y >= 0 && y < (long long)rows) {
                    const size_t idx = (size_t)y * cols * C + c * C;
                    for (size_t ch = 0; ch < C; ++ch) acc[ch] += temp[idx + ch] * ky[t];
                }
            }
            for (size_t ch = 0; ch < C; ++ch) dst[row_offset + c * C + ch] = acc[ch];
        }
    }
}

KO_API int ko_separable_filter_f32(const float* src, float* dst, size_t rows, size_t cols, size_t C, const float* kx,
                                   size_t kxn, const float* ky, size_t kyn, int mt) {
    if (kxn == 0 || kyn == 0) return -1;  // InvalidKernelLength
    separable_filter_impl(src, dst, rows, cols, C, kx, kxn, ky, kyn, mt != 0);
    return 0;
}

// Resolve gaussian_blur's (kernel_size, sigma) exactly as ops.rs:122-153.  Returns 0 and
// the resolved values, or -1 (InvalidSigmaValue).
KO_API int ko_gaussian_resolve(size_t kx_in, size_t ky_in, float sx_in, float sy_in, size_t* kx, size_t* ky,
                               float* sx, float* sy) {
    size_t kernel_x = kx_in, kernel_y = ky_in;
    float sigma_x = sx_in, sigma_y = sy_in;
    if (sigma_y <= 0.0f) sigma_y = sigma_x;
    auto auto_k = [](float s) -> size_t {
        const float v = 2.0f * roundf(4.0f * s) + 1.0f;
        size_t k = v <= 0.0f ? 0 : (size_t)v;
        return k | 1;
    };
    if (kernel_x == 0 && sigma_x > 0.0f) kernel_x = auto_k(sigma_x);
    if (kernel_y == 0 && sigma_y > 0.0f) kernel_y = auto_k(sigma_y);
    if (!(kernel_x > 0 && kernel_x % 2 == 1 && kernel_y > 0 && kernel_y % 2 == 1)) return -1;
    sigma_x = std::max(sigma_x, 0.0f);
    sigma_y = std::max(sigma_y, 0.0f);
    if (sigma_x == 0.0f) sigma_x = ((float)kernel_x - 1.0f) / 8.0f;
    if (sigma_y == 0.0f) sigma_y = ((float)kernel_y - 1.0f) / 8.0f;
    *kx = kernel_x; *ky = kernel_y; *sx = sigma_x; *sy = sigma_y;
    return 0;
}

KO_API int ko_gaussian_blur_f32(const float* src, float* dst, size_t rows, size_t cols, size_t C, size_t kx_in,
                                size_t ky_in, float sx_in, float sy_in, int mt) {
    size_t kxn, kyn;
    float sx, sy;
    if (ko_gaussian_resolve(kx_in, ky_in, sx_in, sy_in, &kxn, &kyn, &sx, &sy) != 0) return -1;
    std::vector<float> kx(kxn), ky(kyn);
    ko_gaussian_kernel_1d(kxn, sx, kx.data());
    ko_gaussian_kernel_1d(kyn, sy, ky.data());
    separable_filter_impl(src, dst, rows, cols, C, kx.data(), kxn, ky.data(), kyn, mt != 0);
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f)#4 — video ENCODE: color/yuv/kernels.rs:1223-1252 (Q8 BT.601 limited coefficients, encode_y / encode_uv),
// :1301-1322 (YUYV: luma per pixel, chroma of the rounded pair average), :1480-1516 + :1563-1576 (NV12: luma per pixel,
// chroma of the rounded 2x2 mean).
// ─────────────────────────────────────────────────────────────────────────────
static inline uint8_t enc_y(int r, int g, int b) { return (uint8_t)std::min(std::max(((66 * r + 129 * g + 25 * b + 128) >> 8) + 16, 0), 255); }
static inline void enc_uv(int r, int g, int b, uint8_t* u, uint8_t* v) {
    *u = (uint8_t)std::min(std::max(((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128, 0), 255);
    *v = (uint8_t)std::min(std::max(((112 * r - 94 * g - 18 * b + 128) >> 8) + 128, 0), 255);
}
KO_API int ko_yuyv_from_rgb_u8(const uint8_t* src, size_t w, size_t h, uint8_t* dst) {
    if (w % 2 != 0) return -1;
    for (size_t row = 0; row < h; ++row)
        for (size_t g = 0; g < w / 2; ++g) {
            const uint8_t* s = src + row * w * 3 + g * 6;
            uint8_t* d = dst + row * w * 2 + g * 4;
            d[0] = enc_y(s[0], s[1], s[2]); d[2] = enc_y(s[3], s[4], s[5]);
            enc_uv((s[0] + s[3] + 1) >> 1, (s[1] + s[4] + 1) >> 1, (s[2] + s[5] + 1) >> 1, &d[1], &d[3]);
        }
    return 0;
}
KO_API int ko_nv12_from_rgb_u8(const uint8_t* src, size_t w, size_t h, uint8_t* dst) {
    if (w % 2 != 0 || h % 2 != 0) return -1;
    uint8_t* yp = dst;
    uint8_t* uv = dst + w * h;
    for (size_t y = 0; y < h; ++y)
        for (size_t x = 0; x < w; ++x) { const uint8_t* s = src + (y * w + x) * 3; yp[y * w + x] = enc_y(s[0], s[1], s[2]); }
    for (size_t cy = 0; cy < h / 2; ++cy)
        for (size_t cx = 0; cx < w / 2; ++cx) {
            const uint8_t* t = src + ((2 * cy) * w + 2 * cx) * 3;
            const uint8_t* b = t + w * 3;
            enc_uv((t[0] + t[3] + b[0] + b[3] + 2) >> 2, (t[1] + t[4] + b[1] + b[4] + 2) >> 2, (t[2] + t[5] + b[2] + b[5] + 2) >> 2,
                   &uv[cy * w + 2 * cx], &uv[cy * w + 2 * cx + 1]);
        }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f)#1 — u8 blurs: filter/ops.rs:22-29 (path selection), :59-98 (box_blur_u8), :639-757 (gaussian_blur_u8),
// :759-770 (quantize_kernel_256), :773-851 + :852-1100 (general Q8 two-pass, replicate border, u8 intermediate),
// :1105-1285 ([1,2,1]/4 binomial path as nested rounding half-adds).
// ─────────────────────────────────────────────────────────────────────────────
// (k * 256 + 0.5) as u8 per tap (saturating), then the centre tap absorbs the rounding so the weights sum to 256
KO_API void ko_quantize_kernel_256(const float* k, size_t n, uint8_t* out) {
    uint32_t sum = 0;
    for (size_t i = 0; i < n; ++i) {
        const float v = k[i] * 256.0f + 0.5f;
        out[i] = std::isnan(v) || v <= 0.0f ? 0 : (v >= 255.0f ? 255 : (uint8_t)v);
        sum += out[i];
    }
    if (sum != 256) {
        int c = (int)out[n / 2] + (256 - (int)(uint16_t)sum);   // i16 arithmetic in the reference; sums stay far below 2^15
        out[n / 2] = (uint8_t)std::min(std::max(c, 0), 255);
    }
}

static inline size_t clampi(long long v, size_t n) { return (size_t)std::min<long long>(std::max<long long>(v, 0), (long long)n - 1); }

// General Q8 path: H pass (acc + 128) >> 8 to u8 with the row replicated left/right, V pass likewise over
// row-clamped H rows.
static void separable_blur_u8(const uint8_t* src, uint8_t* dst, size_t rows, size_t cols, size_t C, const uint8_t* kx, size_t kxn,
                              const uint8_t* ky, size_t kyn) {
    const size_t stride = cols * C, hx = kxn / 2, hy = kyn / 2;
    std::vector<uint8_t> h(rows * stride);
    for (size_t r = 0; r < rows; ++r)
        for (size_t x = 0; x < cols; ++x)
            for (size_t ch = 0; ch < C; ++ch) {
                uint32_t acc = 0;
                for (size_t k = 0; k < kxn; ++k) acc += (uint32_t)src[r * stride + clampi((long long)x + (long long)k - (long long)hx, cols) * C + ch] * kx[k];
                h[r * stride + x * C + ch] = (uint8_t)((acc + 128) >> 8);
            }
    for (size_t r = 0; r < rows; ++r)
        for (size_t j = 0; j < stride; ++j) {
            uint32_t acc = 0;
            for (size_t k = 0; k < kyn; ++k) acc += (uint32_t)h[clampi((long long)r + (long long)k - (long long)hy, rows) * stride + j] * ky[k];
            dst[r * stride + j] = (uint8_t)((acc + 128) >> 8);
        }
}

static inline uint32_t rhadd(uint32_t a, uint32_t b) { return (a + b + 1) >> 1; }
// [1,2,1]/4 as rhadd(rhadd(a,b), rhadd(b,d)), horizontally then vertically, neighbours replicated at the borders
static void binomial3_u8(const uint8_t* src, uint8_t* dst, size_t rows, size_t cols, size_t C) {
    const size_t stride = cols * C;
    std::vector<uint8_t> h(rows * stride);
    for (size_t r = 0; r < rows; ++r)
        for (size_t x = 0; x < cols; ++x)
            for (size_t ch = 0; ch < C; ++ch) {
                const uint32_t a = src[r * stride + clampi((long long)x - 1, cols) * C + ch], b = src[r * stride + x * C + ch],
                               d = src[r * stride + clampi((long long)x + 1, cols) * C + ch];
                h[r * stride + x * C + ch] = (uint8_t)rhadd(rhadd(a, b), rhadd(b, d));
            }
    for (size_t r = 0; r < rows; ++r)
        for (size_t j = 0; j < stride; ++j) {
            const uint32_t a = h[clampi((long long)r - 1, rows) * stride + j], b = h[r * stride + j], d = h[clampi((long long)r + 1, rows) * stride + j];
            dst[r * stride + j] = (uint8_t)rhadd(rhadd(a, b), rhadd(b, d));
        }
}

// returns 0, -1 (InvalidSigmaValue)
KO_API int ko_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, size_t rows, size_t cols, size_t C, size_t kx_in, size_t ky_in,
                               float sx_in, float sy_in) {
    size_t kxn, kyn;
    float sx, sy;
    if (ko_gaussian_resolve(kx_in, ky_in, sx_in, sy_in, &kxn, &kyn, &sx, &sy) != 0) return -1;
    if (kxn == 3 && kyn == 3 && sx >= 0.6f && sx <= 1.2f && sy >= 0.6f && sy <= 1.2f) { binomial3_u8(src, dst, rows, cols, C); return 0; }  // blur_u8_path
    std::vector<float> fx(kxn), fy(kyn);
    ko_gaussian_kernel_1d(kxn, sx, fx.data());
    ko_gaussian_kernel_1d(kyn, sy, fy.data());
    std::vector<uint8_t> ikx(kxn), iky(kyn);
    ko_quantize_kernel_256(fx.data(), kxn, ikx.data());
    ko_quantize_kernel_256(fy.data(), kyn, iky.data());
    separable_blur_u8(src, dst, rows, cols, C, ikx.data(), kxn, iky.data(), kyn);
    return 0;
}

// returns 0, -1 (InvalidSigmaValue: zero or even kernel size)
KO_API int ko_box_blur_u8(const uint8_t* src, uint8_t* dst, size_t rows, size_t cols, size_t C, size_t kx, size_t ky) {
    if (kx == 0 || ky == 0 || kx % 2 == 0 || ky % 2 == 0) return -1;
    std::vector<float> fx(kx, 1.0f / (float)kx), fy(ky, 1.0f / (float)ky);   // filter/kernels.rs:10-13
    std::vector<uint8_t> ikx(kx), iky(ky);
    ko_quantize_kernel_256(fx.data(), kx, ikx.data());
    ko_quantize_kernel_256(fy.data(), ky, iky.data());
    separable_blur_u8(src, dst, rows, cols, C, ikx.data(), kx, iky.data(), ky);
    return 0;
}

KO_API int ko_sobel_f32(const float* src, float* dst, size_t rows, size_t cols, size_t C, size_t ksize, int mt) {
    float kx[5], ky[5];
    if (ko_sobel_kernel_1d(ksize, kx, ky) != 0) return -1;
    const size_t n = rows * cols * C;
    std::vector<float> gx(n, 0.0f), gy(n, 0.0f);
    separable_filter_impl(src, gx.data(), rows, cols, C, kx, ksize, ky, ksize, mt != 0);  // :191-192
    separable_filter_impl(src, gy.data(), rows, cols, C, ky, ksize, kx, ksize, mt != 0);  // :194-195
#pragma omp parallel for schedule(static) if (mt != 0)
    for (long long i = 0; i < (long long)n; ++i) dst[i] = sqrtf(gx[i] * gx[i] + gy[i] * gy[i]);  // :197-200
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// a10: normalize — normalize.rs:56-87 (mean/std), :123-146 (find_min_max),
// :191-222 (min_max), :235-263 + :407-421 (normalize_rgb_u8; AVX2 leaf uses
// fmadd on the npixels&~7 bulk).
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_normalize_mean_std_f32(const float* src, float* dst, size_t npixels, size_t C, const float* mean,
                                      const float* stdv) {
#pragma omp parallel for schedule(static) if (npixels >= 65536)
    for (long long i = 0; i < (long long)npixels; ++i)
        for (size_t c = 0; c < C; ++c) dst[i * C + c] = (src[i * C + c] - mean[c]) / stdv[c];
}

KO_API int ko_find_min_max_f32(const float* src, size_t n, float* mn, float* mx) {
    if (n == 0) return -1;
    float lo = src[0], hi = src[0];
    for (size_t i = 0; i < n; ++i) {
        if (src[i] < lo) lo = src[i];
        if (src[i] > hi) hi = src[i];
    }
    *mn = lo; *mx = hi;
    return 0;
}

KO_API int ko_normalize_min_max_f32(const float* src, float* dst, size_t n, float mn, float mx) {
    float min_val, max_val;
    if (ko_find_min_max_f32(src, n, &min_val, &max_val) != 0) return -1;
#pragma omp parallel for schedule(static) if (n >= 65536)
    for (long long i = 0; i < (long long)n; ++i)
        dst[i] = (src[i] - min_val) * (mx - mn) / (max_val - min_val) + mn;
    return 0;
}

KO_API void ko_normalize_rgb_u8(const uint8_t* src, float* dst, size_t npixels, const float scale[3],
                                const float offset[3], int leaf) {
    size_t bulk = 0;
    if (leaf == KO_LEAF_X86_AVX2_FMA || leaf == KO_LEAF_AARCH64_NEON) bulk = npixels & ~(size_t)7;
#pragma omp parallel for schedule(static) if (npixels >= 1024 * 1024)
    for (long long i = 0; i < (long long)npixels; ++i)
        for (int c = 0; c < 3; ++c) {
            const float v = (float)src[i * 3 + c];
            dst[i * 3 + c] = ((size_t)i < bulk) ? fmaf(v, scale[c], offset[c]) : v * scale[c] + offset[c];
        }
}

// ─────────────────────────────────────────────────────────────────────────────
// a11: std_mean — core.rs:42-67.  f64 accumulation of integer-valued terms is
// exact below 2^53, so integer sums reproduce it.  Returns raw sums too.
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_std_mean_u8_c3(const uint8_t* src, size_t npixels, double stdv[3], double mean[3],
                              uint64_t sums[6]) {
    double sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    for (size_t i = 0; i < npixels; ++i)
        for (int c = 0; c < 3; ++c) {
            const double p = (double)src[i * 3 + c];
            sum[c] += p;
            sq[c] += p * p;  // powi(2)
        }
    const double n = (double)npixels;
    for (int c = 0; c < 3; ++c) {
        mean[c] = sum[c] / n;
        stdv[c] = std::sqrt(sq[c] / n - mean[c] * mean[c]);
        if (sums) { sums[c] = (uint64_t)sum[c]; sums[3 + c] = (uint64_t)sq[c]; }
    }
}

// ─────────────────────────────────────────────────────────────────────────────
// a12: camera preprocess.  The NV12/YUYV path has NO CPU implementation in the
// reference; the CUDA source string preprocess.rs:430-647 is the spec, restated
// here op for op (compiled there with fmad=false, IEEE div).  Affine::new
// preprocess.rs:349-370.
// ─────────────────────────────────────────────────────────────────────────────
struct ko_preprocess_desc {
    float scale_x, scale_y, pad_x, pad_y;
    int32_t src_w, src_h, src_pitch, src_bpp, fmt;  // fmt: 0 RGB-order, 1 BGR-order, 2 gray, 3 NV12, 4 YUYV
    int32_t dst_w, dst_h;
    float mean[3], inv_std[3];
    float pad_value;
    int32_t sampling;  // 0 nearest, 1 bilinear
};

enum { KO_RESIZE_LETTERBOX = 0, KO_RESIZE_STRETCH = 1 };

KO_API void ko_preprocess_affine(int mode, size_t sw, size_t sh, size_t dw, size_t dh, float out[4]) {
    if (mode == KO_RESIZE_LETTERBOX) {
        const float s = std::min((float)dw / (float)sw, (float)dh / (float)sh);
        out[0] = s; out[1] = s;
        out[2] = ((float)dw - (float)sw * s) * 0.5f;
        out[3] = ((float)dh - (float)sh * s) * 0.5f;
    } else {
        out[0] = (float)dw / (float)sw; out[1] = (float)dh / (float)sh; out[2] = 0.0f; out[3] = 0.0f;
    }
}

static inline void yuv_to_rgbf(int yv, int u, int v, float px[3]) {  // :501-508
    const int yy = std::max(yv - 16, 0) * 1220542;
    u -= 128; v -= 128;
    px[2] = (float)std::min(std::max((yy + 2116026 * u + (1 << 19)) >> 20, 0), 255);
    px[1] = (float)std::min(std::max((yy + (-409993) * u + (-852492) * v + (1 << 19)) >> 20, 0), 255);
    px[0] = (float)std::min(std::max((yy + 1673527 * v + (1 << 19)) >> 20, 0), 255);
}

static inline void fetch_px(const uint8_t* src, int x, int y, const ko_preprocess_desc& d, float px[3]) {  // :510-530
    if (d.fmt <= 1) {
        const uint8_t* p = src + (long long)y * d.src_pitch + x * d.src_bpp;
        if (d.fmt == 0) { px[0] = (float)p[0]; px[1] = (float)p[1]; px[2] = (float)p[2]; }
        else { px[0] = (float)p[2]; px[1] = (float)p[1]; px[2] = (float)p[0]; }
    } else if (d.fmt == 2) {
        const float v = (float)src[(long long)y * d.src_pitch + x];
        px[0] = v; px[1] = v; px[2] = v;
    } else if (d.fmt == 3) {
        const int yv = src[(long long)y * d.src_w + x];
        const uint8_t* uv = src + (long long)d.src_w * d.src_h + (long long)(y >> 1) * d.src_w + (x >> 1) * 2;
        yuv_to_rgbf(yv, uv[0], uv[1], px);
    } else {
        const uint8_t* grp = src + (long long)y * d.src_pitch + (x >> 1) * 4;
        const int yv = grp[(x & 1) ? 2 : 0];
        yuv_to_rgbf(yv, grp[1], grp[3], px);
    }
}

// f32 -> binary16 bits, RNE — preprocess.rs:461-484 (f2h).
KO_API uint16_t ko_f2h(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int exp = (int)((x >> 23) & 0xFFu) - 127 + 15;
    uint32_t man = x & 0x7FFFFFu;
    if (exp >= 31) {
        // Mirrors the reference exactly: any nonzero mantissa sets the quiet bit, also for a
        // FINITE f32 that overflows binary16 (the reference's own rule, preprocess.rs:467-471).
        const uint32_t nan_bit = (man != 0u) ? 0x0200u : 0u;
        return (uint16_t)(sign | 0x7C00u | nan_bit);
    }
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        const uint32_t shift = (uint32_t)(14 - exp);
        uint16_t h = (uint16_t)(sign | (man >> shift));
        const uint32_t rem = man & ((1u << shift) - 1u);
        const uint32_t mid = 1u << (shift - 1u);
        if (rem > mid || (rem == mid && (h & 1u))) h++;
        return h;
    }
    uint16_t h = (uint16_t)(sign | ((uint32_t)exp << 10) | (man >> 13));
    const uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return h;
}

// One frame.  out_f16 != 0 -> dst is uint16_t (binary16 bits).  If `touched` is non-null
// (src-buffer-sized byte mask) every source byte addressed by a tap is marked — used to
// count the algorithmic bytes of SURVEY §8(d) exactly.
KO_API int ko_preprocess_frame(const uint8_t* src, void* dst, const ko_preprocess_desc* dp, int out_f16,
                               uint8_t* touched) {
    const ko_preprocess_desc d = *dp;
    if (d.sampling != KO_NEAREST && d.sampling != KO_BILINEAR && d.sampling != KO_LANCZOS) return -1;
    const int pixels = d.dst_w * d.dst_h;
    // preprocess.rs:481-488 (kernel source): 1-D Lanczos-3 weight with libm sinf — the ONE place on this path where
    // the reference calls a transcendental; host sinf and the CUDA math library may differ in the last ulp, so the
    // Lanczos preprocess is checked within 1e-4 against this restatement and bit-for-bit against the reference's own
    // kernel on the GPU (tests/test_ref_gpu_kernels.py).
    auto lanczos_w = [](float dd) -> float {
        const float ad = fabsf(dd);
        if (ad < 1e-6f) return 1.0f;
        if (ad >= 3.0f) return 0.0f;
        const float pd = 3.14159265358979f * dd;
        return 3.0f * sinf(pd) * sinf(pd / 3.0f) / (pd * pd);
    };
    float* dst32 = (float*)dst;
    uint16_t* dst16 = (uint16_t*)dst;
    auto mark = [&](int x, int y) {
        if (!touched) return;
        if (d.fmt <= 1) { for (int c = 0; c < 3; ++c) touched[(long long)y * d.src_pitch + x * d.src_bpp + c] = 1; }
        else if (d.fmt == 2) touched[(long long)y * d.src_pitch + x] = 1;
        else if (d.fmt == 3) {
            touched[(long long)y * d.src_w + x] = 1;
            const long long o = (long long)d.src_w * d.src_h + (long long)(y >> 1) * d.src_w + (x >> 1) * 2;
            touched[o] = 1; touched[o + 1] = 1;
        } else {
            const long long o = (long long)y * d.src_pitch + (x >> 1) * 4;
            touched[o + ((x & 1) ? 2 : 0)] = 1; touched[o + 1] = 1; touched[o + 3] = 1;
        }
    };
#pragma omp parallel for schedule(static) if (pixels >= 65536 && !touched)
    for (int i = 0; i < pixels; ++i) {
        const int ox = i % d.dst_w, oy = i / d.dst_w;
        const float sx = ((float)ox - d.pad_x) / d.scale_x;  // plan_pixel :437-448
        const float sy = ((float)oy - d.pad_y) / d.scale_y;
        const bool inside = !(sx < 0.0f || sy < 0.0f || sx >= (float)d.src_w || sy >= (float)d.src_h);
        float px[3];
        if (inside) {
            if (d.sampling == KO_BILINEAR) {  // :534-554
                int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
                const float ax = sx - (float)x0, ay = sy - (float)y0;
                const int x1 = std::min(x0 + 1, d.src_w - 1), y1 = std::min(y0 + 1, d.src_h - 1);
                x0 = std::max(x0, 0); y0 = std::max(y0, 0);
                float t00[3], t10[3], t01[3], t11[3];
                fetch_px(src, x0, y0, d, t00); fetch_px(src, x1, y0, d, t10);
                fetch_px(src, x0, y1, d, t01); fetch_px(src, x1, y1, d, t11);
                mark(x0, y0); mark(x1, y0); mark(x0, y1); mark(x1, y1);
                for (int c = 0; c < 3; ++c) {
                    const float top = t00[c] + (t10[c] - t00[c]) * ax;
                    const float bot = t01[c] + (t11[c] - t01[c]) * ax;
                    px[c] = top + (bot - top) * ay;
                }
            } else if (d.sampling == KO_LANCZOS) {  // :565-590 sample_lanczos
                const int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
                float acc[3] = {0.0f, 0.0f, 0.0f};
                float wsum = 0.0f;
                for (int j = -2; j <= 3; ++j) {
                    const int yj = y0 + j;
                    const float wyv = lanczos_w(sy - (float)yj);
                    const int yc = std::min(std::max(yj, 0), d.src_h - 1);
                    for (int ii = -2; ii <= 3; ++ii) {
                        const int xi = x0 + ii;
                        const float w = wyv * lanczos_w(sx - (float)xi);
                        const int xc = std::min(std::max(xi, 0), d.src_w - 1);
                        float t[3];
                        fetch_px(src, xc, yc, d, t);
                        mark(xc, yc);
                        for (int c = 0; c < 3; ++c) acc[c] += w * t[c];
                        wsum += w;
                    }
                }
                px[0] = acc[0] / wsum; px[1] = acc[1] / wsum; px[2] = acc[2] / wsum;
            } else {  // :556-563
                const int xn = std::min(std::max((int)roundf(sx), 0), d.src_w - 1);
                const int yn = std::min(std::max((int)roundf(sy), 0), d.src_h - 1);
                fetch_px(src, xn, yn, d, px);
                mark(xn, yn);
            }
        } else {
            px[0] = px[1] = px[2] = d.pad_value;
        }
        const float o0 = (px[0] / 255.0f - d.mean[0]) * d.inv_std[0];  // BODY :610-612
        const float o1 = (px[1] / 255.0f - d.mean[1]) * d.inv_std[1];
        const float o2 = (px[2] / 255.0f - d.mean[2]) * d.inv_std[2];
        if (out_f16) { dst16[i] = ko_f2h(o0); dst16[pixels + i] = ko_f2h(o1); dst16[2 * pixels + i] = ko_f2h(o2); }
        else { dst32[i] = o0; dst32[pixels + i] = o1; dst32[2 * pixels + i] = o2; }
    }
    return 0;
}

// Preprocessor::run_cpu for C==3 RGB + bilinear (preprocess.rs:933-1020): the host path —
// fused bilinear into the (integer) content box, pad fill, row placement.
KO_API int ko_preprocess_cpu_rgb_bilinear(const uint8_t* src, size_t sw, size_t sh, float* dst, size_t dw, size_t dh,
                                          int mode, const float mean[3], const float inv_std[3], float pad_value,
                                          int leaf) {
    float a[4];
    ko_preprocess_affine(mode, sw, sh, dw, dh, a);
    float scale[3], bias[3], pad[3];
    for (int c = 0; c < 3; ++c) {
        scale[c] = inv_std[c] / 255.0f;
        bias[c] = -mean[c] * inv_std[c];
        pad[c] = pad_value * scale[c] + bias[c];
    }
    size_t cw = dw, ch = dh;
    if (mode == KO_RESIZE_LETTERBOX) {
        cw = std::min(std::max((size_t)roundf((float)sw * a[0]), (size_t)1), dw);
        ch = std::min(std::max((size_t)roundf((float)sh * a[1]), (size_t)1), dh);
    }
    const size_t px0 = (dw - cw) / 2, py0 = (dh - ch) / 2;
    const bool padded = px0 != 0 || py0 != 0 || cw != dw || ch != dh;
    if (!padded) return ko_resize_normalize_u8_to_f32_chw_bilinear(src, sw, sh, dst, dw, dh, scale, bias, leaf);
    std::vector<float> content(3 * ch * cw);
    ko_resize_normalize_u8_to_f32_chw_bilinear(src, sw, sh, content.data(), cw, ch, scale, bias, leaf);
    const size_t pixels = dw * dh;
    for (int c = 0; c < 3; ++c) {
        float* plane = dst + c * pixels;
        std::fill(plane, plane + pixels, pad[c]);
        for (size_t y = 0; y < ch; ++y)
            std::memcpy(plane + (py0 + y) * dw + px0, content.data() + c * ch * cw + y * cw, cw * sizeof(float));
    }
    return 0;
}

// Count distinct source elements addressed by ≥1 tap for the generic bilinear samplers —
// used only to state algorithmic bytes.  kind: 0 = resize a1 (half-pixel,
// clamped), 1 = fused a2.
KO_API size_t ko_count_touched_resize(size_t sw, size_t sh, size_t dw, size_t dh, int kind) {
    std::vector<uint8_t> tx(sw, 0), ty(sh, 0);
    auto axis = [&](size_t s, size_t d, std::vector<uint8_t>& t) {
        for (size_t i = 0; i < d; ++i) {
            size_t i0, i1;
            if (kind == 0) {
                const float a = (float)s / (float)d, b = 0.5f * a - 0.5f;
                float v = a * (float)i + b;
                v = std::min(std::max(v, 0.0f), (float)(s - 1));
                i0 = (size_t)v; i1 = std::min(i0 + 1, s - 1);
            } else {
                float f = ((float)i + 0.5f) * ((float)s / (float)d) - 0.5f;
                f = std::max(f, 0.0f);
                i0 = std::min((size_t)f, s - 1); i1 = std::min(i0 + 1, s - 1);
            }
            t[i0] = 1; t[i1] = 1;
        }
    };
    axis(sw, dw, tx); axis(sh, dh, ty);
    size_t nx = 0, ny = 0;
    for (auto v : tx) nx += v;
    for (auto v : ty) ny += v;
    return nx * ny;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f) #4: Gaussian pyramids — pyramid.rs:22-250 (pyrup_f32: polyphase [1,4,6,4,1]/8 with its own border
// rule), :252-427 (reflect_101, pyrdown_f32: 25 taps `sum += v * (ky*kx)` unfused, BORDER_REFLECT_101),
// :469-650 (pyrdown_u8: u16 horizontal sums, (sum + 128) >> 8), :656-840 (pyrup_u8: (p + 6c + n + 4) >> 3 /
// (c + n + 1) >> 1 with reflect_101 neighbours).
// ─────────────────────────────────────────────────────────────────────────────
static inline int reflect_101(int p, int len) {  // pyramid.rs:252-269
    if (len == 1) return 0;
    if (p < 0) p = -p;
    const int period = 2 * (len - 1);
    p %= period;
    if (p >= len) p = period - p;
    return p;
}

KO_API int ko_pyrdown_f32(const float* src, size_t sw, size_t sh, size_t C, float* dst) {
    const size_t dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    static const float k1[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    float kw[25];
    for (int y = 0, i = 0; y < 5; ++y) for (int x = 0; x < 5; ++x) kw[i++] = k1[y] * k1[x];
    // reflect_101 is the identity on interior coordinates, so the reference's centre / border split is one loop here
    for (size_t dy = 0; dy < dh; ++dy)
        for (size_t dx = 0; dx < dw; ++dx)
            for (size_t c = 0; c < C; ++c) {
                float sum = 0.0f;
                int k = 0;
                for (int ky = 0; ky < 5; ++ky) {
                    const size_t sy = (size_t)reflect_101((int)(dy * 2) + ky - 2, (int)sh);
                    for (int kx = 0; kx < 5; ++kx) {
                        const size_t sx = (size_t)reflect_101((int)(dx * 2) + kx - 2, (int)sw);
                        sum += src[(sy * sw + sx) * C + c] * kw[k++];
                    }
                }
                dst[(dy * dw + dx) * C + c] = sum;
            }
    return 0;
}

KO_API int ko_pyrup_f32(const float* src, size_t sw, size_t sh, size_t C, float* dst) {
    const float SCALE_EVEN = 0.125f, SCALE_ODD = 0.5f, W_CENTER = 6.0f, W_NEIGHBOR = 1.0f, W_BORDER = 7.0f;
    const size_t dw = sw * 2, stride = dw * C;
    std::vector<float> buf(dw * sh * C);
    for (size_t y = 0; y < sh; ++y) {  // horizontal pass, pyramid.rs:22-90
        const float* s = src + y * sw * C;
        float* d = buf.data() + y * stride;
        if (sw == 1) {
            for (size_t k = 0; k < C; ++k) { d[k] = s[k]; if (k + C < stride) d[k + C] = s[k]; }
            continue;
        }
        for (size_t k = 0; k < C; ++k) {
            const float l = s[k], r = s[C + k];
            d[k] = (W_CENTER * l + 2.0f * r) * SCALE_EVEN;
            d[C + k] = (l + r) * SCALE_ODD;
        }
        for (size_t x = 1; x + 1 < sw; ++x)
            for (size_t k = 0; k < C; ++k) {
                const float p = s[(x - 1) * C + k], c = s[x * C + k], n = s[(x + 1) * C + k];
                d[2 * x * C + k] = (W_NEIGHBOR * p + W_CENTER * c + W_NEIGHBOR * n) * SCALE_EVEN;
                d[2 * x * C + C + k] = (c + n) * SCALE_ODD;
            }
        const size_t lx = sw - 1;
        for (size_t k = 0; k < C; ++k) {
            const float p = s[(lx - 1) * C + k], c = s[lx * C + k];
            d[2 * lx * C + k] = (W_NEIGHBOR * p + W_BORDER * c) * SCALE_EVEN;
            d[2 * lx * C + C + k] = c;
        }
    }
    for (size_t y = 0; y < sh; ++y) {  // vertical pass, pyramid.rs:98-160
        size_t rt, rc, rb;
        if (sh == 1) { rt = rc = rb = 0; }
        else if (y == 0) { rt = 0; rc = 0; rb = 1; }
        else if (y == sh - 1) { rt = sh - 2; rc = sh - 1; rb = sh - 1; }
        else { rt = y - 1; rc = y; rb = y + 1; }
        const float *t = buf.data() + rt * stride, *c = buf.data() + rc * stride, *b = buf.data() + rb * stride;
        float* even = dst + (2 * y) * stride;
        float* odd = even + stride;
        for (size_t i = 0; i < stride; ++i) {
            if (y == 0) {
                even[i] = (W_CENTER * c[i] + 2.0f * b[i]) * SCALE_EVEN;
                odd[i] = (c[i] + b[i]) * SCALE_ODD;
            } else if (y == sh - 1) {
                even[i] = (W_NEIGHBOR * t[i] + W_BORDER * c[i]) * SCALE_EVEN;
                odd[i] = c[i];
            } else {
                even[i] = (W_NEIGHBOR * t[i] + W_CENTER * c[i] + W_NEIGHBOR * b[i]) * SCALE_EVEN;
                odd[i] = (c[i] + b[i]) * SCALE_ODD;
            }
        }
    }
    return 0;
}

KO_API int ko_pyrdown_u8(const uint8_t* src, size_t sw, size_t sh, size_t C, uint8_t* dst) {
    const size_t dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    std::vector<uint16_t> buf(dw * sh * C);
    for (size_t y = 0; y < sh; ++y)
        for (size_t dx = 0; dx < dw; ++dx) {
            size_t ix[5];
            for (int t = 0; t < 5; ++t) ix[t] = (size_t)reflect_101((int)(dx * 2) + t - 2, (int)sw) * C;
            for (size_t k = 0; k < C; ++k) {
                const uint8_t* r = src + y * sw * C + k;
                buf[(y * dw + dx) * C + k] = (uint16_t)(r[ix[0]] + 4 * r[ix[1]] + 6 * r[ix[2]] + 4 * r[ix[3]] + r[ix[4]]);
            }
        }
    const size_t stride = dw * C;
    for (size_t dy = 0; dy < dh; ++dy) {
        size_t off[5];
        for (int t = 0; t < 5; ++t) off[t] = (size_t)reflect_101((int)(dy * 2) + t - 2, (int)sh) * stride;
        for (size_t i = 0; i < stride; ++i) {
            const uint32_t sum = (uint32_t)buf[off[0] + i] + 4u * buf[off[1] + i] + 6u * buf[off[2] + i] + 4u * buf[off[3] + i] + buf[off[4] + i];
            dst[dy * stride + i] = (uint8_t)std::min<uint32_t>((sum + 128u) >> 8, 255u);
        }
    }
    return 0;
}

KO_API int ko_pyrup_u8(const uint8_t* src, size_t sw, size_t sh, size_t C, uint8_t* dst) {
    const size_t dw = sw * 2, stride = dw * C;
    std::vector<uint8_t> buf(dw * sh * C);
    for (size_t y = 0; y < sh; ++y)
        for (size_t x = 0; x < sw; ++x) {
            const size_t ip = (size_t)reflect_101((int)x - 1, (int)sw) * C, in = (size_t)reflect_101((int)x + 1, (int)sw) * C;
            for (size_t k = 0; k < C; ++k) {
                const uint8_t* r = src + y * sw * C + k;
                const uint16_t c = r[x * C], p = r[ip], n = r[in];
                buf[y * stride + 2 * x * C + k] = (uint8_t)((p + 6 * c + n + 4) >> 3);
                if ((2 * x + 1) * C < stride) buf[y * stride + (2 * x + 1) * C + k] = (uint8_t)((c + n + 1) >> 1);
            }
        }
    for (size_t y = 0; y < sh; ++y) {
        const uint8_t* p = buf.data() + (size_t)reflect_101((int)y - 1, (int)sh) * stride;
        const uint8_t* c = buf.data() + y * stride;
        const uint8_t* n = buf.data() + (size_t)reflect_101((int)y + 1, (int)sh) * stride;
        uint8_t* even = dst + 2 * y * stride;
        uint8_t* odd = even + stride;
        for (size_t i = 0; i < stride; ++i) {
            even[i] = (uint8_t)(((uint16_t)p[i] + 6 * (uint16_t)c[i] + n[i] + 4) >> 3);
            odd[i] = (uint8_t)(((uint16_t)c[i] + n[i] + 1) >> 1);
        }
    }
    return 0;
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f) #2: undistort maps — calibration/distortion.rs:60-103 (distort_point_polynomial, f64, unfused),
// :135-150 (generate_correction_map_polynomial: per destination pixel, cast to f32).
// intr = {fx, fy, cx, cy}; dist = {k1..k6, p1, p2}.
// ─────────────────────────────────────────────────────────────────────────────
KO_API void ko_distort_point_polynomial(double x, double y, const double intr[4], const double dist[8], double out[2]) {
    const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
    const double k1 = dist[0], k2 = dist[1], k3 = dist[2], k4 = dist[3], k5 = dist[4], k6 = dist[5], p1 = dist[6], p2 = dist[7];
    x = (x - cx) / fx;
    y = (y - cy) / fy;
    const double r2 = x * x + y * y;
    const double r4 = r2 * r2;
    const double r6 = r4 * r2;
    const double kr = (1.0 + k1 * r2 + k2 * r4 + k3 * r6) / (1.0 + k4 * r2 + k5 * r4 + k6 * r6);
    const double x_2 = 2.0 * x, y_2 = 2.0 * y;
    const double xy_2 = x_2 * y;
    const double xd = x * kr + xy_2 * p1 + p2 * (r2 + x_2 * x);
    const double yd = y * kr + p1 * (r2 + y_2 * y) + xy_2 * p2;
    out[0] = fx * xd + cx;
    out[1] = fy * yd + cy;
}

KO_API void ko_generate_correction_map_polynomial(const double intr[4], const double dist[8], size_t w, size_t h, float* map_x, float* map_y) {
    for (size_t y = 0; y < h; ++y)
        for (size_t x = 0; x < w; ++x) {
            double o[2];
            ko_distort_point_polynomial((double)x, (double)y, intr, dist, o);
            map_x[y * w + x] = (float)o[0];
            map_y[y * w + x] = (float)o[1];
        }
}

// ─────────────────────────────────────────────────────────────────────────────
// §8(f) #3: the cuda/fusion.rs stage vocabulary — ReadU8RgbBilinear (:520-585: half-pixel `a*d + b`, max(0),
// weights first, four-term sum), Normalize (:592-620: v*s + b), RgbToGray (:624-642: 0.299x + 0.587y + 0.114z
// replicated to all three lanes), WriteChwF32 / WriteC1F32 (:645-690).  Restates the generated kernel body (and the
// tests' own `cpu_reference`, :705-760).  maps: 0 none, 1 Normalize, 2 RgbToGray, 3 Normalize->RgbToGray,
// 4 RgbToGray->Normalize.  sink: 0 CHW (3 planes), 1 C1 (the .x lane).
// ─────────────────────────────────────────────────────────────────────────────
KO_API int ko_fused_pipeline_u8(const uint8_t* src, size_t sw, size_t sh, size_t dw, size_t dh, int maps, const float scale[3],
                                const float bias[3], int sink, float* dst) {
    if (maps < 0 || maps > 4 || sink < 0 || sink > 1) return -1;
    const float ax = (float)sw / (float)dw, ay = (float)sh / (float)dh;
    const float bx = 0.5f * ax - 0.5f, by = 0.5f * ay - 0.5f;
    const size_t plane = dw * dh;
    for (size_t y = 0; y < dh; ++y)
        for (size_t x = 0; x < dw; ++x) {
            const float sxf = std::max(ax * (float)x + bx, 0.0f), syf = std::max(ay * (float)y + by, 0.0f);
            const size_t sx0 = std::min((size_t)sxf, sw - 1), sy0 = std::min((size_t)syf, sh - 1);
            const size_t sx1 = std::min(sx0 + 1, sw - 1), sy1 = std::min(sy0 + 1, sh - 1);
            const float wx = sxf - (float)sx0, wy = syf - (float)sy0;
            const float w00 = (1.0f - wy) * (1.0f - wx), w01 = (1.0f - wy) * wx, w10 = wy * (1.0f - wx), w11 = wy * wx;
            float v[3];
            for (int c = 0; c < 3; ++c) {
                auto p = [&](size_t yy, size_t xx) { return (float)src[(yy * sw + xx) * 3 + c]; };
                v[c] = w00 * p(sy0, sx0) + w01 * p(sy0, sx1) + w10 * p(sy1, sx0) + w11 * p(sy1, sx1);
            }
            auto norm = [&]() { for (int c = 0; c < 3; ++c) v[c] = v[c] * scale[c] + bias[c]; };
            auto gray = [&]() { const float g = 0.299f * v[0] + 0.587f * v[1] + 0.114f * v[2]; v[0] = v[1] = v[2] = g; };
            if (maps == 1) norm();
            else if (maps == 2) gray();
            else if (maps == 3) { norm(); gray(); }
            else if (maps == 4) { gray(); norm(); }
            const size_t di = y * dw + x;
            if (sink == 0) { dst[di] = v[0]; dst[di + plane] = v[1]; dst[di + 2 * plane] = v[2]; }
            else dst[di] = v[0];
        }
    return 0;
}
