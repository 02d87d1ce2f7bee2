"""Generates tests/golden/*.npz.  Run HERE (the build container), not on the GPU box:
it reads /root/reference (data fixture dog-rgb8.png) and uses cv2, neither of which the
tests may depend on at run time.

    python tests/golden/make_fixtures.py

Fixtures:
  nv12_cv2.npz    — NV12 frames + cv2.cvtColor(COLOR_YUV2RGB_NV12) outputs.  The reference states
                    its Q20 BT.601-limited decode is bit-identical to cv2; this pins
                    the oracle's a9 path against an implementation that is not ours.
  yuyv_cv2.npz    — same for YUYV (COLOR_YUV2RGB_YUYV).
  dog_cfg1.npz    — BASELINE.json configs[0]: dog-rgb8.png (258x195, lossless) → f32/255 →
                    gray_from_rgb → resize 128x128 bilinear, outputs produced by the oracle
                    (scalar leaf).  Pins config 1's input without needing the reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    import cv2

    from oracle import oracle as o

    rng = np.random.default_rng(20260922)
    # NV12: include the reference's own generators (yuv/kernels.rs:2078-2081, :2112-2122)
    cases = {}
    for name, (w, h) in {"a": (4, 4), "b": (64, 6), "c": (70, 4), "d": (128, 96)}.items():
        if name == "a":
            y = np.array([(v * 9 + 16) & 0xFF for v in range(w * h)], np.uint8)
            uv = np.array([(v * 5 + 100) & 0xFF for v in range(w * h // 2)], np.uint8)
        elif name in ("b", "c"):
            y = np.array([(i * 7 + 16) % 240 for i in range(w * h)], np.uint8)
            uv = np.array([(i * 5 + 90) % 250 for i in range(w * h // 2)], np.uint8)
        else:
            y = rng.integers(0, 256, w * h, dtype=np.uint8)
            uv = rng.integers(0, 256, w * h // 2, dtype=np.uint8)
        raw = np.concatenate([y, uv])
        rgb = cv2.cvtColor(raw.reshape(h * 3 // 2, w), cv2.COLOR_YUV2RGB_NV12)
        cases[f"{name}_raw"] = raw
        cases[f"{name}_rgb"] = rgb
        cases[f"{name}_wh"] = np.array([w, h])
    np.savez_compressed(os.path.join(HERE, "nv12_cv2.npz"), **cases)

    cases = {}
    for name, (w, h) in {"a": (2, 1), "b": (64, 5), "c": (130, 7)}.items():
        raw = rng.integers(0, 256, w * h * 2, dtype=np.uint8)
        if name == "a":
            raw = np.array([16, 128, 16, 128], np.uint8)  # packed422_known_gray yuv/kernels.rs:2068
        rgb = cv2.cvtColor(raw.reshape(h, w, 2), cv2.COLOR_YUV2RGB_YUYV)
        cases[f"{name}_raw"] = raw
        cases[f"{name}_rgb"] = rgb
        cases[f"{name}_wh"] = np.array([w, h])
    np.savez_compressed(os.path.join(HERE, "yuyv_cv2.npz"), **cases)

    png = "/root/reference/tests/data/dog-rgb8.png"
    bgr = cv2.imread(png, cv2.IMREAD_COLOR)
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    assert rgb.shape == (195, 258, 3), rgb.shape
    f = (rgb.astype(np.float32) * np.float32(1.0 / 255.0)).astype(np.float32)  # cast_and_scale(1/255)
    gray = o.gray_from_rgb_f32(f, o.LEAF_SCALAR)
    small = o.resize_f32(gray, 128, 128, o.BILINEAR)
    np.savez_compressed(os.path.join(HERE, "dog_cfg1.npz"), rgb=rgb, gray=gray, resized=small)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
