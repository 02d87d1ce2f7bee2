"""Run ONE hot-path op a few times (for ncu captures): python tools/run_op.py <op> [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kornia_rs_b200 as kb

op = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
MEAN, STD = kb.IMAGENET_MEAN, kb.IMAGENET_STD


def nv12_frames(n, w=1920, h=1080):
    raw = torch.randint(0, 256, (n, w * h * 3 // 2), dtype=torch.uint8, device=dev, generator=g)
    return raw, [raw[i] for i in range(n)]


if op == "cfg3a":
    raw, frames = nv12_frames(64)
    pre = kb.Preprocessor.builder().source_format(kb.SourceFormat.Nv12).mode(kb.ResizeMode.Stretch).normalize(kb.Normalize.imagenet()).build_cuda()
    dst = torch.empty((64, 3, 1080, 1920), dtype=torch.float32, device=dev)
    fn = lambda: pre.run_raw_batch(frames, 1920, 1080, dst)
elif op == "cfg3b":
    raw, frames = nv12_frames(64)
    pre = kb.Preprocessor.builder().source_format(kb.SourceFormat.Nv12).mode(kb.ResizeMode.Letterbox).normalize(kb.Normalize.imagenet()).build_cuda()
    dst = torch.empty((64, 3, 640, 640), dtype=torch.float32, device=dev)
    fn = lambda: pre.run_raw_batch(frames, 1920, 1080, dst)
elif op in ("gauss", "sobel", "warp", "affine", "resize_f32", "gray", "normalize"):
    n, w, h = 8, 3840, 2160
    src = kb.Image(torch.rand((n, h, w, 3), dtype=torch.float32, device=dev, generator=g))
    dst = kb.Image.zeros_cuda(kb.ImageSize(w, h), 3, torch.float32, dev, batch=n)
    if op == "gauss":
        fn = lambda: kb.imgproc.gaussian_blur(src, dst, (5, 5), (1.5, 1.5))
    elif op == "sobel":
        fn = lambda: kb.imgproc.sobel(src, dst, 3)
    elif op == "warp":
        H = [1.02, 0.03, -40.0, -0.03, 1.01, 25.0, 2.0e-6, 1.2e-6, 1.0]
        fn = lambda: kb.imgproc.warp_perspective(src, dst, H, kb.InterpolationMode.Bilinear)
    elif op == "affine":
        M = kb.imgproc.get_rotation_matrix2d((w / 2, h / 2), 30.0, 1.0)
        fn = lambda: kb.imgproc.warp_affine(src, dst, M, kb.InterpolationMode.Bilinear)
    elif op == "resize_f32":
        small = kb.Image.zeros_cuda(kb.ImageSize(1280, 720), 3, torch.float32, dev, batch=n)
        fn = lambda: kb.imgproc.resize(src, small, kb.InterpolationMode.Bilinear)
    elif op == "gray":
        gray = kb.Image.zeros_cuda(kb.ImageSize(w, h), 1, torch.float32, dev, batch=n)
        fn = lambda: kb.imgproc.gray_from_rgb(src, gray)
    else:
        fn = lambda: kb.imgproc.normalize_mean_std(src, dst, MEAN, STD)
elif op == "rgb_nv12":
    raw = torch.randint(0, 256, (64, 1920 * 1080 * 3 // 2), dtype=torch.uint8, device=dev, generator=g)
    rgb = kb.Image.zeros_cuda(kb.ImageSize(1920, 1080), 3, torch.uint8, dev, batch=64)
    fn = lambda: kb.imgproc.rgb_from_nv12(raw, rgb)
elif op == "remap":
    n, w, h = 8, 3840, 2160
    src = kb.Image(torch.rand((n, h, w, 3), dtype=torch.float32, device=dev, generator=g))
    dst = kb.Image.zeros_cuda(kb.ImageSize(w, h), 3, torch.float32, dev, batch=n)
    yy, xx = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32), torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
    r2 = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / float(w * w)
    mx = kb.Image((w / 2 + (xx - w / 2) * (1 + 0.05 * r2)).unsqueeze(-1).contiguous()); my = kb.Image((h / 2 + (yy - h / 2) * (1 + 0.05 * r2)).unsqueeze(-1).contiguous())
    fn = lambda: kb.imgproc.remap(src, dst, mx, my, kb.InterpolationMode.Bilinear)
elif op == "fused_general":
    nb = 16
    src = torch.randint(0, 256, (nb, 2160, 3840, 3), dtype=torch.uint8, device=dev, generator=g)
    p = kb.imgproc.NormalizeParams.from_mean_std(MEAN, STD)
    dst = torch.empty((nb, 3, 900, 1600), dtype=torch.float32, device=dev)
    fn = lambda: kb.imgproc.resize_normalize_to_tensor_u8_to_f32_bilinear(src, 1600, 900, p.scale, p.bias, out=dst)
elif op == "blur_u8":
    u8 = kb.Image(torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device=dev, generator=g))
    o8 = kb.Image.zeros_cuda(kb.ImageSize(3840, 2160), 3, torch.uint8, dev, batch=8)
    fn = lambda: kb.imgproc.gaussian_blur_u8(u8, o8, (5, 5), (1.5, 1.5))
elif op == "warp_u8":
    u8 = kb.Image(torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device=dev, generator=g))
    o8 = kb.Image.zeros_cuda(kb.ImageSize(3840, 2160), 3, torch.uint8, dev, batch=8)
    H = [1.02, 0.03, -40.0, -0.03, 1.01, 25.0, 2.0e-6, 1.2e-6, 1.0]
    fn = lambda: kb.imgproc.warp_perspective_u8(u8, o8, H)
elif op == "std_mean":
    u8 = kb.Image(torch.randint(0, 256, (32, 2160, 3840, 3), dtype=torch.uint8, device=dev, generator=g))
    fn = lambda: kb.imgproc.std_mean_sums(u8)
elif op == "cfg2":
    nb = int(os.environ.get("KB_BATCH", "16"))
    src = torch.randint(0, 256, (nb, 2160, 3840, 3), dtype=torch.uint8, device=dev, generator=g)
    p = kb.imgproc.NormalizeParams.from_mean_std(MEAN, STD)
    dst = torch.empty((nb, 3, 720, 1280), dtype=torch.float32, device=dev)
    fn = lambda: kb.imgproc.resize_normalize_to_tensor_u8_to_f32_bilinear(src, 1280, 720, p.scale, p.bias, out=dst)
else:
    raise SystemExit(f"unknown op {op}")

for _ in range(iters):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
print(op, "ms/iter", e0.elapsed_time(e1) / iters)
