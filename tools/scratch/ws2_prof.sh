python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, kornia_rs_b200 as kb
kb._lib.set_knob("warp.path", 3)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
n, w, h = 8, 3840, 2160
src = kb.Image(torch.rand((n, h, w, 3), dtype=torch.float32, device=dev, generator=g))
dst = kb.Image.zeros_cuda(kb.ImageSize(w, h), 3, torch.float32, dev, batch=n)
H = [1.02, 0.03, -40.0, -0.03, 1.01, 25.0, 2.0e-6, 1.2e-6, 1.0]
for _ in range(4):
    kb.imgproc.warp_perspective(src, dst, H, kb.InterpolationMode.Bilinear)
torch.cuda.synchronize()
print(kb._lib.last_kernel())
PY
