#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
python -m pytest tests -m gpu -x -q -k "resize_fast_u8 or q14" 2>&1 | tail -2
python - <<'PY'
import torch, kornia_rs_b200 as kb
dev=torch.device("cuda:0")
n=64
src=kb.Image(torch.randint(0,256,(n,2160,3840,3),dtype=torch.uint8,device=dev))
dst=kb.Image.zeros_cuda(kb.ImageSize(1280,720),3,torch.uint8,dev,batch=n)
fn=lambda: kb.imgproc.resize_fast_u8(src,dst,kb.InterpolationMode.Bilinear)
for _ in range(5): fn()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): fn()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/30
print(f"u8 bilinear 4K->720p x{n}: {ms:.4f} ms  {n*1280*720/1e6/ms*1e3:.0f} Mpix/s  alg(4/9 src + dst) {(n*(3840*2160*3*4/9+1280*720*3))/ms/1e6:.0f} GB/s")
PY
