#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
python -m pytest tests -m gpu -x -q -k "blur_u8 or box_blur or resize_fast_u8" 2>&1 | tail -3
python - <<'PY'
import torch, kornia_rs_b200 as kb
dev=torch.device("cuda:0")
n,w,h=16,3840,2160
src=kb.Image(torch.randint(0,256,(n,h,w,3),dtype=torch.uint8,device=dev))
dst=kb.Image.zeros_cuda(kb.ImageSize(w,h),3,torch.uint8,dev,batch=n)
for name,fn in [("gaussian_blur_u8 5x5",lambda: kb.imgproc.gaussian_blur_u8(src,dst,(5,5),(1.5,1.5))),("gaussian_blur_u8 3x3 binomial",lambda: kb.imgproc.gaussian_blur_u8(src,dst,(3,3),(1.0,1.0))),("gaussian_blur_u8 7x7",lambda: kb.imgproc.gaussian_blur_u8(src,dst,(7,7),(2.0,2.0)))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f"{name} 4K x{n}: {ms:.4f} ms  {n*w*h/1e6/ms*1e3:.0f} Mpix/s  src+dst {(2*n*w*h*3)/ms/1e6:.0f} GB/s")
PY
