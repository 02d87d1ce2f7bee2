#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "remap" 2>&1 | tail -3
python - <<'PY'
import torch, numpy as np, kornia_rs_b200 as kb
dev=torch.device("cuda:0")
n,w,h=8,3840,2160
y,x=torch.meshgrid(torch.arange(h,device=dev,dtype=torch.float32),torch.arange(w,device=dev,dtype=torch.float32),indexing="ij")
r2=((x-w/2)**2+(y-h/2)**2)/(w*w)
mx=kb.Image((w/2+(x-w/2)*(1+0.05*r2)).unsqueeze(-1).contiguous()); my=kb.Image((h/2+(y-h/2)*(1+0.05*r2)).unsqueeze(-1).contiguous())
for name,dt,c in [("remap f32",torch.float32,3),("remap_u8",torch.uint8,3)]:
    src=kb.Image(torch.rand((n,h,w,c),device=dev) if dt==torch.float32 else torch.randint(0,256,(n,h,w,c),dtype=torch.uint8,device=dev))
    dst=kb.Image.zeros_cuda(kb.ImageSize(w,h),c,dt,dev,batch=n)
    fn=(lambda: kb.imgproc.remap(src,dst,mx,my,kb.InterpolationMode.Bilinear)) if dt==torch.float32 else (lambda: kb.imgproc.remap_u8(src,dst,mx,my,kb.InterpolationMode.Bilinear))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    es=4 if dt==torch.float32 else 1
    print(f"{name} 4K x{n}: {ms:.4f} ms  {n*w*h/1e6/ms*1e3:.0f} Mpix/s  src+dst+maps {(2*n*w*h*c*es+n*w*h*8)/ms/1e6:.0f} GB/s")
    del src,dst
PY
