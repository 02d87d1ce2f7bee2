#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "resize_fast_u8 or q14" 2>&1 | tail -3
python - <<'PY'
import torch, kornia_rs_b200 as kb
dev=torch.device("cuda:0")
for (sw,sh,dw,dh,n,mode) in [(3840,2160,1920,1080,32,"Bilinear"),(1920,1080,3840,2160,16,"Bilinear"),(3840,2160,1280,720,64,"Nearest"),(3840,2160,1280,720,64,"Bilinear")]:
    src=kb.Image(torch.randint(0,256,(n,sh,sw,3),dtype=torch.uint8,device=dev))
    dst=kb.Image.zeros_cuda(kb.ImageSize(dw,dh),3,torch.uint8,dev,batch=n)
    fn=lambda: kb.imgproc.resize_fast_u8(src,dst,getattr(kb.InterpolationMode,mode))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/30
    print(f"u8 {mode} {sw}x{sh}->{dw}x{dh} x{n}: {ms:.4f} ms  dst {n*dw*dh/1e6/ms*1e3:.0f} Mpix/s  src+dst {(n*sw*sh*3+n*dw*dh*3)/ms/1e6:.0f} GB/s")
PY
