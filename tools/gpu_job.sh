#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "u8" 2>&1 | tail -3
python - <<'PY'
import torch, kornia_rs_b200 as kb
dev=torch.device("cuda:0")
n,w,h=16,3840,2160
src=kb.Image(torch.randint(0,256,(n,h,w,3),dtype=torch.uint8,device=dev))
dst=kb.Image.zeros_cuda(kb.ImageSize(w,h),3,torch.uint8,dev,batch=n)
H=[1.02,0.03,-40.0,-0.03,1.01,25.0,2.0e-6,1.2e-6,1.0]
M=kb.imgproc.get_rotation_matrix2d((w/2,h/2),30.0,1.0)
for name,fn in [("warp_perspective_u8",lambda: kb.imgproc.warp_perspective_u8(src,dst,H)),("warp_affine_u8 rot30",lambda: kb.imgproc.warp_affine_u8(src,dst,M)),("warp_affine_u8 shift",lambda: kb.imgproc.warp_affine_u8(src,dst,[1,0,5.5,0,1,-3.25]))]:
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/30
    print(f"{name} 4K x{n}: {ms:.4f} ms  {n*w*h/1e6/ms*1e3:.0f} Mpix/s  src+dst {(2*n*w*h*3)/ms/1e6:.0f} GB/s")
PY
