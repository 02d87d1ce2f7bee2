#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fused or resize_normalize or config2" 2>&1 | tail -2
python - <<'PY'
import torch, kornia_rs_b200 as kb
dev=torch.device("cuda:0")
p = kb.imgproc.NormalizeParams.from_mean_std(kb.IMAGENET_MEAN, kb.IMAGENET_STD)
for (sw,sh,dw,dh,n) in [(3840,2160,1920,1080,32),(1920,1080,1280,720,64),(3840,2160,1280,720,64),(1920,1080,640,360,128),(3840,2160,960,540,64),(1920,1080,1920,1080,32)]:
    src=torch.randint(0,256,(n,sh,sw,3),dtype=torch.uint8,device=dev)
    dst=torch.empty((n,3,dh,dw),dtype=torch.float32,device=dev)
    fn=lambda: kb.imgproc.resize_normalize_to_tensor_u8_to_f32_bilinear(src,dw,dh,p.scale,p.bias,out=dst)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/30
    print(f"{sw}x{sh}->{dw}x{dh} x{n}: {ms:.4f} ms  dst {n*dw*dh/1e6/ms*1e3:.0f} Mpix/s  full-src+dst bytes {(n*sw*sh*3+n*dw*dh*12)/ms/1e6:.0f} GB/s")
    del src,dst
PY
