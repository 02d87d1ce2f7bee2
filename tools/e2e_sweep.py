"""Sweep the host-pipeline ring (chunk frames x depth) for config 2: python tools/e2e_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kornia_rs_b200 as kb

dev = torch.device("cuda:0")
SW, SH, DW, DH, B = 3840, 2160, 1280, 720, 64
p = kb.imgproc.NormalizeParams.from_mean_std(kb.IMAGENET_MEAN, kb.IMAGENET_STD)
hs = torch.randint(0, 256, (B, SH, SW, 3), dtype=torch.uint8).pin_memory()
hd = torch.empty((B, 3, DH, DW), dtype=torch.float32, pin_memory=True)
st = torch.cuda.current_stream(dev)
for chunk in (4, 8, 16, 32):
    for depth in (2, 3, 4):
        pipe = kb.imgproc.HostPipeline(dev, src_chunk_bytes=chunk * SW * SH * 3, dst_chunk_bytes=chunk * 3 * DW * DH * 4, depth=depth)
        fn = lambda: kb.imgproc.resize_normalize_to_tensor_u8_to_f32_bilinear(hs, DW, DH, p.scale, p.bias, out=hd, pipeline=pipe)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(6):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        up, down = pipe.last_transfer()
        print(f"chunk={chunk:2d} depth={depth}: {ms:7.3f} ms/step  {B*DW*DH/1e6/ms*1e3:7.0f} Mpix/s  up {up/ms/1e6:5.1f} GB/s  down {down/ms/1e6:5.1f} GB/s", flush=True)
        pipe.close()
