"""CPU-only tests: the C-ABI library loads and exports every symbol include/kornia_b200.h declares, the
host-side helpers of the ABI reproduce the reference's arithmetic (checked against the oracle), argument
validation fails with the reference's messages BEFORE any CUDA call, and the Python mirror's host logic.
No compute kernels are launched here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(kb):
    hdr = open(os.path.join(ROOT, "include", "kornia_b200.h")).read()
    declared = sorted(set(re.findall(r"KB200_API\s+[\w\s\*]+?\b(kb200_\w+)\s*\(", hdr)))
    assert len(declared) >= 38
    from kornia_rs_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (kb200_\w+)", out))
    assert not [s for s in declared if s not in exported]
    l = _lib.lib()
    for s in declared:
        assert hasattr(l, s)
    assert kb.native_version() == 100
    # only kb200_* is exported from the product library (no accidental oracle / helper leakage)
    others = [ln for ln in out.splitlines() if " T " in ln and "kb200_" not in ln]
    assert not others, others[:5]


def test_product_does_not_touch_the_oracle():
    """The product package must never import / link the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "kornia-rs_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.\.?oracle|kornia_oracle|oracle/|ko_[a-z_]+\(")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f"{f} references the oracle"
    from kornia_rs_b200 import _lib

    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "kornia_oracle" not in ldd


def test_host_helpers_match_oracle(kb, oracle):
    ip = kb.imgproc
    for m in ([1, 0, 0, 0, 1, 0], [0.7, 0.2, 3.0, -0.1, 1.3, -2.0], [2, 0, 0, 0, 0, 0]):
        assert ip.invert_affine_transform(m) == pytest.approx(oracle.invert_affine_transform(m).tolist(), nan_ok=True, abs=0)
    for ang in (0.0, 30.0, 90.0, 45.5, -120.0):
        assert ip.get_rotation_matrix2d((3.5, 7.25), ang, 1.25) == oracle.get_rotation_matrix2d((3.5, 7.25), ang, 1.25).tolist()
    h = [1.02, 0.03, -5.0, -0.01, 0.99, 2.0, 0.00005, 0.00003, 1.0]
    assert ip.invert_homography(h) == oracle.invert_homography(h).tolist()
    assert ip.invert_homography([0] * 9) is None and ip.invert_homography([1, 2, 3, 2, 4, 6, 3, 6, 9]) is None
    for k, s in ((5, 0.5), (3, 0.25), (13, 1.5), (31, 4.0)):
        assert ip.gaussian_kernel_1d(k, s) == oracle.gaussian_kernel_1d(k, s).tolist()
    assert ip.gaussian_kernel_1d(5, 0.5) == [float(np.float32(v)) for v in
                                             (0.00026386508, 0.10645077, 0.78657067, 0.10645077, 0.00026386508)]  # filter/kernels.rs:201
    from kornia_rs_b200 import _lib

    l = _lib.lib()
    for args in ((0, 0, 1.5, 0.0), (3, 3, 0.0, 0.0), (5, 0, 0.7, 2.0), (0, 7, 0.3, -1.0)):
        kx, ky, sx, sy = C.c_uint32(), C.c_uint32(), C.c_float(), C.c_float()
        assert l.kb200_gaussian_resolve(*args, C.byref(kx), C.byref(ky), C.byref(sx), C.byref(sy)) == 0
        assert (kx.value, ky.value, sx.value, sy.value) == oracle.gaussian_resolve(*args)
    kx, ky, sx, sy = C.c_uint32(), C.c_uint32(), C.c_float(), C.c_float()
    assert l.kb200_gaussian_resolve(2, 3, 1.0, 1.0, C.byref(kx), C.byref(ky), C.byref(sx), C.byref(sy)) == _lib.ERR_INVALID_KERNEL
    for mode, geom in ((0, (1920, 1080, 640, 640)), (1, (1920, 1080, 640, 640)), (0, (4, 4, 8, 4)), (0, (23, 17, 8, 6))):
        a = (C.c_float * 4)()
        l.kb200_preprocess_affine(mode, *geom, a)
        assert tuple(a) == oracle.preprocess_affine(mode, *geom)
    sums = (C.c_uint64 * 6)(445, 449, 453, 84489, 85383, 86285)
    std, mean = (C.c_double * 3)(), (C.c_double * 3)()
    l.kb200_std_mean_finalize(sums, 4, std, mean)
    assert list(std) == [93.5183805462862] * 3 and list(mean) == [111.25, 112.25, 113.25]  # core.rs:27-40


def test_argument_validation_without_a_gpu(kb):
    """Validation errors are produced before any CUDA call, with the reference's wording."""
    from kornia_rs_b200 import _lib

    l = _lib.lib()
    buf = (C.c_float * 64)()
    p = C.addressof(buf)

    def err():
        return _lib.last_error()

    assert l.kb200_resize_bilinear_f32_c3(None, p, 64, p, 64, 0, 4, 2, 2, 1, 0) == _lib.ERR_INVALID_ARGUMENT
    assert err() == "image dimensions must be non-zero"  # cuda/resize.rs:515-519
    assert l.kb200_resize_bilinear_f32_c3(None, p, 64, p, 5, 4, 4, 2, 2, 1, 0) == _lib.ERR_SLICE_TOO_SMALL
    assert err() == "device slice 'dst' length 5 < required 12"  # SliceTooSmall{what,got,need}
    assert l.kb200_resize_bilinear_f32_c3(None, None, 64, p, 64, 4, 4, 2, 2, 1, 0) == _lib.ERR_INVALID_ARGUMENT
    assert l.kb200_resize_bilinear_normalize_f32_c3(None, p, 64, p, 64, 2, 2, 2, 2, 1, _lib.f3([0, 0, 0]), _lib.f3([1, 0, 1]), 0) \
        == _lib.ERR_INVALID_ARGUMENT
    assert err() == "std must be non-zero for all channels"  # cuda/resize.rs:606-610
    assert l.kb200_warp_perspective_f32_c3(None, p, 64, p, 64, 2, 2, 2, 2, 1, _lib.f3([1, 2, 3, 2, 4, 6, 3, 6, 9], 9), 1) \
        == _lib.ERR_SINGULAR_MATRIX
    assert err() == "homography matrix is singular (|det| < 1e-10)"  # cuda/warp_perspective.rs:408-410
    assert l.kb200_warp_affine_f32_c3(None, p, 64, p, 64, 2, 2, 2, 2, 1, _lib.f3([1, 0, 0, 0, 1, 0], 6), 7) == _lib.ERR_UNSUPPORTED   # 0..3 = Nearest/Bilinear/Bicubic/Lanczos
    assert l.kb200_sobel_f32(None, p, 64, p, 64, 2, 2, 3, 1, 7) == _lib.ERR_INVALID_KERNEL
    assert l.kb200_separable_filter_f32(None, p, 64, p, 64, None, _lib.f3([1] * 3), 0, _lib.f3([1] * 3), 3, 2, 2, 3, 1) == _lib.ERR_INVALID_KERNEL
    assert l.kb200_resize_bilinear_u8(None, p, 64, p, 64, 1, 4, 2, 2, 3, 1) == _lib.ERR_INVALID_ARGUMENT  # needs >= 2x2
    assert l.kb200_rgb_from_nv12_u8(None, p, 64, p, 64, 5, 4, 1) == _lib.ERR_INVALID_ARGUMENT
    d = _lib.PreprocessDesc()
    d.scale_x = d.scale_y = 1.0
    d.src_w, d.src_h, d.src_pitch, d.src_bpp, d.fmt, d.dst_w, d.dst_h, d.sampling = 8, 5, 8, 1, 3, 4, 4, 1
    ptrs = (C.c_void_p * 1)(p)
    lens = (C.c_size_t * 1)(60)
    assert l.kb200_preprocess_f32(None, C.byref(d), ptrs, lens, 1, p, 64) == _lib.ERR_INVALID_SOURCE  # odd NV12 height
    d.src_h = 6
    assert l.kb200_preprocess_src_bytes(C.byref(d)) == 72
    assert l.kb200_preprocess_f32(None, C.byref(d), ptrs, lens, 1, p, 64) == _lib.ERR_INVALID_SOURCE
    assert "got 60 bytes, need 72" in err()  # preprocess.rs:1287-1300
    d.sampling = 2   # Bicubic: UnsupportedSampling in the reference too (preprocess.rs:1044-1051); 3 = Lanczos is supported
    assert l.kb200_preprocess_f32(None, C.byref(d), ptrs, lens, 1, p, 64) == _lib.ERR_UNSUPPORTED
    assert l.kb200_status_name(_lib.ERR_SLICE_TOO_SMALL) == b"KB200_ERR_SLICE_TOO_SMALL"


def test_python_mirror_host_logic(kb):
    SF = kb.SourceFormat
    assert SF.Nv12.buffer_len(8, 6) == 72 and SF.Yuyv.buffer_len(8, 6) == 96 and SF.Rgba8.buffer_len(8, 6) == 192
    assert not SF.Nv12.dims_ok(8, 5) and SF.Yuyv.dims_ok(8, 5) and not SF.Yuyv.dims_ok(7, 5)
    assert SF.from_name("NV12") is SF.Nv12 and SF.from_name("bgr") is SF.Bgr8 and SF.from_name("xyz") is None
    assert [f.fmt_code() for f in (SF.Rgb8, SF.Rgba8, SF.Bgr8, SF.Bgra8, SF.Gray8, SF.Nv12, SF.Yuyv)] == [0, 0, 1, 1, 2, 3, 4]
    mean, inv = kb.Normalize.imagenet().mean_inv_std()
    assert inv == tuple(float(np.float32(1.0) / np.float32(s)) for s in kb.IMAGENET_STD)
    assert kb.Normalize.UnitScale().mean_inv_std() == ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))
    with pytest.raises(kb.PreprocessError) as e:
        kb.Preprocessor.builder().normalize(kb.Normalize.MeanStd([0.5] * 3, [0.0, 0.2, 0.2])).build_cuda()
    assert e.value.kind == "InvalidNormalize"
    with pytest.raises(kb.PreprocessError) as e:
        kb.Preprocessor.builder().sampling(kb.InterpolationMode.Bicubic).build_cuda()
    assert e.value.kind == "UnsupportedSampling"
    # Image container rules
    with pytest.raises(kb.ImageError) as e:
        kb.Image(kb.ImageSize(2, 2), [1.0] * 13, channels=3, dtype=torch.float32)
    assert e.value.kind == "InvalidChannelShape"
    im = kb.Image(kb.ImageSize(2, 3), list(range(18)), dtype=torch.float32)
    assert (im.rows(), im.cols(), im.num_channels(), im.batch) == (3, 2, 3, 1) and im.size() == kb.ImageSize(2, 3)
    with pytest.raises(kb.ImageError) as e:
        kb.Image(torch.zeros(4, 4, 3).transpose(0, 1))
    assert e.value.kind == "ImageDataNotContiguous"
    # host operands: typed error, never a CPU fallback
    with pytest.raises(kb.ImageError) as e:
        kb.imgproc.resize(im, im, kb.InterpolationMode.Bilinear)
    assert e.value.kind == "UnsupportedDevice"
    with pytest.raises(kb.ImageError) as e:
        kb.imgproc.resize(im, im, kb.InterpolationMode.Lanczos)
    assert e.value.kind == "UnsupportedDevice"   # every sampler is built for device images; host images have no CPU path here
    with pytest.raises(kb.ImageError) as e:
        kb.imgproc.sobel_kernel_1d(7)
    assert e.value.kind == "InvalidKernelLength"
    p = kb.imgproc.NormalizeParams.from_mean_std([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    want_scale = (np.float32(1.0) / (np.array([0.229, 0.224, 0.225], np.float32) * np.float32(255.0))).tolist()
    assert p.scale == want_scale


def test_shard_ranges(kb):
    d = kb.dist
    for n, ws in ((64, 8), (512, 8), (10, 4), (3, 8), (0, 2), (257, 3)):
        shards = [d.shard_range(n, r, ws) for r in range(ws)]
        assert shards[0].start == 0 and shards[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
        counts = [s.count for s in shards]
        assert max(counts) - min(counts) <= 1 and sum(counts) == n
    assert d.shard_range(512, 3, 8).count == 64 and d.shard_range(128, 7, 8) == d.Shard(7, 8, 112, 128)
    with pytest.raises(ValueError):
        d.shard_range(4, 4, 4)
    assert d.world_size() == 1 and d.rank() == 0
    assert d.broadcast_params({"h": [1.0, 0.1, 1e-6]}) == {"h": [float(np.float32(v)) for v in (1.0, 0.1, 1e-6)]}


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import kornia_rs_b200 as kb
from oracle import oracle as o
dev = kb.dist.init_from_env("gloo")
r, ws = kb.dist.rank(), kb.dist.world_size()
assert ws == 2 and dev.type == "cpu"
# 1) ONE broadcast of the parameter block from rank 0
layout = {{"homography": [0.0] * 9, "mean": [0.0] * 3, "inv_std": [0.0] * 3}}
if r == 0:
    layout = {{"homography": [1.02, 0.03, -40.0, -0.03, 1.01, 25.0, 2.0e-6, 1.2e-6, 1.0], "mean": [0.485, 0.456, 0.406],
              "inv_std": [float(np.float32(1) / np.float32(s)) for s in (0.229, 0.224, 0.225)]}}
got = kb.dist.broadcast_params(layout)
assert got["homography"] == [float(np.float32(v)) for v in (1.02, 0.03, -40.0, -0.03, 1.01, 25.0, 2.0e-6, 1.2e-6, 1.0)], got
# 2) shard a batch of 5 images; each rank computes its shard's exact integer sums (oracle stands in for the
#    GPU kernel on this CPU-only host); ONE all-reduce gives the global std_mean
n, w, h = 5, 31, 17
imgs = [o.pattern_u8(w * h * 3, 1000 + i).reshape(h, w, 3) for i in range(n)]
sh = kb.dist.shard_range(n, r, ws)
local = np.zeros(6, np.int64)
for i in range(sh.start, sh.stop):
    local += o.std_mean(imgs[i])[2].astype(np.int64)
total = kb.dist.all_reduce_sums(torch.from_numpy(local))
std, mean = kb.imgproc.std_mean_finalize(total.tolist(), n * w * h)
ostd, omean, osums = o.std_mean(np.concatenate(imgs, 0))
assert total.tolist() == [int(v) for v in osums]
assert std == ostd.tolist() and mean == omean.tolist()
assert kb.dist.max_over_ranks(float(r + 1)) == 2.0 and kb.dist.sum_over_ranks(1.0) == 2.0
kb.dist.barrier()
sys.stdout.write("RANK%dOK %d-%d\n" % (r, sh.start, sh.stop)); sys.stdout.flush()
"""


def test_world_size_2_gloo(tmp_path):
    """The N>1 host path (sharding + the two collectives) on CPU with gloo, world_size 2."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "RANK0OK 0-3" in r.stdout and "RANK1OK 3-5" in r.stdout, r.stdout


def test_cfg3b_algorithmic_bytes(oracle):
    """SURVEY §8(d) asks for the distinct source bytes of config 3b counted exactly in the oracle.  1080p NV12 →
    640x640 letterbox is an exact 3:1 decimation (sx = 3·ox, sy = 3·(oy-140) in f32): the reference kernel addresses
    the 2x2 taps {3ox, 3ox+1} x {3oy', 3oy'+1} of every content pixel (the +1 taps carry weight 0)."""
    w, h = 1920, 1080
    raw = ((np.arange(w * h * 3 // 2, dtype=np.int64) * 7 + 13) % 251).astype(np.uint8)
    cfg = oracle.PreprocessCfg(mode=oracle.LETTERBOX, fmt=oracle.FMT_NV12)
    assert oracle.preprocess_affine(oracle.LETTERBOX, w, h, 640, 640) == (float(np.float32(640) / np.float32(1920)),) * 2 + (0.0, 140.0)
    out, touched = oracle.preprocess_frame(raw, cfg, w, h, 640, 640, count_touched=True)
    # Y: 2 of 3 columns x 2 of 3 rows = 1280 x 720 bytes; UV: every pair of the 360+... rows {(3k)>>1, (3k+1)>>1} = 540 rows x 1920 B
    assert touched == 1280 * 720 + 1920 * 540 == 1958400      # bench.py's algorithmic source bytes for cfg 3b
    assert out.shape == (3, 640, 640)
    # (the CUDA kernel fetches only the weight-carrying tap: 640*360 Y + 640*360*2 UV = 691,200 B — fewer than algorithmic)
    # config 2: 4/9 of the source is addressed (taps at rows/cols 3d+1, 3d+2)
    assert oracle.count_touched_resize(3840, 2160, 1280, 720, 1) * 3 == 3840 * 2160 * 3 * 4 // 9


def test_bench_lcg_pattern_matches_reference_generator(oracle):
    """bench.py generates the per-frame LCG patterns with torch (vectorised closed form); it must equal the
    reference's sequential generator (cuda/color/mod.rs:303-316) for any seed."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n, seed in ((7, 1), (15, 2), (16, 0x12345678), (5000, 0x12345678 + 63), (100003, 0xFFFFFFFF)):
        got = bench.lcg_pattern_u8(n, seed, "cpu").numpy()
        np.testing.assert_array_equal(got, oracle.pattern_u8(n, seed))


def test_resize_row_plan_matches_tapped_rows(kb):
    """kb200_resize_row_plan (host-only) reports a window that holds every row the f32 half-pixel sampler taps with a
    non-zero weight, and is minimal for integer downscales."""
    import numpy as np

    plan = kb.imgproc.resize_row_plan

    assert plan(2160, 720) == (3, 1, 1)      # BASELINE config 2: wy == 0 on every row
    assert plan(2160, 540) == (4, 1, 2)
    assert plan(100, 20) == (5, 2, 1)
    assert plan(216, 36) == (6, 2, 2)
    assert plan(2160, 1080) == (1, 0, 1)     # exact 2x is the box path: all rows
    assert plan(1080, 720) == (1, 0, 1)
    assert plan(720, 2160) == (1, 0, 1)
    assert plan(0, 0) == (1, 0, 1)
    for sh, dh in [(2160, 720), (2160, 540), (100, 20), (216, 36), (63, 21), (45, 9), (1080, 720), (75, 41)]:
        P, F, K = plan(sh, dh)
        d = np.arange(dh, dtype=np.float32)
        f = np.maximum((d + np.float32(0.5)) * (np.float32(sh) / np.float32(dh)) - np.float32(0.5), np.float32(0))
        y0 = np.minimum(f.astype(np.uint32), sh - 1)
        wy = f - y0.astype(np.float32)
        y1 = np.where(wy == 0, y0, np.minimum(y0 + 1, sh - 1))
        for y in np.concatenate([y0, y1]):
            assert F <= int(y) % P < F + K, (sh, dh, int(y), (P, F, K))


def test_ctypes_table_matches_the_header(kb):
    """Every entry point the header declares has a ctypes signature in _lib.py, and nothing is bound that the header does
    not declare — the binding the tests exercise is the documented ABI, all of it."""
    hdr = open(os.path.join(ROOT, "include", "kornia_b200.h")).read()
    declared = set(re.findall(r"KB200_API\s+[\w\s\*]+?\b(kb200_\w+)\s*\(", hdr))
    from kornia_rs_b200 import _lib

    bound = set(re.findall(r'"(kb200_\w+)":', open(_lib.__file__).read()))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name).argtypes is not None, name


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): exactly one JSON line on stdout with
    the contract's keys; it needs no GPU."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_rust_binding_covers_the_header():
    """bindings/rust/src/lib.rs cannot be compiled here (no Rust toolchain): at least keep its `extern "C"` block in
    lock-step with include/kornia_b200.h — every declared entry point has an FFI declaration, and nothing extra."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "kornia_b200.h")).read()
    rs = open(os.path.join(root, "bindings", "rust", "src", "lib.rs")).read()
    declared = set(re.findall(r"KB200_API\s+[\w\s\*]+?\b(kb200_\w+)\s*\(", hdr))
    ffi = set(re.findall(r"pub fn (kb200_\w+)\s*\(", rs.split("extern \"C\" {", 1)[1].split("\n    }\n", 1)[0]))
    assert declared - ffi == set(), f"missing in the Rust FFI block: {sorted(declared - ffi)}"
    assert ffi - declared == set(), f"not in the header: {sorted(ffi - declared)}"
    assert "#[repr(C)]\n    #[derive(Clone, Copy, Debug, Default)]\n    pub struct kb200_preprocess_desc" in rs   # repr(C) on the descriptor
