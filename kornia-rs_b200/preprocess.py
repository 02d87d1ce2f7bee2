"""`Preprocessor` — the fused camera preprocess (a12), mirroring preprocess.rs:67-149, :654-1282.

Built once via `Preprocessor.builder()…build_cuda()`, then applied to any number of frames of any
resolution: the target H/W is read from the destination tensor each call.  Camera formats
(NV12 / YUYV / Gray) go through `run_raw*`; colour decode is fused into the resample taps.  Unlike the
reference (one launch per frame, preprocess.rs:1277-1280) a batch is ONE launch.

Only the CUDA preprocessor exists here (`build()` = the reference's CPU preprocessor is out of scope:
it raises).  Errors mirror `PreprocessError` (preprocess.rs:252-332).
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Sequence

import torch

from . import _lib
from .image import Image, InterpolationMode

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # preprocess.rs:93
IMAGENET_STD = (0.229, 0.224, 0.225)   # preprocess.rs:95


class PreprocessError(Exception):
    """preprocess.rs:252-332; `kind` names the variant."""

    def __init__(self, kind: str, message: str, **fields):
        super().__init__(message)
        self.kind = kind
        self.fields = fields


class ResizeMode(enum.Enum):  # preprocess.rs:67
    Letterbox = 0
    Stretch = 1


class SourceFormat(enum.Enum):  # preprocess.rs:130
    Rgb8 = "rgb8"
    Bgr8 = "bgr8"
    Rgba8 = "rgba8"
    Bgra8 = "bgra8"
    Gray8 = "gray8"
    Nv12 = "nv12"
    Yuyv = "yuyv"

    def fmt_code(self) -> int:  # :153-161
        return {"rgb8": 0, "rgba8": 0, "bgr8": 1, "bgra8": 1, "gray8": 2, "nv12": 3, "yuyv": 4}[self.value]

    def bpp(self) -> int:  # :164-171
        return {"rgb8": 3, "bgr8": 3, "rgba8": 4, "bgra8": 4, "gray8": 1, "nv12": 1, "yuyv": 2}[self.value]

    def pitch(self, w: int) -> int:
        return w * self.bpp()

    def buffer_len(self, w: int, h: int) -> int:  # :177-186
        chroma = w * h // 2 if self is SourceFormat.Nv12 else 0
        return self.pitch(w) * h + chroma

    def dims_ok(self, w: int, h: int) -> bool:  # :188-195
        if self is SourceFormat.Nv12:
            return w % 2 == 0 and h % 2 == 0
        if self is SourceFormat.Yuyv:
            return w % 2 == 0
        return True

    def interleaved(self) -> bool:  # :211-216
        return self in (SourceFormat.Rgb8, SourceFormat.Bgr8, SourceFormat.Rgba8, SourceFormat.Bgra8)

    @staticmethod
    def from_name(name: str):  # :221-233
        table = {"rgb": "Rgb8", "rgb8": "Rgb8", "bgr": "Bgr8", "bgr8": "Bgr8", "rgba": "Rgba8", "rgba8": "Rgba8",
                 "bgra": "Bgra8", "bgra8": "Bgra8", "gray": "Gray8", "gray8": "Gray8", "nv12": "Nv12", "yuyv": "Yuyv"}
        key = table.get(name.lower())
        return None if key is None else SourceFormat[key]


@dataclass(frozen=True)
class Normalize:  # preprocess.rs:77
    mean: tuple | None = None
    std: tuple | None = None

    @staticmethod
    def UnitScale() -> "Normalize":
        return Normalize()

    @staticmethod
    def MeanStd(mean, std) -> "Normalize":
        return Normalize(tuple(mean), tuple(std))

    @staticmethod
    def imagenet() -> "Normalize":
        return Normalize(IMAGENET_MEAN, IMAGENET_STD)

    def mean_inv_std(self):  # :113-125, f32 arithmetic for 1/std
        if self.mean is None:
            return (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
        import math

        std32 = torch.tensor(self.std, dtype=torch.float32)
        mean32 = torch.tensor(self.mean, dtype=torch.float32)
        if any((not math.isfinite(s)) or s <= 0.0 for s in std32.tolist()) or any(not math.isfinite(m) for m in mean32.tolist()):
            raise PreprocessError("InvalidNormalize", f"invalid normalize: mean {self.mean} must be finite, std {self.std} must be finite and > 0",
                                  mean=self.mean, std=self.std)
        inv = (torch.tensor(1.0, dtype=torch.float32) / std32).tolist()
        return tuple(mean32.tolist()), tuple(inv)


@dataclass
class PitchedSurface:  # preprocess.rs:380-392
    data: torch.Tensor
    width: int
    height: int
    row_pitch: int
    channels: int

    def validate(self) -> None:  # :829-842
        if self.channels not in (3, 4):
            raise PreprocessError("UnsupportedChannels", f"unsupported source channel count {self.channels} (expected 3 or 4)")
        if (self.row_pitch < self.width * self.channels or self.data.numel() < self.row_pitch * self.height
                or self.width == 0 or self.height == 0):
            raise PreprocessError("InvalidSurface", "invalid pitched surface (need pitch >= width*channels and len >= pitch*height)")


class PreprocessorBuilder:  # preprocess.rs:654-779
    def __init__(self):
        self._mode = ResizeMode.Letterbox
        self._normalize = Normalize.UnitScale()
        self._pad_value = 114
        self._sampling = InterpolationMode.Bilinear
        self._source_format = SourceFormat.Rgb8

    def source_format(self, fmt: SourceFormat):
        self._source_format = fmt
        return self

    def mode(self, mode: ResizeMode):
        self._mode = mode
        return self

    def normalize(self, normalize: Normalize):
        self._normalize = normalize
        return self

    def pad_value(self, pad_value: int):
        self._pad_value = int(pad_value) & 0xFF
        return self

    def sampling(self, sampling: InterpolationMode):
        self._sampling = sampling
        return self

    def _validated(self):
        if self._sampling not in (InterpolationMode.Nearest, InterpolationMode.Bilinear, InterpolationMode.Lanczos):
            raise PreprocessError("UnsupportedSampling", f"unsupported sampling mode {self._sampling!r} (expected Nearest, Bilinear, or Lanczos)")
        return self._normalize.mean_inv_std()

    def build(self):
        """The reference's CPU preprocessor (host image → host tensor) is out of this tier's scope."""
        self._validated()
        raise PreprocessError("NotDeviceImage", "CPU preprocessor not built in this tier: use build_cuda() (device operands)")

    def build_cuda(self, device: str | torch.device | None = None) -> "Preprocessor":
        mean, inv_std = self._validated()
        _lib.lib()  # fail loudly if the native library is missing
        return Preprocessor(self._mode, self._sampling, self._source_format, mean, inv_std, float(self._pad_value),
                            torch.device(device) if device is not None else None)


class Preprocessor:
    def __init__(self, mode, sampling, source_format, mean, inv_std, pad_value, device):
        self._mode, self._sampling, self._fmt = mode, sampling, source_format
        self._mean, self._inv_std, self._pad = mean, inv_std, pad_value
        self._device = device

    @staticmethod
    def builder() -> PreprocessorBuilder:
        return PreprocessorBuilder()

    @staticmethod
    def letterbox(device=None) -> "Preprocessor":
        return PreprocessorBuilder().mode(ResizeMode.Letterbox).build_cuda(device)

    @staticmethod
    def stretch(device=None) -> "Preprocessor":
        return PreprocessorBuilder().mode(ResizeMode.Stretch).build_cuda(device)

    @staticmethod
    def with_mode(mode: ResizeMode, device=None) -> "Preprocessor":
        return PreprocessorBuilder().mode(mode).build_cuda(device)

    def mode(self) -> ResizeMode:
        return self._mode

    # ── helpers ─────────────────────────────────────────────────────────────
    def _desc(self, sw, sh, pitch, bpp, fmt_code, dw, dh) -> _lib.PreprocessDesc:
        lim = 2 ** 31 - 1
        if sw > lim or sh > lim or pitch > lim or dw * dh > lim:  # :1336-1339
            raise PreprocessError("DimensionsTooLarge", "dimensions exceed the 32-bit CUDA kernel index limit")
        a = (C.c_float * 4)()
        _lib.lib().kb200_preprocess_affine(self._mode.value, sw, sh, dw, dh, a)
        d = _lib.PreprocessDesc()
        d.scale_x, d.scale_y, d.pad_x, d.pad_y = a[0], a[1], a[2], a[3]
        d.src_w, d.src_h, d.src_pitch, d.src_bpp, d.fmt = sw, sh, pitch, bpp, fmt_code
        d.dst_w, d.dst_h = dw, dh
        d.mean = (C.c_float * 3)(*self._mean)
        d.inv_std = (C.c_float * 3)(*self._inv_std)
        d.pad_value = self._pad
        d.sampling = {InterpolationMode.Nearest: 0, InterpolationMode.Bilinear: 1, InterpolationMode.Lanczos: 3}[self._sampling]
        return d

    @staticmethod
    def _validate_dst(dst: torch.Tensor, expected_n: int, f16: bool) -> None:
        shape = tuple(dst.shape)
        if dst.dim() != 4 or shape[1] != 3 or (expected_n == 1 and shape[0] != 1):  # :805-808
            raise PreprocessError("BadOutputShape", f"destination tensor must be [1, 3, H, W], got {list(shape)}")
        if shape[0] != expected_n:
            raise PreprocessError("BatchMismatch", f"destination batch dim {shape[0]} != frame count {expected_n}",
                                  dst_n=shape[0], frames=expected_n)
        if not dst.is_cuda:
            raise PreprocessError("NotDeviceTensor", "CUDA preprocessor requires a device-resident destination tensor")
        want = torch.float16 if f16 else torch.float32
        if dst.dtype != want or not dst.is_contiguous():
            raise PreprocessError("BadOutputShape", f"destination tensor must be contiguous {want}, got {dst.dtype}")

    def _launch(self, desc, frames: Sequence[torch.Tensor], dst: torch.Tensor, f16: bool) -> None:
        dev = dst.device
        for f in frames:
            if not f.is_cuda:
                raise PreprocessError("NotDeviceImage", "CUDA preprocessor requires a device-resident source image")
            if f.device != dev:
                raise PreprocessError("Cuda", "CUDA error: source and destination are on different CUDA devices")
        _lib.set_device(dev.index)
        n = len(frames)
        ptrs = (C.c_void_p * n)(*[f.data_ptr() for f in frames])
        lens = (C.c_size_t * n)(*[f.numel() * f.element_size() for f in frames])
        l = _lib.lib()
        fn = l.kb200_preprocess_f16 if f16 else l.kb200_preprocess_f32
        st = fn(torch.cuda.current_stream(dev).cuda_stream, C.byref(desc), ptrs, lens, n, dst.data_ptr(), dst.numel())
        if st != _lib.OK:
            kind = {_lib.ERR_INVALID_SOURCE: "InvalidRawSource", _lib.ERR_DIMS_TOO_LARGE: "DimensionsTooLarge"}.get(st, "Cuda")
            raise PreprocessError(kind, ("CUDA error: " if kind == "Cuda" else "") + _lib.last_error())

    def _validate_typed_format(self, c: int) -> None:  # :913-923
        f = self._fmt
        ok = (c == 4) if f in (SourceFormat.Rgba8, SourceFormat.Bgra8) else f.interleaved()
        if not ok:
            raise PreprocessError("FormatNeedsRawBuffer", f"source format {f!r} needs run_raw (raw device buffer), not the typed run()")

    def _validate_raw(self, got: int, w: int, h: int) -> None:  # :1287-1300
        f = self._fmt
        need = f.buffer_len(w, h)
        if not f.dims_ok(w, h) or got < need:
            raise PreprocessError("InvalidRawSource", f"invalid raw source for {f!r} at {w}x{h} (got {got} bytes, need {need})",
                                  format=f, width=w, height=h, got=got, need=need)

    # ── typed entries (interleaved RGB/BGR[A] images) ───────────────────────
    def _run_typed(self, src: Image, dst: torch.Tensor, f16: bool) -> None:
        c = src.num_channels()
        if c not in (3, 4):
            raise PreprocessError("UnsupportedChannels", f"unsupported source channel count {c} (expected 3 or 4)")
        self._validate_dst_shape_only(dst)
        self._validate_typed_format(c)
        if not src.is_device:
            raise PreprocessError("NotDeviceImage", "CUDA preprocessor requires a device-resident source image")
        self._validate_dst(dst, 1, f16)
        if src.dtype != torch.uint8 or src.is_batched:
            raise PreprocessError("UnsupportedChannels", "typed run() takes one u8 HWC image")
        w, h = src.width(), src.height()
        desc = self._desc(w, h, w * c, c, self._fmt.fmt_code(), dst.shape[3], dst.shape[2])
        self._launch(desc, [src.data], dst, f16)

    @staticmethod
    def _validate_dst_shape_only(dst: torch.Tensor) -> None:
        shape = tuple(dst.shape)
        if dst.dim() != 4 or shape[1] != 3 or shape[0] != 1:
            raise PreprocessError("BadOutputShape", f"destination tensor must be [1, 3, H, W], got {list(shape)}")

    def run(self, src: Image, dst: torch.Tensor) -> None:
        """preprocess.rs:887 — `src` u8 HWC C∈{3,4} → `dst` [1,3,H,W] f32."""
        self._run_typed(src, dst, False)

    def run_f16(self, src: Image, dst: torch.Tensor) -> None:
        """preprocess.rs:1086."""
        self._run_typed(src, dst, True)

    # ── pitched surfaces ────────────────────────────────────────────────────
    def _run_surface(self, src: PitchedSurface, dst: torch.Tensor, f16: bool) -> None:
        src.validate()
        self._validate_dst(dst, 1, f16)
        if not self._fmt.interleaved():  # :1150-1156
            raise PreprocessError("FormatNeedsRawBuffer", f"source format {self._fmt!r} needs run_raw (raw device buffer), not the typed run()")
        desc = self._desc(src.width, src.height, src.row_pitch, src.channels, self._fmt.fmt_code(), dst.shape[3], dst.shape[2])
        self._launch(desc, [src.data], dst, f16)

    def run_surface(self, src: PitchedSurface, dst: torch.Tensor) -> None:
        """preprocess.rs:1112."""
        self._run_surface(src, dst, False)

    def run_surface_f16(self, src: PitchedSurface, dst: torch.Tensor) -> None:
        self._run_surface(src, dst, True)

    # ── raw camera buffers ──────────────────────────────────────────────────
    def _run_raw_batch(self, frames: Sequence[torch.Tensor], src_w: int, src_h: int, dst: torch.Tensor, f16: bool,
                       single: bool) -> None:
        self._validate_dst(dst, 1 if single else len(frames), f16)
        for f in frames:
            self._validate_raw(f.numel() * f.element_size(), src_w, src_h)
        fmt = self._fmt
        desc = self._desc(src_w, src_h, fmt.pitch(src_w), fmt.bpp(), fmt.fmt_code(), dst.shape[3], dst.shape[2])
        self._launch(desc, list(frames), dst, f16)

    def run_raw(self, src: torch.Tensor, src_w: int, src_h: int, dst: torch.Tensor) -> None:
        """preprocess.rs:1184 — one raw device frame in the builder's SourceFormat → [1,3,H,W]."""
        self._run_raw_batch([src], src_w, src_h, dst, False, True)

    def run_raw_f16(self, src: torch.Tensor, src_w: int, src_h: int, dst: torch.Tensor) -> None:
        self._run_raw_batch([src], src_w, src_h, dst, True, True)

    def run_raw_batch(self, frames: Sequence[torch.Tensor], src_w: int, src_h: int, dst: torch.Tensor) -> None:
        """preprocess.rs:1234 — N same-sized raw frames → [N,3,H,W], ONE launch (≤256 frames per launch)."""
        self._run_raw_batch(frames, src_w, src_h, dst, False, False)

    def run_raw_batch_f16(self, frames: Sequence[torch.Tensor], src_w: int, src_h: int, dst: torch.Tensor) -> None:
        self._run_raw_batch(frames, src_w, src_h, dst, True, False)

    def run_raw_host(self, host_frames: torch.Tensor, src_w: int, src_h: int, dst_size: tuple[int, int], out: torch.Tensor | None = None,
                     f16: bool = False, pipeline=None) -> torch.Tensor:
        """HOST camera frames in, HOST tensor out — what the reference's Python `Preprocessor` does around its kernel (pinned
        staging, upload, launch: kornia-py/src/cuda_ext/mod.rs:700-760), as one call: `host_frames` is a [N, frame_bytes] uint8
        host tensor (page-locked for overlap), the result a [N,3,H,W] f32 / f16 host tensor (pinned when allocated here).
        Chunks ride a `HostPipeline` ring: upload -> ONE fused launch per chunk -> download.  Enqueue-only: synchronise the
        current stream of the pipeline's device before reading the result."""
        from . import imgproc

        t = host_frames
        if t.is_cuda:
            raise PreprocessError("NotDeviceImage", "run_raw_host takes host frames; use run_raw_batch / run_raw_strided for device frames")
        if t.dtype != torch.uint8 or t.dim() != 2 or not t.is_contiguous():
            raise PreprocessError("InvalidRawSource", "host frames must be a contiguous [N, frame_bytes] uint8 tensor")
        n, stride = t.shape
        self._validate_raw(stride, src_w, src_h)
        dw, dh = dst_size
        want = torch.float16 if f16 else torch.float32
        if out is None:
            out = torch.empty((n, 3, dh, dw), dtype=want, pin_memory=True)
        if out.is_cuda or out.dtype != want or not out.is_contiguous() or tuple(out.shape) != (n, 3, dh, dw):
            raise PreprocessError("BadOutputShape", f"destination must be a contiguous host {want} tensor of shape {[n, 3, dh, dw]}")
        fmt = self._fmt
        desc = self._desc(src_w, src_h, fmt.pitch(src_w), fmt.bpp(), fmt.fmt_code(), dw, dh)
        pipe = imgproc._pipeline_for(pipeline)
        _lib.set_device(pipe.device.index)
        st = _lib.lib().kb200_preprocess_host(pipe._h, torch.cuda.current_stream(pipe.device).cuda_stream, C.byref(desc), t.data_ptr(), t.numel(),
                                             stride, n, out.data_ptr(), out.numel(), 1 if f16 else 0)
        if st != _lib.OK:
            kind = {_lib.ERR_INVALID_SOURCE: "InvalidRawSource"}.get(st, "Cuda")
            raise PreprocessError(kind, _lib.last_error())
        return out

    def run_raw_strided(self, base: torch.Tensor, frame_stride: int, batch: int, src_w: int, src_h: int,
                        dst: torch.Tensor, f16: bool = False) -> None:
        """Frames at `base + i*frame_stride` bytes (a capture ring buffer): no pointer table."""
        self._validate_dst(dst, batch, f16)
        fmt = self._fmt
        self._validate_raw(min(frame_stride, base.numel()) if batch > 1 else base.numel(), src_w, src_h)
        desc = self._desc(src_w, src_h, fmt.pitch(src_w), fmt.bpp(), fmt.fmt_code(), dst.shape[3], dst.shape[2])
        if not base.is_cuda:
            raise PreprocessError("NotDeviceImage", "CUDA preprocessor requires a device-resident source image")
        dev = dst.device
        _lib.set_device(dev.index)
        l = _lib.lib()
        fn = l.kb200_preprocess_strided_f16 if f16 else l.kb200_preprocess_strided_f32
        st = fn(torch.cuda.current_stream(dev).cuda_stream, C.byref(desc), base.data_ptr(), base.numel(), frame_stride, batch,
                dst.data_ptr(), dst.numel())
        if st != _lib.OK:
            kind = {_lib.ERR_INVALID_SOURCE: "InvalidRawSource"}.get(st, "Cuda")
            raise PreprocessError(kind, _lib.last_error())
