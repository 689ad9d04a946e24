"""Device runtime: thin Python over the C ABI (include/tmx.h).  PyTorch supplies
device memory and the current CUDA stream only; every arithmetic op on the hot
path is a libtmx kernel.  Nothing here falls back to torch ops or the CPU."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

LRELU_ALPHA = 0.2


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Act:
    """A device activation in one or both internal layouts.
    f32: torch.float32 [N,H,W,C] (NHWC) ; hi/lo: torch.bfloat16 [N,H+2,W+2,C]
    (SPLIT_BF16_HALO: x ~= hi + lo, halo materialised: 'reflect' for a 3x3
    consumer, 'replicate' for a consumer that reads through upscale2d)."""
    __slots__ = ('n', 'h', 'w', 'c', 'f32', 'hi', 'lo', 'halo')

    def __init__(self, n, h, w, c, f32=None, hi=None, lo=None, halo='reflect'):
        self.n, self.h, self.w, self.c = n, h, w, c
        self.f32, self.hi, self.lo, self.halo = f32, hi, lo, halo

    @property
    def shape_nchw(self):
        return (self.n, self.c, self.h, self.w)


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                                     # pragma: no cover - older/newer torch without the accessor
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


HALO_KINDS = {'reflect': 0, 'replicate': 1, 'zero': 2}     # tmx_split_halo_pack


class Runtime:
    """One per (process, device).  Holds the tmx handle; launches on torch's
    current stream so that CUDA-graph capture and stream semantics are torch's."""
    _instances = {}

    @classmethod
    def get(cls, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('texturemixer_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in cls._instances:
            cls._instances[idx] = cls(idx)
        return cls._instances[idx]

    def __init__(self, index):
        self.lib = _lib.load()
        self.index = index
        self.device = torch.device('cuda', index)
        h = C.c_void_p()
        _lib.check(self.lib.tmx_create(index, C.byref(h)), 'tmx_create')
        self.handle = h
        sm, major, minor = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.tmx_device_info(h, C.byref(sm), C.byref(major), C.byref(minor)))
        self.sm_count, self.cc = sm.value, (major.value, minor.value)
        # kernel selection override for tests/benchmarks: auto | ffma | tc | tc_k32
        self.conv_algo = os.environ.get('TMX_CONV_ALGO', 'auto')
        # bench.py: CUDA events around each conv launch -> [(tag, start, end)], tag = (algo, k, cin, cout, h, w)
        self.profile_kernels = False
        self.kernel_events = []

    # ------------------------------------------------------------------ helpers
    def stream(self):
        # torch's CURRENT stream of this device as a raw cudaStream_t (the private accessor is ~20x cheaper than
        # building a torch.cuda.Stream object; it is called once per launch, ~1 800 times per train step)
        return C.c_void_p(_raw_stream(self.index))

    def launch_count(self):
        v = C.c_uint64()
        _lib.check(self.lib.tmx_launch_count(self.handle, C.byref(v)))
        return int(v.value)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ layout
    def nchw_to_nhwc(self, x, out=None, c_off=0, c_total=None, bcast_hw=None):
        """x: NCHW f32 device tensor ([N,C,1,1] with bcast_hw=(H,W)) -> NHWC slice."""
        assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
        n, c = x.shape[0], x.shape[1]
        h, w = (x.shape[2], x.shape[3]) if bcast_hw is None else bcast_hw
        c_total = c if c_total is None else c_total
        if out is None:
            out = self.empty(n, h, w, c_total)
        _lib.check(self.lib.tmx_nchw_to_nhwc(self.handle, _ptr(x), _ptr(out), n, c, h, w, c_off, c_total,
                                             0 if bcast_hw is None else 1, self.stream()), 'tmx_nchw_to_nhwc')
        return out

    def nhwc_to_nchw(self, x, c_off=0, c=None):
        n, h, w, c_total = x.shape
        c = c_total - c_off if c is None else c
        out = self.empty(n, c, h, w)
        _lib.check(self.lib.tmx_nhwc_to_nchw(self.handle, _ptr(x), _ptr(out), n, c, h, w, c_off, c_total,
                                             self.stream()), 'tmx_nhwc_to_nchw')
        return out

    def planes(self, n, hp, wp, c):
        """One bf16 plane [n,hp,wp,c] with 64 zeroed elements of slack behind it: the X-MERGED operand view of the
        16-channel layers (TMX_CONV_XMERGE) reads up to 3 pixels past the last one (times a zero weight)."""
        numel = n * hp * wp * c
        buf = torch.empty(numel + 64, dtype=torch.bfloat16, device=self.device)
        if c == 16:                 # only the 16-channel layers are ever read X-MERGED (use_xmerge)
            buf[numel:].zero_()
        return buf[:numel].view(n, hp, wp, c)

    def split_pack(self, act, halo='reflect'):
        """Make sure `act` carries split planes with the requested halo kind."""
        if act.hi is None or act.halo != halo:
            if act.f32 is None:
                self.split_unpack(act)
            act.hi = self.planes(act.n, act.h + 2, act.w + 2, act.c)
            act.lo = self.planes(act.n, act.h + 2, act.w + 2, act.c)
            act.halo = halo
            _lib.check(self.lib.tmx_split_halo_pack(self.handle, _ptr(act.f32), _ptr(act.hi), _ptr(act.lo), act.n,
                                                    act.h, act.w, act.c, HALO_KINDS[halo], self.stream()),
                       'tmx_split_halo_pack')
        return act

    def split_unpack(self, act):
        if act.f32 is None:
            assert act.hi is not None
            act.f32 = self.empty(act.n, act.h, act.w, act.c)
            _lib.check(self.lib.tmx_split_halo_unpack(self.handle, _ptr(act.hi), _ptr(act.lo), _ptr(act.f32), act.n,
                                                      act.h, act.w, act.c, self.stream()), 'tmx_split_halo_unpack')
        return act

    # ------------------------------------------------------------------ conv
    def choose_algo(self, cin, cout, k, up2, hw=None):
        """Tensor cores for every conv whose channel counts fit the UMMA tile
        (multiples of 16) and whose contraction is at least one 3x3x16 window;
        the CUDA-core kernel keeps the rest (and is the exact-fp32 cross-check)."""
        mode = self.conv_algo
        tc_ok = cin % 16 == 0 and cout % 16 == 0 and (k == 3 or not up2)
        if hw is not None and (hw[0] < 2 or hw[1] < 2):
            tc_ok = False              # the halo layout needs at least 2x2 stored pixels
        if mode == 'ffma' or not tc_ok:
            return _lib.ALGO_FFMA
        if mode == 'tc':
            return _lib.ALGO_TC
        if mode == 'tc_k32':
            return _lib.ALGO_TC_K32
        if k * k * cin >= 144 or (k == 1 and cin >= 64):
            return _lib.ALGO_TC
        return _lib.ALGO_FFMA

    def use_xmerge(self, cin, k, up2):
        return cin == 16 and k == 3 and not up2 and not os.environ.get('TMX_NO_XMERGE')

    def prepare_weights_xmerge(self, w, wscale, cout):
        hi = self.empty(cout, 192, dtype=torch.bfloat16)
        lo = self.empty(cout, 192, dtype=torch.bfloat16)
        _lib.check(self.lib.tmx_conv_weights_prepare_xmerge(self.handle, _ptr(w), float(wscale), int(w.shape[2]), cout,
                                                            _ptr(hi), _ptr(lo), self.stream()),
                   'tmx_conv_weights_prepare_xmerge')
        return hi, lo

    def prepare_weights(self, w, wscale, k, cin, cout, up2_phase=False, cin_pad=None):
        rows = cout * 4 if up2_phase else cout
        cin_pad = cin if cin_pad is None else cin_pad
        hi = self.empty(rows, k * k * cin_pad, dtype=torch.bfloat16)
        lo = self.empty(rows, k * k * cin_pad, dtype=torch.bfloat16)
        _lib.check(self.lib.tmx_conv_weights_prepare(self.handle, _ptr(w), float(wscale), k, cin, cin_pad, cout,
                                                     int(up2_phase), _ptr(hi), _ptr(lo), self.stream()),
                   'tmx_conv_weights_prepare')
        return hi, lo

    def conv2d(self, x, w, bias, wscale, k, cout, lrelu=False, residual=None, up2=False, want_f32=True,
               want_split=False, up2_out=False, halo_out='reflect', algo=None, prepared=None, torgb=None,
               halo_in=None, alpha=None, per_sample_weights=False):
        """y = [residual +] lrelu(wscale*conv(x, w) + bias) on an Act.  `w` is the raw
        HWIO variable; `prepared` an optional (w_hi, w_lo) pair for the TC kernel
        (sub-pixel planes when up2).  `torgb` = (w_rgb [Cout,C], b_rgb, wscale, C, tanh)
        fuses the 1x1 image head behind the conv (TC only) and returns (out, images).
        `per_sample_weights`: `prepared` holds one weight set per image, [N*Cout][k*k*Cin] (TMX_CONV_W_PER_SAMPLE)."""
        cin = x.c
        h, w_ = (x.h * 2, x.w * 2) if up2 else (x.h, x.w)
        if algo is None:
            algo = self.choose_algo(cin, cout, k, up2, (x.h, x.w))
        d = _lib.ConvDesc(N=x.n, H=h, W=w_, Cin=cin, Cout=cout, k=k, flags=0, algo=algo, wscale=float(wscale),
                          lrelu_alpha=LRELU_ALPHA if alpha is None else float(alpha))     # alpha 0 = ReLU (VGG-19)
        io = _lib.ConvIO()
        flags = 0
        if lrelu:
            flags |= _lib.CONV_LRELU
        if residual is not None:
            flags |= _lib.CONV_RESIDUAL
            io.residual = residual.data_ptr()
        if up2:
            flags |= _lib.CONV_UP2_IN
        if per_sample_weights:
            assert prepared is not None and algo != _lib.ALGO_FFMA
            flags |= _lib.CONV_W_PER_SAMPLE
        io.bias = None if bias is None else bias.data_ptr()
        keep = []
        images = None
        if algo == _lib.ALGO_FFMA:
            if torgb is not None:
                raise RuntimeError('conv2d: the fused ToRGB head needs the tensor-core kernel')
            if halo_in not in (None, 'reflect'):
                raise RuntimeError('conv2d: the CUDA-core kernel pads by REFLECT only (halo_in=%r)' % halo_in)
            self.split_unpack(x)
            io.x_f32 = x.f32.data_ptr()
            io.w = w.data_ptr()
            out = Act(x.n, h, w_, cout, f32=self.empty(x.n, h, w_, cout))
            io.y_f32 = out.f32.data_ptr()
        else:
            # halo_in='zero': the SAME (zero) padding of the fused_scale convs instead of the REFLECT default
            self.split_pack(x, halo_in or ('replicate' if up2 else 'reflect'))
            io.x_hi, io.x_lo = x.hi.data_ptr(), x.lo.data_ptr()
            xmerge = self.use_xmerge(cin, k, up2) and x.hi.untyped_storage().nbytes() >= (x.hi.numel() + 64) * 2 \
                and not per_sample_weights
            if xmerge:
                flags |= _lib.CONV_XMERGE
                if prepared is None or prepared[0].shape[1] != 192:
                    prepared = self.prepare_weights_xmerge(w, wscale, cout)
            elif prepared is not None and prepared[0].shape[1] == 192 and k * k * cin != 192:
                prepared = None          # X-MERGED planes handed in, but this call cannot use them
            if prepared is None:
                prepared = self.prepare_weights(w, wscale, k, cin, cout, up2_phase=up2)
            keep.append(prepared)
            io.w_hi, io.w_lo = prepared[0].data_ptr(), prepared[1].data_ptr()
            out = Act(x.n, h, w_, cout)
            if torgb is not None:
                w_rgb, b_rgb, ws_rgb, c_rgb, tanh_rgb = torgb
                images = self.empty(x.n, c_rgb, h, w_)
                flags |= _lib.CONV_TORGB
                d.rgb_cout, d.rgb_tanh, d.rgb_wscale = c_rgb, int(tanh_rgb), float(ws_rgb)
                io.rgb_w = w_rgb.data_ptr()
                io.rgb_b = None if b_rgb is None else b_rgb.data_ptr()
                io.y_rgb = images.data_ptr()
            if want_f32 or (not want_split and torgb is None):
                out.f32 = self.empty(x.n, h, w_, cout)
                io.y_f32 = out.f32.data_ptr()
            if want_split:
                ho, wo = (h * 2, w_ * 2) if up2_out else (h, w_)
                out.hi = self.planes(x.n, ho + 2, wo + 2, cout)
                out.lo = self.planes(x.n, ho + 2, wo + 2, cout)
                if halo_out == 'zero':       # SAME-padded consumer: the border pixels' threads write the zero ring
                    flags |= _lib.CONV_HALO_ZERO
                out.halo = halo_out
                io.y_hi, io.y_lo = out.hi.data_ptr(), out.lo.data_ptr()
                if up2_out:
                    flags |= _lib.CONV_UP2_OUT
                if halo_out == 'replicate':
                    flags |= _lib.CONV_HALO_REPLICATE
        d.flags = flags
        if self.profile_kernels:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(self.lib.tmx_conv2d_fwd(self.handle, C.byref(d), C.byref(io), self.stream()), 'tmx_conv2d_fwd')
        if self.profile_kernels:
            e1.record()
            self.kernel_events.append((('ffma' if algo == _lib.ALGO_FFMA else 'tc', k, cin, cout, x.h, x.w), e0, e1))
        if want_split and out.hi is None:
            self.split_pack(out, halo_out)
        return (out, images) if torgb is not None else out

    # ------------------------------------------------------------------ backward
    def transpose_weights(self, prepared, rows, taps, kdim):
        """Forward planes [rows][taps*kdim] -> data-gradient planes [kdim][taps*rows] (taps flipped)."""
        hi = self.empty(kdim, taps * rows, dtype=torch.bfloat16)
        lo = self.empty(kdim, taps * rows, dtype=torch.bfloat16)
        _lib.check(self.lib.tmx_conv_weights_transpose(self.handle, _ptr(prepared[0]), _ptr(prepared[1]), rows, taps,
                                                       kdim, _ptr(hi), _ptr(lo), self.stream()),
                   'tmx_conv_weights_transpose')
        return hi, lo

    def conv_dgrad(self, dz, n, h, w, cin, cout, k, wt):
        """dz = (hi, lo) planes on the zero-ringed grid [n][h+4][w+4][cout]; wt = transposed weight planes.
        -> fp32 g on the grid [n][h+4][w+4][cin] (ring = what the padding adjoint folds back)."""
        g = self.empty(n, h + 4, w + 4, cin)
        _lib.check(self.lib.tmx_conv2d_dgrad(self.handle, n, h, w, cin, cout, k, _ptr(dz[0]), _ptr(dz[1]),
                                             _ptr(wt[0]), _ptr(wt[1]), _ptr(g), self.stream()), 'tmx_conv2d_dgrad')
        return g

    def conv_dgrad_gp(self, dz, n, h, w, cin, cout, k, wt, fold, add=None, y_f32=None, y_hi=None, want_f32=False,
                      dbias=None, alpha=None, want_planes=True):
        """tmx_conv2d_dgrad_gp: the data gradient of a conv AND the grad_prepare of its input activation [n,h,w,cin]
        (padding adjoint `fold`, addend, mask of the layer that produced the activation, bias gradient, planes) in one
        tensor-core kernel plus a border pass.  Returns (planes, f32 or None), or None when the library does not serve
        the shape that way (nothing was launched: run conv_dgrad + grad_prepare)."""
        d = _lib.GradDesc(N=n, H=h, W=w, C=cin, src_kind=0, fold=fold, mask_kind=0, phase_pack=0,
                          alpha=LRELU_ALPHA if alpha is None else float(alpha), dbias_scale=1.0)
        io = _lib.GradIO()
        if add is not None:
            io.add = add.data_ptr()
        if y_f32 is not None:
            d.mask_kind, io.y_mask = 1, y_f32.data_ptr()
        elif y_hi is not None:
            d.mask_kind, io.y_mask = 2, y_hi.data_ptr()
        g = self.empty(n, h + 4, w + 4, cin)
        planes = f32 = None
        if want_planes:
            planes = (self.empty(n, h + 4, w + 4, cin, dtype=torch.bfloat16),
                      self.empty(n, h + 4, w + 4, cin, dtype=torch.bfloat16))
            io.dz_hi, io.dz_lo = planes[0].data_ptr(), planes[1].data_ptr()
        if want_f32:
            f32 = self.empty(n, h, w, cin)
            io.dz_f32 = f32.data_ptr()
        if dbias is not None:
            io.dbias = dbias.data_ptr()
        served = C.c_int(0)
        _lib.check(self.lib.tmx_conv2d_dgrad_gp(self.handle, n, h, w, cin, cout, k, _ptr(dz[0]), _ptr(dz[1]),
                                                _ptr(wt[0]), _ptr(wt[1]), _ptr(g), C.byref(d), C.byref(io),
                                                C.byref(served), self.stream()), 'tmx_conv2d_dgrad_gp')
        if not served.value:
            return None
        return planes, f32

    def conv_wgrad(self, x, dz, n, h, w, cin, cout, k, wscale, dw):
        """dw (HWIO fp32 view of the gradient buffer) += wscale * x^T dz; x = (hi, lo) forward input planes
        [n][h+2][w+2][cin], dz = (hi, lo) planes on the zero-ringed grid [n][h+4][w+4][cout]."""
        # planes from Runtime.planes() carry 64 elements of slack: the thin layers may read their taps PACKED
        flags = 0
        if k == 3 and cin in (16, 32) and all(t.untyped_storage().nbytes() - t.storage_offset() * 2 >=
                                              (t.numel() + 64) * 2 for t in x):
            flags = _lib.WGRAD_X_SLACK
        nbytes = C.c_size_t()
        _lib.check(self.lib.tmx_conv2d_wgrad_workspace_bytes(self.handle, n, h, w, cin, cout, k, flags, C.byref(nbytes)),
                   'tmx_conv2d_wgrad_workspace_bytes')
        ws = self.empty(max(1, nbytes.value // 4))
        _lib.check(self.lib.tmx_conv2d_wgrad(self.handle, n, h, w, cin, cout, k, float(wscale), _ptr(x[0]), _ptr(x[1]),
                                             _ptr(dz[0]), _ptr(dz[1]), _ptr(dw), _ptr(ws), flags, self.stream()),
                   'tmx_conv2d_wgrad')
        return dw

    def grad_prepare(self, g, n, h, w, c, src_kind, fold=2, add=None, y_f32=None, y_hi=None, want_planes=True,
                     want_f32=False, dbias=None, dbias_scale=1.0, phase_pack=False, alpha=None):
        """See tmx_grad_prepare (include/tmx.h).  Returns (planes or None, f32 or None)."""
        d = _lib.GradDesc(N=n, H=h, W=w, C=c, src_kind=src_kind, fold=fold, mask_kind=0, phase_pack=int(phase_pack),
                          alpha=LRELU_ALPHA if alpha is None else float(alpha), dbias_scale=float(dbias_scale))
        io = _lib.GradIO()
        io.g = g.data_ptr()
        if add is not None:
            io.add = add.data_ptr()
        if y_f32 is not None:
            d.mask_kind, io.y_mask = 1, y_f32.data_ptr()
        elif y_hi is not None:
            d.mask_kind, io.y_mask = 2, y_hi.data_ptr()
        planes = f32 = None
        if want_planes:
            if phase_pack:   # half-resolution grid; its ring is never written by the kernel: start from zeros
                shape = (n, h // 2 + 4, w // 2 + 4, 4 * c)
                planes = (torch.zeros(shape, dtype=torch.bfloat16, device=self.device),
                          torch.zeros(shape, dtype=torch.bfloat16, device=self.device))
            else:
                planes = (self.empty(n, h + 4, w + 4, c, dtype=torch.bfloat16),
                          self.empty(n, h + 4, w + 4, c, dtype=torch.bfloat16))
            io.dz_hi, io.dz_lo = planes[0].data_ptr(), planes[1].data_ptr()
        if want_f32:
            f32 = self.empty(n, h, w, c)
            io.dz_f32 = f32.data_ptr()
        if dbias is not None:
            io.dbias = dbias.data_ptr()
        _lib.check(self.lib.tmx_grad_prepare(self.handle, C.byref(d), C.byref(io), self.stream()), 'tmx_grad_prepare')
        return planes, f32

    # ------------------------------------------------------------------ pointwise
    def fromrgb(self, x_nchw, w, bias, wscale, cout, lrelu=True):
        n, cin, h, w_ = x_nchw.shape
        out = Act(n, h, w_, cout, f32=self.empty(n, h, w_, cout))
        _lib.check(self.lib.tmx_fromrgb_fwd(self.handle, _ptr(x_nchw), _ptr(w), _ptr(bias), float(wscale),
                                            _ptr(out.f32), n, cin, h, w_, cout, int(lrelu), LRELU_ALPHA,
                                            self.stream()), 'tmx_fromrgb_fwd')
        return out

    def torgb(self, x, w, bias, wscale, cout, apply_tanh):
        self.split_unpack(x)
        out = self.empty(x.n, cout, x.h, x.w)
        _lib.check(self.lib.tmx_torgb_fwd(self.handle, _ptr(x.f32), _ptr(w), _ptr(bias), float(wscale), _ptr(out),
                                          x.n, x.h, x.w, x.c, cout, int(apply_tanh), self.stream()), 'tmx_torgb_fwd')
        return out

    def avgpool2(self, x, pack=None):
        """downscale2d by 2.  `pack` = halo kind ('reflect' | 'replicate' | 'zero'): the consumer is a tensor-core conv -
        the pooled map is also written as split planes with that halo by the same kernel (tmx_avgpool2_pack)."""
        self.split_unpack(x)
        if pack is not None and x.h >= 4 and x.w >= 4 and x.c % 8 == 0 and not os.environ.get('TMX_NO_POOL_PACK'):
            out = Act(x.n, x.h // 2, x.w // 2, x.c, f32=self.empty(x.n, x.h // 2, x.w // 2, x.c))
            out.hi = self.planes(out.n, out.h + 2, out.w + 2, out.c)
            out.lo = self.planes(out.n, out.h + 2, out.w + 2, out.c)
            out.halo = pack
            _lib.check(self.lib.tmx_avgpool2_pack(self.handle, _ptr(x.f32), _ptr(out.f32), _ptr(out.hi), _ptr(out.lo),
                                                  x.n, x.h, x.w, x.c, HALO_KINDS[pack], self.stream()),
                       'tmx_avgpool2_pack')
            return out
        out = Act(x.n, x.h // 2, x.w // 2, x.c, f32=self.empty(x.n, x.h // 2, x.w // 2, x.c))
        _lib.check(self.lib.tmx_avgpool2_fwd(self.handle, _ptr(x.f32), _ptr(out.f32), x.n, x.h, x.w, x.c,
                                             self.stream()), 'tmx_avgpool2_fwd')
        return out

    def mbstd(self, x, group_size):
        """minibatch_stddev_layer: Act [N,H,W,C] -> Act [N,H,W,C+1 padded to a multiple of 64]; the logical
        channel count (C+1) is kept in `.c_logical` for the consumer's weight layout."""
        self.split_unpack(x)
        c_total = (x.c + 1 + 63) // 64 * 64          # 64-channel K chunks for the consumer conv
        g = min(group_size, x.n)
        out = Act(x.n, x.h, x.w, c_total, f32=self.empty(x.n, x.h, x.w, c_total))
        stat = self.empty(x.n // g)
        _lib.check(self.lib.tmx_mbstd_fwd(self.handle, _ptr(x.f32), _ptr(out.f32), _ptr(stat), x.n, x.h, x.w, x.c,
                                          c_total, group_size, self.stream()), 'tmx_mbstd_fwd')
        return out, stat

    def dense(self, x, w, bias, wscale, lrelu):
        """x [N,K] fp32 device tensor, w [K,Cout] raw variable -> [N,Cout]."""
        n, k = x.shape
        cout = w.shape[1]
        nbytes = C.c_size_t()
        _lib.check(self.lib.tmx_dense_workspace_bytes(n, k, cout, C.byref(nbytes)))
        ws = self.empty(max(1, nbytes.value // 4))
        out = self.empty(n, cout)
        _lib.check(self.lib.tmx_dense_fwd(self.handle, _ptr(x), _ptr(w), _ptr(bias), float(wscale), _ptr(out), _ptr(ws),
                                          n, k, cout, int(lrelu), LRELU_ALPHA, self.stream()), 'tmx_dense_fwd')
        return out

    def bias_act(self, x, bias, lrelu):
        """[lrelu](x + bias) on an Act's fp32 map (tmx_bias_act)."""
        self.split_unpack(x)
        out = Act(x.n, x.h, x.w, x.c, f32=self.empty(x.n, x.h, x.w, x.c))
        _lib.check(self.lib.tmx_bias_act(self.handle, _ptr(x.f32), _ptr(bias), _ptr(out.f32), x.n * x.h * x.w, x.c,
                                         int(lrelu), LRELU_ALPHA, self.stream()), 'tmx_bias_act')
        return out

    # ------------------------------------------------------------------ windows (crop-aware train step)
    def window(self, x, win, nhwc=False):
        """x[..., oy:oy+h, ox:ox+w] (NCHW) or x[:, oy:oy+h, ox:ox+w, :] (NHWC) as a new contiguous tensor.  `win` =
        (oy, ox, h, w); when it carries a `.dev` int32 device view {oy, ox} the kernel reads the offset from there
        (identical launch every step -> CUDA-graph replay, see loss.Window)."""
        oy, ox, h, w = win
        if nhwc:
            n, H, W, c = x.shape
            A, B, out = n, c, self.empty(n, h, w, c)
        else:
            n, c, H, W = x.shape
            A, B, out = n * c, 1, self.empty(n, c, h, w)
        dev = getattr(win, 'dev', None)
        _lib.check(self.lib.tmx_window_copy(self.handle, _ptr(x.contiguous()), _ptr(out), A, H, W, B, h, w, int(oy),
                                            int(ox), _ptr(dev), 0, self.stream()), 'tmx_window_copy')
        return out

    def window_embed(self, d, win, H, W, nhwc=False):
        """Adjoint of `window`: `d` placed at the window's offset inside a zero [.., H, W] tensor."""
        oy, ox, h, w = win
        if nhwc:
            n, dh, dw, c = d.shape
            A, B, out = n, c, self.empty(n, H, W, c)
        else:
            n, c, dh, dw = d.shape
            A, B, out = n * c, 1, self.empty(n, c, H, W)
        assert (dh, dw) == (h, w), ((dh, dw), tuple(win))
        dev = getattr(win, 'dev', None)
        _lib.check(self.lib.tmx_window_copy(self.handle, _ptr(d.contiguous()), _ptr(out), A, H, W, B, h, w, int(oy),
                                            int(ox), _ptr(dev), 1, self.stream()), 'tmx_window_copy')
        return out

    # ------------------------------------------------------------------ latent blend
    def latent_blend(self, srcs, H, W, mode, idx_h=None, idx_w=None, ramps_h=None, ramps_w=None, t=None,
                     pin_rows=0, pin_cols=0, src_reverse=0, math_f32=False, out_nchw=True, out_nhwc=None,
                     c_off=0, c_total=None):
        """srcs: list of NCHW f32 device tensors ([N,C,h,w] or [N,C,1,1] broadcast)."""
        k = len(srcs)
        n, c, h, w = srcs[0].shape
        bcast = (h == 1 and w == 1 and (H > 1 or W > 1))
        d = _lib.BlendDesc(N=n, C=c, h=h, w=w, H=H, W=W, K=k, mode=mode, math_f32=int(math_f32),
                           src_bcast=int(bcast), src_reverse=src_reverse, pin_rows=pin_rows, pin_cols=pin_cols,
                           c_off=c_off, C_total=c if c_total is None else c_total)
        io = _lib.BlendIO()
        for i, s in enumerate(srcs):
            assert s.dtype == torch.float32 and s.is_contiguous() and tuple(s.shape) == (n, c, h, w)
            io.src[i] = s.data_ptr()
            if idx_h is not None and idx_h[i] is not None:
                assert idx_h[i].dtype == torch.int32 and tuple(idx_h[i].shape) == (n, H)
                io.idx_h[i] = idx_h[i].data_ptr()
            if idx_w is not None and idx_w[i] is not None:
                assert idx_w[i].dtype == torch.int32 and tuple(idx_w[i].shape) == (n, W)
                io.idx_w[i] = idx_w[i].data_ptr()
            if ramps_h is not None:
                assert ramps_h[i].dtype == torch.float64 and ramps_h[i].numel() == H
                assert ramps_w[i].dtype == torch.float64 and ramps_w[i].numel() == W
                io.ramp_h[i] = ramps_h[i].data_ptr()
                io.ramp_w[i] = ramps_w[i].data_ptr()
        if t is not None:
            assert t.dtype == torch.float32 and t.numel() == n
            io.t = t.data_ptr()
        res_nchw = None
        if out_nchw:
            res_nchw = self.empty(n, c, H, W)
            io.out_nchw = res_nchw.data_ptr()
        if out_nhwc is not None:
            io.out_nhwc = out_nhwc.data_ptr()
        _lib.check(self.lib.tmx_latent_blend(self.handle, C.byref(d), C.byref(io), self.stream()),
                   'tmx_latent_blend')
        return res_nchw


# ---------------------------------------------------------------------- host sampler
def perm_indices_from_uniforms(u, length, levels, count):
    """Index vectors of the hierarchical swap permutation (run.py:107-182, 436-507)
    from pre-drawn uniforms.  Host only.  -> (int32 [count, length], consumed)."""
    lib = _lib.load()
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty((count, length), np.int32)
    used = C.c_int64()
    _lib.check(lib.tmx_perm_indices_from_uniforms(u.ctypes.data_as(C.POINTER(C.c_double)), u.size, length, levels,
                                                  count, out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(used)),
               'tmx_perm_indices_from_uniforms')
    return out, int(used.value)


def uniforms_per_matrix(length, levels):
    return sum(2 * (length >> lvl) for lvl in range(levels) if (length >> lvl) > 1)
