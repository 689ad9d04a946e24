#!/usr/bin/env python
"""One launch (after one warm-up) of every kernel FAMILY of the hot path at the shape it has in the train step /
cfg 4, for `ncu --set full` captures and for a CUDA-event roofline table.

    python profiles/kernel_targets.py                 # CUDA-event table -> stdout (achieved TFLOP/s | GB/s, frac)
    ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|conv_wgrad_kernel|grad_prepare_kernel|latent_blend_kernel|adam_kernel|window_copy_kernel' \
        -o gpurun_out/r02_kernels python profiles/kernel_targets.py --once

Algorithmic work per launch: convs 2*9*Cin*Cout FLOP per output pixel (bf16x3 executes 3x);
HBM-bound kernels: the bytes each must read and write once."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from texturemixer_b200 import _lib, interp  # noqa: E402
from texturemixer_b200.runtime import Act, Runtime  # noqa: E402


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d['bf16_tflops'], 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--once', action='store_true', help='one warm-up + one launch per target (profiler runs)')
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    rt = Runtime.get(0)
    dev = rt.device
    g = torch.Generator(device='cpu').manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    targets = []

    def rand(*shape):
        return torch.randn(*shape, generator=g).to(dev)

    def conv_case(tag, n, h, w, cin, cout, residual=False, lrelu=True, f32_out=False):
        x = rt.split_pack(Act(n, h, w, cin, f32=rand(n, h, w, cin)))
        wv, b = rand(3, 3, cin, cout), rand(cout)
        ws = float(np.sqrt(2.0 / (9 * cin)))
        xm = rt.use_xmerge(cin, 3, False)
        prepared = rt.prepare_weights_xmerge(wv, ws, cout) if xm else rt.prepare_weights(wv, ws, 3, cin, cout)
        res = rand(n, h, w, cout) if residual else None
        flops = 2.0 * 9 * cin * cout * n * h * w
        # activation planes in (2 x 2 B) + planes out (2 x 2 B) [+ fp32 out 4 B + residual 4 B]
        byts = n * h * w * (4 * cin + 4 * cout + (4 * cout if f32_out or residual else 0) + (4 * cout if residual else 0))
        targets.append((tag, 'tensor', flops, byts,
                        lambda: rt.conv2d(x, wv, b, ws, 3, cout, lrelu=lrelu, residual=res, want_f32=f32_out or residual,
                                          want_split=True, prepared=prepared, algo=_lib.ALGO_TC)))
        return x, wv, ws, prepared

    def grad_cases(tag, n, h, w, cin, cout, x, wv, ws, prepared):
        dy = rand(n, h, w, cout)
        y = rt.conv2d(x, wv, None, ws, 3, cout, lrelu=True, want_f32=False, want_split=True, prepared=prepared,
                      algo=_lib.ALGO_TC)
        dz, _ = rt.grad_prepare(dy, n, h, w, cout, src_kind=1, y_hi=y.hi, want_planes=True)
        flops = 2.0 * 9 * cin * cout * n * h * w
        if prepared[0].shape[1] != 9 * cin:          # x-merged forward planes: the data gradient uses the plain ones
            prepared = rt.prepare_weights(wv, ws, 3, cin, cout)
        wt = rt.transpose_weights(prepared, cout, 9, cin)
        gh = (h + 4) * (w + 4)
        targets.append((tag + ' dgrad (LIN mode)', 'tensor', flops, n * gh * (4 * cout + 4 * cin),
                        lambda: rt.conv_dgrad(dz, n, h, w, cin, cout, 3, wt)))
        dw = torch.zeros(3, 3, cin, cout, device=dev)
        targets.append((tag + ' wgrad', 'tensor', flops, n * h * w * 4 * cin + n * gh * 4 * cout,
                        lambda: rt.conv_wgrad((x.hi, x.lo), dz, n, h, w, cin, cout, 3, ws, dw)))
        gg = rt.conv_dgrad(dz, n, h, w, cin, cout, 3, wt)
        # grad_prepare of the layer below: dgrad grid (4 B) + hi-plane mask (2 B) in, hi/lo planes (4 B) out
        xb = rt.split_pack(Act(n, h, w, cin, f32=rand(n, h, w, cin)))
        dbias = torch.zeros(cin, device=dev)
        targets.append((tag + ' grad_prepare (fold + mask + bias grad + split) c%d' % cin, 'hbm', 0.0,
                        n * gh * cin * 4 + n * h * w * cin * 2 + n * gh * cin * 4,
                        lambda: rt.grad_prepare(gg, n, h, w, cin, src_kind=0, fold=0, y_hi=xb.hi, want_planes=True,
                                                dbias=dbias)))
        # the two launches above as ONE (tmx_conv2d_dgrad_gp: grad_prepare in the data-gradient epilogue + border pass)
        dbias2 = torch.zeros(cin, device=dev)
        targets.append((tag + ' dgrad + grad_prepare fused (GP epilogue + border kernel)', 'tensor', flops,
                        n * gh * 4 * cout + n * h * w * cin * 2 + n * gh * cin * 4,
                        lambda: rt.conv_dgrad_gp(dz, n, h, w, cin, cout, 3, wt, 0, y_hi=xb.hi, dbias=dbias2)))

    n = 32
    t = conv_case('conv fwd 64x64 256->256 (Residual_0: planes out)', n, 64, 64, 256, 256)
    conv_case('conv fwd 64x64 256->256 (Residual_1: + fp32 residual in, fp32 + planes out)', n, 64, 64, 256, 256,
              residual=True, lrelu=False)
    grad_cases('64x64 256->256', n, 64, 64, 256, 256, *t)
    t = conv_case('conv fwd 128x128 16->16 (thin, X-MERGED)', n, 128, 128, 16, 16, f32_out=True)
    grad_cases('128x128 16->16', n, 128, 128, 16, 16, *t)
    t = conv_case('conv fwd 128x128 16->32 (thin, X-MERGED)', n, 128, 128, 16, 32, f32_out=True)
    grad_cases('128x128 16->32', n, 128, 128, 16, 32, *t)
    t = conv_case('conv fwd 64x64 32->64', n, 64, 64, 32, 64, f32_out=True)
    grad_cases('64x64 32->64', n, 64, 64, 32, 64, *t)

    # latent blend of cfg 4: 4 sources -> 4x4 tile grid, per-source gathers + 4-corner matte, batch 16
    nb, S = 16, 4
    srcs = [rand(nb, 128, 32, 32) for _ in range(4)]
    np.random.seed(0)
    ih = [interp.sample_permutation_indices(nb, 32 * S, 5) for _ in range(4)]
    iw = [interp.sample_permutation_indices(nb, 32 * S, 5) for _ in range(4)]
    pr, pc = interp._corner_pins(S, S)
    out_bytes = nb * 128 * (32 * S) ** 2 * 4
    L = 32 * S
    rh, rw = interp.linkern_ramps(L, L, 32)
    rh = [torch.from_numpy(np.ascontiguousarray(r)).to(dev) for r in rh]
    rw = [torch.from_numpy(np.ascontiguousarray(r)).to(dev) for r in rw]
    ihd = [torch.from_numpy(np.ascontiguousarray(i, dtype=np.int32)).to(dev) for i in ih]
    iwd = [torch.from_numpy(np.ascontiguousarray(i, dtype=np.int32)).to(dev) for i in iw]
    targets.append(('latent_blend (cfg 4: 4 sources, gather + re-pin + matte) [16,128,128,128]', 'hbm', 0.0,
                    out_bytes + 4 * nb * 128 * 32 * 32 * 4,
                    lambda: rt.latent_blend(srcs, L, L, _lib.BLEND_MATTE, idx_h=ihd, idx_w=iwd, ramps_h=rh, ramps_w=rw,
                                            pin_rows=pr, pin_cols=pc)))
    # window copy (crop-aware canvases): [32,128,96,96] -> 64x64 window at a device-resident offset
    canvas = rand(32, 128, 96, 96)
    off = torch.tensor([13, 22], dtype=torch.int32, device=dev)
    from texturemixer_b200.loss import Window
    win = Window(13, 22, 64, 64, dev=off)
    targets.append(('window_copy NCHW [32,128,96,96] -> 64x64', 'hbm', 0.0, 2 * 32 * 128 * 64 * 64 * 4,
                    lambda: rt.window(canvas, win)))
    # Adam on the E/G bucket (15.78 M parameters): w, g, v in + w, m, v out (beta1 = 0: m is not read)
    npar = 15_780_000 // 4 * 4
    w_, g_, m_, v_ = (rand(npar) for _ in range(4))
    v_.abs_()
    powers = torch.tensor([0.0, 0.99], device=dev)
    mark = torch.zeros(1, device=dev)
    targets.append(('adam (fused 1/world scale + skip mark + TF1 Adam), 15.78 M parameters', 'hbm', 0.0, npar * 24,
                    lambda: _lib.check(rt.lib.tmx_adam_update(rt.handle, C.c_void_p(w_.data_ptr()), C.c_void_p(g_.data_ptr()),
                                                              C.c_void_p(m_.data_ptr()), C.c_void_p(v_.data_ptr()), npar,
                                                              0.0015, 0.0, 0.99, 1e-8, 1.0, C.c_void_p(powers.data_ptr()),
                                                              None, C.c_void_p(mark.data_ptr()), rt.stream()))))

    # dense head of D_patch (8192 -> 512, batch 32): weight-streaming bound (16.8 MB of fp32 weights per launch)
    xd, wd, bd = rand(32, 8192), rand(8192, 512), rand(512)
    yd = rt.dense(xd, wd, bd, 0.0156, True)
    dyd = rand(32, 512)
    dwd, dbd = torch.zeros(8192, 512, device=dev), torch.zeros(512, device=dev)
    dxd = torch.empty(32, 8192, device=dev)
    wbytes = 8192 * 512 * 4
    P = lambda t_: C.c_void_p(t_.data_ptr())                                                     # noqa: E731
    targets.append(('dense fwd 8192->512 batch 32 (split-K partial + finish)', 'hbm', 0.0, wbytes + 32 * 8192 * 4,
                    lambda: rt.dense(xd, wd, bd, 0.0156, True)))
    targets.append(('dense wgrad 8192->512 batch 32', 'hbm', 0.0, 2 * wbytes + 32 * 8192 * 4,
                    lambda: _lib.check(rt.lib.tmx_dense_wgrad(rt.handle, P(xd), P(dyd), P(yd), P(dwd), P(dbd), 32, 8192, 512,
                                                              0.0156, 1, 0.2, rt.stream()))))
    targets.append(('dense bwd_input 8192->512 batch 32', 'hbm', 0.0, wbytes + 32 * 8192 * 4,
                    lambda: _lib.check(rt.lib.tmx_dense_bwd_input(rt.handle, P(dyd), P(yd), P(wd), 0.0156, P(dxd), 32, 8192,
                                                                  512, 1, 0.2, rt.stream()))))
    # VGG-19 Gram loss: Gram matrices (wgrad kernel, sample = tap) and the per-sample-weight 1x1 conv of their gradient
    import types
    from texturemixer_b200.vgg import GramLoss
    me = types.SimpleNamespace(rt=rt, use_tc=True)
    for (c_, hw_) in ((64, 128), (256, 32)):
        fa = rt.split_pack(Act(n, hw_, hw_, c_, f32=torch.relu(rand(n, hw_, hw_, c_))), 'zero')
        Sg = rand(n, c_, c_)
        gf = 2.0 * c_ * c_ * hw_ * hw_ * n
        targets.append(('gram fwd (tensor cores) [32,%d,%d,%d]' % (c_, hw_, hw_), 'tensor', gf, n * hw_ * hw_ * c_ * 4,
                        (lambda a=fa: GramLoss._gram_of(me, a))))
        targets.append(('gram bwd (per-sample-weight 1x1 conv) [32,%d,%d,%d]' % (c_, hw_, hw_), 'tensor', gf,
                        n * hw_ * hw_ * c_ * 8, (lambda a=fa, S_=Sg: GramLoss._feature_gradient(me, a, S_))))

    hbm, bf16, src = peaks()
    rows = []
    for tag, bound, flops, byts, fn in targets:
        if args.only and args.only not in tag:
            continue
        fn()
        torch.cuda.synchronize()
        reps = 1 if args.once else 10
        ms = []
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t_ms = float(np.median(ms))
        if bound == 'tensor':
            ach = flops / t_ms / 1e9
            rows.append((tag, t_ms, '%.1f TFLOP/s algorithmic (x3 executed: %.1f = %.3f of bf16 burst peak), %.0f GB/s algorithmic bytes'
                         % (ach, 3 * ach, 3 * ach / bf16, byts / t_ms / 1e6)))
        else:
            ach = byts / t_ms / 1e6
            rows.append((tag, t_ms, '%.0f GB/s = %.3f of HBM peak (%.1f MB algorithmic)' % (ach, ach / hbm, byts / 1e6)))
    print('peaks: hbm %.1f GB/s, bf16 %.1f TFLOP/s burst - %s; L2 flushed before every timed launch' % (hbm, bf16, src))
    for tag, t_ms, txt in rows:
        print('%9.1f us  %-88s %s' % (t_ms * 1e3, tag, txt))


if __name__ == '__main__':
    main()
