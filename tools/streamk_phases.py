"""Phase stamps of the stream-K convolution kernel (gemm_streamk.cu) for the UNet's Conv3d shapes, next to the plain
kernel's back-to-back time: where a stream-K CTA's cycles go (setup, mainloop per k-block, dump, wait, epilogue).

  python tools/streamk_phases.py [--cps 36]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cps", type=int, default=36)
args = ap.parse_args()
L.init()
dev = "cuda"


def run(tag, samples, D, H, W, C, N, cps, ctas):
    M = samples * D * H * W
    a = torch.randn(M, C, device=dev).bfloat16()
    w = (torch.randn(N, 27 * C, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev)
    stamps = torch.zeros(16, device=dev, dtype=torch.int64)

    def call(dbg):
        L.check(L.lib().pd_op_conv_gemm_streamk_phases(L.ptr(a), L.ptr(w), samples, D, H, W, C, 3, 3, 3, N, L.ptr(bias), None,
                                                       L.ptr(out), L.ptr(out), None, None, None, cps, dbg, L.ptr(stamps),
                                                       L.stream_ptr()))
        torch.cuda.synchronize()
        return stamps.cpu().tolist()

    call(-20)
    ns = call(-50)[15]
    flops = 2.0 * M * N * C * 27
    # plain kernel, same shape, back to back
    st9 = torch.zeros(16, device=dev, dtype=torch.int64)
    pargs = (L.ptr(a), L.ptr(w), samples, D, H, W, C, 3, 3, 3, N, L.ptr(bias), L.ptr(out), L.ptr(out), None, 0, 0, 0,
             L.ptr(st9), L.stream_ptr())
    for _ in range(3):
        L.check(L.lib().pd_op_conv_gemm_phases(*pargs))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        L.check(L.lib().pd_op_conv_gemm_phases(*pargs))
    e1.record()
    torch.cuda.synchronize()
    plain_us = e0.elapsed_time(e1) * 1000 / 50
    print(f"{tag}: stream-K {cps}/sample {ns / 1e3:.1f} us/launch ({flops / ns * 1e-3:.0f} TF/s) | plain {plain_us:.1f} us/launch "
          f"({flops / plain_us * 1e-6:.0f} TF/s)")
    for c in ctas:
        s = call(c)
        z = s[0]
        rel = lambda i: (s[i] - z) if s[i] else None
        print(f"  cta{c:3d}: setup={rel(1)} first_stage@{rel(2)} seg0_issued@{rel(3)} acc0@{rel(4)} dumped@{rel(5)} "
              f"seg1_issued@{rel(6)} acc1@{rel(7)} partials@{rel(8)} epi_done@{rel(9)} exit@{rel(10)} "
              f"= {s[12] - s[11]} ns")


run("L0 conv3d 256->256 B=4", 4, 13, 16, 16, 256, 256, args.cps, (0, 1, 17, 35, 143))
run("L0 conv3d 256->256 B=2", 2, 13, 16, 16, 256, 256, args.cps, (0, 17))
run("L1 conv3d 512->512 B=4", 4, 13, 8, 8, 512, 512, args.cps, (0, 1, 17, 35, 143))
run("L1 conv3d 512->512 B=2", 2, 13, 8, 8, 512, 512, args.cps, (0, 17))
