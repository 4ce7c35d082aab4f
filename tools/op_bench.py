#!/usr/bin/env python
"""Per-kernel device times of the ConvNet step (CUDA events on the launching stream, after warm-up,
L2 flushed between iterations by overwriting a 256 MB buffer).  Writes gpurun_out/op_bench.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_distributed_train_b200 import _C  # noqa: E402
from pytorch_distributed_train_b200.utils import l2_flush  # noqa: E402

dev = torch.device("cuda", 0)
B = 100


def bench(name, fn, iters=30, flush=True):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            l2_flush(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return {"op": name, "us_median": ts[len(ts) // 2], "us_min": ts[0], "flush": flush}


def main():
    torch.manual_seed(0)
    x1 = torch.rand(B, 28, 28, 1, device=dev)
    w1, b1 = torch.randn(16, 1, 5, 5, device=dev) * 0.1, torch.zeros(16, device=dev)
    w2, b2 = torch.randn(32, 16, 5, 5, device=dev) * 0.05, torch.zeros(32, device=dev)
    y1, st1 = _C.conv5x5_fwd(x1, w1, b1, True, "simt")
    g1, be1 = torch.ones(16, device=dev), torch.zeros(16, device=dev)
    g2, be2 = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    a1, sv1 = _C.bn_relu_pool_fwd(y1, st1, g1, be1, None, None, None, 0.1, 1e-5, False)
    y2, st2 = _C.conv5x5_fwd(a1, w2, b2, True, "tcgen05")
    a2, sv2 = _C.bn_relu_pool_fwd(y2, st2, g2, be2, None, None, None, 0.1, 1e-5, True)
    flat = a2.reshape(B, -1)
    wf, bf = torch.randn(10, 1568, device=dev) * 0.01, torch.zeros(10, device=dev)
    logits = _C.linear_fwd(flat, wf, bf)
    tgt = torch.randint(0, 10, (B,), device=dev)
    loss, probs = _C.cross_entropy_fwd(logits, tgt)
    one = torch.ones((), device=dev)
    dlog = _C.cross_entropy_bwd(probs, tgt, one)
    dwf, dbf = torch.empty_like(wf), torch.empty_like(bf)
    dflat = _C.linear_bwd(dlog, flat, wf, True, dwf, dbf)
    d2 = dflat.view(B, 32, 7, 7)
    sums2, dg2, db2 = _C.bn_relu_pool_bwd_reduce(d2, y2, sv2, g2, be2, True)
    dy2 = _C.bn_relu_pool_bwd_apply(d2, y2, sv2, g2, be2, sums2, st2[64:], True)
    dw2, dbb2 = torch.empty_like(w2), torch.empty_like(b2)
    da1 = _C.conv5x5_dgrad(dy2, w2, "tcgen05")
    sums1, dg1, db1 = _C.bn_relu_pool_bwd_reduce(da1, y1, sv1, g1, be1, False)
    dy1 = _C.bn_relu_pool_bwd_apply(da1, y1, sv1, g1, be1, sums1, st1[32:], False)
    dw1, dbb1 = torch.empty_like(w1), torch.empty_like(b1)
    params = [w1, b1, g1, be1, w2, b2, g2, be2, wf, bf]
    grads = [torch.randn_like(p) for p in params]
    ops = [
        ("conv1_fwd_simt", lambda: _C.conv5x5_fwd(x1, w1, b1, True, "simt")),
        ("bn_relu_pool1_fwd", lambda: _C.bn_relu_pool_fwd(y1, st1, g1, be1, None, None, None, 0.1, 1e-5, False)),
        ("conv2_fwd_tcgen05(+repack)", lambda: _C.conv5x5_fwd(a1, w2, b2, True, "tcgen05")),
        ("conv2_fwd_tma_im2col(+repack)", lambda: _C.conv5x5_fwd(a1, w2, b2, True, "tma")),
        ("conv2_dgrad_tma_im2col(+repack)", lambda: _C.conv5x5_dgrad(dy2, w2, "tma")),
        ("conv2_wgrad_tcgen05(+fold)", lambda: _C.conv5x5_wgrad(dy2, a1, dw2, dbb2, "tcgen05")),
        ("conv2_wgrad_simt(+fold)", lambda: _C.conv5x5_wgrad(dy2, a1, dw2, dbb2, "simt")),
        ("conv2_fwd_simt", lambda: _C.conv5x5_fwd(a1, w2, b2, True, "simt")),
        ("bn_relu_pool2_fwd", lambda: _C.bn_relu_pool_fwd(y2, st2, g2, be2, None, None, None, 0.1, 1e-5, True)),
        ("linear_fwd", lambda: _C.linear_fwd(flat, wf, bf)),
        ("cross_entropy_fwd", lambda: _C.cross_entropy_fwd(logits, tgt)),
        ("cross_entropy_bwd", lambda: _C.cross_entropy_bwd(probs, tgt, one)),
        ("linear_bwd", lambda: _C.linear_bwd(dlog, flat, wf, True, dwf, dbf)),
        ("bn_relu_pool2_bwd_reduce", lambda: _C.bn_relu_pool_bwd_reduce(d2, y2, sv2, g2, be2, True)),
        ("bn_relu_pool2_bwd_apply", lambda: _C.bn_relu_pool_bwd_apply(d2, y2, sv2, g2, be2, sums2, st2[64:], True)),
        ("conv2_dgrad_tcgen05(+repack)", lambda: _C.conv5x5_dgrad(dy2, w2, "tcgen05")),
        ("conv2_dgrad_simt", lambda: _C.conv5x5_dgrad(dy2, w2, "simt")),
        ("conv2_wgrad(+fold)", lambda: _C.conv5x5_wgrad(dy2, a1, dw2, dbb2, "auto")),
        ("bn_relu_pool1_bwd_reduce", lambda: _C.bn_relu_pool_bwd_reduce(da1, y1, sv1, g1, be1, False)),
        ("bn_relu_pool1_bwd_apply", lambda: _C.bn_relu_pool_bwd_apply(da1, y1, sv1, g1, be1, sums1, st1[32:], False)),
        ("conv1_wgrad(+fold)", lambda: _C.conv5x5_wgrad(dy1, x1, dw1, dbb1, "auto")),
        ("sgd_multi(10 tensors)", lambda: _C.sgd_multi(params, grads, [], 1e-4, None, 0.0, 0.0, 0.0, False, False, False)),
        ("empty_launch_floor(torch.zero_ 1 elem)", lambda: one.zero_()),
    ]
    rows = []
    for name, fn in ops:
        r = bench(name, fn)
        r["us_warm_min"] = bench(name, fn, flush=False)["us_min"]
        rows.append(r)
        print(f"{r['us_median']:9.2f} us (cold L2)  {r['us_warm_min']:9.2f} us (warm)  {name}", flush=True)
    tot = sum(r["us_median"] for r in rows if "simt" not in r["op"] or "conv1" in r["op"])
    print("sum of the step's kernels (cold):", round(tot, 1), "us")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "op_bench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
