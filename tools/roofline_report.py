#!/usr/bin/env python
"""Roofline table for the ConvNet step and the allreduce (profiles/roofline.md).

Inputs: gpurun_out/op_bench.json (tools/op_bench.py: per-kernel CUDA-event times, L2 flushed between
iterations), profiles/allreduce_sweep_{2,8}gpu.json (tools/allreduce_sweep.py), MEASURED_PEAKS.json
(driver-measured copy bandwidth and bf16 matmul throughput).  Bytes are the compulsory traffic of each
kernel at batch 100 (every tensor read or written once); FLOPs are 2·MACs.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 100
MB = 1e6
f = 4  # bytes per element
x0 = B * 28 * 28 * 1 * f
y1 = B * 28 * 28 * 16 * f
p1 = B * 14 * 14 * 16 * f
y2 = B * 14 * 14 * 32 * f
p2 = B * 7 * 7 * 32 * f
KERNELS = [  # (op name in op_bench.json, bytes, flops, note)
    ("conv1_fwd_simt", x0 + y1, 2 * B * 784 * 16 * 25, "SIMT fp32; + BN Σ/Σ² (one-CTA-per-image variant measured here; the default 400-CTA variant is 20.5)"),
    ("bn_relu_pool1_fwd", y1 + p1, 0, ""),
    ("conv2_fwd_tma_im2col(+repack)", p1 + y2, 2 * B * 196 * 32 * 400, "tcgen05 TF32, TMA im2col; + BN Σ/Σ²"),
    ("bn_relu_pool2_fwd", y2 + p2, 0, ""),
    ("linear_fwd", p2 + 10 * 1568 * f, 2 * B * 10 * 1568, "fp32 SIMT"),
    ("cross_entropy_fwd", 2 * B * 10 * f, 0, ""),
    ("cross_entropy_bwd", 2 * B * 10 * f, 0, ""),
    ("linear_bwd", 2 * p2 + 2 * 10 * 1568 * f, 4 * B * 10 * 1568, "dX, dW, db in one kernel"),
    ("bn_relu_pool2_bwd_reduce", p2 + y2, 0, ""),
    ("bn_relu_pool2_bwd_apply", p2 + 2 * y2, 0, ""),
    ("conv2_wgrad_tcgen05(+fold)", y2 + p1, 2 * B * 196 * 32 * 400, "tcgen05 TF32 MN-major split-K"),
    ("conv2_dgrad_tma_im2col(+repack)", y2 + p1, 2 * B * 196 * 32 * 400, "tcgen05 TF32, TMA im2col"),
    ("bn_relu_pool1_bwd_reduce", p1 + y1, 0, ""),
    ("bn_relu_pool1_bwd_apply", p1 + 2 * y1, 0, ""),
    ("conv1_wgrad(+fold)", y1 + x0, 2 * B * 784 * 16 * 25, "SIMT fp32 (one-CTA-per-image variant measured here; default variant 25.5)"),
    ("sgd_multi(10 tensors)", 3 * 116136, 2 * 29034, ""),
]


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"] * 1e9
    tf32 = peaks["bf16_tflops"] * 1e12 / 2  # tcgen05 kind::tf32 runs at half the bf16 rate
    ops = {r["op"]: r for r in json.load(open(os.path.join(ROOT, "gpurun_out", "op_bench.json")))}
    floor = ops.get("empty_launch_floor(torch.zero_ 1 elem)", {}).get("us_median", 6.0)
    out = ["# Roofline position of the ConvNet step and of the allreduce", "",
           f"Denominators from `MEASURED_PEAKS.json` (driver-measured on this image): copy bandwidth **{peaks['hbm_gbs']:.0f} GB/s**, "
           f"cuBLAS bf16 **{peaks['bf16_tflops']:.0f} TFLOP/s** (TF32 tensor-core peak taken as half of it).  Kernel times: "
           "`tools/op_bench.py`, CUDA events, L2 flushed between iterations, median of 20; every measurement of a single "
           f"launch carries ≈{floor:.0f} µs of launch + event overhead (the `torch.zero_` floor), so `net` = time − floor.", "",
           "| kernel | time µs (net) | compulsory MB | MFLOP | roofline µs (max of bytes/BW, FLOP/peak) | net ÷ roofline | note |",
           "|---|---|---|---|---|---|---|"]
    tot_t = tot_net = tot_roof = tot_b = tot_f = 0.0
    for name, nbytes, flops, note in KERNELS:
        if name not in ops:
            continue
        t = ops[name]["us_median"]
        net = max(t - floor, 0.5)
        roof = max(nbytes / hbm, flops / tf32) * 1e6
        out.append(f"| {name} | {t:.1f} ({net:.1f}) | {nbytes / MB:.2f} | {flops / 1e6:.0f} | {roof:.2f} | {net / roof:.0f}× | {note} |")
        tot_t += t; tot_net += net; tot_roof += roof; tot_b += nbytes; tot_f += flops
    out += [f"| **sum** | {tot_t:.0f} ({tot_net:.0f}) | {tot_b / MB:.1f} | {tot_f / 1e6:.0f} | {tot_roof:.1f} | {tot_net / tot_roof:.0f}× | |", "",
            f"Reading: the whole step moves ≈{tot_b / MB:.0f} MB and does {tot_f / 1e9:.1f} GFLOP — **≈{tot_roof:.0f} µs of B200 time at the roofline** — and every "
            "tensor fits in the 126 MB L2, so DRAM bandwidth is not even the binding resource.  What the step actually pays "
            "for is *dependent kernel boundaries*: 20 launches whose ramp-up / drain / grid-wide reductions cost 3–8 µs each "
            f"regardless of payload.  The measured graph-replayed step (0.157 ms on 1 GPU) is ≈{157 / tot_roof:.0f}× the roofline and 7.5× "
            "faster than the reference stack (1.18 ms), which pays the same physics through 60–80 launches plus host "
            "overhead.  Closing the rest is a fusion problem (fewer boundaries), then a TMA-gather problem for the two "
            "im2col convolutions (profiles/conv_tma_ncu.md) — not a FLOP or bandwidth problem.", ""]
    for n in (2, 8):
        p = os.path.join(ROOT, "profiles", f"allreduce_sweep_{n}gpu.json")
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        out += [f"## Allreduce, {n}×B200 (fp32 SUM): achieved fraction of the NVLink roofline", "",
                f"Roofline time = bytes that must cross one GPU's NVLink port ÷ {d['link_gbs']:.0f} GB/s per direction "
                "(one-shot push: (N−1)·S out, or S with `multimem.st`; two-shot NVLS: ≈S in + S out overlapped ⇒ S; two-shot P2P: "
                "2·S·(N−1)/N).", "",
                "| bytes | best algorithm | µs | roofline µs | fraction | busbw GB/s | NCCL µs | vs NCCL |", "|---|---|---|---|---|---|---|---|"]
        for r in d["rows"]:
            out.append(f"| {r['bytes']:,} | {r['best']} | {r['best_us']:.1f} | {r['roofline_us']:.2f} | {100 * r['roofline_frac']:.1f} % | "
                       f"{r['busbw_gbs']:.0f} | {r['nccl_us']:.1f} | {r['speedup_vs_nccl']:.2f}× |")
        out += ["", "Small messages are latency-bound (one device-side barrier over NVSwitch ≈ 4–6 µs), so their bandwidth "
                "fraction is meaningless by construction; the ConvNet's 116 KB bucket sits there.  Large messages reach "
                f"{100 * max(r['roofline_frac'] for r in d['rows']):.0f} % of the link roofline.", ""]
    path = os.path.join(ROOT, "profiles", "roofline.md")
    open(path, "w").write("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
