#!/usr/bin/env python
"""Turn the artefacts of tools/trip_evidence.sh (gpurun_out/) into the tracked summaries under profiles/r2/:

  op_bench.md      per-kernel device times, ours next to the library ops of the reference stack (tools/op_bench.py)
  ncu_fused.md     `ncu --set full` of the step's kernels: launch shape, registers / smem, throughput and stall metrics,
                   and where each kernel sits against the measured roofline (MEASURED_PEAKS.json)
  launches.md      the kernels one training step launches (ncu launch list)

Runs on the CPU-only dev box (ncu -i reads the reports).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2")
GO = os.path.join(ROOT, "gpurun_out")
B = 100
f = 4

# compulsory global traffic (every tensor read or written once) and FLOPs (2·MAC) of the step's kernels at batch 100
x0, y1, p1f, y2, p2 = B * 784 * f, B * 784 * 16 * f, B * 324 * 16 * f, B * 196 * 32 * f, B * 1568 * f
dy2f = B * 324 * 32 * f
wts = 116136
WORK = {
    "convnet_fwd_kernel": (x0 + y1 + p1f + y2 + p2 + wts, 2 * B * 784 * 16 * 25 + 2 * B * 196 * 32 * 400 + 2 * B * 10 * 1568,
                           "conv1 + BN1/ReLU/pool + conv2 (tcgen05) + BN2/ReLU/pool + fc"),
    "convnet_l2_bwd_kernel": (2 * p2 + y2 + dy2f + p1f + 51200 + 2 * 62720, 2 * B * 196 * 32 * 400 + 4 * B * 10 * 1568,
                              "classifier bwd + pool/ReLU/BN2 bwd + conv2 dgrad (tcgen05)"),
    "convnet_l1_bwd_kernel": (p1f + y1 + x0 + dy2f + p1f, 2 * B * 784 * 16 * 25 + 2 * B * 196 * 32 * 400,
                              "pool/ReLU/BN1 bwd + conv1 wgrad (mma.sync) + conv2 wgrad (tcgen05, TMA)"),
    "conv5x5_wgrad_win_kernel": (dy2f + p1f, 2 * B * 196 * 32 * 400, "(stand-alone variant, PDT_WGRAD_MERGED=0) conv2 wgrad, TMA-materialised tap pairs"),
    "linear_bwd_kernel": (2 * p2 + 2 * 10 * 1568 * f, 4 * B * 10 * 1568, "(stand-alone variant, PDT_FC_MERGED=0) fc dX, dW, db"),
}
METRICS = [
    ("gpu__time_duration.sum", "duration under ncu"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__shared_mem_per_block_static", "static smem / CTA"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("lts__t_bytes.sum", "L2 traffic"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__t_sectors.sum", "L2 sectors (32 B)"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum", "tcgen05 TF32 tensor ops"),
    ("sm__ops_path_tensor_op_hmma_src_tf32_dst_fp32_sparsity_off.sum", "mma.sync TF32 tensor ops"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stalled warps per issue: CTA barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stalled warps per issue: long scoreboard (L2 / global)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stalled warps per issue: short scoreboard (smem / MUFU)"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stalled warps per issue: memory barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stalled warps per issue: fixed-latency wait"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stalled warps per issue: sleeping (nanosleep / mbarrier try_wait)"),
]


def ncu_raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    rows = [x for x in rows if len(x) > 20]
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return f"{v} {u}".strip()
    if x == int(x) and abs(x) < 1e7:
        return f"{int(x)} {u}".strip()
    return f"{x:.3g} {u}".strip()


def op_bench_md():
    path = os.path.join(GO, "op_bench.json")
    if not os.path.exists(path):
        return
    rows = json.load(open(path))
    out = ["# Per-kernel device time: this framework next to the library ops of the reference stack (B200, batch 100)", "",
           "`tools/op_bench.py`: every row is a CUDA graph of 20 back-to-back launches of that op (CUDA events around the replay, ÷ 20) — the",
           "cost of the op inside a captured training step, launch gap included, host overhead excluded, for BOTH arms; `cold` is one eager",
           "launch after a 256 MB write (> 126 MB L2).  Library rows are the ATen / cuDNN (benchmark mode, TF32 allowed) / cuBLAS ops that the",
           "reference's `nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d / Linear / CrossEntropyLoss / SGD` dispatch to, on the same shapes.", "",
           "| arm | op | µs in graph | µs cold eager |", "|---|---|---|---|"]
    for r in rows:
        g = f"{r['us_in_graph']:.2f}" if r.get("us_in_graph") is not None else "n/a"
        out.append(f"| {r['arm']} | {r['op']} | {g} | {r['us_cold_eager']:.2f} |")
    ours = sum(r["us_in_graph"] for r in rows if r["arm"] == "ours" and r.get("us_in_graph") and not r["op"].startswith("(variant)"))
    lib = {r["op"]: r["us_in_graph"] for r in rows if r["arm"] == "library" and r.get("us_in_graph")}
    lib_step = sum(v for k, v in lib.items() if "fwd+bwd" in k or k.startswith("SGD"))
    out += ["", f"Sum of our step's kernels: **{ours:.1f} µs**; the library's forward+backward rows + SGD (the same work): **{lib_step:.1f} µs** "
            f"({lib_step / ours:.1f}×).", ""]
    open(os.path.join(OUT, "op_bench.md"), "w").write("\n".join(out))


def ncu_md():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tf32 = peaks["hbm_gbs"] * 1e9, peaks["bf16_tflops"] * 1e12 / 2
    ob = {}
    p = os.path.join(GO, "op_bench.json")
    if os.path.exists(p):
        ob = {r["op"]: r for r in json.load(open(p)) if r["arm"] == "ours"}
    out = ["# `ncu --set full` of the training step's kernels (B200, batch 100, one capture per kernel, `--clock-control none`)", "",
           "Reports: `gpurun_out/prof_<kernel>.ncu-rep` (scratch; re-create with `tools/trip_evidence.sh`).  Durations under ncu include its",
           "per-kernel serialisation; the in-graph times of `op_bench.md` are the ones that add up to the step.", "",
           f"Roofline denominators (`MEASURED_PEAKS.json`): copy bandwidth {peaks['hbm_gbs']:.0f} GB/s, cuBLAS bf16 {peaks['bf16_tflops']:.0f} TFLOP/s",
           "(TF32 tensor-core peak taken as half of it).", ""]
    roof = ["| kernel | contents | compulsory MB | MFLOP | roofline µs | measured µs (ncu) | measured ÷ roofline |", "|---|---|---|---|---|---|---|"]
    for k, (nbytes, flops, what) in WORK.items():
        rep = os.path.join(GO, f"prof_{k}.ncu-rep")
        if not os.path.exists(rep):
            continue
        m = ncu_raw(rep)
        if m is None:
            continue
        dur = float(m["gpu__time_duration.sum"][0].replace(",", ""))
        if m["gpu__time_duration.sum"][1].startswith("ns"):
            dur /= 1e3
        r_us = max(nbytes / hbm, flops / tf32) * 1e6
        roof.append(f"| `{k}` | {what} | {nbytes / 1e6:.2f} | {flops / 1e6:.0f} | {r_us:.2f} | {dur:.1f} | {dur / r_us:.0f}× |")
        out += [f"## `{k}`", "", "| metric | value |", "|---|---|"]
        for name, label in METRICS:
            if name in m:
                out.append(f"| {label} (`{name}`) | {fmt(*m[name])} |")
        out.append("")
    out += ["## Where the kernels sit against the roofline", ""] + roof + [
        "",
        "Reading: at batch 100 the whole step moves ~20 MB and computes ~0.6 GFLOP — 3 µs of HBM time and 0.7 µs of tensor-core time.",
        "Every kernel is 10-40× above its roofline: the step is bound by *dependent phases*, not by bandwidth or FLOPs.  Each cooperative",
        "kernel is a chain of short phases (load, reduce, grid barrier ≈ 1.8 µs, fold, MMA, epilogue) in which one CTA per image keeps",
        "≤ 25 warps on an SM; the stall tables above show it (CTA-barrier and long-scoreboard waits dominate, achieved occupancy 10-40 %).",
        "That is why the optimisation work of this round went into removing phases (launches 20 → 3, grid barriers merged and their shadows",
        "filled, the conv2 weight gradient hidden behind layer-1 backward, the classifier / loss / optimizer riding on the layer kernels)",
        "rather than into the inner loops; `fused_trace.log` has the per-phase timelines (eager and inside the replayed graph).", ""]
    open(os.path.join(OUT, "ncu_fused.md"), "w").write("\n".join(out))


def launches_md():
    path = os.path.join(GO, "launches_fused.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    out = ["# Kernels launched by training steps (eager loop, `tools/one_step.py`, ncu launch list with `gpu__time_duration`)", "",
           "| # | kernel | grid | block | ns under ncu |", "|---|---|---|---|---|"]
    for r in rows:
        name = r[4].split("(")[0].replace("pdt::<unnamed>::", "").replace("void ", "")
        out.append(f"| {r[0]} | `{name[:80]}` | {r[8]} | {r[7]} | {r[-1]} |")
    out += ["", "`at::…FillFunctor` / `MulFunctor` rows are the eager loop's loss-seed `ones_like` and `zero_grad`; a captured step "
            "(`engine.GraphedTrainStep`) has neither.", ""]
    open(os.path.join(OUT, "launches.md"), "w").write("\n".join(out))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    op_bench_md()
    ncu_md()
    launches_md()
    print("wrote", sorted(os.listdir(OUT)))
