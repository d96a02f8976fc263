"""Turns the ncu artefacts of a round into the tracked summaries under profiles/:
  python tools/ncu_summarize.py launches gpurun_out/r2k/launches.csv profiles/ncu_launches_r2_summary.md r2
  python tools/ncu_summarize.py full gpurun_out/r2k/top.ncu-rep profiles/ncu_top_kernels_r2.json profiles/dram_traffic.json cfg2
The "full" summary keeps, per kernel: duration, DRAM bytes, occupancy, registers, issue utilisation, the instruction
mix by pipe (inst_executed_pipe_*), lane utilisation (thread_inst_executed_per_inst_executed) and the warp stall
reasons per issued instruction — what is needed to decompose "issue-bound"."""
import csv
import io
import json
import os
import subprocess
import sys

STEP_KERNELS = ["precull_kernel", "preprocess_kernel", "count_large_kernel", "select_grad_rows_kernel", "tile_meta_kernel", "emit_kernel", "emit_large_kernel", "tile_sort_pack_small_kernel",
                "tile_sort_pack_mid_kernel", "tile_sort_pack_long_kernel", "blend_fwd_kernel", "blend_fwd_packed_kernel",
                "blend_bwd_kernel", "blend_bwd_packed_kernel", "preprocess_bwd_kernel", "adam_rows_all_kernel"]


def short(name):
    n = name.split("(")[0].split("::")[-1]
    n = n.replace("void ", "").strip()
    return n.split("<")[0]


def launches(src, dst, tag):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    seq = [(short(r[ki]), float(r[vi].replace(",", "")) / 1000.0) for r in rows[1:]]
    # the last complete step: from the last cull kernel (the first launch of a forward pass) that is followed by an adam kernel
    first = "precull_kernel" if any(k == "precull_kernel" for k, _ in seq) else "preprocess_kernel"
    starts = [i for i, (k, _) in enumerate(seq) if k == first]
    step = None
    for s in reversed(starts):
        tail = seq[s:]
        names = [k for k, _ in tail]
        if "adam_rows_all_kernel" in names:
            step = tail[: names.index("adam_rows_all_kernel") + 1]
            break
    tot = sum(t for _, t in step)
    out = [f"# ncu launch list, round {tag} (cfg2: 500k SH-3 Gaussians, 1200x680, one mapping step = fwd + bwd + masked Adam)", "",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv python bench.py --steps 2 --warmup 1 "
           "--no-cpu-baseline --no-cfg5 --no-shim` (one steady-state step; cold-cache, serialised: compare SHARES with the live "
           "CUDA-event stage times of the bench line).", "", "| # | kernel | us | share |", "|---|---|---|---|"]
    for i, (k, t) in enumerate(step):
        out.append(f"| {i} | {k} | {t:.2f} | {100 * t / tot:.1f}% |")
    out.append(f"| | total | {tot:.1f} | |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.max.pct_of_peak_sustained_active", "smsp__issue_active.min.pct_of_peak_sustained_active"]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def full(rep, dst, traffic_dst, cfg):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    idx = {w: h.index(w) for w in ["Kernel Name"] + KEEP if w in h}
    stall_idx = {w: i for i, w in enumerate(h) if w.startswith(STALL_PREFIX) and w.endswith("_per_issue_active.ratio")}
    kernels, traffic = [], {}
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        d = {w: r[i] for w, i in idx.items()}
        d["Kernel Name"] = short(d["Kernel Name"])
        stalls = {w[len(STALL_PREFIX):-len("_per_issue_active.ratio")]: float(r[i]) for w, i in stall_idx.items() if r[i]}
        d["stall_reasons_per_issued_instruction"] = {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])
                                                     if v >= 0.05}
        kernels.append(d)

        def to_bytes(key):
            v, u = float(d[key].replace(",", "")), units[idx[key]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        stage = d["Kernel Name"].replace("_kernel", "").replace("_small", "").replace("_packed", "")
        traffic[stage] = traffic.get(stage, 0) if False else int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
    json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:... python bench.py --steps 2 --warmup 1 "
                          "--no-cpu-baseline --no-cfg5 --no-shim --config " + cfg,
               "units": {w: units[i] for w, i in idx.items()}, "kernels": kernels}, open(dst, "w"), indent=1)
    if traffic_dst:
        old = json.load(open(traffic_dst)) if os.path.exists(traffic_dst) else {}
        old.setdefault(cfg, {}).update(traffic)
        json.dump(old, open(traffic_dst, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"], k.get("gpu__time_duration.sum"), k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
              k.get("sm__warps_active.avg.pct_of_peak_sustained_active"), k.get("launch__registers_per_thread"),
              k.get("smsp__thread_inst_executed_per_inst_executed.ratio"))


def dram(src, traffic_dst, cfg):
    """metrics-only pass (gpu__time_duration, dram__bytes_read / _write per launch) -> per-kernel DRAM bytes of `cfg`"""
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    ii = h.index("ID")
    per = {}
    for r in rows[1:]:
        if not r[mi].startswith("dram__bytes"):
            continue
        v = float(r[vi].replace(",", "")) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(r[ui].lower(), 1)
        per.setdefault((r[ii], short(r[ki])), 0.0)
        per[(r[ii], short(r[ki]))] += v
    traffic = {}
    for (_id, k), v in per.items():       # the largest launch of each kernel (the second, "redo" launches are near-empty)
        name = k.replace("_kernel", "").replace("_small", "")
        traffic[name] = max(traffic.get(name, 0), int(v))
    old = json.load(open(traffic_dst)) if os.path.exists(traffic_dst) else {}
    old[cfg] = traffic
    json.dump(old, open(traffic_dst, "w"), indent=1)
    print(cfg, traffic)


if __name__ == "__main__":
    if sys.argv[1] == "dram":
        dram(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "2")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else "cfg2")
