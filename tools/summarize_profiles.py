"""Turn gpurun_out/*.ncu-rep and launches.csv into the small text summaries committed under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "lts__t_bytes.sum", "smsp__inst_executed.sum"]


def units(v, u):
    return v


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return []
    hdr, unit = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append({h: (v, u) for h, u, v in zip(hdr, unit, r)})
    return out


def summarize_rep(name):
    rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
    if not os.path.exists(rep):
        return None
    launches = raw(rep)
    lines = ["# %s -- ncu --set full --clock-control none (cold-cache, serialised); source: gpurun_out/%s.ncu-rep" % (name, name)]
    res = []
    for L in launches:
        kn = L.get("Kernel Name", ("?", ""))[0]
        lines.append("\n## %s  grid=%s block=%s" % (kn[:110], L.get("Grid Size", ("", ""))[0], L.get("Block Size", ("", ""))[0]))
        d = {}
        for k in KEYS:
            if k in L:
                lines.append("%-80s %s %s" % (k, L[k][0], L[k][1]))
                d[k] = L[k]
        st = [(k, float(v[0])) for k, v in L.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v[0]]
        st.sort(key=lambda kv: -kv[1])
        lines.append("top stall reasons (warps per issue-active cycle): " +
                     ", ".join("%s=%.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v) for k, v in st[:6]))
        res.append((kn, d))
    with open(os.path.join(OUT, "%s_ncu_%s.txt" % (TAG, name.replace("prof_", ""))), "w") as f:
        f.write("\n".join(lines) + "\n")
    return res


def to_ms(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)


def summarize_launches():
    path = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(r[ki], to_ms(r[vi], r[ui])) for r in data if len(r) > vi]
    marks = [i for i, (n, _) in enumerate(seq) if "sgd_kernel" in n]
    if len(marks) >= 2:
        step = seq[marks[-2] + 1:marks[-1] + 1]     # one complete training step: after an optimizer kernel up to the next
    else:
        step = seq
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        key = n.split("(")[0].replace("void ", "")[:90]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    own = sum(v[1] for k, v in agg.items() if "edb::" in k)
    with open(os.path.join(OUT, "%s_launches_step.txt" % TAG), "w") as f:
        f.write("# one training step (B=128, bf16; the last of `python tools/one_step.py 4`, or of bench.py in earlier rounds) under\n"
                "# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised:\n"
                "# compare SHARES).  %d launches, %.3f ms summed; editor_b200 kernels: %d launches, %.1f %% of the time.\n"
                % (len(step), tot, sum(v[0] for k, v in agg.items() if "edb::" in k), 100 * own / tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-92s n=%4d %9.3f ms %5.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))


def main():
    os.makedirs(OUT, exist_ok=True)
    summarize_launches()
    g = summarize_rep("prof_gemm")
    for n in ("prof_attn_fwd", "prof_attn_bwd", "prof_sfts", "prof_ln", "prof_augment", "prof_gemm_fc1", "prof_gemm_fc2d",
              "prof_gemm_proj", "prof_gemm_fc1d"):
        summarize_rep(n)
    if g:
        tr = []
        for kn, d in g:
            rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(d["dram__bytes_read.sum"][1], 1)
            wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(d["dram__bytes_write.sum"][1], 1)
            tr.append(rd + wr)
        json.dump({"kernel": "gemm_bf16_kernel", "captured_launches": len(tr), "dram_bytes_per_launch": tr,
                   "mean_dram_bytes_per_launch": sum(tr) / len(tr)}, open(os.path.join(OUT, "%s_gemm_traffic.json" % TAG), "w"), indent=1)
    for f in ("bench_own.json", "bench_ref.json", "bench_rgbnt100.json", "bench_msvr310_fp32.json", "pytest_gpu.log",
              "gemm_bench.txt", "attn_bench.txt", "ln_bench.txt", "step_trace.txt"):
        src = os.path.join(ROOT, "gpurun_out", f)
        if os.path.exists(src):
            open(os.path.join(OUT, "%s_%s" % (TAG, f)), "w").write(open(src).read())


if __name__ == "__main__":
    main()
