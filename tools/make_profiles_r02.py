#!/usr/bin/env python
"""Round-2 profile summaries from the raw ncu files of the final build (gpurun_out/*<tag>*) -> profiles/r02_*.
    python tools/make_profiles_r02.py [tag]    (default r3b; needs `ncu` on PATH to read the .ncu-rep; no GPU)"""
import collections, csv, gzip, json, os, re, shutil, subprocess, sys
TAG = sys.argv[1] if len(sys.argv) > 1 else "r3b"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarise(rep, title, traffic=None):
    hdr, units, rows = raw_rows(rep)
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    out = [title, ""]
    for r in rows:
        short = re.sub(r"\(CUtensorMap.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("snacb::", "")
        out.append(f"== {short}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"   {w:72s} {r[i]:>16s} {units[i]}")
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        by = float(r[ir]) * MULT[units[ir]] + float(r[iw]) * MULT[units[iw]]
        tm = float(r[it]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[it], 1e-6)
        out.append(f"   {'=> DRAM bytes per launch / achieved':72s} {by / 1e9:10.3f} GB   {by / tm / 1e12:6.2f} TB/s (peak measured 6.55)")
        st = sorted(((float(r[hdr.index(h)]), h) for h in stalls), reverse=True)[:6]
        out.append("   top stalls (warps per issue-active cycle): " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in st))
        if traffic is not None:
            traffic[short] = by
    return "\n".join(out) + "\n"


def main():
    # ---- launch list
    src = os.path.join(G, f"launches_{TAG}.csv")
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("snacb::", "")
        ns = float(r[vi]); tot += ns
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ns
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-extra",
             "(round-2 final build; the first 400 launches = warm-up + timed full-window steps + part of the sliced leg; cold-cache, serialised: compare SHARES)",
             f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}"]
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n:62s} {c:8d} {ns / 1e6:10.3f} {ns / tot * 100:6.1f}% {ns / c / 1e3:9.1f}")
    chain = sum(ns for n, (c, ns) in agg.items() if n.startswith("k_chain"))
    lines.append(f"k_chain share of the listed launches: {chain / tot * 100:.1f} %")
    open(os.path.join(P, "r02_ncu_launch_summary.txt"), "w").write("\n".join(lines) + "\n")
    with open(src, "rb") as f, gzip.open(os.path.join(P, "r02_ncu_launches.csv.gz"), "wb") as g:
        shutil.copyfileobj(f, g)
    # ---- full captures
    traffic = {}
    txt = summarise(os.path.join(G, f"prof_chain_{TAG}.ncu-rep"),
                    "ncu --set full --clock-control none --import-source on, the three k_chain launches of one decode of 1024 four-frame windows\n"
                    "(full windows, SNACB_NO_TRIM=1, fp16 operands, round-2 final build): python tests/gpu_one.py 1024 fp16 1\n"
                    "(per-launch times under ncu are cold-cache and serialised; shares, not absolutes)", traffic)
    open(os.path.join(P, "r02_chain_ncu.txt"), "w").write(txt)
    json.dump({"chain": {"1024": sum(traffic.values()) / len(traffic)}, "_per_kernel_bytes": traffic,
               "_what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the three k_chain launches of a step), "
                        "ncu --set full, B = 1024 full windows, round-2 final build"}, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    txt = summarise(os.path.join(G, f"prof_mem_{TAG}.ncu-rep"),
                    "ncu --set full --clock-control none, the kernels outside the chain in one decode of 1024 four-frame windows (full windows, fp16,\n"
                    "round-2 final build): k_vq_stem (token unpack fused), stem / ConvTranspose / NoiseBlock GEMMs (k_gemm_tc), block-0 ResidualUnits\n"
                    "(k_resunit2), k_convt_ph (block 2), k_convt_res (block 3), k_tail_bulk.  Achieved HBM bandwidth per kernel against the measured 6.55 TB/s.")
    open(os.path.join(P, "r02_membound_ncu.txt"), "w").write(txt)
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
