#!/usr/bin/env python
"""Turn the raw files a GPU run left under gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag>        # e.g. r01p: bench_<x>.log, launches_<tag>.csv, prof_chain_<tag>.ncu-rep ...

Needs `ncu` on PATH to read the .ncu-rep (no GPU needed)."""
import collections
import csv
import gzip
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


def main():
    tag = sys.argv[1]
    letter = tag[-1]
    os.makedirs(P, exist_ok=True)
    # ---- bench line + stage profile + configs
    b = last_json(os.path.join(G, f"bench_{letter}.log"))
    json.dump(b, open(os.path.join(P, "r01_bench.json"), "w"), indent=1)
    shutil.copy(os.path.join(G, f"profile_{tag}.json"), os.path.join(P, "r01_stage_profile.json"))
    if os.path.exists(os.path.join(G, "configs.json")):
        shutil.copy(os.path.join(G, "configs.json"), os.path.join(P, "r01_configs.json"))
    # ---- ncu launch list
    src = os.path.join(G, f"launches_{tag}.csv")
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("snacb::", "")
        ns = float(r[vi]); tot += ns
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ns
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency",
             "(first 600 launches = warm-up + timed full-window steps + part of the sliced-mode leg; cold-cache, serialised: compare shares)",
             f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}"]
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n:58s} {c:8d} {ns / 1e6:10.3f} {ns / tot * 100:6.1f}% {ns / c / 1e3:9.1f}")
    open(os.path.join(P, "r01_ncu_launch_summary.txt"), "w").write("\n".join(lines) + "\n")
    with open(src, "rb") as f, gzip.open(os.path.join(P, "r01_ncu_launches.csv.gz"), "wb") as g:
        shutil.copyfileobj(f, g)
    # ---- ncu --set full of the chain kernels
    rep = os.path.join(G, f"prof_chain_{tag}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
            "sm__cycles_elapsed.max"]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    out, traffic = [], {}
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    for r in rows[2:]:
        short = r[hdr.index("Kernel Name")].split("(")[0]
        out.append(f"== {short}")
        for w in want:
            if w in hdr:
                i = hdr.index(w); out.append(f"   {w:75s} {r[i]:>16s} {units[i]}")
        st = sorted(((float(r[hdr.index(h)]), h) for h in stalls), reverse=True)[:7]
        out.append("   top stalls (warps stalled per issue-active cycle): " +
                   ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in st))
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic[short] = float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]]
    txt = ("ncu --set full --clock-control none, k_chain launches of one decode of 1024 four-frame windows (full windows, no trimming),\n"
           "fp16 operands; command: SNACB_NO_TRIM=1 ncu ... -k regex:k_chain -c 3 python tests/gpu_one.py 1024 fp16 1\n"
           "(per-launch times under ncu are cold-cache and serialised; shares, not absolutes)\n\n" + "\n".join(out) + "\n")
    open(os.path.join(P, "r01_chain_ncu.txt"), "w").write(txt)
    json.dump({"chain": {"1024": sum(traffic.values()) / len(traffic)}, "_per_kernel_bytes": traffic,
               "_what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the three k_chain launches of a step), "
                        "ncu --set full, B = 1024 full windows"}, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    # ---- in-kernel phase timing
    ph = ("In-kernel clock64 phase timing of k_chain (CTA 0, thread 0; SNACB_CHAIN_PROF=1), B = 1024 four-frame windows, fp16.\n"
          "cycles per tile: nz+ld = noise values + wait for the tile's TMA load; noise = NoiseBlock MMA + epilogue; per ResidualUnit L0..L2:\n"
          "pre = fetch of span neighbours + barrier, pro = in-place prologue, sync = barrier skew, mma+epi = MMA wait + epilogue;\n"
          "store/load issue = TMA stores of the tile and refills for the next one.\n\n--- full windows (SNACB_NO_TRIM=1)\n")
    keep = lambda f: "".join(l for l in open(os.path.join(G, f), errors="replace") if l.startswith(("chain", "k_chain", "   CTA")))
    ph += keep("chain_phases_full.txt")
    ph += "\n--- sliced call (trimmed; 'tiles' in the header is computed for the untrimmed range, divide accordingly)\n"
    ph += keep("chain_phases_sliced.txt")
    open(os.path.join(P, "r01_chain_phases.txt"), "w").write(ph)
    sw = b["sliding_window_mode"]
    print(f"value {b['value']:.0f} audio-s/s, {b['ms_per_step']:.2f} ms/step, e2e {b['e2e']['value']:.0f}; sliced {sw['windows_per_s']:.0f} windows/s "
          f"({sw['ms_per_step']:.2f} ms), roofline frac {b['roofline']['frac']:.3f}, share {b['roofline']['share_of_step']:.3f}")
    print("\n".join(lines[:10]))


if __name__ == "__main__":
    main()
