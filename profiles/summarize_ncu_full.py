#!/usr/bin/env python
"""Summarise one `ncu --set full` capture of the kernels of ONE evaluation: per-kernel key metrics (text) and the DRAM
traffic per phase (JSON, read by bench.py for roofline.traffic).
usage: summarize_ncu_full.py rep.ncu-rep out_summary.txt out_traffic.json "<command line of the capture>"
"""
import csv, io, json, subprocess, sys

rep, out_txt, out_json = sys.argv[1:4]
cmdline = sys.argv[4] if len(sys.argv) > 4 else ""
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct"]
PHASE = [("k_tc_w_prep", "w_prep"), ("k_tc_w_tiles", "w_prep"), ("k_tc_w_border", "w_prep"), ("k_tc_anchor_rows", "gemm_fwd"), ("k_w_prep", "w_prep"), ("k_entity_build", "entity_build"), ("k_tc_gemm<0", "gemm_fwd"),
         ("k_gemm_fwd", "gemm_fwd"), ("k_score", "score"), ("k_tc_gemm<1", "gemm_dw"), ("k_gemm_dw", "gemm_dw"),
         ("k_tc_gemm<2", "gemm_ge"), ("k_gemm_ge", "gemm_ge"), ("k_entity_pull", "entity_pull"), ("k_word", "word_pull"),
         ("k_finalize", "finalize"), ("k_final_reduce", "finalize")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
# keep the last COMPLETE evaluation of the capture: from a weight-prep launch up to (not including) the next one
names = [r[col["Kernel Name"]] for r in rows[2:]]
starts = [i for i, n in enumerate(names) if "w_prep" in n or "k_tc_w_tiles" in n]
if len(starts) >= 2:
    rows = rows[:2] + rows[2 + starts[-2]: 2 + starts[-1]]
traffic = {}
with open(out_txt, "w") as f:
    f.write(cmdline + "\n(one launch of every kernel of one evaluation; serialised, caches flushed between replays: compare shares, not absolutes)\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("(int)", "").replace("(bool)", "")
        f.write(f"\n{name}\n")
        for m in METRICS:
            if m in col:
                f.write(f"    {m:75s} {r[col[m]]:>12s} {units[col[m]]}\n")
        b = sum(float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        t = t / 1000 if units[col["gpu__time_duration.sum"]] in ("ns", "nsecond") else t
        f.write(f"    {'traffic = dram read+write (bytes)':75s} {b:12.0f}\n")
        ph = next((p for k, p in PHASE if name.startswith(k)), None)
        if ph:
            e = traffic.setdefault(ph, {"dram_bytes_per_eval": 0.0, "time_us": 0.0, "kernels": []})
            e["dram_bytes_per_eval"] += b; e["time_us"] += t; e["kernels"].append(name)
json.dump(traffic, open(out_json, "w"), indent=1)
print(open(out_txt).read()[:400])
