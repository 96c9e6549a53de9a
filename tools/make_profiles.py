#!/usr/bin/env python
"""Turn the raw gpurun_out/ artefacts of tools/gpu_round.sh <tag> into the committed summaries under profiles/."""
import csv, json, subprocess, sys, os, shutil
tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if len(r) == len(hdr) and r[0].isdigit()]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg, order = {}, []
for r in data:
    short = r[ki].split("(")[0][-90:]
    if short not in agg:
        agg[short] = [0, 0.0, r[gi], r[bi]]; order.append(short)
    agg[short][0] += 1; agg[short][1] += float(r[vi].replace(",", ""))
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write(f"# ncu launch list ({tag}): `python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES).\n")
    f.write("First 400 launches of the process: corpus generation (torch), ingest, then the searches.\n\n| kernel | launches | total us | avg us | grid | block |\n|---|---|---|---|---|---|\n")
    for k in order:
        n, t, g, b = agg[k]
        f.write(f"| `{k}` | {n} | {t/1e3:.1f} | {t/1e3/n:.2f} | {g} | {b} |\n")
    prs = {k: v for k, v in agg.items() if "prs::" in k and "ingest" not in k}
    tot = sum(v[1] for v in prs.values())
    f.write("\nShare of a search step (libprs kernels only, ingest excluded):\n\n")
    for k, v in prs.items():
        f.write(f"- `{k}`: {100*v[1]/tot:.1f} % ({v[1]/1e3/v[0]:.1f} us per launch)\n")
out = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"]
mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
with open(f"profiles/{tag}_scan_kernel_ncu_full.md", "w") as f:
    f.write(f"# ncu --set full of the scan kernel ({tag})\n\n`ncu --set full --clock-control none --import-source on -k regex:flat_scan -s 4 -c 2` on "
            "`python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline` (1M x 768 fp16, B=64, k=10, IP).\n"
            "Times under ncu are replayed/cold: the bench's CUDA-event time is the number of record.\n\n")
    for r in rows[2:]:
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            if w in hdr:
                f.write(f"| {w} | {r[hdr.index(w)]} | {rows[1][hdr.index(w)]} |\n")
        f.write("\n")
    r = rows[2]
    traffic = float(r[hdr.index("dram__bytes_read.sum")]) * mul[rows[1][hdr.index("dram__bytes_read.sum")]] + \
              float(r[hdr.index("dram__bytes_write.sum")]) * mul[rows[1][hdr.index("dram__bytes_write.sum")]]
    f.write(f"traffic per launch = dram read + write = {traffic/1e9:.4f} GB vs algorithmic 1.5360 GB (ratio {traffic/1.536e9:.4f}).\n")
json.dump({"fp16_1000000x768_b64_tcgen05": traffic, "_source": f"profiles/{tag}_scan_kernel_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)"},
          open("profiles/traffic.json", "w"), indent=1)
shutil.copy(f"gpurun_out/bench_{tag}.json", f"profiles/{tag}_bench.json")
for extra in ("aux_sparse_1m.json", "aux_sparse_10m.json", "aux_pool.json"):
    if os.path.exists(f"gpurun_out/{extra}"):
        shutil.copy(f"gpurun_out/{extra}", f"profiles/{tag}_{extra}")
print("profiles written for", tag)
