"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_r1.ncu-rep profiles/r1_top_kernels.txt
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v
    # kernels of the library (anonymous namespace of csrc/*.cu) vs. everything else: the cuBLAS DGEMM and the
    # DFMA probe are the roofline-denominator measurements of bench.py, torch kernels are its setup
    ours = {k: v for k, v in agg.items() if "<unnamed>::" in k and "probe_fp64" not in k}
    tot = sum(v[1] for v in ours.values())
    with open(dst, "w") as out:
        out.write(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({src})\n")
        out.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        out.write("# share = of the library's own kernels (the hot path); probe / cuBLAS / torch setup kernels listed below\n")
        out.write(f"{'avg us':>10} {'count':>6} {'share':>7}  kernel\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            out.write(f"{v[1] / v[0]:10.1f} {v[0]:6d} {100 * v[1] / tot:6.1f}%  {k}\n")
        out.write("# not part of the step (roofline probes, torch setup):\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if k not in ours:
                out.write(f"{v[1] / v[0]:10.1f} {v[0]:6d}      --  {k}\n")


WANT_SECTIONS = {"GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis",
                 "Scheduler Statistics", "Warp State Statistics", "Launch Statistics", "Occupancy"}
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__waves_per_multiprocessor",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active"]


def full(src, dst, max_ids=6):
    det = subprocess.run(["ncu", "-i", src, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(det.splitlines()))
    ix = {h: i for i, h in enumerate(rows[0])}
    with open(dst, "w") as out:
        out.write(f"# ncu --set full --clock-control none summary of {src}\n")
        rr = list(csv.reader(raw.splitlines()))
        rix = {h: i for i, h in enumerate(rr[0])}
        for r in rr[2 : 2 + max_ids]:
            out.write(f"\n== raw: ID {r[rix['ID']]} {r[rix['Kernel Name']][:80]}\n")
            for m in RAW:
                if m in rix:
                    out.write(f"   {m:90s} {r[rix[m]]} {rr[1][rix[m]]}\n")
        seen = None
        for r in rows[1:]:
            if len(r) < 15 or int(r[ix["ID"]]) >= max_ids:
                continue
            if r[ix["ID"]] != seen:
                seen = r[ix["ID"]]
                out.write(f"\n== details: ID {seen} {r[ix['Kernel Name']][:80]}  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}\n")
            if r[ix["Section Name"]] in WANT_SECTIONS and r[ix["Metric Name"]].strip():
                out.write(f"   {r[ix['Section Name']][:30]:30s} {r[ix['Metric Name']]:45s} {r[ix['Metric Value']]} {r[ix['Metric Unit']]}\n")
            if r[ix["Rule Name"]] in ("CPIStall", "SOLBottleneck", "UncoalescedGlobalAccess", "SharedMemoryConflicts", "WorkloadImbalance"):
                out.write(f"   RULE {r[ix['Rule Name']]}: {r[ix['Rule Description']][:260]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
