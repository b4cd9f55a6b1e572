"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_r1.ncu-rep profiles/r1_top_kernels.txt
  python tools/ncu_summary.py dram profiles/r2_ncu_dram.json BATCH rep1.ncu-rep [rep2.ncu-rep ...]
  python tools/ncu_summary.py sass eftpipe_b200/libeftb200.so profiles/r2_sass_summary.txt
  python tools/ncu_summary.py source gpurun_out/r2_resum_kernel.ncu-rep profiles/r2_resum_kernel_source.txt
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


STAGE_OF = {"resum_kernel": "resum", "antidiag_kernel": "antidiag", "ap_geom_kernel": "ap", "ap_apply_kernel": "ap",
            "regroup_kernel": "spectral", "like_gram_kernel": "likelihood", "like_vectors_kernel": "likelihood",
            "group_kernel": "group", "front_tails_kernel": "front"}


def dram(dst, batch, *reps):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel of every stage, from `ncu --set full`
    captures: what bench.py reports as `roofline.traffic` (scaled to its batch)"""
    import json
    import os

    stages, kernels = {}, {}
    for src in reps:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        ix = {h: i for i, h in enumerate(rr[0])}
        unit = lambda m: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[rr[1][ix[m]]]
        for r in rr[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
            byt = float(r[ix["dram__bytes_read.sum"]]) * unit("dram__bytes_read.sum") + \
                float(r[ix["dram__bytes_write.sum"]]) * unit("dram__bytes_write.sum")
            ent = dict(kernel=name, dram_bytes_per_launch=byt, batch=int(batch), capture=os.path.basename(src),
                       duration_us=float(r[ix["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[rr[1][ix["gpu__time_duration.sum"]]],
                       fp64_pipe_pct=float(r[ix["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]),
                       dmma_pipe_pct=float(r[ix["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"]])
                       if "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active" in ix else None,
                       registers=int(float(r[ix["launch__registers_per_thread"]])))
            kernels.setdefault(name, ent)
            st = STAGE_OF.get(name)
            if st and (st not in stages or byt > stages[st]["dram_bytes_per_launch"]):
                stages[st] = ent
    with open(dst, "w") as out:
        json.dump(dict(capture="ncu --set full --clock-control none, " + ", ".join(os.path.basename(r) for r in reps), stages=stages,
                       kernels=kernels), out, indent=1)


def sass(lib, dst):
    """per-kernel opcode counts of the built library: the evidence for DMMA (FP64 tensor core), UBLKCP (TMA bulk copy),
    UTMALDG (TMA tensor map), LDGSTS (cp.async), SYNCS (mbarrier)"""
    import re

    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    want = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "RED", "ATOMG", "MUFU"]
    cur, counts, order = None, {}, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            mm = re.search(r"(?:\(anonymous namespace\)::)?([A-Za-z_][A-Za-z_0-9]*(?:<[^>]*>)?)\(", dem.replace("void ", ""))
            cur = mm.group(1) if mm else dem[:90]
            while cur in counts:
                cur += "'"
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["_total"] += 1
    with open(dst, "w") as out:
        out.write(f"# static SASS opcode counts per kernel of {lib} (cuobjdump -sass, sm_100a)\n")
        out.write("# DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind), UBLKCP = TMA bulk copy (cp.async.bulk),\n")
        out.write("# UTMALDG = TMA tensor-map load, LDGSTS = cp.async, SYNCS = mbarrier\n")
        out.write(f"{'kernel':52s} {'total':>6} " + " ".join(f"{w:>7}" for w in want) + "\n")
        tot = collections.Counter()
        for k in order:
            c = counts[k]
            out.write(f"{k:52s} {c['_total']:6d} " + " ".join(f"{c[w]:7d}" for w in want) + "\n")
            tot.update(c)
        out.write(f"{'ALL':52s} {tot['_total']:6d} " + " ".join(f"{tot[w]:7d}" for w in want) + "\n")


def source(src, dst):
    """warp-stall samples and executed-instruction mix of one kernel from the source page of an `--import-source on` capture"""
    import re

    txt = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    si, so, ie = col["# Samples"], col["Source"], col["Instructions Executed"]
    stalls = collections.Counter()
    ops, ops_s = collections.Counter(), collections.Counter()
    for r in data:
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h and r[col[h]].isdigit():
                stalls[h[6:]] += int(r[col[h]])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[so])
        if m:
            ops[m.group(2)] += int(r[ie])
            ops_s[m.group(2)] += int(r[si])
    T, TI = sum(int(r[si]) for r in data), sum(ops.values())
    with open(dst, "w") as out:
        out.write(f"# {rows[0][1] if len(rows[0]) > 1 else ''}\n# source page of {src}: {T} warp-stall samples, {TI} warp instructions executed\n")
        out.write("\nstall reason (all samples)      share\n")
        for k, v in stalls.most_common(12):
            out.write(f"  {k:28s} {100 * v / T:5.1f}%\n")
        out.write("\nopcode        instructions   samples\n")
        for k, v in ops.most_common(16):
            out.write(f"  {k:10s} {100 * v / TI:9.1f}%  {100 * ops_s[k] / T:7.1f}%\n")
        out.write("\ntop instructions by samples (not DFMA / DMMA)\n")
        items = sorted(((int(r[si]), n, r[so].strip()[:70], int(r[ie])) for n, r in enumerate(data) if "DFMA" not in r[so] and "DMMA" not in r[so]), reverse=True)
        for sct, n, t, e in items[:20]:
            out.write(f"  {100 * sct / T:5.2f}%  #{n:<6d} executed {e:>10d}  {t}\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "dram":
        dram(sys.argv[2], sys.argv[3], *sys.argv[4:])
    elif mode == "sass":
        sass(sys.argv[2], sys.argv[3])
    elif mode == "source":
        source(sys.argv[2], sys.argv[3])
    else:
        if mode == "full" and len(sys.argv) > 4:
            full(sys.argv[2], sys.argv[3], max_ids=int(sys.argv[4]))
        else:
            {"launches": launches, "full": full}[mode](sys.argv[2], sys.argv[3])
