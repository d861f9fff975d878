"""Summarise an ncu report (--set full --import-source on): headline metrics, per-barrier wait samples and the
hottest SASS lines.  Usage: python profiles/ncu_stalls.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2:]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]
for v in vals:
    print("kernel:", v[hdr.index("Kernel Name")][:100])
    for w in want:
        if w in hdr:
            print(f"  {w:80s} {v[hdr.index(w)]} {rows[1][hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# one section per profiled launch: a title row, a header row, then the SASS lines
sections, cur = [], None
for r in rows:
    if len(r) > 10 and "# Samples" in r:
        cur = dict(h=r, data=[])
        sections.append(cur)
    elif cur is not None and len(r) > 10:
        cur["data"].append(r)
for si, sec in enumerate(sections):
    h, data = sec["h"], sec["data"]
    ix = {k: i for i, k in enumerate(h)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]]) for r in data) for k in st}
    print(f"==== launch {si}: total samples", tot)
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print(f"  {k:26s}{v:8d} {100 * v / max(tot, 1):5.1f}%")
    print("wait loops (mbarrier try_wait): address, operand, executions, samples in the loop")
    for i, r in enumerate(data):
        if "TRYWAIT" in r[ix["Source"]]:
            s = sum(int(x[ix["# Samples"]]) for x in data[i:i + 8])
            print("  ", r[ix["Address"]][-5:], r[ix["Source"]].strip()[40:80], r[ix["Instructions Executed"]], s)
    print("hottest lines")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
        s = {k: int(r[ix[k]]) for k in st}
        m = max(s, key=s.get)
        print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), m[6:].ljust(14), r[ix["Source"]].strip()[:90])
