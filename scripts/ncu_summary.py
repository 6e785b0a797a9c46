"""Summarise an .ncu-rep (raw + source pages) into text: headline metrics, opcode mix, stall mix."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_lsu.sum"]
units = rows[1] if len(rows) > 2 else [""] * len(hdr)
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h = rows[hi]
isrc, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
byop, stall, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    try:
        n, s = int(r[ie]), int(r[iss] or 0)
    except Exception:
        continue
    t = r[isrc].strip()
    op = re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] if t else "?"
    if "IMAD.MOV" in t:
        op = "IMAD.MOV"
    byop[op] += n; stall[op] += s; tot += n
print(f"\nexecuted warp-instructions: {tot}")
ts = max(1, sum(stall.values()))
for op, n in byop.most_common(22):
    print(f"  {op:10s} {100*n/tot:5.1f}% of issued   {100*stall[op]/ts:5.1f}% of stall samples")
# stall reasons if present
stall_cols = [c for c in h if c.startswith("stall_")]
if stall_cols:
    agg = collections.Counter()
    for r in rows[hi + 1:]:
        for c in stall_cols:
            try:
                agg[c] += int(r[h.index(c)] or 0)
            except Exception:
                pass
    print("\nstall reasons (samples):")
    for c, n in agg.most_common(10):
        print(f"  {c:32s} {100*n/max(1,sum(agg.values())):5.1f}%")
