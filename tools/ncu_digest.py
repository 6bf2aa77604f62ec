"""Digest of an .ncu-rep: key throughput metrics per kernel + stall breakdown + hottest SASS lines.
    python tools/ncu_digest.py report.ncu-rep [--top 25]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    print(d.get("Kernel Name", "?")[:110], d.get("Grid Size"), d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"  {k:85s} {d[k]:>14s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name",')
for blk in blocks[1:]:
    lines = blk.split("\n")
    name = lines[0][:100]
    rr = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    if len(rr) < 2:
        continue
    h = rr[0]
    ix = {x: i for i, x in enumerate(h)}
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    data = [r for r in rr[1:] if len(r) == len(h) and r[ix["# Samples"]].isdigit()]
    # drop the idle warps parked at the final barrier
    data2 = [r for r in data if "EXIT" not in r[ix["Source"]]]
    samp = sum(int(r[ix["# Samples"]]) for r in data2)
    tot = {s_: sum(int(r[ix[s_]]) for r in data2 if r[ix[s_]].isdigit()) for s_ in stalls}
    print("-" * 100)
    print(name, " samples (excl. EXIT):", samp, " SASS lines:", len(data))
    print("  " + "  ".join(f"{k[6:]}={100 * v / max(samp, 1):.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data2, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
        st = sorted([(int(r[ix[s_]]), s_[6:]) for s_ in stalls if r[ix[s_]].isdigit()], reverse=True)[:2]
        print(f"  {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']][:64]:64s} {st}")
