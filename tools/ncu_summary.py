"""Turns an .ncu-rep (brought back in gpurun_out/) into the text summary committed under
profiles/: headline metrics from the raw page + the top stalled SASS lines of the source page.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_k1_1cta.txt
"""
import csv
import io
import re
import subprocess
import sys

PAT = re.compile(
    r"^(gpu__time_duration\.sum|sm__cycles_elapsed\.avg\.per_second|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)"
    r"|sm__throughput\.avg\.pct|dram__bytes_(read|write)\.sum$|dram__bytes_read\.sum\.per_second|gpu__dram_throughput\.avg\.pct"
    r"|lts__throughput\.avg\.pct|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|launch__registers_per_thread$"
    r"|launch__grid_size|launch__block_size|launch__cluster|launch__shared_mem_per_block_dynamic|sm__warps_active\.avg\.per_cycle_active"
    r"|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum$|smsp__inst_executed\.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$"
    r"|sm__cycles_active\.avg$|sm__cycles_elapsed\.max$|launch__occupancy_limit_(registers|shared_mem)|launch__waves_per_multiprocessor"
    r"|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active)")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep}", "# (--set full --clock-control none; replayed, cold-cache: compare shares, not absolutes)", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## kernel: {name}")
        for h, u, v in zip(hdr, units, r):
            if PAT.search(h):
                lines.append(f"{h:78s} {v:>22s} {u}")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    his = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    names = [r[hdr.index("Kernel Name")] for r in rows[2:]] if "Kernel Name" in hdr else []
    for n, hi in enumerate(his):
        h = srows[hi]
        end = his[n + 1] if n + 1 < len(his) else len(srows)
        data = [r for r in srows[hi + 1:end] if len(r) == len(h) and (r[h.index("# Samples")] or "0").isdigit()]
        isamp, iex, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
        tot = sum(int(r[isamp] or 0) for r in data) or 1
        lines.append(f"## top stalled SASS instructions (warp-state samples) — kernel {n}: {(names[n] if n < len(names) else '?')[:60]}")
        lines.append(f"{'share':>7s} {'samples':>9s} {'executed':>11s}  sass")
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:12]:
            lines.append(f"{100.0 * int(r[isamp] or 0) / tot:6.2f}% {r[isamp]:>9s} {r[iex]:>11s}  {r[isrc][:100]}")
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"wrote {out} ({len(lines)} lines)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
