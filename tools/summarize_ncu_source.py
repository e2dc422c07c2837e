"""Per-source-line warp-stall samples of one kernel launch from an `ncu --set full --import-source on` report
(`ncu -i X.ncu-rep --page source --csv`): which lines of the kernel the sampled warps were stalled on.

    python tools/summarize_ncu_source.py gpurun_out/prof.ncu-rep [launch-index] > profiles/ncu_source_<tag>.txt
"""
import csv
import io
import subprocess
import sys


def main(path, which):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    # the CSV is a sequence of blocks: "File Path", "Function Name", header row, rows...; one block per (launch, file)
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "File Path":
            cur = {"file": row[1], "func": None, "hdr": None, "rows": []}
            blocks.append(cur)
        elif row[0] == "Function Name" and cur is not None:
            cur["func"] = row[1]
        elif row[0] == "Line No" and cur is not None:
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None:
            cur["rows"].append(row)
    # group consecutive blocks by kernel launch: a new launch starts when the same file repeats for the same function
    launches, seen, func = [], None, None
    for b in blocks:
        key = (b["func"], b["file"])
        if seen is None or key in seen or b["func"] != func:
            launches.append([])
            seen, func = set(), b["func"]
        seen.add(key)
        launches[-1].append(b)
    print("# %s: %d kernel launches with source pages; showing launch %d" % (path, len(launches), which))
    total = 0
    lines = []
    for b in launches[which]:
        h = b["hdr"]
        i_line, i_src, i_stall, i_inst = h.index("Line No"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        for r in b["rows"]:
            if r[i_line] and r[i_line] != "-":                      # a source line (SASS rows have an empty line number)
                try:
                    st, ins = int(r[i_stall]), int(r[i_inst])
                except ValueError:
                    continue
                total += st
                if st:
                    lines.append((st, ins, b["file"].split("/")[-1], r[i_line], r[i_src].strip()[:110]))
    print("# function: %s" % launches[which][0]["func"])
    print("# %d warp-stall samples in total; lines with >= 1.5 %% of them:" % total)
    for st, ins, f, ln, src in sorted(lines, reverse=True):
        if st >= 0.015 * total:
            print("%5.1f%%  %6d samples  %8d warp-instr  %s:%s  %s" % (100.0 * st / max(total, 1), st, ins, f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
