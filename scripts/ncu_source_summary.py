"""Per-phase stall summary of one kernel from a .ncu-rep captured with --set full --import-source on:
the SASS is cut at its TMEM / mbarrier / TMA instructions and the warp-state samples of each segment are summed
by stall reason.  Usage: python scripts/ncu_source_summary.py report.ncu-rep [min_instruction_index]"""
import csv
import re
import subprocess
import sys


def main(path, lo):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    smp = lambda i: int(data[i][ix["# Samples"]])
    total = sum(smp(i) for i in range(len(data)))
    print(f"{len(data)} SASS instructions, {total} warp-state samples")
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    print("all:", ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    marks = [i for i, r in enumerate(data)
             if re.search(r"TRYWAIT|LDTM|STTM|UTMASTG|UTMALDG|USETMAXREG|BAR\.SYNC|SYNCS\.ARRIVE|UTCHMMA|EXIT", r[ix["Source"]])
             and int(r[ix["Instructions Executed"]]) > 0]
    prev = 0
    for i in marks:
        if i >= lo:
            seg = sum(smp(j) for j in range(prev, i + 1))
            d = {s: sum(int(data[j][ix[s]]) for j in range(prev, i + 1)) for s in stalls}
            top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(d.items(), key=lambda x: -x[1])[:4] if v > 0)
            print(f"{i:5d} samples {seg:5d} ({100 * seg / total:4.1f}%) exec {data[i][ix['Instructions Executed']]:>8s}  "
                  f"{data[i][ix['Source']].strip()[:52]:52s} {top}")
        prev = i + 1


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
