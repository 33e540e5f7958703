"""Per-opcode totals from the source page of an ncu report: instructions, shared wavefronts, L1 tag requests, samples.
Usage: python tools/ncu_opcodes.py report.ncu-rep n_cells [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cells = sys.argv[1], float(sys.argv[2])
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the CSV is a sequence of kernels: a "Kernel Name" row, a header row, then instruction rows
    k = 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name":
            name, hdr = rows[k][1], rows[k + 1]
            j = k + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                j += 1
            if pat is None or pat.search(name):
                report(name, hdr, rows[k + 2:j], cells)
            k = j
        else:
            k += 1


def report(name, hdr, rows, cells):
    ix = {h: i for i, h in enumerate(hdr)}
    tot = {}
    nsamp = 0
    for r in rows:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].strip().split()
        op = src[1] if src[0].startswith("@") else src[0]
        parts = op.split(".")
        key = parts[0] + ("." + parts[1] if len(parts) > 1 and parts[0] in ("LDS", "STS", "LDG", "RED", "REDG", "LDGSTS", "ATOMG") else "")
        t = tot.setdefault(key, [0, 0, 0, 0, 0, 0])
        t[0] += int(r[ix["Instructions Executed"]])
        t[1] += int(r[ix["L1 Wavefronts Shared"]])
        t[2] += int(r[ix["L1 Wavefronts Shared Excessive"]])
        t[3] += int(r[ix["L1 Tag Requests Global"]])
        t[4] += int(r[ix["L2 Theoretical Sectors Global"]])
        t[5] += int(r[ix["# Samples"]])
        nsamp += int(r[ix["# Samples"]])
    print("===", name[:110])
    print("%-10s %10s %10s %10s %10s %10s %8s" % ("opcode", "inst/cell", "shwf/cell", "excess", "tags/cell", "l2sec/cell", "samples%"))
    for key, v in sorted(tot.items(), key=lambda kv: -kv[1][5])[:22]:
        print("%-10s %10.2f %10.2f %10.2f %10.2f %10.2f %8.1f" % (key, v[0] / cells, v[1] / cells, v[2] / cells, v[3] / cells, v[4] / cells,
                                                                   100.0 * v[5] / max(nsamp, 1)))


if __name__ == "__main__":
    main()
