"""Top SASS lines of an `ncu --page source --csv` dump by executed instructions and by stall samples.
Usage: python tools/ncu_source_top.py src.csv [top]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    with open(path, newline="") as f:
        lines = f.readlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"Address"'))
    rows = list(csv.DictReader(lines[start:]))
    tot_inst = sum(int(r["Instructions Executed"] or 0) for r in rows)
    tot_samp = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"{len(rows)} SASS lines, {tot_inst} warp instructions, {tot_samp} samples")
    stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    print("-- by samples")
    for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
        st = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:3]
        print(f'{r["Address"][-5:]} {int(r["# Samples"] or 0):6d} samp {int(r["Instructions Executed"] or 0):9d} exec  '
              f'{r["Source"][:60]:60s} ' + " ".join(f"{c[6:]}={n}" for n, c in st if n))
    print("-- by executed instructions")
    for r in sorted(rows, key=lambda r: -int(r["Instructions Executed"] or 0))[:top]:
        print(f'{r["Address"][-5:]} {int(r["Instructions Executed"] or 0):9d} exec {int(r["# Samples"] or 0):6d} samp  {r["Source"][:80]}')


if __name__ == "__main__":
    main()


def segments(path, marks=("LDTM", "STTM", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "SYNCS.ARRIVE", "BAR.SYNC", "STG", "VOTE", "UTCHMMA", "UTCBAR", "USETMAXREG", "UBLKCP")):
    """samples / executed instructions between landmark instructions, in program order"""
    with open(path, newline="") as f:
        lines = f.readlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"Address"'))
    rows = list(csv.DictReader(lines[start:]))
    seg = ex = 0
    segstart = 0
    for i, r in enumerate(rows):
        s = int(r["# Samples"] or 0)
        if any(m in r["Source"] for m in marks):
            print(f"  [{seg:6d} samp {ex:10d} exec {i - segstart:4d} instrs] -> {r['Address'][-5:]} {r['Source'][:70]} ({s} samp, {r['Instructions Executed']} exec)")
            seg = ex = 0
            segstart = i
        seg += s
        ex += int(r["Instructions Executed"] or 0)
    print(f"  [{seg:6d} samp {ex:10d} exec] tail")
