"""Executed (dynamic) SASS instruction histogram by opcode from an .ncu-rep captured with --import-source on: which
instruction kinds fill the issue slots of a kernel, per pipe class.  No line table needed.
usage: python tools/dyn_by_opcode.py <rep.ncu-rep> [top=30]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, k = None, 0
byop, byfull = collections.Counter(), collections.Counter()
lanes = collections.Counter()
name = ""
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        if k > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        ins = re.sub(r"^@!?U?P\w+\s+", "", d["Source"].strip())
        full = ins.split()[0].rstrip(";")
        n = int(d["Instructions Executed"])
        byop[full.split(".")[0]] += n
        byfull[full] += n
        lanes[full.split(".")[0]] += n * float(d.get("Avg. Threads Executed") or 0)
tot = sum(byop.values())
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
f64 = sum(byop[o] for o in FP64)
print(f"executed warp instructions {tot}; FP64 pipe (DFMA DMUL DADD DSETP) {f64} = {100 * f64 / tot:.1f} %")
print("-- by base opcode")
for op, n in byop.most_common(top):
    print(f"{n:12d} {100 * n / tot:5.1f}%  lanes {lanes[op] / max(n, 1):5.2f}  {op}")
print("-- by full opcode (non-FP64)")
for op, n in [x for x in byfull.most_common() if x[0].split(".")[0] not in FP64][:top]:
    print(f"{n:12d} {100 * n / tot:5.1f}%  {op}")
