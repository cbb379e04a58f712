"""Dynamic (executed) SASS instruction counts per source line from an .ncu-rep with source counters, joined with the
line table of the library that was profiled (same build!).
usage: python tools/dyn_by_line.py <rep.ncu-rep> <lib.so> <mangled kernel> [opcode-regex] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
opre = re.compile(sys.argv[4]) if len(sys.argv) > 4 else re.compile(".")
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, dyn, k = None, [], 0
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
        dyn.append((d["Source"].strip(), int(d["Instructions Executed"]), int(d["# Samples"])))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "kernels" in f][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(f".text.{kern}:")][0]
end = next(i for i, l in enumerate(txt) if i > start and l.startswith("//--------------------- "))
cur, static = None, []
stack = []
for l in txt[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    else:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m:
            static.append((cur, m.group(1).strip()))
print("dynamic listing", len(dyn), "static listing", len(static))
assert len(dyn) == len(static), "the library is not the build that was profiled"
byline = collections.Counter()
tot = 0
for (src, ins), (dsrc, n, smp) in zip(static, dyn):
    op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
    assert op.split(".")[0] == re.sub(r"^@!?U?P\d+\s+", "", dsrc).split()[0].split(".")[0], (ins, dsrc)
    if opre.search(ins):
        byline[src] += n
        tot += n
print("matching executed warp instructions:", tot)
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aerobulk_b200", "csrc")
for (src, n) in byline.most_common(top):
    text = ""
    if src and os.path.exists(os.path.join(root, src[0])):
        text = open(os.path.join(root, src[0])).read().split("\n")[src[1] - 1].strip()[:110]
    print(f"{n:11d} {100 * n / tot:5.1f}%  {src[0] if src else '?'}:{src[1] if src else 0}  {text}")
