"""Executed SASS instructions of one kernel attributed to the STATEMENTS of a chosen function (full inline chains from
`nvdisasm -gi`), from an .ncu-rep captured with --import-source on (same build of the library!).

For every statement of the solver (the frame inlined directly into the kernel) it prints: warp instructions executed,
share, FP64-pipe instructions, average active lanes, and the lane-slots lost to divergence (n * (32 - lanes)) --
where the time goes and where the divergence is.

usage: python tools/dyn_by_callsite.py <rep.ncu-rep> <lib.so> <mangled kernel> [depth=1] [top=40]
  depth 1: statements of the function inlined into the kernel (solve_*); 2: one level deeper; 0: kernel statements"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 1
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")

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
        dyn.append((d["Source"].strip(), int(d["Instructions Executed"]), float(d["Avg. Threads Executed"] or 0)))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "kernels" in f][0]
txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(f".text.{kern}:")][0]
end = next(i for i, l in enumerate(txt) if i > start and l.startswith("//--------------------- "))
chain, static, fresh = [], [], True
for l in txt[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        static.append((tuple(chain), m.group(1).strip()))
        fresh = True
assert len(dyn) == len(static), (len(dyn), len(static), "the library is not the build that was profiled")

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aerobulk_b200", "csrc")
agg = collections.defaultdict(lambda: [0, 0, 0.0])
tot = tot64 = 0
lost_tot = 0.0
for (ch, ins), (dsrc, n, thr) in zip(static, dyn):
    op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
    # chain is innermost .. outermost (kernel); pick the frame `depth` below the kernel
    key = ch[max(0, len(ch) - 1 - depth)] if ch else None
    a = agg[key]
    a[0] += n
    if op in FP64:
        a[1] += n
        tot64 += n
    a[2] += n * (32.0 - thr)
    tot += n
    lost_tot += n * (32.0 - thr)
print(f"executed warp instructions {tot}, FP64-pipe {tot64} ({100 * tot64 / tot:.1f} %), lane-slots lost {100 * lost_tot / (32 * tot):.1f} %")
print(f"{'instr':>11s} {'share':>6s} {'fp64%':>6s} {'lanes':>6s} {'lost%':>6s}  statement")
for key, (n, n64, lost) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key and os.path.exists(os.path.join(root, key[0])):
        text = open(os.path.join(root, key[0])).read().split("\n")[key[1] - 1].strip()[:100]
    print(f"{n:11d} {100 * n / tot:5.1f}% {100 * n64 / max(n, 1):5.1f}% {32 - lost / max(n, 1):6.2f} {100 * lost / max(lost_tot, 1):5.1f}%  "
          f"{key[0] if key else '?'}:{key[1] if key else 0}  {text}")
