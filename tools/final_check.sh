#!/bin/bash
# End-of-round check on a GPU box: GPU test suite, smoke, both bench arms (outputs under gpurun_out/).
tag=${1:-final}
python -m pytest tests -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("value %.0f Mpt/s | e2e %.0f | e2e_pageable %.0f | launches %d | clocks %s | roofline %s %.3f (%s) | parity %s" % (
    d["value"] / 1e6, d["e2e"]["value"] / 1e6, (d["e2e_pageable"]["value"] or 0) / 1e6, d["gpu_launches"], d["clocks"],
    d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"]["variant"], d["parity"]))
print([(k["variant"], round(k["avg_launch_ms"], 2), round(k["fp64_frac"], 3)) for k in d["kernels"]])
PY
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/bench_ref_$tag.json
tail -c 400 gpurun_out/bench_ref_$tag.json
