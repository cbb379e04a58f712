#!/bin/bash
# Evidence on N GPUs of one box (gpurun --gpus N): concurrent PCIe probe, the multi-device tests, bench.py under torchrun.
#   tools/multi_gpu_report.sh <N> <tag> [bench args]
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR tools/pcie_probe_ranks.py > gpurun_out/pcie_probe_${N}gpu_${TAG}.txt 2>gpurun_out/pcie_probe_${N}gpu_${TAG}.err
cat gpurun_out/pcie_probe_${N}gpu_${TAG}.txt
python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/multi_device_tests_${N}gpu_${TAG}.txt
$TR bench.py --gpus $N "$@" > gpurun_out/bench_${N}gpu_${TAG}.json 2>gpurun_out/bench_${N}gpu_${TAG}.err
tail -c 400 gpurun_out/bench_${N}gpu_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_${TAG}.json").read().strip().splitlines()[-1])
    print("N=$N value %.3f Gpt/s  ms/step %.2f  e2e %.3f Gpt/s  parity %s" % (d["value"]/1e9, d["ms_per_step"], (d["e2e"]["value"] or 0)/1e9, d["parity"].get("max_scaled_err_all_ranks", d["parity"].get("max_scaled_err"))))
    print([round(k["avg_launch_ms"],2) for k in d["kernels"]])
except Exception as e:
    print("bench line unreadable:", e)
PY
