#!/bin/bash
# round-2 GPU job C (8 GPUs): what the host side can deliver to 1/2/4/8 GPUs at once, the bench under torchrun, the C++ dispatcher
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi topo -m > $O/topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > $O/lscpu.txt 2>&1
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n tools/pcie_bw.py --secs 2 >> $O/pcie_bw.jsonl 2>> $O/pcie_bw.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/pcie_bw.py --secs 2 --no-bind >> $O/pcie_bw.jsonl 2>> $O/pcie_bw.err
cat $O/pcie_bw.jsonl | cut -c1-600
timeout 300 tests/cpp/_bin/multi_test 0 > $O/multi_test8.log 2>&1; tail -3 $O/multi_test8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 600 $O/bench_n8.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
    print("N=8 value", l["value"], "e2e", l["e2e"]["value"], "verified", l["verified"])
    print(json.dumps(l["e2e"]["per_rank"]))
except Exception as e:
    print("bench_n8 failed", e)
PY
