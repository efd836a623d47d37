#!/bin/bash
# round-2 GPU job E: the integer key-switching path -- its own tests first (transforms, path vs oracle), then the key-switching
# tests of the existing suite (all run the integer path by default at n >= 4096), then a short device-resident bench per algorithm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_ksint.py -q -x > $O/r2e_ksint.log 2>&1
tail -25 $O/r2e_ksint.log
python -m pytest tests/test_gpu_chunks.py tests/test_gpu_configs.py -q -x > $O/r2e_chunks.log 2>&1
tail -8 $O/r2e_chunks.log
for A in 1 0; do
  timeout 900 python bench.py --batch 256 --steps 2 --warmup 2 --ks-algo $A --no-cpu-baseline --no-e2e --no-configs > $O/bench_r2e_algo$A.json 2> $O/bench_r2e_algo$A.err
  tail -c 300 $O/bench_r2e_algo$A.err
  python - $A <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/bench_r2e_algo%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("algo", sys.argv[1], "value", round(l["value"], 1), "verified", l["verified"] and l["verified"]["ok"], "chunk", l["config"]["ciphertexts_per_key_pass"])
    for kk in l["roofline"]["kernels"]:
        print("   ", kk)
    for kk in l["roofline"]["alu"]["kernels"]:
        print("   ", {a: (round(b, 3) if isinstance(b, float) else b) for a, b in kk.items()})
    print("   step alu frac", l["roofline"]["alu"]["step_frac_of_alu_ceiling"], "ceil", l["roofline"]["alu"]["ceilings_warp_ops_per_s"])
except Exception as e:
    print("bench failed", e)
PY
done
