#!/bin/bash
# quick GPU iteration: the integer path's tests + one device-resident bench line (batch 256) with the per-kernel table
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-q}
python -m pytest tests/test_gpu_ksint.py tests/test_gpu_chunks.py -q -x > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log
timeout 900 python bench.py --batch 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-configs > $O/bench_$T.json 2> $O/bench_$T.err
tail -c 300 $O/bench_$T.err
python - $T <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(l["value"], 1), "verified", l["verified"] and l["verified"]["ok"], "chunk", l["config"]["ciphertexts_per_key_pass"])
    alu = {e["kernel"]: e for e in l["roofline"]["alu"]["kernels"]}
    for kk in l["roofline"]["kernels"]:
        a = alu.get(kk["name"], {})
        print("    %-22s %8.2f ms  share %.3f  alu %.2f" % (kk["name"], kk["ms"], kk["share"], a.get("frac_of_alu_ceiling", 0)))
    print("    step alu frac", round(l["roofline"]["alu"]["step_frac_of_alu_ceiling"], 3))
except Exception as e:
    print("bench failed", e)
PY
