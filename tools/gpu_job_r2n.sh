#!/bin/bash
# round-2 GPU job N: the whole GPU suite, smoke(), and the default bench line (all configs, e2e, CPU baseline) with the integer path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2n_tests.log 2>&1; tail -6 $O/r2n_tests.log
SB200_KS_FUSE_CRT=0 python -m pytest tests/test_gpu_ksint.py tests/test_gpu_configs.py -q > $O/r2n_tests_unfused.log 2>&1; tail -3 $O/r2n_tests_unfused.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2n_smoke.log 2>&1; tail -2 $O/r2n_smoke.log
timeout 1800 python bench.py > $O/bench_r2n.json 2> $O/bench_r2n.err
tail -c 400 $O/bench_r2n.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_r2n.json").read().strip().splitlines()[-1])
    print("value", l["value"], "e2e", l["e2e"]["value"], "verified", l["verified"], "cpu", l["cpu_baseline"] and l["cpu_baseline"]["value"])
    print("roofline frac", l["roofline"]["frac"], "kernel", l["roofline"]["kernel"], "alu step", l["roofline"]["alu"]["step_frac_of_alu_ceiling"])
    for kk in l["roofline"]["kernels"]:
        print("    ", kk)
    print("configs", json.dumps(l.get("configs"))[:3000])
except Exception as e:
    print("bench_r2n failed", e)
PY
