#!/bin/bash
# round-2 GPU job F: integer key-switching path, second iteration (key tile in shared memory, one-word Barrett, fused outer-inverse +
# reconstruction): its tests in both reconstruction modes, the whole GPU suite, short device-resident bench lines per variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
SB200_KS_FUSE_CRT=1 python -m pytest tests/test_gpu_ksint.py -q -x > $O/r2f_ksint_fused.log 2>&1; tail -4 $O/r2f_ksint_fused.log
python -m pytest tests -m gpu -q -x > $O/r2f_tests.log 2>&1; tail -8 $O/r2f_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --batch 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-configs > $O/bench_r2f_$name.json 2> $O/bench_r2f_$name.err
  tail -c 300 $O/bench_r2f_$name.err
  python - $name <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/bench_r2f_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(l["value"], 1), "verified", l["verified"] and l["verified"]["ok"], "chunk", l["config"]["ciphertexts_per_key_pass"])
    alu = {e["kernel"]: e for e in l["roofline"]["alu"]["kernels"]}
    for kk in l["roofline"]["kernels"]:
        a = alu.get(kk["name"], {})
        print("    %-22s %8.2f ms  share %.3f  alu %.2f" % (kk["name"], kk["ms"], kk["share"], a.get("frac_of_alu_ceiling", 0)))
    print("    step alu frac", round(l["roofline"]["alu"]["step_frac_of_alu_ceiling"], 3))
except Exception as e:
    print("bench failed", e)
PY
}
run default X=1
run fused SB200_KS_FUSE_CRT=1
run nosmem SB200_KS_MAC_SMEM=0
