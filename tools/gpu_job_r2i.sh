#!/bin/bash
# round-2 GPU job I: is the product kernel bound by address translation?  Same prime count, rows 8x shorter (n = 8192), 8x the batch.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_ksint.py -q -x > $O/r2i_ksint.log 2>&1; tail -3 $O/r2i_ksint.log
SB200_KS_FUSE_CRT=1 python -m pytest tests/test_gpu_ksint.py -q -x > $O/r2i_ksint_f.log 2>&1; tail -3 $O/r2i_ksint_f.log
run() { # name, args...
  name=$1; shift
  timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-configs "$@" > $O/bench_r2i_$name.json 2> $O/bench_r2i_$name.err
  tail -c 300 $O/bench_r2i_$name.err
  python - $name <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/bench_r2i_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
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
run n65536 --batch 256
run n8192 --workload ckks_n8192_k32 --batch 2048 --no-verify
