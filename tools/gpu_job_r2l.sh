#!/bin/bash
# round-2 GPU job L: key-tile product kernel with the digit ring, reconstruction kernel with next-prime prefetch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_ksint.py tests/test_gpu_chunks.py -q -x > $O/r2l_ksint.log 2>&1; tail -3 $O/r2l_ksint.log
SB200_KS_MAC_TILE=0 python -m pytest tests/test_gpu_ksint.py -q -x > $O/r2l_ksint_notile.log 2>&1; tail -3 $O/r2l_ksint_notile.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --batch 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-configs > $O/bench_r2l_$name.json 2> $O/bench_r2l_$name.err
  tail -c 300 $O/bench_r2l_$name.err
  python - $name <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/bench_r2l_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
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
run tile X=1
run tile_fused SB200_KS_FUSE_CRT=1
