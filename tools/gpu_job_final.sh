#!/bin/bash
# final verification of a round: whole GPU suite, smoke(), the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-final}
python -m pytest tests -m gpu -q > $O/${T}_tests.log 2>&1; tail -4 $O/${T}_tests.log
timeout 900 tests/cpp/_bin/shim_test > $O/${T}_shim.log 2>&1; tail -3 $O/${T}_shim.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 1800 python bench.py > $O/bench_$T.json 2> $O/bench_$T.err
tail -c 300 $O/bench_$T.err
python - $T <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("value", l["value"], "e2e", l["e2e"]["value"], "verified", l["verified"]["ok"], "cpu", l["cpu_baseline"] and l["cpu_baseline"]["value"], "launches", l["gpu_launches"])
    r = l["roofline"]
    print("roofline frac", r["frac"], r["kernel"], "traffic", r["traffic"], "ratio", r["traffic_ratio"], "alu step", r["alu"]["step_frac_of_alu_ceiling"])
    c = l["configs"]
    print("cfg2", c["cfg2"]["value"], c["cfg2"]["e2e"]["value"], "cfg3", c["cfg3"]["value"], c["cfg3"]["e2e"]["value"], "cfg4", c["cfg4"]["value"], c["cfg4"]["e2e"]["value"])
    print("cpp", json.dumps(c["cpp"])[:900])
    print("ntt", l["ntt"]["achieved_GBps"], l["ntt"]["frac_of_hbm_peak"])
except Exception as e:
    print("bench failed", e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err; tail -c 500 $O/bench_ref_$T.json
