#!/bin/bash
# round-2 GPU job A: parity of the new kernels, A/B of the limb multiply-accumulate, microbenchmarks, first full bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
(cd tools && ./bfly2 > ../$O/bfly2b.txt 2>&1)
python -m pytest tests -m gpu -q > $O/r2_tests2.log 2>&1
tail -25 $O/r2_tests2.log
# quick A/B at the headline shape (device-resident only)
python bench.py --batch 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/ab_limb.json 2> $O/ab_limb.err
SB200_NO_LIMB_MAC=1 python bench.py --batch 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/ab_nolimb.json 2> $O/ab_nolimb.err
python - <<'PY'
import json
for f in ("ab_limb", "ab_nolimb"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(l["value"], 1), "ct/s", [(k["name"], k["ms"], k["share"]) for k in l["roofline"]["kernels"][:4]])
        print("   alu", [(e["kernel"], round(e.get("achieved_clk_per_warp_bfly", 0), 1), round(e.get("floor_clk_per_warp_bfly", 0), 1)) for e in l["roofline"]["alu"]["kernels"][:4]])
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
# the full default line (verification, e2e, cpu baseline, cfg2-4, C++ harness)
timeout 900 python bench.py > $O/bench_r2a.json 2> $O/bench_r2a.err
tail -c 1500 $O/bench_r2a.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_r2a.json").read().strip().splitlines()[-1])
    print("value", l["value"], "e2e", l["e2e"]["value"], "verified", l["verified"], "cpu", l["cpu_baseline"] and l["cpu_baseline"]["value"])
    print("roofline", {k: l["roofline"][k] for k in ("kernel", "achieved", "frac", "B_reuse", "key_passes_per_step")})
    print("configs", json.dumps(l["configs"])[:1500])
except Exception as e:
    print("bench_r2a failed", e)
PY
