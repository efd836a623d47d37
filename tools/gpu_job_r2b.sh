#!/bin/bash
# round-2 GPU job B: parity of the final kernels, bench line, launch list and full ncu capture of one key-switching chunk
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_tests3.log 2>&1
tail -5 $O/r2_tests3.log
for gib in 8 48; do
  python bench.py --batch 512 --steps 2 --warmup 2 --scratch-gib $gib --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/quick_$gib.json 2> $O/quick_$gib.err
done
python - <<'PY'
import json
for f in ("quick_8", "quick_48"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(l["value"], 1), "ct/s chunk", l["config"]["ciphertexts_per_key_pass"], [(k["name"], k["ms"], k["share"]) for k in l["roofline"]["kernels"][:3]])
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
(cd tools && ./bfly2 > ../$O/bfly2c.txt 2>&1)
timeout 1200 python bench.py > $O/bench_r2b.json 2> $O/bench_r2b.err
tail -c 800 $O/bench_r2b.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_r2b.json").read().strip().splitlines()[-1])
    print("value", l["value"], "e2e", l["e2e"]["value"], "verified", l["verified"]["ok"], l["verified"]["indices"], "cpu", l["cpu_baseline"] and l["cpu_baseline"]["value"])
    print("roofline", {k: l["roofline"][k] for k in ("kernel", "achieved", "frac", "B_reuse", "key_passes_per_step")}, "alu step frac", l["roofline"]["alu"]["step_frac_of_alu_ceiling"])
except Exception as e:
    print("bench_r2b failed", e)
PY
# launch list of two steps (shares of the step) and a full capture of the kernels of one chunk
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches.csv python bench.py --batch 26 --steps 2 --warmup 1 --scratch-gib 8 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/ncu_list.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ks_|ckks_tensor|ntt_' -s 9 -c 9 -o $O/r02_keyswitch python bench.py --batch 13 --steps 1 --warmup 1 --scratch-gib 8 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/ncu_full.log 2>&1
ls -la $O/r02_keyswitch.ncu-rep
