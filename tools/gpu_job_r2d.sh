#!/bin/bash
# round-2 GPU job D: the pieces added after job B (symmetric / public-key encryption, CKKS encoder, their C++ stand-ins), then the
# whole GPU suite, smoke() and the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "encrypt_zero or ckks_encoder" > $O/r2d_new.log 2>&1
tail -15 $O/r2d_new.log
timeout 600 tests/cpp/_bin/shim_test > $O/r2d_shim.log 2>&1; tail -5 $O/r2d_shim.log
python -m pytest tests -m gpu -q > $O/r2d_tests.log 2>&1
tail -8 $O/r2d_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2d_smoke.log 2>&1; tail -2 $O/r2d_smoke.log
timeout 1500 python bench.py > $O/bench_r2d.json 2> $O/bench_r2d.err
tail -c 600 $O/bench_r2d.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_r2d.json").read().strip().splitlines()[-1])
    print("value", l["value"], "e2e", l["e2e"]["value"], "verified", l["verified"]["ok"], "cpu", l["cpu_baseline"] and l["cpu_baseline"]["value"])
    print("configs", json.dumps(l.get("configs"))[:1500])
except Exception as e:
    print("bench_r2d failed", e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_r2d.json 2> $O/bench_ref_r2d.err; tail -c 400 $O/bench_ref_r2d.json
# DRAM traffic of the two dominant kernels at the chunk size the benchmark runs (110 ciphertexts per key pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks_local_mac|ks_digit_col' -s 2 -c 2 -o $O/r02_chunk110 python bench.py --batch 110 --steps 1 --warmup 1 --scratch-gib 64 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/ncu_chunk110.log 2>&1
ls -la $O/r02_chunk110.ncu-rep
ncu -i $O/r02_chunk110.ncu-rep --page raw --csv > $O/r02_chunk110_raw.csv 2>/dev/null; wc -c $O/r02_chunk110_raw.csv
