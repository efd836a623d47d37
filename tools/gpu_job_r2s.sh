#!/bin/bash
# round-2 GPU job S: the committed profile artifacts of the integer key-switching path
#  (1) ncu --set full of every kernel of one key-switching chunk at the benchmark's chunk size (436 ciphertexts); the report stays
#      on the box (too large for gpurun_out), its raw page comes back as CSV
#  (2) ncu launch list (durations only) of a short bench run
#  (3) [first run of this job] compute-sanitizer memcheck + racecheck over the integer path's small parity cases:
#      compute-sanitizer --tool memcheck  python -m pytest tests/test_gpu_ksint.py -q -x -k "4096 or bfv or bgv"
#      compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ksint.py -q -x -k "4096-bits0 or transforms_roundtrip_and_convolution and 4096"
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:'ks32_|ckks_tensor|ntt_' -s 10 -c 10 -o /tmp/r02_ksint_chunk436 -f \
     python bench.py --batch 436 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2s_ncu_full.log 2>&1
ls -la /tmp/r02_ksint_chunk436.ncu-rep
ncu -i /tmp/r02_ksint_chunk436.ncu-rep --page raw --csv > $O/r02_ksint_chunk436_raw.csv 2>/dev/null; wc -c $O/r02_ksint_chunk436_raw.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ksint_launches.csv \
     python bench.py --batch 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2s_ncu_list.log 2>&1
wc -l $O/r02_ksint_launches.csv
