#!/bin/bash
# round-2 GPU job J: ncu (full set, source counters) of the current product and reconstruction kernels at a 64-ciphertext chunk
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks32_mac|ks32_crt' -s 2 -c 2 -o $O/r2j_mac_crt -f \
     python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2j_ncu.log 2>&1
ls -la $O/r2j_mac_crt.ncu-rep
