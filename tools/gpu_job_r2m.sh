#!/bin/bash
# round-2 GPU job M: ncu (full set, source counters) of the key-tile product kernel and the local passes at a 64-ciphertext chunk
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks32_mac|ks32_inv_local|ks32_fwd_local' -s 3 -c 3 -o $O/r2m_mac_local -f \
     python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2m_ncu.log 2>&1
ls -la $O/r2m_mac_local.ncu-rep
