#!/bin/bash
# round-2 GPU job Q: ncu of the fused reconstruction kernel and the 64-bit result transform at a 64-ciphertext chunk
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks32_crt|ntt_fwd_local|ntt_fwd_col' -s 3 -c 3 -o $O/r2q_crt_res -f \
     python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2q_ncu.log 2>&1
ls -la $O/r2q_crt_res.ncu-rep
