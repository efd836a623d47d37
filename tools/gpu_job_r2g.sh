#!/bin/bash
# round-2 GPU job G: ncu captures of the integer path's kernels (one launch each) at a 64-ciphertext chunk
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for V in 1 0; do
  SB200_KS_MAC_SMEM=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks32_mac' -s 1 -c 1 -o $O/r2g_mac_smem$V -f \
     python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2g_ncu_mac$V.log 2>&1
  ls -la $O/r2g_mac_smem$V.ncu-rep
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks32_crt|ks32_inv_local|ks32_fwd_local|ks32_fwd_outer|ks32_inv_outer|ckks_tensor' -s 6 -c 6 -o $O/r2g_rest -f \
     python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-verify > $O/r2g_ncu_rest.log 2>&1
ls -la $O/r2g_rest.ncu-rep
