#!/usr/bin/env bash
# Round-2 evidence capture (run under gpurun): ncu launch list of one warm denoise step, --set full captures of the fused
# kernels and of gemm_tc_kernel, compact metric tables, racecheck of the fused path.  Numbers printed under ncu are never
# bench values.
set -u
export MCM_DUAL=0 MCM_GRAPH=0
O=gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_step_t2m.csv \
    python tools/profile_step.py t2m 256 > $O/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'fused_block_kernel|sa_front_kernel|sa_tail_kernel' -c 3 \
    -f -o $O/r02_fused_full python tools/profile_step.py t2m 256 > $O/r02_ncu_fused.log 2>&1
ncu -i $O/r02_fused_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_metrics_table.py \
    "ncu --set full --clock-control none, t2m B=256 T=196, single stream (MCM_DUAL=0, eager): first launch of sa_front_kernel, sa_tail_kernel, fused_block_kernel of a warm denoise step (round 2 build)" \
    > $O/r02_fused_kernels_metrics.csv
ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 8 \
    -f -o $O/r02_gemm_tc_full python tools/profile_step.py t2m 256 > $O/r02_ncu_gemm.log 2>&1
ncu -i $O/r02_gemm_tc_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_metrics_table.py \
    "ncu --set full --clock-control none, t2m B=256 T=196, single stream: the first 8 gemm_tc_kernel launches of a warm denoise step (timestep embed 0/2, AdaLN emb GEMM, joint_embed, then layer 0: token-softmax context GEMM ...) (round 2 build)" \
    > $O/r02_gemm_tc_kernel_metrics.csv
ncu -i $O/r02_fused_full.ncu-rep --page source --csv --kernel-name fused_block_kernel 2>/dev/null | head -c 3000000 > $O/r02_fused_block_source.csv
rm -f $O/r02_gemm_tc_full.ncu-rep
ls -la $O | tail -12
