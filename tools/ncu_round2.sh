#!/bin/bash
# Round-2 ncu evidence, one gpurun call (1 GPU):  bash tools/ncu_round2.sh
#   1. launch list of the default offline-64 step (per-launch device time: compare SHARES with bench.py's roofline.gemm_ms_per_step)
#   2. --set full of the dominant kernel (gemm2_bf16_kernel<256>): DRAM bytes per launch (roofline.traffic), tensor-pipe utilisation
#   3. section captures of the two flash-attention forward kernels (--set full's SASS patching hangs on setmaxnreg kernels)
#   4. launch list of one eager training step
# Each capture runs only after the same command has exited 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CMD="python bench.py --workload offline64 --steps 1 --warmup 1 --no-cpu-baseline"
TR="python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline"
SEC="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy"

$CMD > $O/r02_plain_offline.log 2>&1 || { echo "plain offline run failed"; tail -5 $O/r02_plain_offline.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r02_ncu_launches_offline64_final.csv $CMD > $O/r02_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 40 -c 4 -f -o $O/r02_ncu_gemm2 $CMD > $O/r02_ncu2.log 2>&1
echo "gemm2 set full rc=$?"
timeout 600 ncu $SEC --clock-control none -k regex:attn_vit2_kernel -s 3 -c 2 -f -o $O/r02_ncu_attn_vit2 $CMD > $O/r02_ncu3.log 2>&1
echo "attn_vit2 sections rc=$?"
timeout 600 ncu $SEC --clock-control none -k regex:attn_gqa2_kernel -s 3 -c 2 -f -o $O/r02_ncu_attn_gqa2 $CMD > $O/r02_ncu4.log 2>&1
echo "attn_gqa2 sections rc=$?"
SLB_TRAIN_GRAPHS=0 $TR > $O/r02_plain_train.log 2>&1 || { echo "plain train run failed"; tail -5 $O/r02_plain_train.log; exit 1; }
SLB_TRAIN_GRAPHS=0 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file $O/r02_ncu_launches_train.csv $TR > $O/r02_ncu5.log 2>&1
echo "train launch list rc=$?"
ls -la $O | grep r02_ncu
