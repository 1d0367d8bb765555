#!/bin/bash
# Collects the ncu evidence kept under profiles/ (run on the GPU box through gpurun, one GPU).
# usage: bash scripts/run_profiles.sh <tag>     -> gpurun_out/<tag>_*
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
RAW="--page raw --csv"
METRICS="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed"

# 1) launch list of one prefill batch + 3 decode steps (eager, every kernel a separate launch)
timeout 600 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv \
  --log-file $OUT/${TAG}_launches_profile_step.csv python scripts/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1

# 2) full captures of the dominant kernels (one launch each, after warm-up launches)
cap() {  # name, kernel regex, script args...
  local name=$1 kern=$2; shift 2
  timeout 300 $NCU --set full --import-source on -k regex:$kern -s 3 -c 1 -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
  ncu -i $OUT/${TAG}_$name.ncu-rep $RAW > $OUT/${TAG}_$name.raw.csv 2>/dev/null
}
cap gemm_gateup gemm_pair_kernel python scripts/prof_gemm.py 8192
cap attn_prefill_8x1024 attn_prefill python scripts/prof_attn.py prefill 8 1024
cap attn_prefill_2x4096 attn_prefill python scripts/prof_attn.py prefill 2 4096
cap attn_decode_64x1088 attn_decode python scripts/prof_attn.py decode 64 1088
cap attn_decode_160x1088 attn_decode python scripts/prof_attn.py decode 160 1088
cap gemm_decode_qkv_m128 gemm_pair_decode python scripts/prof_gemm_decode.py 128 6144 4096
cap gemm_decode_gateup_m128 gemm_pair_decode python scripts/prof_gemm_decode.py 128 28672 4096
ls -la $OUT
