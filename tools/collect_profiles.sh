#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/collect_profiles.sh'): regenerates the raw material of profiles/r02_* into
# gpurun_out/r02/.  One GPU.  tools/ncu_summarize.py turns the ncu outputs into the committed summaries afterwards.
set -u
O=gpurun_out/r02
mkdir -p $O
T="timeout 600"
$T python bench.py --steps 30 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
$T python bench.py --dtype fp32 --no-mad --steps 10 --warmup 3 > $O/bench_1gpu_fp32.json 2> $O/bench_1gpu_fp32.err
$T python tools/stage_times.py > $O/stage_times.txt 2>&1
$T python tools/ablate_lanes.py > $O/ablate_lanes.txt 2>&1
$T python tools/bench_gemm.py > $O/bench_gemm.txt 2>&1
$T python tools/bench_ffn.py > $O/bench_ffn.txt 2>&1
$T python tools/hbm_table.py > $O/hbm_bound_kernels.txt 2>&1
$T python tools/trace_gemm.py 73584 288 288 3 1 ln > $O/trace_conv288.txt 2>&1
$T python tools/trace_ffn.py 36864 256 > $O/trace_ffn.txt 2>&1
M=gpu__time_duration.sum
$T ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --metrics $M,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --profile-from-start off -k regex:'gemm_tc_kernel|ffn_tc_kernel' --csv --log-file $O/gemm_metrics.csv python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ffn_tc_kernel -c 2 -f -o $O/ffn_full python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 22 -c 8 -f -o $O/gemm_heads_full python tools/one_step.py > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'preattn|local_attn_mma|xattn_mma|head_out_mma|tcn_fused|decode|nms' -c 10 -f -o $O/other_full python tools/one_step.py > /dev/null 2>&1
$T python tools/sweep.py charades > $O/charades_1gpu.jsonl 2> $O/charades.err
$T python tools/sweep.py lengths > $O/sweep_lengths.jsonl 2> $O/sweep_lengths.err
$T python tools/sweep.py nms --cpu-reference > $O/sweep_nms.jsonl 2> $O/sweep_nms.err
$T python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_cpu.json 2> $O/bench_reference_cpu.err
ls -la $O
