#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/collect_profiles.sh'): regenerates the raw material of profiles/r02_* into
# gpurun_out/r02/.  One GPU.  tools/ncu_summarize.py turns the ncu outputs into the committed summaries afterwards.
set -u
O=gpurun_out/r02
mkdir -p $O
T="timeout 600"
$T python bench.py --steps 48 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
$T python bench.py --dtype fp32 --no-mad --steps 10 --warmup 3 > $O/bench_1gpu_fp32.json 2> $O/bench_1gpu_fp32.err
# the same step with full-width launches and 4 lanes (the configuration of the first half of the round)
DECAF_LANE_GEMM_SMS=0 $T python bench.py --no-cpu-baseline --no-mad --lanes 4 --steps 48 --warmup 5 > $O/bench_1gpu_fullwidth_4lanes.json 2> /dev/null
$T python tools/timeline.py --videos 32 > $O/timeline_48sms_8lanes.txt 2>&1
DECAF_LANE_GEMM_SMS=0 $T python tools/timeline.py --lanes 4 --videos 32 > $O/timeline_148sms_4lanes.txt 2>&1
$T python tools/stage_times.py > $O/stage_times.txt 2>&1
$T python tools/ablate_lanes.py > $O/ablate_lanes.txt 2>&1
$T python tools/bench_gemm.py > $O/bench_gemm.txt 2>&1
$T python tools/bench_ffn.py > $O/bench_ffn.txt 2>&1
$T python tools/hbm_table.py > $O/hbm_bound_kernels.txt 2>&1
M=gpu__time_duration.sum
$T ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --metrics $M,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --profile-from-start off -k regex:'gemm_tc_kernel|ffn_tc_kernel' --csv --log-file $O/gemm_metrics.csv python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ffn_tc_kernel -c 2 -f -o $O/ffn_full python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 22 -c 8 -f -o $O/gemm_heads_full python tools/one_step.py > /dev/null 2>&1
DECAF_TEXT_TC=0 $T ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'preattn|local_attn_mma|xattn_mma_kernel|xattn_pack|head_out_mma|tcn_fused|decode_kernel|softnms|saliency_kernel|select_kernel|map_combine|adaln' -c 24 -f -o $O/other_full python tools/one_step.py > /dev/null 2>&1
# summaries on the box: the .ncu-rep files together exceed what gpurun copies back (64 MiB)
python tools/ncu_summarize.py full $O/ffn_full.ncu-rep > $O/ffn_ncu_full_summary.txt 2>&1
python tools/ncu_summarize.py full $O/gemm_heads_full.ncu-rep > $O/gemm_ncu_full_summary.txt 2>&1
python tools/ncu_summarize.py full $O/other_full.ncu-rep > $O/other_kernels_ncu_full_summary.txt 2>&1
python tools/ncu_summarize.py shares $O/launches_step.csv > $O/step_kernel_shares.txt 2>&1
python tools/ncu_summarize.py traffic $O/gemm_metrics.csv > $O/gemm_traffic.json 2>&1
rm -f $O/gemm_heads_full.ncu-rep $O/other_full.ncu-rep
$T python tools/sweep.py charades > $O/charades_1gpu.jsonl 2> $O/charades.err
$T python tools/sweep.py lengths > $O/sweep_lengths.jsonl 2> $O/sweep_lengths.err
$T python tools/sweep.py nms --cpu-reference > $O/sweep_nms.jsonl 2> $O/sweep_nms.err
$T python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_cpu.json 2> $O/bench_reference_cpu.err
ls -la $O
