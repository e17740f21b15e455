# Evidence capture on one B200 (run under gpurun): bench lines, reference arm, the ncu launch list of the bench command,
# ncu --set full of the dominant kernels (after the un-profiled runs exited 0), DRAM / L2 traffic of the K@W kernel at the
# full cfg3 size, full-size configs.  Usage: bash tools/final_capture.sh <tag>
TAG=${1:-r02}
set -x
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ncu_launches_$TAG.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/profile_target.py kmatw_tc > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kmatw_tc2_kernel -c 1 -s 1 -o gpurun_out/kmatw_tc2_$TAG python tools/profile_target.py kmatw_tc > gpurun_out/ncu_kmatw_tc2_$TAG.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:kmatw_tc2_kernel -c 1 -s 1 --csv --log-file gpurun_out/ncu_kmatw_tc2_traffic_$TAG.csv python bench.py --steps 1 --warmup 1 --no-also > gpurun_out/ncu_kmatw_tc2_traffic_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tf32_kernel -c 1 -s 1 -o gpurun_out/gram_cfg2_$TAG python bench.py --workload gram --steps 1 --warmup 1 --no-also > gpurun_out/ncu_gram_$TAG.log 2>&1
python tools/configs_full.py all > gpurun_out/configs_full_$TAG.jsonl 2>&1
python tools/quickbench.py all > gpurun_out/quickbench_$TAG.jsonl 2>&1
python tools/profile_target.py kgrad_tc > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kgrad_tc_kernel -c 1 -s 1 -o gpurun_out/kgrad_tc_$TAG python tools/profile_target.py kgrad_tc > gpurun_out/ncu_kgrad_tc_$TAG.log 2>&1
tail -c 300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/configs_full_$TAG.jsonl
