# End-of-session evidence capture on one B200 (run under gpurun): bench lines, reference arm, the ncu launch list of the
# bench command, ncu --set full of the two dominant kernels (after the un-profiled runs exited 0), full-size configs.
set -x
python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err || exit 1
python bench.py --impl reference > gpurun_out/bench_ref_final4.json 2> gpurun_out/bench_ref_final4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ncu_launches_final4.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_final4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tf32_kernel -c 1 -s 1 -o gpurun_out/gram_cfg2_final4 python bench.py --steps 1 --warmup 1 --no-also > gpurun_out/ncu_gram_final4.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kmatw_tc_kernel -c 1 -s 1 --csv --log-file gpurun_out/ncu_kmatw_traffic_final4.csv python bench.py --workload kmatw --steps 1 --warmup 1 > gpurun_out/ncu_kmatw_traffic_final4.log 2>&1
python tools/profile_target.py kmatw_tc > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kmatw_tc_kernel -c 1 -s 1 -o gpurun_out/kmatw_tc_final4 python tools/profile_target.py kmatw_tc > gpurun_out/ncu_kmatw_tc_final4.log 2>&1
python tools/configs_full.py all > gpurun_out/configs_full_final4.jsonl 2>&1
python tools/quickbench.py all > gpurun_out/quickbench_final4.jsonl 2>&1
tail -c 300 gpurun_out/bench_final4.json; tail -3 gpurun_out/configs_full_final4.jsonl
