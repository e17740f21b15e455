set -x
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
python bench.py --impl reference > gpurun_out/bench_ref_final2.json 2> gpurun_out/bench_ref_final2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ncu_launches_final2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_final2.log 2>&1
python tools/profile_target.py kmatw_tc > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kmatw_tc_kernel -c 1 -s 1 -o gpurun_out/kmatw_tc_final2 python tools/profile_target.py kmatw_tc > gpurun_out/ncu_kmatw_tc_final2.log 2>&1
python tools/configs_full.py all > gpurun_out/configs_full_final2.jsonl 2>&1
tail -c 300 gpurun_out/bench_final2.json; tail -3 gpurun_out/configs_full_final2.jsonl
