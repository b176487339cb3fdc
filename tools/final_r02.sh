set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/final_tests.log
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo bench_rc=$?
python bench.py --impl reference > gpurun_out/bench_r02_final_reference.json 2>/dev/null; echo ref_rc=$?
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:resize|fast_|select_kernel|blur_|describe|match_kernel|filter_kernel|five_point|score_kernel|pose_|finalize|repack|clear_detect|adaptive|bgr2gray|popc_rate' -c 3000 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launches_ncu.log 2>&1; echo ncu_rc=$?
python tools/summarize_launches.py gpurun_out/launches_r02_final.csv > gpurun_out/launches_r02_final_summary.md 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
