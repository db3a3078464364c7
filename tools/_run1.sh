mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; cut -c1-700 gpurun_out/bench_r1f.json; tail -2 gpurun_out/bench_r1f.err
echo "=== bench reference"; timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_r1f.json 2> gpurun_out/bench_ref_r1f.err; cut -c1-900 gpurun_out/bench_ref_r1f.json; tail -2 gpurun_out/bench_ref_r1f.err
echo "=== ncu launch list"; timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python tools/profile_forward.py ncu > gpurun_out/ncu_r1f.log 2>&1; wc -l gpurun_out/launches_r1f.csv
