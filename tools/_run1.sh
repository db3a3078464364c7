mkdir -p gpurun_out
echo "=== sched test"; timeout 100 python -m pytest tests/test_pipelines_gpu.py -x -q -k scheduler 2>&1 | tail -2
echo "=== breakdown"; timeout 300 python tools/pipeline_breakdown.py 2>&1 | tail -6
echo "=== bench"; timeout 600 python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; tail -c 3000 gpurun_out/bench_r1e.json; tail -3 gpurun_out/bench_r1e.err
