mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "=== 2-GPU bench"; timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_r1f.json 2> gpurun_out/bench_2gpu_r1f.err; cut -c1-900 gpurun_out/bench_2gpu_r1f.json; tail -3 gpurun_out/bench_2gpu_r1f.err
echo "=== 2-GPU reference arm"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref_r1f.json 2> gpurun_out/bench_2gpu_ref_r1f.err; cut -c1-400 gpurun_out/bench_2gpu_ref_r1f.json; tail -2 gpurun_out/bench_2gpu_ref_r1f.err
