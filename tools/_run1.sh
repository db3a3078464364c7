mkdir -p gpurun_out
echo "=== kernels"; timeout 240 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -6
echo "=== chunk rate"; timeout 120 python tools/gemm_chunk_rate.py 2>&1 | head -14
echo "=== timeline pf pair"; MODE=pf timeout 60 python tools/gemm_timeline.py 2048 1280 1280 2>&1 | tail -14
echo "=== full unet"; timeout 200 python tools/diag_models.py full_unet full_vae 2>&1 | tail -8
echo "=== full unet nopair"; B200_NO_PAIR=1 timeout 200 python tools/diag_models.py full_unet 2>&1 | tail -4
