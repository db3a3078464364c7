echo "=== attn tests"; timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -3
echo "=== attn bench"; timeout 120 python tools/bench_attention.py 2>&1 | tail -6
echo "=== full unet"; timeout 200 python tools/diag_models.py full_unet 2>&1 | tail -4
