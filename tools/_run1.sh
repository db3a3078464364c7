echo "=== gn/ln tests"; timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "gn or ln or group_norm" 2>&1 | tail -4
echo "=== models"; timeout 300 python -m pytest tests/test_models_gpu.py -x -q 2>&1 | tail -3
echo "=== full unet"; timeout 200 python tools/diag_models.py full_unet full_vae 2>&1 | tail -7
