timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -x -q 2>&1 | tail -n 2
echo "== 1gpu"; timeout 300 python bench.py --steps 20 --warmup 6 2>/dev/null | tail -n 1 | grep -o "ms_per_step\": [0-9.]*\|final_loss\": [0-9.]*"
