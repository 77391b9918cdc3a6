timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -x -q 2>&1 | tail -n 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_pipeline.py --boundary fused 2>&1 | grep CHECK
echo "== 2gpu"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 6 2>/dev/null | tail -n 1 | grep -o "ms_per_step\": [0-9.]*\|flag_wait_errors\": [0-9]*\|final_loss\": [0-9.]*"
echo "== 1gpu"; timeout 300 python bench.py --steps 20 --warmup 6 2>/dev/null | tail -n 1 | grep -o "ms_per_step\": [0-9.]*\|final_loss\": [0-9.]*"
