mkdir -p gpurun_out/n8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 6 > gpurun_out/n8/final_n8.json 2> gpurun_out/n8/final_n8.err
tail -n 1 gpurun_out/n8/final_n8.json | grep -o "ms_per_step\": [0-9.]*\|\"value\": [0-9.]*\|flag_wait_errors\": [0-9]*\|layers_per_stage\": \[[0-9, ]*\]"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 6 > gpurun_out/n8/final_n4.json 2> gpurun_out/n8/final_n4.err
tail -n 1 gpurun_out/n8/final_n4.json | grep -o "ms_per_step\": [0-9.]*\|\"value\": [0-9.]*\|flag_wait_errors\": [0-9]*"
