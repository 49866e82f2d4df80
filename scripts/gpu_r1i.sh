# call I (2 GPUs): train + fwd bench under torchrun, world size 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/i_smi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/i_bench_train_n2.json 2> gpurun_out/i_bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --mode fwd > gpurun_out/i_bench_fwd_n2.json 2> gpurun_out/i_bench_fwd_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/i_bench_ref_n2.json 2> gpurun_out/i_bench_ref_n2.err
tail -c 1500 gpurun_out/i_bench_train_n2.json; grep -E "^\[bench\]|Error|error" gpurun_out/i_bench_train_n2.err | tail -5
