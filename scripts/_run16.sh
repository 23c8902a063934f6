timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err
tail -c 1500 gpurun_out/bench_r2_8gpu.json; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_r2_8gpu.err | tail -n 3
