timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/bench_r2_4gpu.json 2> gpurun_out/bench_r2_4gpu.err
tail -c 1500 gpurun_out/bench_r2_4gpu.json; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_r2_4gpu.err | tail -n 3
