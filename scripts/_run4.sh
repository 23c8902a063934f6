timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/gpu_c5.py --rows 2000000 > gpurun_out/c5_2gpu.json 2> gpurun_out/c5_2gpu.err
tail -3 gpurun_out/c5_2gpu.json; tail -5 gpurun_out/c5_2gpu.err
