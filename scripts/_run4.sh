for ex in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/gpu_c5.py --rows 2000000 --exchange $ex > gpurun_out/c5_2gpu_$ex.json 2> gpurun_out/c5_2gpu_$ex.err
tail -n 2 gpurun_out/c5_2gpu_$ex.json; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/c5_2gpu_$ex.err | tail -n 5
done
