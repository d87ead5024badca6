# round-2 GPU job 63: the bench on 4 GPUs (the one N of the driver's scaling run not yet exercised)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/j63_bench4.log 2> gpurun_out/j63_bench4.err
echo "bench4 rc=$?"; tail -c 300 gpurun_out/j63_bench4.err; cut -c1-250 gpurun_out/j63_bench4.log | tail -1
