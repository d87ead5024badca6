# round-2 GPU job 19: the bench as the driver launches it on 2 GPUs (library communicator, pipelined steps) + reference arm
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/j19_bench2.log 2> gpurun_out/j19_bench2.err
echo "bench2 rc=$?"; tail -c 1500 gpurun_out/j19_bench2.err; cut -c1-1500 gpurun_out/j19_bench2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/j19_ref2.log 2> gpurun_out/j19_ref2.err
echo "ref2 rc=$?"; cut -c1-600 gpurun_out/j19_ref2.log
