# round-2 GPU job 55: the bench on 8 GPUs with the final tree (own arm + reference arm)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/j55_bench8.log 2> gpurun_out/j55_bench8.err
echo "bench8 rc=$?"; tail -c 300 gpurun_out/j55_bench8.err; cut -c1-200 gpurun_out/j55_bench8.log | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/j55_ref8.log 2> gpurun_out/j55_ref8.err
echo "ref8 rc=$?"; cut -c1-250 gpurun_out/j55_ref8.log | tail -1
