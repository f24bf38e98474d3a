# usage: bash profiles/tools/bench_multigpu.sh N  -- the driver's launch line for N GPUs, default workload (+ sharded_cluster)
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@" 2>gpurun_out/bench_${N}gpu.err | tail -1
