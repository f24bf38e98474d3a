set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_multigpu.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5 --steps 2000 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2_c5_${N}gpu.json
python -c "
import json,sys; d=json.load(open('gpurun_out/r2_c5_${N}gpu.json')); print('C5 x$N', d['ms_per_step']*1e3, 'us/step', d['value'], d['e2e']['ms_per_step'])"
DRSIM_NO_SHARD_KERNEL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --steps 2000 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5 x$N four-kernel', d['ms_per_step']*1e3, 'us/step')"
