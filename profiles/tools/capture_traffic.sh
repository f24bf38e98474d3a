# run on the GPU box: steady-state DRAM traffic of the step kernels of c4 / c3 / c5 (see ncu_traffic.py) + launch lists
set -x
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for wl in c4 c3; do
  # in-kernel step loop: launches 13.. of `--steps 640` are full 64-step blocks (10 warm-up steps, an 8-step and a 46-step
  # launch come first)
  ncu --metrics $M --cache-control none --clock-control none -k regex:'k_fused' -s 13 -c 4 --csv \
      --log-file gpurun_out/traffic_$wl.csv python bench.py --workload $wl --steps 640 --no-cpu --no-c5 --no-rollout \
      > gpurun_out/traffic_$wl.log 2>&1
  # one launch per step
  DRSIM_NO_STREAM=1 ncu --metrics $M --cache-control none --clock-control none -k regex:'k_fused' -s 40 -c 6 --csv \
      --log-file gpurun_out/traffic_${wl}_launch.csv python bench.py --workload $wl --steps 60 --no-cpu --no-c5 --no-rollout \
      > gpurun_out/traffic_${wl}_launch.log 2>&1
done
ncu --metrics $M --cache-control none --clock-control none -k regex:'k_shard' -s 40 -c 6 --csv --log-file gpurun_out/traffic_c5.csv \
    python bench.py --workload c5 --steps 60 --no-cpu --no-c5 --no-rollout > gpurun_out/traffic_c5.log 2>&1
python profiles/tools/ncu_traffic.py c4=gpurun_out/traffic_c4.csv:64 c3=gpurun_out/traffic_c3.csv:64 \
    c4_per_step_launch=gpurun_out/traffic_c4_launch.csv c3_per_step_launch=gpurun_out/traffic_c3_launch.csv c5=gpurun_out/traffic_c5.csv
cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
# launch list of the default bench command (kernel shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 200 --warmup 3 --no-cpu > gpurun_out/r2_launches.log 2>&1
