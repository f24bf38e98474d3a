# run on the GPU box: steady-state DRAM traffic of the step kernels of c4 / c3 / c5 (see ncu_traffic.py) + launch lists
set -x
for wl in c4 c3 c5; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
      -k regex:'k_fused|k_shard' -s 40 -c 6 --csv --log-file gpurun_out/traffic_$wl.csv \
      python bench.py --workload $wl --steps 60 --no-cpu --no-c5 --no-rollout > gpurun_out/traffic_$wl.log 2>&1
done
python profiles/tools/ncu_traffic.py c4=gpurun_out/traffic_c4.csv c3=gpurun_out/traffic_c3.csv c5=gpurun_out/traffic_c5.csv
cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
# launch list of the default bench command (kernel shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/r2_launches.log 2>&1
