# run on the GPU box (1 GPU): the round-2 evidence set -> gpurun_out/
set -x
# the DRAM-traffic capture first: it stamps profiles/r2_traffic.json with the source hash the bench lines then report against
bash profiles/tools/capture_traffic.sh > gpurun_out/capture_traffic.log 2>&1
python bench.py > gpurun_out/r2_bench_c4_1gpu.json 2> gpurun_out/r2_bench_c4_1gpu.err
# the driver's own command line (20 steps: ONE launch of the in-kernel step loop)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_1gpu_steps20.json 2> /dev/null
for wl in c3 c5 c2 c1; do
  python bench.py --workload $wl --no-c5 > gpurun_out/r2_bench_${wl}_1gpu.json 2> gpurun_out/r2_bench_${wl}_1gpu.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
python profiles/tools/time_dropin.py 10 100 1000 > gpurun_out/r2_dropin_times.txt 2>&1
python profiles/tools/time_actor.py > gpurun_out/r2_actor_times.txt 2>&1
python profiles/tools/actor_phases.py c4 > gpurun_out/r2_actor_phases_c4.txt 2>&1
python profiles/tools/actor_phases.py c3 > gpurun_out/r2_actor_phases_c3.txt 2>&1
python profiles/tools/shard_phases.py 1000000 1 > gpurun_out/r2_shard_phases_1gpu.txt 2>&1
# one full ncu capture of each dominant kernel (source-level), after the plain runs above exited 0
ncu --set full --clock-control none --import-source on -k regex:k_shard -s 30 -c 1 -o gpurun_out/r2_kshard_c5 -f python bench.py --workload c5 --steps 60 --no-cpu > /dev/null 2>&1
# (launch 11 of `--steps 20` is the timed launch: 20 steps inside one kernel)
ncu --set full --clock-control none --import-source on -k regex:k_fused_tma -s 11 -c 1 -o gpurun_out/r2_kfused_c4 -f python bench.py --workload c4 --steps 20 --no-cpu --no-c5 --no-rollout > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_actor3x -s 2 -c 1 -o gpurun_out/r2_actor3x_c4 -f python profiles/tools/actor_once.py c4 > /dev/null 2>&1
for r in r2_kshard_c5 r2_kfused_c4 r2_actor3x_c4; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; python profiles/tools/ncu_top.py gpurun_out/$r.ncu-rep 25 > gpurun_out/$r.top.txt 2>&1; done
python profiles/tools/sass_hist.py > gpurun_out/r2_sass_histograms.md 2>&1
ls -la gpurun_out | tail -40
