set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q -m gpu -k "actor" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python profiles/tools/time_actor.py 2>&1 | tee gpurun_out/r2b_actor_times.txt
timeout 300 python profiles/tools/actor_phases.py c4 > gpurun_out/r2b_actor_phases_c4.txt 2>&1
timeout 300 python profiles/tools/actor_phases.py c3 > gpurun_out/r2b_actor_phases_c3.txt 2>&1
