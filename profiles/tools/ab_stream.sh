# in-kernel tape loop of k_fused_tma (StepIn::stream_steps) against one launch per step, same drsim_run_tape call
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q -m gpu -k "tape_stream or run_equals" 2>&1 | tail -8
for wl in "$@"; do
echo "--- $wl stream"; bash profiles/tools/quick_bench.sh $wl
echo "--- $wl per-step launches"; DRSIM_NO_STREAM=1 bash profiles/tools/quick_bench.sh $wl
done
