#!/usr/bin/env bash
# GPU work queued at the end of round 1 (the budget ran out before it could run). Each block is one `gpurun` call;
# everything is logged IN FULL under gpurun_out/ (a `| tail` on the command line lost the evidence of the one DDP-graph
# trial of round 1) and every risky command has its own short timeout.
#
#   bash tools/next_round_gpu_checklist.sh 1     # run block 1, etc.
set -u
GPURUN=/usr/local/graft/bin/gpurun
case "${1:-}" in
1)  # antialiased bicubic kernel: first GPU run. Green -> drop the gate in embeddings._get_embedding and
    # the skipif in tests/test_gpu_preprocess_aa.py.
    $GPURUN --timeout 300 -- 'PVR_EXPERIMENTAL_AA=1 timeout 200 compute-sanitizer --error-exitcode 9 python tools/sanitize_aa.py > gpurun_out/aa_sanitizer.log 2>&1; PVR_EXPERIMENTAL_AA=1 timeout 200 python -m pytest tests/test_gpu_preprocess_aa.py -q -m gpu > gpurun_out/aa_tests.log 2>&1; tail -5 gpurun_out/aa_sanitizer.log gpurun_out/aa_tests.log'
    ;;
2)  # north-star argmax criterion on a trained policy (currently xfail(strict=False): XPASS -> remove the marker)
    $GPURUN --timeout 300 -- 'timeout 200 python -m pytest tests/test_gpu_policy.py -q -m gpu -k trained_policy -rxX > gpurun_out/argmax_trained.log 2>&1; tail -5 gpurun_out/argmax_trained.log'
    ;;
3)  # ncu captures that round 1 still owes (profiles/README.md): bicubic preprocessing, one BC step with the wavefront
    $GPURUN --timeout 600 -- 'ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -c 2 -o gpurun_out/prof_r02_preprocess_bicubic python bench.py --workload mae_base --no-bc --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_mae.log 2>&1; ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bc_r02.csv python tools/profile_bc.py 2 > gpurun_out/ncu_bc.log 2>&1; ls -la gpurun_out | tail -5'
    ;;
4)  # whole-step CUDA graph under data parallelism (hung in round 1). Hypothesis: the process group's async error
    # handling / watchdog. 2 GPUs, eager first (reference number), then the forced graph, 90 s cap, full log.
    $GPURUN --gpus 2 --timeout 200 -- 'TORCH_NCCL_ASYNC_ERROR_HANDLING=0 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/profile_bc_ddp.py 50 eager graph > gpurun_out/bc_ddp_graph_r02.log 2>&1; grep "^world" gpurun_out/bc_ddp_graph_r02.log; tail -3 gpurun_out/bc_ddp_graph_r02.log'
    ;;
*)  sed -n 2,9p "$0" ;;
esac
