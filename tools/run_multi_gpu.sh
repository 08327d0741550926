set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -s 2>&1 | tail -8 > gpurun_out/r02_dp_tests.log; cat gpurun_out/r02_dp_tests.log
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err
  tail -c 400 gpurun_out/r02_bench_${n}gpu.json; echo
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload config3 --steps 30 --warmup 5 > gpurun_out/r02_bench_config3_8gpu.json 2> gpurun_out/r02_bench_config3_8gpu.err
tail -c 600 gpurun_out/r02_bench_config3_8gpu.json; echo
timeout 200 python bench.py --steps 50 --warmup 5 --no_cpu_baseline --no_ref_cuda > gpurun_out/r02_bench_1gpu_same_box.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_1gpu_same_box.json
