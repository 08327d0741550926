# round-2 evidence run (1 GPU): bench lines, launch list, ncu --set full of the dominant kernels.  Outputs under gpurun_out/.
cd $GRAFT_REPO_ROOT
o=gpurun_out
python bench.py > $o/r02_bench_default.json 2> $o/r02_bench_default.err
python bench.py --impl reference --steps 10 --warmup 3 > $o/r02_bench_reference.json 2>> $o/r02_bench_default.err
python bench.py --pool 2 --no_cpu_baseline --no_ref_cuda > $o/r02_bench_pool2.json 2>/dev/null
python bench.py --fused --no_cpu_baseline --no_ref_cuda > $o/r02_bench_fused.json 2>/dev/null
python bench.py --temp 0.05 --no_cpu_baseline --no_ref_cuda > $o/r02_bench_temp005.json 2>/dev/null
python bench.py --workload config3 --no_cpu_baseline > $o/r02_bench_config3_1gpu.json 2>/dev/null
python bench.py --workload config4 > $o/r02_bench_config4.json 2> $o/r02_bench_config4.err
python bench.py --workload config5 > $o/r02_bench_config5.json 2> $o/r02_bench_config5.err
python tools/trace_step.py > $o/r02_trace_step.log 2>&1
python tools/trace_step.py --fused >> $o/r02_trace_step.log 2>&1
python tools/gemm_timeline.py > $o/r02_gemm_timeline.log 2>&1
python tools/fused_timeline.py > $o/r02_fused_timeline.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02_launches.csv python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_ref_cuda > $o/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"sampler_kernel|sampler_bwd_saved|gemm_tf32|adamax_partials|recon" --launch-skip 30 -c 10 -o $o/r02_step_kernels python tools/trace_step.py --steps 3 > $o/ncu_full.log 2>&1
tail -c 300 $o/r02_bench_default.json
