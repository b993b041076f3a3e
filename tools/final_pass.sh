timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_final.log 2>&1; tail -3 gpurun_out/pytest_r02_final.log
timeout 300 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 600 gpurun_out/bench_r02_final.json
NB200_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-breakdown --no-cpu-baseline --no-verify > /dev/null 2> gpurun_out/prove_trace_r02_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-breakdown --no-cpu-baseline --no-verify > gpurun_out/b_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_ -c 6 -f -o gpurun_out/ncu_fft_r02f python tools/fft_probe.py 20 128 > gpurun_out/ncu_fft_r02f.log 2>&1
ls -la gpurun_out/ | tail -8
