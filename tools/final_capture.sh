# Round-end evidence on one B200: [test] full GPU test suite, bench line, reference arm, clocks, ncu launch list,
# ncu --set full of every kernel of the step, FPS phase trace, overlap probe.   bash tools/final_capture.sh [test]
mkdir -p gpurun_out
if [ "$1" = "test" ]; then timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt; fi
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
timeout 300 python tools/kernel_clocks.py 2>/dev/null | tail -1 > gpurun_out/kernel_clocks.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-widened > gpurun_out/ncu_launch.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'knn_prepare|fps_grid|knn_search|encoder_|group_linear' -s 21 -c 7 \
    -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-widened > gpurun_out/ncu_full.log 2>&1
if [ -f ppt_b200/libppt_b200_trace.so ]; then PPT_B200_LIB=$PWD/ppt_b200/libppt_b200_trace.so timeout 120 python tools/fps_trace.py > gpurun_out/fps_trace.txt 2>&1; fi
timeout 120 python tools/overlap_probe.py > gpurun_out/overlap_probe.txt 2>&1
python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['clocks'], d['roofline']['frac'])"
