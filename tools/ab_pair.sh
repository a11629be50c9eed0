# A/B of the stage-2 variants on one B200 (PPT_STAGE2_PAIR=1 pair kernel, 0 single-CTA kernel); optional ncu capture.
if [ "$1" = "test" ]; then timeout 400 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu 2>&1 | tail -3; fi
for p in 1 0; do PPT_STAGE2_PAIR=$p timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print(round(d['value']), d['roofline']['phase_ms'])
except Exception as e: print('ERR', l[:500])
"; done
if [ "$2" = "ncu" ]; then
timeout 500 ncu --set full --clock-control none --import-source on -k regex:stage2_pair -s 3 -c 1 -o gpurun_out/pair python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/pair.log 2>&1
fi
