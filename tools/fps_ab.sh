for v in trace tracebase; do echo "== $v"; PPT_B200_LIB=$PWD/ppt_b200/libppt_b200_$v.so python tools/fps_trace.py 2>&1 | tail -13; done
for k in U S; do for v in "" _base _rot _rotmicro _norot; do echo -n "variant [$v] "; PPT_B200_LIB=$PWD/ppt_b200/libppt_b200$v.so python tools/time_fps.py $k 2>&1 | tail -1; done; done
timeout 300 python -m pytest tests/test_gpu_geometry.py -q -m gpu -x -k "fps or group_pipeline" 2>&1 | tail -3
