# A/B of fps_grid_kernel builds on one box: bash tools/fps_ab.sh "" _w32 _w8 ...   (suffixes of ppt_b200/libppt_b200*.so)
for v in "$@"; do for k in U S; do echo -n "variant [$v] "; PPT_B200_LIB=$PWD/ppt_b200/libppt_b200$v.so python tools/time_fps.py $k 2>&1 | tail -1; done; done
