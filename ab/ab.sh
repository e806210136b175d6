python tools/gpu_check.py 1 64 200 8 | tail -2
for lib in ab/libagar_base.so a.i.gar_b200/libagar_b200.so; do for th in 64 128; do
 echo -n "$lib threads=$th W=8: "; AGAR_B200_LIB=$PWD/$lib AGAR_SIMPLE_THREADS=$th python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'])"
done; done
for lib in ab/libagar_base.so a.i.gar_b200/libagar_b200.so; do
 echo -n "$lib 262144 envs W=1: "; AGAR_B200_LIB=$PWD/$lib python bench.py --envs 262144 --frames 200 --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tile_width'])"
 echo -n "$lib 16384 envs: "; AGAR_B200_LIB=$PWD/$lib python bench.py --envs 16384 --frames 400 --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['tile_width'])"
done
