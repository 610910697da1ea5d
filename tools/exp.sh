run() { python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], round(d['roofline']['frac'],4))"; }
for lib in lib_u2_c4 lib_u2_c3 lib_u4_c3 lib_u4_c2; do echo "== $lib"; CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/$lib.so run; done
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/lib_u2_c4.so python -m pytest tests -m gpu -x -q 2>&1 | tail -2
