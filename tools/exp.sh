run() { python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], round(d['roofline']['frac'],4), d['roofline']['kernel'])"; }
for lib in lib_u3_c3 lib_u6_c3 lib_u8_c2 lib_u4_c4 lib_u6_c2; do echo "== $lib"; CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/$lib.so run; done
echo "== default cfg2"; run
echo "== default cfg4a"; run --workload cfg4a
echo "== default cfg5 (1 GPU)"; run --workload cfg5
for iv in 1024 4096; do echo "== default cfg2 item_vecs=$iv"; CMGPU_ITEM_VECS=$iv run; done
