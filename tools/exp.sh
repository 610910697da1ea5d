run() { python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline "$@" > /tmp/b.out 2> /tmp/b.err; if [ -s /tmp/b.out ]; then tail -1 /tmp/b.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['kernel'], r['launches_per_step'], round(r['ms_per_launch']*1000,2),'us/launch')"; else tail -6 /tmp/b.err; fi; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default cfg2"; run
for lib in libw_u4_c2 libw_u2_c3 libw_u2_c2 libw_u3_c2; do echo "== cfg4a $lib"; CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/$lib.so run --workload cfg4a; done
echo "== default cfg3 (graph)"; run --workload cfg3
echo "== default cfg5"; run --workload cfg5
