run() { python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e "$@" > /tmp/b.out 2> /tmp/b.err; if [ -s /tmp/b.out ]; then tail -1 /tmp/b.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['kernel'])"; else tail -8 /tmp/b.err; fi; }
echo "== cfg3 default (64 regs, spills)"; run --workload cfg3
echo "== cfg3 g8c3"; CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/lib_g8c3.so run --workload cfg3
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
