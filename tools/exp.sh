run() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e "$@" > /tmp/b.out 2> /tmp/b.err; if [ -s /tmp/b.out ]; then tail -1 /tmp/b.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['kernel'])"; else tail -8 /tmp/b.err; fi; }
echo "== cfg5x"; run --workload cfg5x
echo "== cfg5x copy"; run --workload cfg5x --mode copy
echo "== cfg5 copy"; run --workload cfg5 --mode copy
echo "== cfg5"; run --workload cfg5
