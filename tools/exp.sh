run() { python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline "$@" > /tmp/b.out 2> /tmp/b.err; if [ -s /tmp/b.out ]; then tail -1 /tmp/b.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['kernel'], r['launches_per_step'], round(r['ms_per_launch']*1000,2),'us/launch')"; else tail -6 /tmp/b.err; fi; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default cfg2"; run
echo "== default cfg3"; run --workload cfg3
echo "== default cfg4a"; run --workload cfg4a
echo "== default cfg5"; run --workload cfg5
