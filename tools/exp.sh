run() { python bench.py --steps 50 --warmup 3 --no-cpu-baseline "$@" > /tmp/b.out 2> /tmp/b.err; if [ -s /tmp/b.out ]; then tail -1 /tmp/b.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; e=d['e2e'] or {}; print(round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['kernel'], r['launches_per_step'], round(r['ms_per_launch']*1000,2),'us/launch', '| e2e', round(e.get('value',0)), e.get('parity_spot_check'))"; else tail -8 /tmp/b.err; fi; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== cfg4b"; run --workload cfg4b
echo "== cfg4a"; run --workload cfg4a
