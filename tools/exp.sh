run() { python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e "$@" > /tmp/b.out 2> /tmp/b.err; if [ -s /tmp/b.out ]; then tail -1 /tmp/b.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['kernel'])"; else tail -8 /tmp/b.err; fi; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg2 cfg3 cfg4a cfg5 cfg6ch; do echo "== $w"; run --workload $w; done
