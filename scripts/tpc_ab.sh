# A/B of the chunk-length cap (new default) against the uncapped planner's choice, one box
for pass in 1 2; do
  for T in 0 53; do
    echo "bench IRR_TILES_PER_CHUNK=$T"
    IRR_TILES_PER_CHUNK=$T python bench.py --steps 30 --warmup 5 --no-sweep 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','verified')}, d['roofline']['frac'], d['e2e']['value'])"
  done
done
for T in 0 64; do echo "midq IRR_TILES_PER_CHUNK=$T"; IRR_TILES_PER_CHUNK=$T python scripts/midq.py 3072 4096 8192 2>&1 | cut -c1-90; done
