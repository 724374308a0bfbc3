timeout 900 python -m pytest tests -m gpu -q -k "fir or smoke or blocks" 2>&1 | tail -3
for c in c1 c3; do timeout 300 python bench.py --config $c --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c95-125; done
