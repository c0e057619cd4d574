# K1b configuration sweep (G candidates per CTA, T threads); prints step and per-kernel times (one population at a time)
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --lanes 1 2>/dev/null > /tmp/o.json; python -c "import json; d=json.load(open('/tmp/o.json')); print('$1', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])"; }
LTK_K1_G=4 LTK_K1_THREADS=256 run "K1b G=4 T=256"
LTK_K1_G=2 LTK_K1_THREADS=128 run "K1b G=2 T=128"
LTK_K1_G=2 LTK_K1_THREADS=64 run "K1b G=2 T=64"
