# K1 configuration sweep (G candidates per CTA, T threads of K1b); prints step and per-kernel times
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/o.json; python -c "import json; d=json.load(open('/tmp/o.json')); print('$1', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])"; }
run "K1 default"
LTK_K1_THREADS=256 run "K1b T=256 (G=4)"
LTK_K1_G=8 LTK_K1_THREADS=256 run "K1b G=8 T=256"
LTK_K1=old run "K1 old pair"
