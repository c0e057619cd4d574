# ns = 10001 (BASELINE config 5 density): K1 variants
run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline --lanes 1 --ns 10001 --candidates 65536 2>/dev/null > /tmp/o.json; python -c "import json; d=json.load(open('/tmp/o.json')); print('$1', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])"; }
LTK_K1=old run "old K1 pair (two-pass)"
LTK_K1_G=2 LTK_K1_THREADS=128 run "K1b G=2 T=128"
LTK_K1_G=1 LTK_K1_THREADS=256 run "K1b G=1 T=256"
LTK_K1_G=1 LTK_K1_THREADS=128 run "K1b G=1 T=128"
