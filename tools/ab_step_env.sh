# whole-step A/B of an environment knob: usage tools/ab_step_env.sh VAR v1 v2 ...   (alternating twice)
var=$1; shift
for rep in 1 2; do for v in "$@"; do echo -n "$var=$v: "; env $var=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vae --no-report-dedup 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', d['clocks']['sm_mhz'], 'MHz', d['roofline']['families_ms_per_step']['igemm'], 'igemm ms')"; done; done
