# A/B of whole-step switches (each line: one bench.py run, device-resident ms/step)
run() { echo -n "$1: "; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vae --no-report-dedup $EXTRA 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step', d['clocks']['sm_mhz'], 'MHz', d['clocks']['power_w_max'], 'W')"; }
run baseline X=1
run gn_stats_fused MFB_FUSE_GN_STATS=1
run pdl MFB_PDL=1
EXTRA=--two-streams run two_streams X=1
EXTRA= run baseline_again X=1
