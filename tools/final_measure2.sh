set -x
O=gpurun_out/final2
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
python bench.py --steps 20 --warmup 5 > $O/n1.json 2> $O/n1.err
for wl in mimc_2p14 aggregation_16; do python bench.py --workload $wl --steps 200 --warmup 20 --inflight 1 > $O/$wl.json 2>/dev/null; done
python bench.py --workload training_8192 --steps 10 --warmup 3 --inflight 8 --no-cpu-baseline > $O/training_8192_x8.json 2>/dev/null
python tools/prove_once.py training_2p16 2 > $O/prove_once.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python tools/prove_once.py training_2p16 2 > $O/ncu1.log 2>&1
python - <<'PY'
import json
for f in ['n1','mimc_2p14','aggregation_16','training_8192_x8']:
    try:
        d=json.load(open(f'gpurun_out/final2/{f}.json')); print(f, d['value'], d['ms_per_step'], d.get('prove_ms'), d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
