#!/bin/bash
# Full GPU suite + every bench configuration at HEAD (one GPU).  Usage: gpurun --timeout 1500 -- 'bash tools/final_run.sh TAG'
TAG=${1:-final}
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 200 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
for c in 1 0; do timeout 300 python bench.py --config $c > $O/${TAG}_cfg$c.json 2> $O/${TAG}_cfg$c.err; tail -c 200 $O/${TAG}_cfg$c.err; done
timeout 400 python bench.py --config 2 --steps 2 --warmup 1 > $O/${TAG}_cfg2.json 2> $O/${TAG}_cfg2.err; tail -c 200 $O/${TAG}_cfg2.err
timeout 300 python bench.py --config 4 --steps 5 --warmup 2 > $O/${TAG}_cfg4.json 2> $O/${TAG}_cfg4.err; tail -c 200 $O/${TAG}_cfg4.err
python - <<EOF2
import json
for f in ("bench","bench_ref","cfg1","cfg0","cfg2","cfg4"):
    try:
        d=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, d.get("value"), d.get("unit"), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"),
              "parity", (d.get("parity") or {}).get("ok"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
EOF2
