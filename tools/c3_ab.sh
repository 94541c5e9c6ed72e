#!/bin/bash
# A/B of the metadata path on BASELINE config 3 (many small files): reference's openat+statx SQE pair vs openat + fstat.
for sv in 1 0; do
  python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --statx-via-ring $sv 2>/dev/null > /tmp/c3_$sv.json
  python - "$sv" <<'PY'
import sys, json
d = json.load(open(f"/tmp/c3_{sys.argv[1]}.json"))
print("statx_via_ring", sys.argv[1], round(d["gib_s"], 2), "GiB/s", [round(x, 4) for x in d["secs_per_step"]], d["verify"])
PY
done
