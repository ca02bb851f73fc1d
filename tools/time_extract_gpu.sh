#!/bin/bash
# run on the GPU box: extractor parity tests, extractor-only timing, ncu launch list -> gpurun_out/$1.csv
tag=${1:-r2_launches_tile}
timeout 400 python -m pytest tests/test_extractor_gpu.py tests/test_oracle_reference_extractor.py tests/test_stereo_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/extract_time.py 320 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$tag.csv python tools/extract_time.py 320 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/$tag.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
tot={}
n=0
for r in rows[1:]:
    if "b2s" in r[ki]:
        nm=r[ki].split("(")[0]
        tot.setdefault(nm,[]).append(float(r[vi]))
for k,v in tot.items():
    half=len(v)//2
    print("%-28s launches/run %d  ms/run %.3f" % (k, half, sum(v[half:])/1e6))
PY
