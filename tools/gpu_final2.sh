#!/bin/bash
TAG=${1:-r2ae}
mkdir -p gpurun_out
for t in memcheck synccheck racecheck; do echo "-- $t"; bash tools/sanitize.sh $t; done 2>&1 | tee gpurun_out/${TAG}_sanitize.txt | tail -12
timeout 900 python tools/measure_traffic.py > gpurun_out/${TAG}_traffic.log 2>&1; tail -2 gpurun_out/${TAG}_traffic.log | cut -c1-160
cp gpurun_out/traffic.json profiles/traffic.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1200 gpurun_out/${TAG}_bench.err
