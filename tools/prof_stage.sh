#!/bin/bash
# per-kernel time of one analyze_cu pipeline pass (serialised under ncu); usage: tools/prof_stage.sh [tag]
tag=${1:-pipe}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/prof_analyze.py 2 > gpurun_out/${tag}.log 2>&1
tail -1 gpurun_out/${tag}.log
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv
