#!/bin/bash
# GPU call with a gate: a 2-minute fixture run of the picture-level pass first; the longer steps only run if it passes.
# usage: tools/gpu_gate.sh TAG "bench args"
tag=${1:-gate}; shift
mkdir -p gpurun_out
timeout 150 python tests/picture_on_device.py --fixture-only > gpurun_out/${tag}_fixture.txt 2>&1
rc=$?
echo "exit $rc" >> gpurun_out/${tag}_fixture.txt
tail -n 3 gpurun_out/${tag}_fixture.txt | cut -c1-300
if [ $rc -ne 0 ]; then echo "GATE FAILED"; exit 0; fi
timeout 420 python -m pytest tests/test_dropin.py tests/test_gpu_picture.py -x -q -m gpu > gpurun_out/${tag}_tests.txt 2>&1
echo "exit $?" >> gpurun_out/${tag}_tests.txt
tail -n 5 gpurun_out/${tag}_tests.txt | cut -c1-300
timeout 600 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "exit $?" >> gpurun_out/${tag}_bench.err
tail -n 3 gpurun_out/${tag}_bench.err | cut -c1-400
cut -c1-4000 gpurun_out/${tag}_bench.json
