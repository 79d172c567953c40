#!/bin/bash
# GPU call with a gate: a short fixture run of the picture-level pass first; the longer steps only run if it passes.
# usage: tools/gpu_gate.sh TAG "bench args"
tag=${1:-gate}; shift
mkdir -p gpurun_out
XB200_SCHED_DEBUG=1 timeout 100 python -u tests/picture_on_device.py --fixture-only > gpurun_out/${tag}_fixture.txt 2>&1
rc=$?
echo "exit $rc" >> gpurun_out/${tag}_fixture.txt
tail -n 12 gpurun_out/${tag}_fixture.txt | cut -c1-300
if [ $rc -ne 0 ]; then
    echo "GATE FAILED; retry with 4 workers"
    XB200_CHAIN_WORKERS=4 XB200_SCHED_DEBUG=1 timeout 100 python -u tests/picture_on_device.py --fixture-only > gpurun_out/${tag}_fixture_w4.txt 2>&1
    echo "exit $?" >> gpurun_out/${tag}_fixture_w4.txt
    tail -n 12 gpurun_out/${tag}_fixture_w4.txt | cut -c1-300
    nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
    exit 0
fi
# second gate: three encoder instances in one process share the device (the reference's API, QCIF-size clip)
python - <<'PY'
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import tracedata
QCIF = {k: v for k, v in tracedata.QCIF.items() if k != "n"}
c, yuv = tracedata.clip_yuv("cif", 20, **QCIF)
yuv.tofile("/dev/shm/gate_qcif.yuv")
PY
XB200_SCHED_DEBUG=1 timeout 120 oracle/_ref/xb200_streams -i /dev/shm/gate_qcif.yuv -w 176 -h 144 -z 20 -n 3 -m 2 -o /dev/shm/gate_s > gpurun_out/${tag}_streams.txt 2> gpurun_out/${tag}_streams.err
rc=$?
echo "exit $rc" >> gpurun_out/${tag}_streams.txt
tail -n 2 gpurun_out/${tag}_streams.txt | cut -c1-600
if [ $rc -ne 0 ]; then echo "GATE 2 FAILED"; tail -n 15 gpurun_out/${tag}_streams.err | cut -c1-300; exit 0; fi
timeout 420 python -m pytest tests/test_dropin.py tests/test_gpu_picture.py -x -q -m gpu > gpurun_out/${tag}_tests.txt 2>&1
echo "exit $?" >> gpurun_out/${tag}_tests.txt
tail -n 5 gpurun_out/${tag}_tests.txt | cut -c1-300
timeout 600 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "exit $?" >> gpurun_out/${tag}_bench.err
tail -n 3 gpurun_out/${tag}_bench.err | cut -c1-400
cut -c1-4000 gpurun_out/${tag}_bench.json
