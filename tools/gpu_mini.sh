#!/bin/bash
# small, instrumented GPU call: fixture -> three streams in one process -> live pictures enqueued ahead -> a small bench; stops at the first failure
tag=${1:-mini}; shift
mkdir -p gpurun_out
step() { # name timeout cmd...
    local name=$1 to=$2; shift 2
    XB200_SCHED_DEBUG=${DBG:-0} timeout $to "$@" > gpurun_out/${tag}_${name}.txt 2> gpurun_out/${tag}_${name}.err
    local rc=$?
    echo "exit $rc" >> gpurun_out/${tag}_${name}.txt
    echo "== $name: exit $rc"; tail -n 4 gpurun_out/${tag}_${name}.txt | cut -c1-1500
    if [ $rc -ne 0 ]; then tail -n 12 gpurun_out/${tag}_${name}.err | cut -c1-300; echo "STOP at $name"; exit 0; fi
}
DBG=1 step fixture 100 python -u tests/picture_on_device.py --fixture-only
python - <<'PY'
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import tracedata
QCIF = {k: v for k, v in tracedata.QCIF.items() if k != "n"}
c, yuv = tracedata.clip_yuv("cif", 20, **QCIF)
yuv.tofile("/dev/shm/gate_qcif.yuv")
PY
DBG=1 step streams 120 oracle/_ref/xb200_streams -i /dev/shm/gate_qcif.yuv -w 176 -h 144 -z 20 -n 3 -m 2 -o /dev/shm/gate_s
DBG=1 step live 120 python -u tests/picture_on_device.py
DBG=1 step bench ${BENCH_TIMEOUT:-300} python -u bench.py "$@"
