#!/bin/bash
# First GPU calls of a next round.  One gpurun call each; outputs land in gpurun_out/, copy what matters to profiles/.
# Always gate long runs behind tools/gpu_mini.sh (fixture -> three streams -> live pictures -> a small bench, stops at the first failure):
# a hang costs the whole time limit of the call.
set -e
G=/usr/local/graft/bin/gpurun

# 1. the gate + a small bench with the scheduler's trace
$G --timeout 700 -- 'BENCH_TIMEOUT=240 tools/gpu_mini.sh gate --streams 4 --frames 17 --steps 1 --warmup 0'

# 2. the open question of round 2: 1 584 device pictures in one context (24 x 33, or two sets of 12 x 33) did not finish.  Start with
#    the scheduler trace of the device-resident arm only, 24 streams x 17 pictures, and watch launch rates
$G --timeout 500 -- 'XB200_SCHED_DEBUG=1 timeout 400 python -u bench.py --streams 24 --frames 17 --steps 1 --warmup 0 > gpurun_out/s24.json 2> gpurun_out/s24.err; tail -5 gpurun_out/s24.err'

# 3. racecheck of the worker grid on the fixture (the 256-thread RDOQ hazard of round 1 is gone with 128-thread teams, but nothing proves it)
$G --timeout 600 -- 'XB200_CHAIN_WORKERS=4 timeout 500 compute-sanitizer --tool racecheck python tests/picture_on_device.py --fixture-only > gpurun_out/racecheck.log 2>&1; tail -20 gpurun_out/racecheck.log'

# 4. the whole GPU suite and the default bench, N = 1, 2, 4, 8
$G --timeout 600 -- 'python -m pytest tests -q -m gpu 2>&1 | tail -5; python bench.py > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json'
