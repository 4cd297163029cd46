# the GPU suite several times on one box: tolerance tests against order-free atomics must not be flaky
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -1; done | tee gpurun_out/repeat_tests.log
