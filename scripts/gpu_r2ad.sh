# round 2, call ad: flakiness check -- the GPU suite five times in a row (parameter gradients / column sums are fp32 atomics)
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 600 python -m pytest tests -m gpu -q --timeout 300 --tb=line -p no:cacheprovider 2>&1 | grep -E "passed|failed|^FAILED|Error" | cut -c1-300 | tail -6
done 2>&1 | tee gpurun_out/r2ad_flaky.log
