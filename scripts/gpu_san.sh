mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool initcheck --track-unused-memory no --print-limit 20 python -m pytest tests/test_gpu_backward.py -q --timeout 800 -k "True-ms3 or True-fewshot" 2>&1 > gpurun_out/san_init.log
grep -E "Uninitialized|ERROR SUMMARY|at 0x| in |passed|failed|AssertionError" gpurun_out/san_init.log | head -40 | cut -c1-260
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_backward.py -q --timeout 800 -k "True-ms3" 2>&1 > gpurun_out/san_race.log
grep -E "hazard|ERROR SUMMARY|RACECHECK SUMMARY| in |passed|failed|AssertionError" gpurun_out/san_race.log | head -40 | cut -c1-260
